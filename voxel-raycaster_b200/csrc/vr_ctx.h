/*
 * vr_ctx.h -- the context object behind the C ABI (internal; include/vr_caster.h is the public face).
 *
 * Plays the role of the reference's CLCaster instance (include/CLCaster.h:93-329): owns the device buffers the reference
 * keeps in `buffer_map`, the settings buffer and the retained (aliased) camera / light pointers.  Shared by vr_capi.cu
 * (single-GPU entry points) and vr_mgpu.cu (multi-GPU frame scheduler).
 */
#ifndef VR_CTX_H
#define VR_CTX_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "vr_kernels.h"
#include "vr_types.h"

struct vr_mgpu;                    /* vr_mgpu.cu */

struct vr_ctx {
    int device = 0;
    unsigned flags = 0;
    std::string err;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    bool timing_pending = false;
    float last_kernel_ms = 0.f;
    unsigned long long launches = 0, frames = 0;

    /* viewport */
    int width = 0, height = 0;
    float *d_ray_table = nullptr;
    uint8_t *d_image[2] = {nullptr, nullptr};
    uint8_t *h_image[2] = {nullptr, nullptr};
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    unsigned long long frame_issued = 0, frame_retired = 0;   /* pipelined frames */
    int cur_image = 0;
    vr_aux *d_aux = nullptr;
    bool aux_on = false;
    int band_rows = 1, band_stride = 1, band_first = 0;
    int tile_world = 1, tile_rank = 0;

    /* dense map */
    int8_t *d_map = nullptr;
    int dim[3] = {0, 0, 0};
    /* reference-format octree (host copy: the start bias is evaluated on the host) */
    std::vector<uint64_t> oct_desc;
    uint64_t oct_root = 0;
    bool has_octree = false;
    /* native 64-tree */
    vr_node *d_nodes = nullptr;
    uint8_t *d_leaf_types = nullptr;
    uint32_t *d_grid = nullptr;        /* top grid of the closed-form walk, built from d_nodes when first needed */
    int grid_shift = 0, grid_bits = 0;
    bool grid_tried = false;
    bool grid_directed = true;         /* option "directed_grid": eight per-octant tables (the default) or the undirected table */
    bool grid_is_directed = false;     /* what d_grid holds */
    bool l2_persist = false;           /* option "l2_persist": access-policy window over d_nodes, re-applied per tree */
    const void *l2_base = nullptr;
    int levels = 0, tree_dim = 0;
    uint64_t n_nodes = 0, n_leaf_types = 0, solid_voxels = 0;
    bool tree_valid = false, tree_from_map = false;
    bool collapse_solid = true;        /* option "collapse_solid": solid subtrees of the 64-tree become single nodes (VR_NODE_SOLID) */
    bool gpu_build = true;             /* assign_map builds the 64-tree with vr_build.cu (option "gpu_build") */
    float build_ms = 0.f, build_masks_ms = 0.f;

    /* retained host pointers (CL_MEM_USE_HOST_PTR semantics) */
    const float *cam_dir = nullptr, *cam_pos = nullptr;
    const float *lights = nullptr;
    int light_count = 0;

    /* atlas */
    cudaArray_t atlas_arr = nullptr;
    cudaTextureObject_t atlas_tex = 0;
    uint8_t *d_atlas = nullptr;
    int atlas_dim[2] = {0, 0}, tile_dim[2] = {0, 0};

    /* settings (ref include/CLCaster.h:303-311) */
    int64_t *settings = nullptr;
    unsigned settings_pos = 0;
    std::map<std::string, unsigned> settings_indices;
    std::map<std::string, std::string> setting_define;   /* setting name -> the define registered with it */
    std::map<std::string, std::string> defines;

    int used_svo = 0;
    int bias[3] = {0, 0, 0};
    vr_launch_options opt = {0, 8, 3, 148, nullptr, 2};      /* walk 2 = the closed-form walk is the default */
    struct cudaGraphicsResource *gl_resource = nullptr;   /* the viewer's texture (vr_gl_register_texture) */
    vr_mgpu *mgpu = nullptr;           /* multi-GPU frame scheduler state (vr_mgpu_init) */
};

/* vr_capi.cu internals the scheduler needs */
int vr_i_fail(vr_ctx *c, const char *fmt, ...);                             /* sets the last error, logs, returns 0 */
int vr_i_launch_frame(vr_ctx *c, uint8_t *image, bool timed);               /* one kernel launch on c->stream */
int vr_i_ensure_tree(vr_ctx *c);
void vr_i_free_tree(vr_ctx *c);
int vr_i_local_rows_padded(const vr_ctx *c);
void vr_i_mgpu_destroy(vr_ctx *c);                                          /* vr_mgpu.cu */

#endif
