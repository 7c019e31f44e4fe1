/*
 * vr_capi.cu -- implementation of the C ABI declared in include/vr_caster.h.
 *
 * The context object plays the role of the reference's CLCaster instance
 * (include/CLCaster.h:93-329): it owns the device buffers the reference keeps in `buffer_map`,
 * the settings buffer, and the retained (aliased) camera / light pointers, and dispatches one
 * kernel per frame on a CUDA stream instead of an OpenCL command queue.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/vr_caster.h"
#include "vr_build.h"
#include "vr_kernels.h"
#include "vr_octree.h"
#include "vr_types.h"
#include "vr_ctx.h"

#define VR_SETTINGS_BUFFER_SIZE 64          /* ref include/CLCaster.h:303 */
#define VR_FILL_RGBA 0x64FFFFFFu            /* (255,255,255,100), ref src/CLCaster.cpp:280-286 */

namespace {

int fail(vr_ctx *c, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    fprintf(stderr, "[vrcaster] ERROR: %s\n", buf);      /* ref Logger::log(..., ERROR) */
    return 0;
}

#define VR_CUDA(c, call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) return fail((c), "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

/* Host -> device copy that the kernels on c->stream are ordered after: the stream is created non-blocking, so a plain
 * cudaMemcpy from pageable memory (which may return once the data is staged, before the DMA has landed) would not be. */
cudaError_t upload(vr_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!bytes) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e;
}

int local_rows_padded(const vr_ctx *c) {
    if (c->height <= 0) return 0;
    const int nb = (c->height + c->band_rows - 1) / c->band_rows;
    int mine = 0;
    for (int b = c->band_first; b < nb; b += c->band_stride) mine++;
    return mine * c->band_rows;
}

/* rows actually owned (last band may be partial) */
int local_rows_exact(const vr_ctx *c) {
    if (c->height <= 0) return 0;
    const int nb = (c->height + c->band_rows - 1) / c->band_rows;
    int rows = 0;
    for (int b = c->band_first; b < nb; b += c->band_stride) {
        const int r0 = b * c->band_rows;
        rows += (c->height - r0 < c->band_rows) ? c->height - r0 : c->band_rows;
    }
    return rows;
}

/* CLCaster::create_viewport ray table, ref src/CLCaster.cpp:244-275 (own restatement) */
void make_ray_table(int w, int h, std::vector<float> &out) {
    out.assign((size_t)4 * w * h, 0.0f);
    const double s = sin(1.57), c = cos(1.57);
    const int hw = w / 2, hh = h / 2;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < 2 * hh; j++) {
        const float py = (float)(j - hh);
        for (int i = 0; i < 2 * hw; i++) {
            const float px = (float)(i - hw);
            const float bx = -800.0f;
            const float nx = (float)((double)py * s + (double)bx * c);
            const float ny = px;
            const float nz = (float)((double)py * c - (double)bx * s);
            const float len = sqrtf(nx * nx + ny * ny + nz * nz);
            float *o = &out[4 * ((size_t)i + (size_t)w * (size_t)j)];
            o[0] = nx / len;
            o[1] = ny / len;
            o[2] = nz / len;
            o[3] = 0.0f;
        }
    }
}

int apply_l2_window(vr_ctx *c);

void free_tree(vr_ctx *c) {
    if (c->d_nodes) cudaFree(c->d_nodes);
    if (c->d_leaf_types) cudaFree(c->d_leaf_types);
    if (c->d_grid) cudaFree(c->d_grid);
    c->d_nodes = nullptr;
    c->d_leaf_types = nullptr;
    c->d_grid = nullptr;
    c->grid_tried = false;
    c->l2_base = nullptr;
    if (c->l2_persist && c->stream) {               /* the access-policy window pointed into the arrays just freed */
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    c->tree_valid = false;
    c->n_nodes = c->n_leaf_types = c->solid_voxels = 0;
}

int upload_tree(vr_ctx *c, vr_native_tree &t, bool from_map) {
    if (c->collapse_solid) vr_native_collapse_solid(t);      /* (a no-op for a tree that holds solid nodes already) */
    free_tree(c);
    VR_CUDA(c, cudaMalloc(&c->d_nodes, t.nodes.size() * sizeof(vr_node)));
    VR_CUDA(c, cudaMalloc(&c->d_leaf_types, t.leaf_types.size()));
    VR_CUDA(c, upload(c, c->d_nodes, t.nodes.data(), t.nodes.size() * sizeof(vr_node)));
    VR_CUDA(c, upload(c, c->d_leaf_types, t.leaf_types.data(), t.leaf_types.size()));
    c->levels = t.levels;
    c->tree_dim = t.dim;
    c->n_nodes = t.nodes.size();
    c->n_leaf_types = t.leaf_types.size();
    c->solid_voxels = t.solid_voxels;
    c->tree_valid = true;
    c->tree_from_map = from_map;
    return 1;
}

bool setting_value(const vr_ctx *c, const char *define, int64_t *out) {
    auto it = c->defines.find(define);
    if (it == c->defines.end() || !c->settings) return false;
    char *end = nullptr;
    const long slot = strtol(it->second.c_str(), &end, 10);
    if (end == it->second.c_str() || slot < 0 || slot >= VR_SETTINGS_BUFFER_SIZE) return false;
    *out = c->settings[slot];
    return true;
}

/* (re)applies or clears the L2 access-policy window of option "l2_persist" for the CURRENT node array */
int apply_l2_window(vr_ctx *c) {
    cudaSetDevice(c->device);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (c->l2_persist && c->d_nodes) {
        int max_win = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
        const bool grid = c->l2_base && c->l2_base == (const void *)c->d_grid;
        size_t bytes = grid ? ((size_t)(c->grid_is_directed ? 32 : 4) << (3 * c->grid_bits)) : c->n_nodes * sizeof(vr_node);
        if (max_win > 0 && bytes > (size_t)max_win) bytes = (size_t)max_win;      /* BFS order: top levels first */
        VR_CUDA(c, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes));
        attr.accessPolicyWindow.base_ptr = grid ? (void *)c->d_grid : (void *)c->d_nodes;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    VR_CUDA(c, cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (!c->l2_persist) cudaCtxResetPersistingL2Cache();
    return 1;
}

int ensure_tree(vr_ctx *c) {
    if (c->tree_valid) return 1;
    if (!c->has_octree) return fail(c, "octree traversal requested but neither a cubic map nor an octree is assigned");
    int64_t octdim = 0;
    if (!setting_value(c, "OCTDIM", &octdim) || octdim < 2)
        return fail(c, "octree assigned but the OCTDIM setting is missing");
    vr_native_tree t;
    try {                                                   /* (no exception may cross the C ABI: a huge OCTDIM means bad_alloc) */
        if (!vr_native_from_ref(c->oct_desc.data(), c->oct_desc.size(), c->oct_root, (int)octdim, nullptr, t))
            return fail(c, "malformed octree descriptor buffer");
    } catch (const std::exception &e) {
        return fail(c, "octree import failed: %s", e.what());
    }
    return upload_tree(c, t, false);
}

/* top grid of the closed-form walk, built once per tree; leaves d_grid null when the tree is too shallow for one */
int ensure_grid(vr_ctx *c) {
    if (c->d_grid || c->grid_tried) return 1;
    c->grid_tried = true;
    cudaError_t e = vr_build_grid_device(c->d_nodes, c->levels, c->tree_dim, c->grid_directed, c->stream, &c->d_grid, &c->grid_shift,
                                         &c->grid_bits, &c->launches);
    c->grid_is_directed = c->grid_directed;
    if (e == cudaErrorMemoryAllocation && c->grid_directed) {
        /* the eight directed tables (32 bytes per block) do not fit: the undirected table is an eighth of that */
        cudaGetLastError();
        e = vr_build_grid_device(c->d_nodes, c->levels, c->tree_dim, false, c->stream, &c->d_grid, &c->grid_shift, &c->grid_bits, &c->launches);
        c->grid_is_directed = false;
    }
    if (e != cudaSuccess && e != cudaErrorInvalidValue) return fail(c, "top grid build failed: %s", cudaGetErrorString(e));
    return 1;
}

/* Builds the launch parameters from the retained pointers and the settings buffer, i.e. what the
 * reference gets for free through CL_MEM_USE_HOST_PTR aliasing (ref src/CLCaster.cpp:137-139,322,1069). */
int build_params(vr_ctx *c, vr_frame_params &P, uint8_t *image, int *use_svo) {
    if (!c->d_ray_table || !c->d_image[0]) return fail(c, "compute: viewport not created");
    if (!c->cam_dir || !c->cam_pos) return fail(c, "compute: camera not assigned");
    if (!c->lights || c->light_count < 1) return fail(c, "compute: lights not assigned");
    if (!c->atlas_tex) return fail(c, "compute: texture atlas not created");

    int64_t v = 0;
    int svo;
    if (setting_value(c, "OCTENABLED", &v)) svo = (v == 0);          /* kernel:359: 0 selects the octree branch */
    else svo = c->d_map ? 0 : 1;
    if (svo) {
        if (!ensure_tree(c)) return 0;
    } else if (!c->d_map) {
        return fail(c, "compute: dense traversal selected (OCTENABLED != 0) but no map is assigned");
    }
    *use_svo = svo;

    memset(&P, 0, sizeof(P));
    P.width = c->width;
    P.height = c->height;
    P.band_rows = c->band_rows;
    P.band_stride = c->band_stride;
    P.band_first = c->band_first;
    P.local_rows = local_rows_padded(c);
    P.tile_world = c->tile_world;
    P.tile_rank = c->tile_rank;
    if (c->tile_world > 1 && (c->band_stride != 1 || c->band_first != 0))
        return fail(c, "compute: tile interleave (vr_set_tiles) and row bands (vr_set_bands) are mutually exclusive");
    P.ray_table = c->d_ray_table;
    P.image = image;
    P.aux = c->aux_on ? c->d_aux : nullptr;
    P.map = c->d_map;
    if (c->d_map) { P.dim[0] = c->dim[0]; P.dim[1] = c->dim[1]; P.dim[2] = c->dim[2]; }
    else P.dim[0] = P.dim[1] = P.dim[2] = c->tree_dim;
    for (int i = 0; i < 3; i++) P.cam_pos[i] = c->cam_pos[i];
    P.trig[0] = sinf(c->cam_dir[0]);
    P.trig[1] = cosf(c->cam_dir[0]);
    P.trig[2] = sinf(c->cam_dir[1]);
    P.trig[3] = cosf(c->cam_dir[1]);

    /* get_oct_vox on the camera voxel, kernel:342-354, evaluated once per frame */
    c->bias[0] = c->bias[1] = c->bias[2] = 0;
    if (c->has_octree) {
        int64_t octdim = 0, root = (int64_t)c->oct_root;
        setting_value(c, "OCTREE_ROOT_INDEX", &root);
        if (!setting_value(c, "OCTDIM", &octdim)) octdim = P.dim[0];
        const int pos[3] = {(int)floorf(c->cam_pos[0]), (int)floorf(c->cam_pos[1]), (int)floorf(c->cam_pos[2])};
        int sub[3], res = 0;
        vr_ref_octree_query(c->oct_desc.data(), c->oct_desc.size(), (uint64_t)root, (int)octdim, pos, sub, &res);
        for (int i = 0; i < 3; i++) c->bias[i] = ((sub[i] - pos[i]) * res) / 2;
    }
    for (int i = 0; i < 3; i++) P.bias[i] = (float)c->bias[i];
    P.cam_on_edge = vr_cam_on_edge(P.cam_pos, P.bias);

    /* the reference binds light_count but reads slot 0 only (host:193, kernel:660-670): LIGHT_COUNT defaults to 1 */
    P.light_count = 1;
    if (setting_value(c, "LIGHT_COUNT", &v) && v > 1)
        P.light_count = (int)(v < c->light_count ? v : c->light_count) < VR_MAX_LIGHTS ? (int)(v < c->light_count ? v : c->light_count) : VR_MAX_LIGHTS;
    for (int l = 0; l < P.light_count; l++) {
        for (int i = 0; i < 4; i++) P.light_rgbi[l][i] = c->lights[10 * l + i];
        for (int i = 0; i < 3; i++) P.light_pos[l][i] = c->lights[10 * l + 4 + i];
    }
    P.atlas_tex = (unsigned long long)c->atlas_tex;
    P.atlas = c->d_atlas;
    P.atlas_dim[0] = c->atlas_dim[0];
    P.atlas_dim[1] = c->atlas_dim[1];
    P.atlas_scale[0] = c->atlas_dim[0] / c->tile_dim[0];
    P.atlas_scale[1] = c->atlas_dim[1] / c->tile_dim[1];
    P.max_distance = 20;                                             /* kernel:326 */
    if (setting_value(c, "MAX_DISTANCE", &v)) P.max_distance = (int)v;
    P.max_bounces = 2;                                               /* kernel:357 */
    if (setting_value(c, "MAX_BOUNCES", &v) && v >= 0 && v <= 64) P.max_bounces = (int)v;
    P.nodes = c->d_nodes;
    P.leaf_types = c->d_leaf_types;
    P.levels = c->levels;
    P.root_shift = 2 * (c->levels - 1);
    P.grid = nullptr;
    P.grid_shift = P.grid_bits = P.grid_dim = P.grid_directed = 0;
    if (svo && (c->tree_dim != P.dim[0] || P.dim[0] != P.dim[1] || P.dim[0] != P.dim[2]))
        return fail(c, "octree traversal needs a cubic power-of-two map");
    return 1;
}

int launch_frame(vr_ctx *c, uint8_t *image, bool timed) {
    vr_frame_params P;
    int use_svo = 0;
    if (!build_params(c, P, image, &use_svo)) return 0;
    if (timed) VR_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    /* the per-axis walk sizes its add chains from a closed-form estimate whose error grows with ulp(t) (vr_count_before:
     * < 0.2 crossings up to 4096^3, where it is tested); beyond 16384^3 the merged walk is used whatever the option says */
    vr_launch_options opt = c->opt;
    if (P.dim[0] > 16384 && opt.walk == 1) opt.walk = 0;
    if (use_svo && opt.walk == 2) {
        /* the closed-form walk reads the top levels of the octree from a flat table (vr_canon.h), derived from the tree
         * in HBM when it is first needed; trees of a single level (maps up to 4^3) and maps beyond 65536^3 take walk 0 */
        if (!ensure_grid(c)) return 0;
        if (!c->d_grid || P.dim[0] > 65536) opt.walk = 0;
        P.grid = c->d_grid;
        P.grid_shift = c->grid_shift;
        P.grid_bits = c->grid_bits;
        P.grid_dim = 1 << c->grid_bits;
        P.grid_directed = c->grid_is_directed ? 1 : 0;
    }
    if (c->l2_persist && use_svo) {
        /* the window covers what the selected walk reads at random: the top grid (closed-form walk) or the node array */
        const void *want = (opt.walk == 2 && c->d_grid) ? (const void *)c->d_grid : (const void *)c->d_nodes;
        if (c->l2_base != want) {                                  /* first frame, or the tree / the walk changed since */
            c->l2_base = want;
            if (!apply_l2_window(c)) return 0;
        }
    }
    VR_CUDA(c, vr_launch_raycast(P, use_svo, c->aux_on ? 1 : 0, c->stream, &c->launches, &opt));
    if (timed) {
        VR_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
        c->timing_pending = true;
    }
    c->used_svo = use_svo;
    c->frames++;
    return 1;
}

int alloc_aux(vr_ctx *c) {
    if (c->d_aux) { cudaFree(c->d_aux); c->d_aux = nullptr; }
    if (c->aux_on && c->width > 0)
        VR_CUDA(c, cudaMalloc(&c->d_aux, sizeof(vr_aux) * (size_t)c->width * (size_t)c->height));
    return 1;
}

}  // namespace

/* internals shared with vr_mgpu.cu (vr_ctx.h) */
int vr_i_fail(vr_ctx *c, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return fail(c, "%s", buf);
}
int vr_i_launch_frame(vr_ctx *c, uint8_t *image, bool timed) { return launch_frame(c, image, timed); }
int vr_i_ensure_tree(vr_ctx *c) { return c->tree_valid ? 1 : ensure_tree(c); }
void vr_i_free_tree(vr_ctx *c) { free_tree(c); }
int vr_i_local_rows_padded(const vr_ctx *c) { return local_rows_padded(c); }


extern "C" {

const char *vr_version(void) { return "voxel-raycaster_b200 0.1 (sm_100a)"; }

const char *vr_last_error(const vr_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vr_init(vr_ctx **out, int device, unsigned flags) {
    if (!out) return 0;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, "no usable CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    vr_ctx *c = new vr_ctx();
    c->flags = flags;
    if (device < 0) {
        c->device = 0;
        vr_load_config(c, nullptr);
    } else {
        c->device = device;
    }
    if (c->device >= count) { fail(c, "device %d out of range (%d devices)", c->device, count); delete c; return 0; }
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev_start) != cudaSuccess || cudaEventCreate(&c->ev_stop) != cudaSuccess) {
        fail(c, "failed to create CUDA stream/events on device %d: %s", c->device, cudaGetErrorString(cudaGetLastError()));
        vr_destroy(c);                              /* releases whatever was created before the failure */
        return 0;
    }
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&c->ev_rendered[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming);
    }
    c->stream = c->own_stream;
    cudaDeviceGetAttribute(&c->opt.num_sms, cudaDevAttrMultiProcessorCount, c->device);
    if (cudaMalloc(&c->opt.counter, sizeof(unsigned int)) != cudaSuccess) { fail(c, "cudaMalloc failed"); vr_destroy(c); return 0; }
    if (!vr_create_settings_buffer(c)) { vr_destroy(c); return 0; }
    *out = c;
    return 1;
}

void vr_destroy(vr_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    vr_i_mgpu_destroy(c);
    if (c->gl_resource) { cudaGraphicsUnregisterResource(c->gl_resource); c->gl_resource = nullptr; }
    vr_release_viewport(c);
    vr_release_map(c);
    vr_release_octree(c);
    free_tree(c);
    if (c->atlas_tex) cudaDestroyTextureObject(c->atlas_tex);
    if (c->atlas_arr) cudaFreeArray(c->atlas_arr);
    if (c->d_atlas) cudaFree(c->d_atlas);
    if (c->opt.counter) cudaFree(c->opt.counter);
    for (int i = 0; i < 2; i++) {
        if (c->ev_rendered[i]) cudaEventDestroy(c->ev_rendered[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete[] c->settings;
    delete c;
}

int vr_load_config(vr_ctx *c, const char *path) {
    if (!c) return 0;
    FILE *f = fopen(path ? path : "device_config.bin", "rb");
    if (!f) return 0;                                     /* ref :500-503: no file -> false, caller picks */
    int ordinal = -1;
    char name[256] = {0};
    const bool ok = fread(&ordinal, sizeof(ordinal), 1, f) == 1 && fread(name, 1, sizeof(name), f) == sizeof(name);
    fclose(f);
    if (!ok || ordinal < 0) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ordinal) != cudaSuccess) return 0;
    if (strncmp(prop.name, name, sizeof(name)) != 0) return 0;     /* ref :527: saved device must still match */
    c->device = ordinal;
    return 1;
}

int vr_save_config(vr_ctx *c, const char *path) {
    if (!c) return 0;
    cudaDeviceProp prop;
    VR_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    FILE *f = fopen(path ? path : "device_config.bin", "wb");
    if (!f) return fail(c, "cannot write device config");
    char name[256] = {0};
    snprintf(name, sizeof(name), "%s", prop.name);
    fwrite(&c->device, sizeof(c->device), 1, f);
    fwrite(name, 1, sizeof(name), f);
    fclose(f);
    return 1;
}

/* ---- viewport ---- */
int vr_create_viewport(vr_ctx *c, int width, int height, float v_fov, float h_fov) {
    (void)v_fov; (void)h_fov;                              /* ignored by the reference too (host:249) */
    if (!c) return 0;
    if (width <= 0 || height <= 0) return fail(c, "create_viewport: bad size %dx%d", width, height);
    cudaSetDevice(c->device);
    vr_release_viewport(c);
    c->width = width;
    c->height = height;
    std::vector<float> table;
    try {                                                   /* (no exception may cross the C ABI) */
        make_ray_table(width, height, table);
    } catch (const std::exception &e) {
        c->width = c->height = 0;
        return fail(c, "create_viewport: %dx%d ray table: %s", width, height, e.what());
    }
    const size_t tbytes = table.size() * sizeof(float), ibytes = (size_t)width * height * 4;
    VR_CUDA(c, cudaMalloc(&c->d_ray_table, tbytes));
    VR_CUDA(c, upload(c, c->d_ray_table, table.data(), tbytes));
    for (int i = 0; i < 2; i++) {
        VR_CUDA(c, cudaMalloc(&c->d_image[i], ibytes));
        VR_CUDA(c, cudaMallocHost(&c->h_image[i], ibytes));
        VR_CUDA(c, vr_launch_fill(reinterpret_cast<uint32_t *>(c->d_image[i]), ibytes / 4, VR_FILL_RGBA, c->stream, &c->launches));
    }
    VR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->cur_image = 0;
    c->frame_issued = c->frame_retired = 0;
    return alloc_aux(c);
}

int vr_release_viewport(vr_ctx *c) {
    if (!c) return 0;
    const bool had = c->d_ray_table != nullptr;
    if (c->d_ray_table) cudaFree(c->d_ray_table);
    c->d_ray_table = nullptr;
    for (int i = 0; i < 2; i++) {
        if (c->d_image[i]) cudaFree(c->d_image[i]);
        if (c->h_image[i]) cudaFreeHost(c->h_image[i]);
        c->d_image[i] = nullptr;
        c->h_image[i] = nullptr;
    }
    if (c->d_aux) cudaFree(c->d_aux);
    c->d_aux = nullptr;
    c->width = c->height = 0;
    return had ? 1 : 0;                                    /* ref release_buffer: false if absent */
}

/* ---- lights / camera: retained pointers ---- */
int vr_assign_lights(vr_ctx *c, const float *packed, int count) {
    if (!c) return 0;
    if (!packed || count < 1) return fail(c, "assign_lights: empty light array");
    c->lights = packed;
    c->light_count = count;
    return 1;
}

int vr_assign_camera(vr_ctx *c, const float *direction2, const float *position3) {
    if (!c) return 0;
    if (!direction2 || !position3) return fail(c, "assign_camera: null pointer");
    c->cam_dir = direction2;
    c->cam_pos = position3;
    return 1;
}

int vr_release_camera(vr_ctx *c) {
    if (!c) return 0;
    const bool had = c->cam_dir != nullptr;
    c->cam_dir = c->cam_pos = nullptr;
    return had ? 1 : 0;
}

/* ---- map ---- */
int vr_assign_map(vr_ctx *c, const int8_t *voxels, int nx, int ny, int nz) {
    if (!c) return 0;
    if (!voxels || nx <= 0 || ny <= 0 || nz <= 0) return fail(c, "assign_map: bad arguments");
    cudaSetDevice(c->device);
    if (c->d_map) vr_release_map(c);                       /* ref store_buffer: silently replaces (host:859-863) */
    const size_t bytes = (size_t)nx * ny * nz;
    VR_CUDA(c, cudaMalloc(&c->d_map, bytes));
    VR_CUDA(c, upload(c, c->d_map, voxels, bytes));
    c->dim[0] = nx; c->dim[1] = ny; c->dim[2] = nz;
    /* traversal structure for the octree branch, built from the same voxels (types included) */
    if (nx == ny && ny == nz && (nx & (nx - 1)) == 0 && c->gpu_build && nx >= 4) {
        /* on the device, from the copy just uploaded (vr_build.cu) */
        vr_device_tree dt;
        memset(&dt, 0, sizeof(dt));
        cudaError_t be = vr_build_tree_device(c->d_map, nx, c->stream, &dt, &c->launches);
        if (be == cudaSuccess && c->collapse_solid) {
            be = vr_collapse_solid_device(&dt, c->stream, &c->launches);
            if (be != cudaSuccess) { cudaFree(dt.nodes); cudaFree(dt.types); }
        }
        if (be != cudaSuccess) {
            /* e.g. out of memory for the builder's workspace: the host builder produces the same arrays */
            cudaGetLastError();
            fprintf(stderr, "[vrcaster] device octree build failed (%s): building on the host\n", cudaGetErrorString(be));
            vr_native_tree t;
            try {
                if (!vr_native_from_dense(voxels, nx, t)) { vr_release_map(c); return fail(c, "assign_map: 64-tree build failed"); }
            } catch (const std::exception &e) {
                vr_release_map(c);
                return fail(c, "assign_map: 64-tree build failed: %s", e.what());
            }
            return upload_tree(c, t, true);
        }
        free_tree(c);
        c->d_nodes = dt.nodes;
        c->d_leaf_types = dt.types;
        c->levels = dt.levels;
        c->tree_dim = nx;
        c->n_nodes = dt.n_nodes;
        c->n_leaf_types = dt.n_types;
        c->solid_voxels = dt.solid_voxels;
        c->tree_valid = true;
        c->tree_from_map = true;
        c->build_ms = dt.total_ms;
        c->build_masks_ms = dt.masks_ms;
    } else if (nx == ny && ny == nz && (nx & (nx - 1)) == 0) {
        vr_native_tree t;
        try {
            if (!vr_native_from_dense(voxels, nx, t)) return fail(c, "assign_map: 64-tree build failed");
        } catch (const std::exception &e) {
            return fail(c, "assign_map: 64-tree build failed: %s", e.what());
        }
        if (!upload_tree(c, t, true)) return 0;
    } else if (c->tree_from_map) {
        free_tree(c);
    }
    return 1;
}

int vr_assign_columns(vr_ctx *c, const int32_t *lo, const int32_t *hi, int dim, int type) {
    if (!c) return 0;
    if (!lo || !hi || dim < 1 || (dim & (dim - 1)) || (type != 5 && type != 6)) return fail(c, "assign_columns: bad arguments");
    cudaSetDevice(c->device);
    if (c->d_map) vr_release_map(c);
    c->build_ms = c->build_masks_ms = 0.f;
    if (c->gpu_build && dim >= 4 && dim <= 4096) {
        /* on the device, from the two column tables (vr_build.cu: only the bricks that hold a voxel are materialised) */
        const size_t bytes = (size_t)dim * dim * sizeof(int32_t);
        int32_t *d_lo = nullptr, *d_hi = nullptr;
        cudaError_t e = cudaMalloc(&d_lo, bytes);
        if (e == cudaSuccess) e = cudaMalloc(&d_hi, bytes);
        if (e == cudaSuccess) e = upload(c, d_lo, lo, bytes);
        if (e == cudaSuccess) e = upload(c, d_hi, hi, bytes);
        vr_device_tree dt;
        memset(&dt, 0, sizeof(dt));
        if (e == cudaSuccess) e = vr_build_tree_columns_device(d_lo, d_hi, dim, (uint8_t)type, c->stream, &dt, &c->launches);
        if (e == cudaSuccess && c->collapse_solid) {
            e = vr_collapse_solid_device(&dt, c->stream, &c->launches);
            if (e != cudaSuccess) { cudaFree(dt.nodes); cudaFree(dt.types); }
        }
        cudaFree(d_lo);
        cudaFree(d_hi);
        if (e == cudaSuccess) {
            free_tree(c);
            c->d_nodes = dt.nodes;
            c->d_leaf_types = dt.types;
            c->levels = dt.levels;
            c->tree_dim = dim;
            c->n_nodes = dt.n_nodes;
            c->n_leaf_types = dt.n_types;
            c->solid_voxels = dt.solid_voxels;
            c->tree_valid = true;
            c->tree_from_map = true;
            c->build_ms = dt.total_ms;
            return 1;
        }
        cudaGetLastError();
        fprintf(stderr, "[vrcaster] device octree build from columns failed (%s): building on the host\n", cudaGetErrorString(e));
    }
    vr_native_tree t;
    try {
        if (!vr_native_from_columns(lo, hi, dim, (uint8_t)type, t)) return fail(c, "assign_columns: 64-tree build failed");
    } catch (const std::exception &e) {
        return fail(c, "assign_columns: 64-tree build failed: %s", e.what());
    }
    return upload_tree(c, t, true);
}

int vr_release_map(vr_ctx *c) {
    if (!c) return 0;
    const bool had = c->d_map != nullptr;
    if (c->d_map) cudaFree(c->d_map);
    c->d_map = nullptr;
    c->dim[0] = c->dim[1] = c->dim[2] = 0;
    if (c->tree_from_map) free_tree(c);
    return had ? 1 : 0;
}

/* ---- octree ---- */
int vr_assign_octree(vr_ctx *c, const uint64_t *descriptors, const uint32_t *attach_lookup, const uint64_t *attach,
                     uint64_t entries, uint64_t root_index) {
    (void)attach_lookup; (void)attach;                     /* all zero in the reference (Octree.cpp:8-9) */
    if (!c) return 0;
    if (!descriptors || !entries || root_index >= entries) return fail(c, "assign_octree: bad arguments");
    try {                                                   /* (no exception may cross the C ABI) */
        c->oct_desc.assign(descriptors, descriptors + entries);
    } catch (const std::exception &e) {
        c->oct_desc.clear();
        c->has_octree = false;
        return fail(c, "assign_octree: %llu entries: %s", (unsigned long long)entries, e.what());
    }
    c->oct_root = root_index;
    c->has_octree = true;
    if (!c->tree_from_map) free_tree(c);                   /* re-import lazily at validate/compute */
    /* ref host:113: registers the root index as a setting (fails quietly when already present) */
    if (!c->defines.count("OCTREE_ROOT_INDEX")) vr_add_to_settings_buffer(c, "octree_root_index", "OCTREE_ROOT_INDEX", (int64_t)root_index);
    else vr_overwrite_setting(c, "octree_root_index", (const int64_t *)&root_index);
    return 1;
}

int vr_release_octree(vr_ctx *c) {
    if (!c) return 0;
    const bool had = c->has_octree;
    c->oct_desc.clear();
    c->oct_desc.shrink_to_fit();
    c->has_octree = false;
    if (!c->tree_from_map) free_tree(c);
    return had ? 1 : 0;
}

/* ---- atlas ---- */
int vr_create_texture_atlas(vr_ctx *c, const uint8_t *rgba, int width, int height, int tile_w, int tile_h) {
    if (!c) return 0;
    if (!rgba || width <= 0 || height <= 0 || tile_w <= 0 || tile_h <= 0) return fail(c, "create_texture_atlas: bad arguments");
    cudaSetDevice(c->device);
    if (c->atlas_tex) { cudaDestroyTextureObject(c->atlas_tex); c->atlas_tex = 0; }
    if (c->atlas_arr) { cudaFreeArray(c->atlas_arr); c->atlas_arr = nullptr; }
    if (c->d_atlas) { cudaFree(c->d_atlas); c->d_atlas = nullptr; }
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    VR_CUDA(c, cudaMallocArray(&c->atlas_arr, &fmt, width, height));
    VR_CUDA(c, cudaMemcpy2DToArray(c->atlas_arr, 0, 0, rgba, (size_t)width * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice));
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = c->atlas_arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;                   /* sampler-less read_imagef: nearest, kernel:652 */
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    VR_CUDA(c, cudaCreateTextureObject(&c->atlas_tex, &res, &td, nullptr));
    VR_CUDA(c, cudaMalloc(&c->d_atlas, (size_t)width * height * 4));
    VR_CUDA(c, upload(c, c->d_atlas, rgba, (size_t)width * height * 4));
    c->atlas_dim[0] = width; c->atlas_dim[1] = height;
    c->tile_dim[0] = tile_w; c->tile_dim[1] = tile_h;
    return 1;
}

/* ---- settings ---- */
int vr_create_settings_buffer(vr_ctx *c) {
    if (!c) return 0;
    delete[] c->settings;
    c->settings = new int64_t[VR_SETTINGS_BUFFER_SIZE]();
    c->settings_pos = 0;
    /* the defines registered through add_to_settings_buffer name slots of the buffer that is gone: with them left in
     * place OCTENABLED would read 0 from the zeroed buffer (= the octree branch) and could not be added again */
    for (auto &kv : c->settings_indices) c->defines.erase(c->setting_define[kv.first]);
    c->setting_define.clear();
    c->settings_indices.clear();
    return 1;
}

int vr_release_settings_buffer(vr_ctx *c) {
    if (!c || !c->settings) return 0;
    delete[] c->settings;
    c->settings = nullptr;
    for (auto &kv : c->settings_indices) c->defines.erase(c->setting_define[kv.first]);
    c->setting_define.clear();
    c->settings_indices.clear();
    c->settings_pos = 0;
    return 1;
}

int vr_add_to_settings_buffer(vr_ctx *c, const char *setting_name, const char *define_name, int64_t value) {
    if (!c || !setting_name || !define_name) return 0;
    if (!c->settings) return fail(c, "Trying to push settings to an uninitialized settings buffer");
    if (c->defines.count(define_name)) return fail(c, "Define name already present in the defines map");
    if (c->settings_pos >= VR_SETTINGS_BUFFER_SIZE)
        return fail(c, "Settings buffer has reached the maximum size of %d elements", VR_SETTINGS_BUFFER_SIZE);
    c->defines[define_name] = std::to_string(c->settings_pos);
    c->settings[c->settings_pos] = value;
    c->settings_indices[setting_name] = c->settings_pos;
    c->setting_define[setting_name] = define_name;
    c->settings_pos++;
    return 1;
}

int vr_overwrite_setting(vr_ctx *c, const char *setting_name, const int64_t *value) {
    if (!c || !setting_name || !value) return 0;
    if (!c->settings) return fail(c, "Trying to push settings to an uninitialized settings buffer");
    auto it = c->settings_indices.find(setting_name);
    if (it == c->settings_indices.end()) return fail(c, "No setting matching [%s]", setting_name);
    c->settings[it->second] = *value;
    return 1;
}

int vr_remove_from_settings_buffer(vr_ctx *c, const char *setting_name) {
    (void)setting_name;
    if (c) c->err = "remove_from_settings_buffer() not implemented";      /* ref host:1074-1078 */
    return 0;
}

int64_t *vr_settings_data(vr_ctx *c) { return c ? c->settings : nullptr; }

int vr_set_define(vr_ctx *c, const char *name, const char *value) {
    if (!c || !name || !value) return 0;
    c->defines[name] = value;
    return 1;
}

int vr_remove_define(vr_ctx *c, const char *name) {
    if (!c || !name) return 0;
    c->defines.erase(name);
    return 1;
}

/* ---- per frame ---- */
int vr_validate(vr_ctx *c) {
    if (!c) return 0;
    if (!c->cam_dir) return fail(c, "Raycaster.validate() failed, camera not initialized");
    if (!c->d_map && !c->has_octree && !c->tree_valid) return fail(c, "Raycaster.validate() failed, map not initialized");
    if (!c->d_image[0]) return fail(c, "Raycaster.validate() failed, viewport_image not initialized");
    if (!c->d_ray_table) return fail(c, "Raycaster.validate() failed, viewport_matrix not initialized");
    cudaSetDevice(c->device);
    int64_t v = 0;
    const bool svo = setting_value(c, "OCTENABLED", &v) ? (v == 0) : (c->d_map == nullptr);
    if (svo && !ensure_tree(c)) return 0;
    return 1;
}

int vr_debug_quick_recompile(vr_ctx *c) { return vr_validate(c); }

int vr_compute_async(vr_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    return launch_frame(c, c->d_image[c->cur_image], true);
}

int vr_sync(vr_ctx *c) {
    if (!c) return 0;
    VR_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->timing_pending) {
        cudaEventElapsedTime(&c->last_kernel_ms, c->ev_start, c->ev_stop);
        c->timing_pending = false;
    }
    return 1;
}

int vr_compute(vr_ctx *c) { return vr_compute_async(c) && vr_sync(c); }

int vr_compute_into(vr_ctx *c, void *device_rgba) {
    if (!c || !device_rgba) return 0;
    cudaSetDevice(c->device);
    return launch_frame(c, static_cast<uint8_t *>(device_rgba), false);
}

int vr_compute_views(vr_ctx *c, const float *cameras, int count, void *device_rgba) {
    if (!c || !cameras || count < 1 || !device_rgba) return 0;
    cudaSetDevice(c->device);
    const float *keep_dir = c->cam_dir, *keep_pos = c->cam_pos;
    const size_t frame_bytes = (size_t)c->width * 4 * (size_t)local_rows_padded(c);
    int ok = 1;
    for (int v = 0; v < count && ok; v++) {
        c->cam_dir = cameras + 5 * (size_t)v;                  /* {inclination, azimuth, x, y, z} */
        c->cam_pos = cameras + 5 * (size_t)v + 2;
        ok = launch_frame(c, static_cast<uint8_t *>(device_rgba) + (size_t)v * frame_bytes, false);
    }
    c->cam_dir = keep_dir;
    c->cam_pos = keep_pos;
    return ok;
}

int vr_read_framebuffer(vr_ctx *c, uint8_t *rgba_out, size_t bytes) {
    if (!c || !rgba_out) return 0;
    if (!c->d_image[0]) return fail(c, "read_framebuffer: viewport not created");
    const size_t have = (size_t)c->width * 4 * (size_t)(c->band_stride > 1 ? local_rows_padded(c) : c->height);
    const size_t full = (size_t)c->width * 4 * (size_t)c->height;
    size_t n = bytes < have ? bytes : have;
    if (n > full) n = full;
    VR_CUDA(c, cudaMemcpyAsync(rgba_out, c->d_image[c->cur_image], n, cudaMemcpyDeviceToHost, c->stream));
    return vr_sync(c);
}

/* ---- CUDA-GL interop for the viewer (cuda_gl_interop.h needs <GL/gl.h>, which this image lacks: the one entry point
 * used is declared here with GLuint / GLenum spelled out) */
extern "C" cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource **resource, unsigned int image, unsigned int target,
                                                   unsigned int flags);

int vr_gl_unregister(vr_ctx *c) {
    if (!c) return 0;
    if (c->gl_resource) {
        cudaSetDevice(c->device);
        cudaGraphicsUnregisterResource(c->gl_resource);
        c->gl_resource = nullptr;
    }
    return 1;
}

int vr_gl_register_texture(vr_ctx *c, uint32_t gl_texture, uint32_t gl_target) {
    if (!c) return 0;
    if (!c->d_image[0]) return fail(c, "gl_register_texture: viewport not created");
    cudaSetDevice(c->device);
    vr_gl_unregister(c);
    const cudaError_t e = cudaGraphicsGLRegisterImage(&c->gl_resource, gl_texture, gl_target, cudaGraphicsRegisterFlagsWriteDiscard);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c->gl_resource = nullptr;
        return fail(c, "gl_register_texture: %s (is an OpenGL context current on this thread, and is texture %u an RGBA8 texture of it?)",
                    cudaGetErrorString(e), gl_texture);
    }
    return 1;
}

int vr_gl_draw(vr_ctx *c) {
    if (!c) return 0;
    if (!c->gl_resource) return fail(c, "gl_draw: no texture registered (vr_gl_register_texture)");
    if (c->band_stride > 1 || c->tile_world > 1) return fail(c, "gl_draw: the context renders a part of the frame only");
    cudaSetDevice(c->device);
    VR_CUDA(c, cudaGraphicsMapResources(1, &c->gl_resource, c->stream));
    cudaArray_t arr = nullptr;
    cudaError_t e = cudaGraphicsSubResourceGetMappedArray(&arr, c->gl_resource, 0, 0);
    if (e == cudaSuccess)
        e = cudaMemcpy2DToArrayAsync(arr, 0, 0, c->d_image[c->cur_image], (size_t)c->width * 4, (size_t)c->width * 4, (size_t)c->height,
                                     cudaMemcpyDeviceToDevice, c->stream);
    const cudaError_t u = cudaGraphicsUnmapResources(1, &c->gl_resource, c->stream);      /* orders GL after the copy */
    if (e != cudaSuccess || u != cudaSuccess) return fail(c, "gl_draw: %s", cudaGetErrorString(e != cudaSuccess ? e : u));
    return 1;
}

int vr_frame_begin(vr_ctx *c) {
    if (!c) return 0;
    if (c->frame_issued - c->frame_retired >= 2) return fail(c, "frame_begin: two frames already in flight");
    cudaSetDevice(c->device);
    const int slot = (int)(c->frame_issued & 1);
    /* the image of two frames ago must have left the device before it is overwritten */
    if (c->frame_issued >= 2) VR_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));
    if (!launch_frame(c, c->d_image[slot], false)) return 0;
    VR_CUDA(c, cudaEventRecord(c->ev_rendered[slot], c->stream));
    VR_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rendered[slot], 0));
    const size_t bytes = (size_t)c->width * 4 * (size_t)(c->band_stride > 1 ? local_rows_padded(c) : c->height);
    VR_CUDA(c, cudaMemcpyAsync(c->h_image[slot], c->d_image[slot], bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    VR_CUDA(c, cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    c->cur_image = slot;
    c->frame_issued++;
    return 1;
}

int vr_frame_end(vr_ctx *c, const uint8_t **rgba) {
    if (!c) return 0;
    if (c->frame_issued == c->frame_retired) return fail(c, "frame_end: no frame in flight");
    const int slot = (int)(c->frame_retired & 1);
    VR_CUDA(c, cudaEventSynchronize(c->ev_copied[slot]));
    if (rgba) *rgba = c->h_image[slot];
    c->frame_retired++;
    return 1;
}

/* ---- extensions ---- */
int vr_set_bands(vr_ctx *c, int band_rows, int stride, int first) {
    if (!c) return 0;
    if (band_rows < 1 || stride < 1 || first < 0 || first >= stride) return fail(c, "set_bands: bad arguments");
    c->band_rows = band_rows;
    c->band_stride = stride;
    c->band_first = first;
    return 1;
}

int vr_set_tiles(vr_ctx *c, int world, int rank) {
    if (!c) return 0;
    if (world < 1 || rank < 0 || rank >= world) return fail(c, "set_tiles: bad arguments");
    c->tile_world = world;
    c->tile_rank = rank;
    return 1;
}

int vr_set_option(vr_ctx *c, const char *name, int64_t value) {
    if (!c || !name) return 0;
    const std::string n(name);
    if (n == "persistent") c->opt.persistent = value != 0;
    else if (n == "refill_min") c->opt.refill_min = value < 1 ? 1 : (value > 32 ? 32 : (int)value);
    else if (n == "ctas_per_sm") c->opt.ctas_per_sm = value < 1 ? 1 : (value > 8 ? 8 : (int)value);
    else if (n == "walk") c->opt.walk = (value == 1 || value == 2) ? (int)value : 0;
    else if (n == "gpu_build") c->gpu_build = value != 0;        /* 0: assign_map builds the 64-tree on the host */
    else if (n == "collapse_solid") c->collapse_solid = value != 0;   /* applies to the trees built from now on */
    else if (n == "directed_grid") {
        /* top grid of the closed-form walk: 1 = one table per direction octant of the ray (default), 0 = one undirected table */
        if (c->grid_directed != (value != 0)) {
            c->grid_directed = value != 0;
            if (c->d_grid) {
                cudaSetDevice(c->device);
                cudaStreamSynchronize(c->stream);
                cudaFree(c->d_grid);
                c->d_grid = nullptr;
                if (c->l2_base && c->l2_persist) c->l2_base = nullptr;
            }
            c->grid_tried = false;
        }
    }
    else if (n == "l2_persist") {
        /* pin the 64-tree nodes in L2 (cudaAccessPolicyWindow) for every kernel launched on the context stream */
        if (!c->d_nodes) return fail(c, "set_option l2_persist: no octree yet");
        c->l2_persist = value != 0;
        c->l2_base = nullptr;                       /* (re)applied at the next frame, for what that frame's walk reads */
        if (!value && !apply_l2_window(c)) return 0;
    }
    else return fail(c, "set_option: unknown option [%s]", name);
    return 1;
}

int vr_local_rows(const vr_ctx *c) { return c ? local_rows_exact(c) : 0; }

int vr_set_stream(vr_ctx *c, void *cuda_stream) {
    if (!c) return 0;
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return 1;
}

int vr_enable_aux(vr_ctx *c, int enable) {
    if (!c) return 0;
    c->aux_on = enable != 0;
    cudaSetDevice(c->device);
    return alloc_aux(c);
}

int vr_read_aux(vr_ctx *c, void *out, size_t bytes) {
    if (!c || !out) return 0;
    if (!c->d_aux) return fail(c, "read_aux: aux records are not enabled");
    const size_t have = sizeof(vr_aux) * (size_t)c->width * (size_t)c->height;
    VR_CUDA(c, cudaMemcpyAsync(out, c->d_aux, bytes < have ? bytes : have, cudaMemcpyDeviceToHost, c->stream));
    return vr_sync(c);
}

void *vr_device_image(vr_ctx *c) { return c ? c->d_image[c->cur_image] : nullptr; }

int vr_read_ray_table(vr_ctx *c, float *out, size_t bytes) {
    if (!c || !out) return 0;
    if (!c->d_ray_table) return fail(c, "read_ray_table: viewport not created");
    const size_t have = sizeof(float) * 4 * (size_t)c->width * (size_t)c->height;
    VR_CUDA(c, cudaMemcpy(out, c->d_ray_table, bytes < have ? bytes : have, cudaMemcpyDeviceToHost));
    return 1;
}

int vr_native_tree_info(vr_ctx *c, uint64_t *node_bytes, uint64_t *type_bytes, int32_t *levels, int32_t *dim) {
    if (!c) return 0;
    if (!c->tree_valid && !ensure_tree(c)) return 0;
    if (node_bytes) *node_bytes = c->n_nodes * sizeof(vr_node);
    if (type_bytes) *type_bytes = c->n_leaf_types;
    if (levels) *levels = c->levels;
    if (dim) *dim = c->tree_dim;
    return 1;
}

uint64_t vr_top_grid_read(vr_ctx *c, uint32_t *host_out, uint64_t capacity, int32_t *grid_shift, int32_t *grid_bits) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (!c->tree_valid && !ensure_tree(c)) return 0;
    if (!ensure_grid(c)) return 0;
    if (!c->d_grid) { fail(c, "top_grid_read: the octree is too shallow for a top grid"); return 0; }
    const uint64_t n = (c->grid_is_directed ? 8ull : 1ull) << (3 * c->grid_bits);     /* directed: the eight tables, octant 0 first */
    if (grid_shift) *grid_shift = c->grid_shift;
    if (grid_bits) *grid_bits = c->grid_bits;
    if (host_out && capacity >= n) {
        if (cudaMemcpy(host_out, c->d_grid, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
            fail(c, "top_grid_read: copy failed");
            return 0;
        }
    }
    return n;
}

int vr_native_tree_copy(vr_ctx *c, void *device_nodes, void *device_types) {
    if (!c || !device_nodes || !device_types) return 0;
    if (!c->tree_valid) return fail(c, "native_tree_copy: no 64-tree built");
    VR_CUDA(c, cudaMemcpyAsync(device_nodes, c->d_nodes, c->n_nodes * sizeof(vr_node), cudaMemcpyDeviceToDevice, c->stream));
    VR_CUDA(c, cudaMemcpyAsync(device_types, c->d_leaf_types, c->n_leaf_types, cudaMemcpyDeviceToDevice, c->stream));
    VR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 1;
}

int vr_assign_native_tree(vr_ctx *c, const void *device_nodes, uint64_t node_bytes, const void *device_types,
                          uint64_t type_bytes, int32_t levels, int32_t dim) {
    if (!c) return 0;
    if (!device_nodes || !device_types || node_bytes < sizeof(vr_node) || node_bytes % sizeof(vr_node) || !type_bytes ||
        levels < 1 || levels > VR_MAX_LEVELS || dim < 1 || (dim & (dim - 1)) || (1 << (2 * levels)) < dim)
        return fail(c, "assign_native_tree: bad arguments");
    cudaSetDevice(c->device);
    free_tree(c);
    VR_CUDA(c, cudaMalloc(&c->d_nodes, node_bytes));
    VR_CUDA(c, cudaMalloc(&c->d_leaf_types, type_bytes));
    VR_CUDA(c, cudaMemcpyAsync(c->d_nodes, device_nodes, node_bytes, cudaMemcpyDeviceToDevice, c->stream));
    VR_CUDA(c, cudaMemcpyAsync(c->d_leaf_types, device_types, type_bytes, cudaMemcpyDeviceToDevice, c->stream));
    VR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->levels = levels;
    c->tree_dim = dim;
    c->n_nodes = node_bytes / sizeof(vr_node);
    c->n_leaf_types = type_bytes;
    c->solid_voxels = type_bytes;
    c->tree_valid = true;
    c->tree_from_map = true;      /* not tied to an assigned reference octree */
    return 1;
}

int vr_device_alloc(vr_ctx *c, size_t bytes, void **device_ptr) {
    if (!c || !device_ptr || !bytes) return 0;
    cudaSetDevice(c->device);
    VR_CUDA(c, cudaMalloc(device_ptr, bytes));
    return 1;
}

int vr_device_free(vr_ctx *c, void *device_ptr) {
    if (!c || !device_ptr) return 0;
    VR_CUDA(c, cudaFree(device_ptr));
    return 1;
}

int vr_ipc_get_handle(vr_ctx *c, void *device_ptr, void *handle64) {
    if (!c || !device_ptr || !handle64) return 0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    VR_CUDA(c, cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle64, &h, sizeof(h));
    return 1;
}

int vr_ipc_open_handle(vr_ctx *c, const void *handle64, void **device_ptr) {
    if (!c || !handle64 || !device_ptr) return 0;
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    VR_CUDA(c, cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 1;
}

int vr_ipc_close_handle(vr_ctx *c, void *device_ptr) {
    if (!c || !device_ptr) return 0;
    VR_CUDA(c, cudaIpcCloseMemHandle(device_ptr));
    return 1;
}

int vr_push_bands(vr_ctx *c, const void *slab, void *frame, void *cuda_stream) {
    if (!c || !slab || !frame) return 0;
    if (c->width <= 0) return fail(c, "push_bands: viewport not created");
    cudaSetDevice(c->device);
    const size_t band_bytes = (size_t)c->band_rows * (size_t)c->width * 4;
    const int bands = local_rows_padded(c) / c->band_rows;
    if (!bands) return 1;
    /* band j of the slab -> frame band j * stride + first: one 2-D copy, pitch = stride bands */
    VR_CUDA(c, cudaMemcpy2DAsync(static_cast<uint8_t *>(frame) + (size_t)c->band_first * band_bytes,
                                 (size_t)c->band_stride * band_bytes, slab, band_bytes, band_bytes, (size_t)bands,
                                 cudaMemcpyDefault, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream));
    return 1;
}

int vr_host_register(vr_ctx *c, void *host_ptr, size_t bytes) {
    if (!c || !host_ptr || !bytes) return 0;
    cudaSetDevice(c->device);
    VR_CUDA(c, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return 1;
}

int vr_host_unregister(vr_ctx *c, void *host_ptr) {
    if (!c || !host_ptr) return 0;
    cudaSetDevice(c->device);
    VR_CUDA(c, cudaHostUnregister(host_ptr));
    return 1;
}

/* On-disk octree: "VR64" | u32 version | i32 dim | i32 levels | u64 nodes | u64 types | vr_node[] | u8[] */
int vr_octree_save(vr_ctx *c, const char *path) {
    if (!c || !path) return 0;
    if (!c->tree_valid && !ensure_tree(c)) return 0;
    cudaSetDevice(c->device);
    std::vector<vr_node> nodes;
    std::vector<uint8_t> types;
    try {                                                   /* (no exception may cross the C ABI) */
        nodes.resize(c->n_nodes);
        types.resize(c->n_leaf_types);
    } catch (const std::exception &e) {
        return fail(c, "octree_save: host copy of the tree: %s", e.what());
    }
    VR_CUDA(c, cudaStreamSynchronize(c->stream));           /* the tree may have just been built on the context's stream */
    VR_CUDA(c, cudaMemcpy(nodes.data(), c->d_nodes, nodes.size() * sizeof(vr_node), cudaMemcpyDeviceToHost));
    VR_CUDA(c, cudaMemcpy(types.data(), c->d_leaf_types, types.size(), cudaMemcpyDeviceToHost));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(c, "octree_save: cannot open %s", path);
    uint32_t version = 1;                                    /* 2: the tree holds solid nodes (VR_NODE_SOLID) */
    for (const vr_node &n : nodes)
        if (n.child_base & VR_NODE_SOLID) { version = 2; break; }
    const int32_t dim = c->tree_dim, levels = c->levels;
    const uint64_t nn = nodes.size(), nt = types.size();
    bool ok = fwrite("VR64", 1, 4, f) == 4 && fwrite(&version, 4, 1, f) == 1 && fwrite(&dim, 4, 1, f) == 1 &&
              fwrite(&levels, 4, 1, f) == 1 && fwrite(&nn, 8, 1, f) == 1 && fwrite(&nt, 8, 1, f) == 1 &&
              fwrite(nodes.data(), sizeof(vr_node), nn, f) == nn && fwrite(types.data(), 1, nt, f) == nt;
    fclose(f);
    return ok ? 1 : fail(c, "octree_save: short write to %s", path);
}

int vr_octree_load(vr_ctx *c, const char *path) {
    if (!c || !path) return 0;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(c, "octree_load: cannot open %s", path);
    char magic[4];
    uint32_t version = 0;
    int32_t dim = 0, levels = 0;
    uint64_t nn = 0, nt = 0;
    bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "VR64", 4) == 0 && fread(&version, 4, 1, f) == 1 && (version == 1 || version == 2) &&
              fread(&dim, 4, 1, f) == 1 && fread(&levels, 4, 1, f) == 1 && fread(&nn, 8, 1, f) == 1 && fread(&nt, 8, 1, f) == 1 &&
              nn >= 1 && nn < (1ull << 32) && nt >= 1 && nt < (1ull << 32) && levels >= 1 && levels <= VR_MAX_LEVELS && dim >= 1 &&
              !(dim & (dim - 1)) && (1ll << (2 * levels)) >= dim;
    /* the arrays must fit in what is left of the file BEFORE anything is allocated for them */
    if (ok) {
        const long at = ftell(f);
        ok = at >= 0 && fseek(f, 0, SEEK_END) == 0;
        const long end = ok ? ftell(f) : -1;
        ok = ok && end >= at && (uint64_t)(end - at) == nn * sizeof(vr_node) + nt && fseek(f, at, SEEK_SET) == 0;
    }
    vr_native_tree t;
    try {
        if (ok) {
            t.nodes.resize(nn);
            t.leaf_types.resize(nt);
            ok = fread(t.nodes.data(), sizeof(vr_node), nn, f) == nn && fread(t.leaf_types.data(), 1, nt, f) == nt;
        }
    } catch (const std::exception &) {
        ok = false;
    }
    fclose(f);
    if (!ok) return fail(c, "octree_load: %s is not a valid octree file", path);
    /* Level by level from the root (nodes are stored in BFS order, the children of a level form one contiguous run):
     * the child pointers of an inner level must stay inside the node array and land on the next level's run, those of
     * the leaf level inside the type array.  A corrupt or truncated file must fail here, not as an out-of-bounds load
     * on the device. */
    {
        uint64_t lo = 0, hi = 1;                                /* nodes [lo, hi) = the current level */
        for (int l = 0; l < levels && ok; l++) {
            const bool leaf = l == levels - 1;
            uint64_t next_lo = ~0ull, next_hi = 0;
            for (uint64_t i = lo; i < hi && ok; i++) {
                vr_node &n = t.nodes[i];
                const uint64_t m = (uint64_t)n.mask_lo | ((uint64_t)n.mask_hi << 32);
                const uint64_t pc = (uint64_t)__builtin_popcountll(m);
                n.aux = vr_node_planes(m);                      /* derived from the mask: never trusted from the file */
                if (n.child_base & VR_NODE_SOLID) {             /* a collapsed solid cube: full mask, a voxel type, no children */
                    const uint32_t ty = n.child_base & ~VR_NODE_SOLID;
                    ok = version == 2 && m == ~0ull && (ty == 5 || ty == 6);
                    continue;
                }
                if (!pc) { ok = nn == 1; continue; }            /* only the root of an all-empty map has no children */
                const uint64_t b = n.child_base;
                if (leaf) { ok = b + pc <= nt; continue; }
                ok = b >= hi && b + pc <= nn;
                next_lo = b < next_lo ? b : next_lo;
                next_hi = b + pc > next_hi ? b + pc : next_hi;
            }
            if (!leaf && ok) {
                if (next_hi == 0) { ok = hi == nn; break; }     /* no children at all: the tree ends here */
                ok = next_lo == hi;                             /* BFS order: the next level starts right after this one */
                lo = next_lo; hi = next_hi;
            } else if (leaf && ok) {
                ok = hi == nn;                                  /* nothing may follow the leaf level */
            }
        }
        if (!ok) return fail(c, "octree_load: corrupt child pointer in %s", path);
    }
    t.levels = levels;
    t.dim = dim;
    t.solid_voxels = nt;
    cudaSetDevice(c->device);
    return upload_tree(c, t, true);
}

int vr_get_stats(vr_ctx *c, vr_stats *out) {
    if (!c || !out) return 0;
    memset(out, 0, sizeof(*out));
    out->kernel_launches = c->launches;
    out->frames = c->frames;
    out->native_nodes = c->n_nodes;
    out->native_bytes = c->n_nodes * sizeof(vr_node) + c->n_leaf_types;
    out->solid_voxels = c->solid_voxels;
    out->levels = c->levels;
    out->used_svo = c->used_svo;
    for (int i = 0; i < 3; i++) out->bias[i] = c->bias[i];
    out->device = c->device;
    out->last_kernel_ms = c->last_kernel_ms;
    out->build_ms = c->build_ms;
    out->build_masks_ms = c->build_masks_ms;
    return 1;
}

int vr_octree_generate(const int8_t *voxels, int dim, uint64_t *out, uint64_t *entries, uint64_t *root_index) {
    if (!voxels || !entries) return 0;
    std::vector<uint64_t> buf;
    uint64_t root = 0;
    try {
        if (!vr_ref_octree_generate(voxels, dim, buf, &root)) return 0;
    } catch (const std::exception &) {
        return 0;
    }
    if (out) {
        if (*entries < buf.size()) { *entries = buf.size(); return 0; }
        memcpy(out, buf.data(), buf.size() * sizeof(uint64_t));
    }
    *entries = buf.size();
    if (root_index) *root_index = root;
    return 1;
}

int vr_octree_get_voxel(const uint64_t *descriptors, uint64_t entries, uint64_t root_index, int dim, const int32_t pos[3],
                        int32_t sub_oct_pos[3], int32_t *resolution) {
    int sub[3] = {0, 0, 0}, res = 0;
    const int p[3] = {pos[0], pos[1], pos[2]};
    const int found = vr_ref_octree_query(descriptors, entries, root_index, dim, p, sub, &res);
    if (sub_oct_pos) { sub_oct_pos[0] = sub[0]; sub_oct_pos[1] = sub[1]; sub_oct_pos[2] = sub[2]; }
    if (resolution) *resolution = res;
    return found;
}

}  // extern "C"
