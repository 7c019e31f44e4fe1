/*
 * vr_caster.h -- C ABI of the B200 voxel ray caster (libvrcaster.so).
 *
 * Drop-in boundary for the reference's `CLCaster` host API (MitchellHansen/voxel-raycaster,
 * include/CLCaster.h:110-179): one entry point per public CLCaster method, same call order, same
 * success convention.  It replaces the OpenCL context / cl_khr_gl_sharing path with hand-written
 * sm_100a CUDA kernels and adds a headless framebuffer for benchmarking.  Plain C types only.
 * "ref" below = /root/reference; the C++ facade with the reference's method names is
 * voxel-raycaster_b200/csrc/CUDACaster.hpp, the ctypes binding voxel-raycaster_b200/caster.py.
 *
 * Conventions (ref src/CLCaster.cpp: every method returns bool, true = success, errors are logged):
 *   every vr_* function returning int returns 1 on success and 0 on failure; vr_last_error()
 *   gives the message.  Not thread-safe, not re-entrant (same as the reference).  The library
 *   fails loudly (vr_init returns 0) when no CUDA device is usable: there is no CPU path.
 */
#ifndef VR_CASTER_H
#define VR_CASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vr_ctx vr_ctx;

/* flags for vr_init */
#define VR_INIT_HEADLESS 1u      /* no GL interop: the framebuffer lives in device memory + pinned host mirror */

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* CLCaster::init (ref include/CLCaster.h:110, src/CLCaster.cpp:14-73): pick the device, create the
 * execution context/queue (here: CUDA stream) and the settings buffer.  device < 0 => device from
 * vr_load_config() or CUDA device 0. */
int vr_init(vr_ctx **out, int device, unsigned flags);
/* CLCaster::~CLCaster (ref src/CLCaster.cpp:5-12) */
void vr_destroy(vr_ctx *ctx);
const char *vr_last_error(const vr_ctx *ctx);
const char *vr_version(void);

/* CLCaster::load_config / save_config (ref include/CLCaster.h:148-151, src/CLCaster.cpp:495-542):
 * remembers the chosen device in `device_config.bin` (here: the CUDA ordinal + name). */
int vr_load_config(vr_ctx *ctx, const char *path);
int vr_save_config(vr_ctx *ctx, const char *path);

/* ---- scene upload ------------------------------------------------------------------------------ */

/* CLCaster::create_viewport / release_viewport (ref include/CLCaster.h:114-115,
 * src/CLCaster.cpp:233-311): builds the per-pixel ray table (focal length 800 px, v_fov/h_fov are
 * accepted and ignored exactly like the reference) and the RGBA8 output image, initialised to
 * (255,255,255,100). */
int vr_create_viewport(vr_ctx *ctx, int width, int height, float v_fov, float h_fov);
int vr_release_viewport(vr_ctx *ctx);

/* CLCaster::assign_lights (ref include/CLCaster.h:119, src/CLCaster.cpp:313-328).  packed = count x
 * 10 floats {r,g,b,i, x,y,z, dx,dy,dz} (ref include/LightController.h:63-73).  The pointer is
 * RETAINED (CL_MEM_USE_HOST_PTR semantics): it is re-read at every vr_compute and must stay valid. */
int vr_assign_lights(vr_ctx *ctx, const float *packed, int count);

/* CLCaster::assign_map / release_map (ref include/CLCaster.h:123-124, src/CLCaster.cpp:76-100):
 * copies the dense char volume, index x + nx*(y + nz*z) (ref src/map/ArrayMap.cpp:39-46). */
int vr_assign_map(vr_ctx *ctx, const int8_t *voxels, int nx, int ny, int nz);
int vr_release_map(vr_ctx *ctx);

/* Extension: a map given as solid z-ranges per column -- lo/hi are dim*dim int32 arrays indexed
 * x + dim*y, column solid for lo <= z <= hi, every solid voxel has value `type` (5 or 6).  Builds the octree
 * without materialising the dim^3 volume (4096^3 would be 64 GiB); only the octree traversal is then available. */
int vr_assign_columns(vr_ctx *ctx, const int32_t *lo, const int32_t *hi, int dim, int type);

/* CLCaster::assign_octree / release_octree (ref include/CLCaster.h:127-128, src/CLCaster.cpp:102-131):
 * copies the reference-format child-descriptor buffer (ref include/map/Octree.h:89-94) and registers
 * the OCTREE_ROOT_INDEX setting.  attach_lookup / attach may be NULL (the reference uploads zeros). */
int vr_assign_octree(vr_ctx *ctx, const uint64_t *descriptors, const uint32_t *attach_lookup,
                     const uint64_t *attach, uint64_t entries, uint64_t root_index);
int vr_release_octree(vr_ctx *ctx);

/* CLCaster::assign_camera / release_camera (ref include/CLCaster.h:131-132, src/CLCaster.cpp:133-155):
 * direction = (inclination, azimuth), position = (x,y,z).  Both pointers are RETAINED and re-read
 * at every vr_compute (ref Camera::get_direction_pointer / get_position_pointer). */
int vr_assign_camera(vr_ctx *ctx, const float *direction2, const float *position3);
int vr_release_camera(vr_ctx *ctx);

/* CLCaster::create_texture_atlas (ref include/CLCaster.h:136, src/CLCaster.cpp:208-222): RGBA8 atlas,
 * copied into a CUDA array + texture object. */
int vr_create_texture_atlas(vr_ctx *ctx, const uint8_t *rgba, int width, int height, int tile_w, int tile_h);

/* ---- settings buffer (ref include/CLCaster.h:157-164, src/CLCaster.cpp:1029-1109) -------------- */
int vr_create_settings_buffer(vr_ctx *ctx);
int vr_release_settings_buffer(vr_ctx *ctx);
/* add_to_settings_buffer(name, define, value): slots are handed out in call order, max 64; a define
 * name already present fails.  Recognised defines: OCTDIM, OCTENABLED (0 = octree traversal, like
 * kernel:359), OCTREE_ROOT_INDEX, and the extensions MAX_DISTANCE (kernel:326 hard-codes 20). */
int vr_add_to_settings_buffer(vr_ctx *ctx, const char *setting_name, const char *define_name, int64_t value);
int vr_overwrite_setting(vr_ctx *ctx, const char *setting_name, const int64_t *value);
int vr_remove_from_settings_buffer(vr_ctx *ctx, const char *setting_name);   /* unimplemented in the ref: 0 */
/* direct view of the 64 x int64 array (the reference aliases it to the device) */
int64_t *vr_settings_data(vr_ctx *ctx);
/* CLCaster::set_define / remove_define (ref include/CLCaster.h:154-155) */
int vr_set_define(vr_ctx *ctx, const char *name, const char *value);
int vr_remove_define(vr_ctx *ctx, const char *name);

/* ---- per frame ---------------------------------------------------------------------------------- */

/* CLCaster::validate (ref include/CLCaster.h:139, src/CLCaster.cpp:157-206): checks that camera, map
 * (or octree) and viewport are present, (re)builds the traversal structure.  Nothing is compiled. */
int vr_validate(vr_ctx *ctx);
/* CLCaster::debug_quick_recompile (ref include/CLCaster.h:169): no JIT here; re-runs vr_validate. */
int vr_debug_quick_recompile(vr_ctx *ctx);

/* CLCaster::compute (ref include/CLCaster.h:142, src/CLCaster.cpp:224-228,946-987): one frame,
 * synchronous (returns after the device finished, like clFinish at :970). */
int vr_compute(vr_ctx *ctx);
/* Same frame without the host sync; vr_sync waits for it. */
int vr_compute_async(vr_ctx *ctx);
int vr_sync(vr_ctx *ctx);
/* Renders into a caller-owned device buffer (local_rows * width * 4 bytes) instead of the internal
 * image, asynchronously on the context stream: used by the multi-GPU gather. */
int vr_compute_into(vr_ctx *ctx, void *device_rgba);

/* View batches (BASELINE config 5): renders `count` frames of the same scene, one per camera, back to back on
 * the context stream into `device_rgba` (count consecutive frames).  cameras = count x 5 floats
 * {inclination, azimuth, x, y, z}; the assigned camera is not touched.  Asynchronous like vr_compute_into. */
int vr_compute_views(vr_ctx *ctx, const float *cameras, int count, void *device_rgba);

/* CLCaster::draw (ref include/CLCaster.h:145): headless replacement.  Copies the last frame to host
 * memory (width*height*4 bytes, or the local slab when banded). */
int vr_read_framebuffer(vr_ctx *ctx, uint8_t *rgba_out, size_t bytes);
/* Pipelined variant for streaming frames to the host: vr_frame_begin renders into the next of two
 * device images and queues its device->host copy on a second stream; vr_frame_end waits for the
 * OLDEST outstanding frame and returns its pinned host pixels (valid until two frames later). */
int vr_frame_begin(vr_ctx *ctx);
int vr_frame_end(vr_ctx *ctx, const uint8_t **rgba);

/* CLCaster::draw(sf::RenderWindow*) with CL/GL sharing (ref src/CLCaster.cpp:330-332; the viewport texture is shared with
 * OpenCL by clCreateFromGLTexture, :840-842, and acquired / released around the kernel, :952, :978): the CUDA-GL interop
 * equivalent.  vr_gl_register_texture registers an RGBA8 OpenGL texture of the viewport's size (the sprite's texture:
 * sf::Texture::getNativeHandle(); target GL_TEXTURE_2D = 0x0DE1) with the context -- it must be called on the thread whose
 * OpenGL context owns the texture; vr_gl_draw maps it, copies the last frame into it device-to-device on the context
 * stream and unmaps it (no host round trip), after which the caller draws its sprite.  Both fail with a message when
 * there is no current OpenGL context.  NOT exercised on a GL machine: this image has no GL / EGL; the
 * GPU suite only checks the failure path. */
int vr_gl_register_texture(vr_ctx *ctx, uint32_t gl_texture, uint32_t gl_target);
int vr_gl_draw(vr_ctx *ctx);
int vr_gl_unregister(vr_ctx *ctx);

/* ---- extensions beyond the reference API ------------------------------------------------------- */

/* Multi-GPU screen-tile split: this context renders only the row bands b with b % stride == first
 * (band = band_rows consecutive rows) into a compact slab.  (1,1,0) = whole frame. */
int vr_set_bands(vr_ctx *ctx, int band_rows, int stride, int first);
int vr_local_rows(const vr_ctx *ctx);
/* Multi-GPU 2-D tile interleave, the alternative to row bands: this context renders the 32x4-pixel CTA tiles (tx, ty)
 * with (tx + ty) % world == rank IN PLACE into the full-size frame passed to vr_compute_into -- its own memory or
 * another GPU's frame mapped with vr_ipc_open_handle, in which case the kernel's RGBA stores travel over NVLink and
 * the frame is assembled by the render kernels themselves (no slab, no gather).  A thin expensive screen feature
 * (the horizon) is spread over all ranks instead of landing on the one or two ranks that own its rows.
 * (1, 0) = off. */
int vr_set_tiles(vr_ctx *ctx, int world, int rank);
/* Kernel variant knobs: "persistent" (0/1: persistent warps with warp-level pixel refill for the octree
 * kernel), "refill_min" (idle lanes before a warp refills, 1..32), "ctas_per_sm" (persistent grid size),
 * "walk" (0 merged / 1 per-axis in-cell walk), "l2_persist" (0/1 access-policy window over the octree nodes),
 * "gpu_build" (1, default: vr_assign_map builds the 64-tree on the device from the uploaded map -- the
 * replacement of Octree::Generate, ref src/map/Octree.cpp:13-43,171-323; 0: on the host). */
int vr_set_option(vr_ctx *ctx, const char *name, int64_t value);
/* Use an externally owned CUDA stream (e.g. the host framework's current stream); NULL restores the own stream. */
int vr_set_stream(vr_ctx *ctx, void *cuda_stream);
/* Per-pixel auxiliary records (hit voxel, face, status, step counts; 32 B/pixel) for parity tests. */
int vr_enable_aux(vr_ctx *ctx, int enable);
int vr_read_aux(vr_ctx *ctx, void *out, size_t bytes);
/* Device pointers of the internal image / ray table (for zero-copy wrapping by the host framework). */
void *vr_device_image(vr_ctx *ctx);
int vr_read_ray_table(vr_ctx *ctx, float *out, size_t bytes);

/* Multi-GPU scene replication ("the octree is broadcast once"): rank 0 builds the 64-tree, every other
 * rank receives the two arrays (e.g. through an NCCL broadcast into device buffers) and adopts them.
 * vr_native_tree_info: sizes in bytes + levels + map edge.  vr_native_tree_copy: device-to-device copy of
 * the arrays into caller buffers.  vr_assign_native_tree: adopt (copy) arrays given as DEVICE pointers. */
int vr_native_tree_info(vr_ctx *ctx, uint64_t *node_bytes, uint64_t *type_bytes, int32_t *levels, int32_t *dim);
int vr_native_tree_copy(vr_ctx *ctx, void *device_nodes, void *device_types);
int vr_assign_native_tree(vr_ctx *ctx, const void *device_nodes, uint64_t node_bytes, const void *device_types,
                          uint64_t type_bytes, int32_t levels, int32_t dim);

/* The top grid the closed-form walk (option walk = 2) reads instead of the upper octree levels: a dense table over the
 * blocks of edge 1 << *grid_shift derived from the 64-tree on the device (csrc/vr_build.cu: vr_build_grid_device; layout
 * in csrc/vr_types.h).  Builds it if necessary and copies it to `host_out` when capacity (entries) suffices.  Returns
 * the number of entries, 0 on failure (e.g. a map of 4^3 or less has no grid).  For tests / inspection.
 * With option "directed_grid" = 1 (the default) the grid is eight tables, one per direction octant of a ray (bit a of the
 * octant number set = the ray moves towards lower coordinates on axis a), 8 << (3 * *grid_bits) entries, octant 0 first;
 * with 0 it is the single undirected table of 1 << (3 * *grid_bits) entries. */
uint64_t vr_top_grid_read(vr_ctx *ctx, uint32_t *host_out, uint64_t capacity, int32_t *grid_shift, int32_t *grid_bits);

/* Multi-GPU frame assembly without SM involvement: every rank pushes its band slab straight into the frame
 * buffer that lives on the root GPU with ONE strided copy-engine transfer over NVLink (cudaMemcpy2DAsync
 * through a CUDA-IPC mapping), so the ray casting kernel of the next frame is not disturbed by a gather kernel.
 * vr_device_alloc / vr_device_free: cudaMalloc'ed memory (framework caching allocators may hand out
 * sub-allocations or VMM memory that legacy IPC cannot export).
 * vr_ipc_get_handle: 64-byte cudaIpcMemHandle of such an allocation (root side).
 * vr_ipc_open_handle / vr_ipc_close_handle: map / unmap it in another process.
 * vr_push_bands: enqueue on `cuda_stream` the copy of this context's slab (band layout of vr_set_bands) into
 * `frame` (device pointer, local or IPC-mapped; rows padded to a multiple of band_rows * stride). */
int vr_device_alloc(vr_ctx *ctx, size_t bytes, void **device_ptr);   /* plain cudaMalloc: IPC-exportable */
int vr_device_free(vr_ctx *ctx, void *device_ptr);
int vr_ipc_get_handle(vr_ctx *ctx, void *device_ptr, void *handle64);
int vr_ipc_open_handle(vr_ctx *ctx, const void *handle64, void **device_ptr);
int vr_ipc_close_handle(vr_ctx *ctx, void *device_ptr);
int vr_push_bands(vr_ctx *ctx, const void *slab, void *frame, void *cuda_stream);
/* ---- Multi-GPU frame scheduler (csrc/vr_mgpu.cu): what CLCaster::run_kernel (src/CLCaster.cpp:946-987) is to one OpenCL
 * device, for the GPUs of one node.  ONE process per GPU, each with its own context that has been given the scene and
 * the viewport like a single-GPU caster (rank 0 the map / octree; the others receive the octree by broadcast); a frame
 * is split into the ray kernel's 32x4-pixel tiles, tile (tx, ty) on rank (tx + ty) mod world.
 *   vr_mgpu_init            collective.  `session`: a name unique to this run, the same on every rank (it names the POSIX
 *                           shared-memory segment through which the ranks find each other; nothing else is needed).
 *                           flags 0: the frame is assembled in the root GPU's memory -- every rank's kernel stores its
 *                           pixels in place over NVLink (CUDA-IPC mapping), no gather.  VR_MGPU_HOST_FRAME: the frame is
 *                           assembled in shared pinned host memory, every rank copies its own row bands there over its
 *                           own PCIe link.  Call after create_viewport.
 *   vr_mgpu_broadcast_octree collective: rank 0's 64-tree to every rank (ncclBroadcast over NVLink), once per scene.
 *   vr_mgpu_frame           every rank: enqueue this rank's share of the next frame (camera / lights / settings are read
 *                           from the retained host pointers, as in vr_compute).  Does not wait for the GPU; blocks only
 *                           while all `ring` (4) frame buffers are still unreleased.  *frame_no = the frame's number.
 *   vr_mgpu_frame_wait      root: returns when every rank has finished frame_no; *rgba = the assembled frame (a device
 *                           pointer, or a host pointer with VR_MGPU_HOST_FRAME), valid until vr_mgpu_frame_release.
 *                           Other ranks: returns when their own share is done.
 *   vr_mgpu_frame_release   root: the frame's buffer may be rendered into again.  (No-op elsewhere.)
 *   vr_mgpu_flush           makes the context's stream (vr_set_stream) wait for the frames enqueued so far: consecutive frames
 *                           run on two alternating streams of the scheduler (the next frame's first CTAs fill the tail
 *                           of the current one), host-frame copies on a third.
 *   vr_mgpu_barrier         collective, CPU only: returns when every rank has called it (a spin on the shared segment; the
 *                           ranks leave it within a microsecond of each other).  Enqueues nothing and waits for no
 *                           stream.  For a frame loop that must start on all GPUs at the same instant (a benchmark's
 *                           timed region, a scene change): a rank that starts late makes the others wait at the frame
 *                           ring.
 *   vr_mgpu_shutdown        collective.
 * Completion is signalled through per-rank counters in the shared segment, stored by the device after the rank's kernel
 * and polled by the CPU with time-outs: no per-frame collective.  All calls return 1 on success, 0 + last error. */
#define VR_MGPU_HOST_FRAME 1u
int vr_mgpu_init(vr_ctx *ctx, const char *session, int world, int rank, unsigned flags);
int vr_mgpu_broadcast_octree(vr_ctx *ctx);
int vr_mgpu_frame(vr_ctx *ctx, uint64_t *frame_no);
int vr_mgpu_frame_wait(vr_ctx *ctx, uint64_t frame_no, const uint8_t **rgba);
int vr_mgpu_frame_release(vr_ctx *ctx, uint64_t frame_no);
int vr_mgpu_flush(vr_ctx *ctx);        /* the context's stream waits for every frame enqueued so far (they run on the scheduler's own streams) */
int vr_mgpu_barrier(vr_ctx *ctx);
int vr_mgpu_shutdown(vr_ctx *ctx);

/* Page-locks host memory the caller owns (e.g. a frame in a POSIX shared-memory segment mapped by every rank) so
 * that vr_push_bands can take it as `frame`: each rank then copies its own bands device->host over its own PCIe
 * link, straight into frame order -- the end-to-end path with a HOST result needs no gather on the device. */
int vr_host_register(vr_ctx *ctx, void *host_ptr, size_t bytes);
int vr_host_unregister(vr_ctx *ctx, void *host_ptr);

/* Octree::Load (declared, never defined in the reference: include/map/Octree.h:38) and its counterpart: the
 * traversal octree of this context to / from a file ("VR64" header + node array + leaf types). */
int vr_octree_save(vr_ctx *ctx, const char *path);
int vr_octree_load(vr_ctx *ctx, const char *path);

typedef struct vr_stats {
    uint64_t kernel_launches;      /* kernels launched by this context so far                */
    uint64_t frames;
    uint64_t native_nodes;         /* 64-tree nodes                                          */
    uint64_t native_bytes;         /* nodes + leaf types                                     */
    uint64_t solid_voxels;
    int32_t levels;
    int32_t used_svo;              /* last frame used the SVO kernel                         */
    int32_t bias[3];               /* last frame's get_oct_vox start bias                    */
    int32_t device;
    float last_kernel_ms;          /* CUDA-event time of the last vr_compute's kernel        */
    float build_ms;                /* last on-device 64-tree build (assign_map), CUDA events */
    float build_masks_ms;          /* ... of which the kernel that reads the N^3 map         */
} vr_stats;
int vr_get_stats(vr_ctx *ctx, vr_stats *out);

/* Octree::Generate equivalent (ref src/map/Octree.cpp:13-43,171-323): reference-format descriptor
 * buffer from a cubic char map.  Two-call pattern: out == NULL returns the required entry count in
 * *entries; otherwise writes up to *entries descriptors.  *root_index receives the root position. */
int vr_octree_generate(const int8_t *voxels, int dim, uint64_t *out, uint64_t *entries, uint64_t *root_index);
/* Octree::GetVoxel / get_oct_vox equivalent on a descriptor buffer (ref src/map/Octree.cpp:45). */
int vr_octree_get_voxel(const uint64_t *descriptors, uint64_t entries, uint64_t root_index, int dim,
                        const int32_t pos[3], int32_t sub_oct_pos[3], int32_t *resolution);

#ifdef __cplusplus
}
#endif
#endif
