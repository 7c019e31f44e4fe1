/*
 * ref_kernel_host.cpp -- TEST INFRASTRUCTURE.  Host side of the reference kernel compiled for the CPU: plays the role
 * of CLCaster::run_kernel (ref src/CLCaster.cpp:946-987: NDRange over (width, height), one work-item per pixel) for
 * the translation unit that #includes the reference's kernels/ray_caster_kernel.cl through cl_shim.h.
 * Built by `make -C oracle ref` into oracle/_ref/; never loaded by the product.
 */
#include "cl_shim.h"

thread_local int vr_global_id[2];

#ifdef VR_REF_MAX_DISTANCE_LIFTED
/* second build only: kernel:326's `int max_distance = 20;` reads this variable instead (the one parameter lift the
 * oracle also makes; every other character of the kernel is the reference's) */
int vr_ref_max_distance = 20;
extern "C" void ref_set_max_distance(int v) { vr_ref_max_distance = v; }
#endif

/* the reference kernel, from where it lies (path given by the build recipe; `(typeN)(` literals rewritten on the fly) */
#include VR_REF_KERNEL

extern "C" int ref_kernel_max_distance(void) {
#ifdef VR_REF_MAX_DISTANCE_LIFTED
    return -1;           /* max_distance comes from ref_set_max_distance() */
#else
    return 20;           /* kernel:326, verbatim */
#endif
}

/* One frame.  Buffers are exactly what CLCaster binds (host:186-202): map chars, map_dim int3, resolution int2, ray
 * table float4 stride, camera dir/pos, lights, RGBA8 image (prefilled by the caller), RGBA8 atlas, atlas/tile dims,
 * the three octree buffers and the 64-slot settings buffer (slot 0 OCTDIM, 1 OCTENABLED, 2 OCTREE_ROOT_INDEX, as
 * src/Application.cpp:35-39 and host:113 register them). */
extern "C" int ref_raycast(int width, int height, const float *ray_table, char *map, const int *map_dim3, const float *cam_dir2,
                           const float *cam_pos3, float *lights, int light_count, uint8_t *rgba, uint8_t *written,
                           uint8_t *atlas, int atlas_w, int atlas_h, int tile_w, int tile_h, unsigned long *oct_desc,
                           unsigned int *oct_lookup, unsigned long *oct_attach, unsigned long *settings, int y0, int y1, int y_stride) {
    int3 map_dim(map_dim3[0], map_dim3[1], map_dim3[2]);
    int2 resolution(width, height);
    float2 cam_dir(cam_dir2[0], cam_dir2[1]);
    float3 cam_pos(cam_pos3[0], cam_pos3[1], cam_pos3[2]);
    int2 atlas_dim(atlas_w, atlas_h), tile_dim(tile_w, tile_h);
    vr_image image = {width, height, rgba, written};
    vr_image atlas_img = {atlas_w, atlas_h, atlas, nullptr};
    static_assert(sizeof(float3) == 16, "float3 must have OpenCL's 16-byte stride");
    float3 *projection = reinterpret_cast<float3 *>(const_cast<float *>(ray_table));
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = y0; y < y1; y += y_stride)
        for (int x = 0; x < width; x++) {
            vr_global_id[0] = x;
            vr_global_id[1] = y;
            raycaster(map, &map_dim, &resolution, projection, &cam_dir, &cam_pos, lights, &light_count, &image, &atlas_img,
                      &atlas_dim, &tile_dim, oct_desc, oct_lookup, oct_attach, settings);
        }
    return 0;
}
