/* vr_build.h -- GPU 64-tree builder (internal; see vr_build.cu). */
#ifndef VR_BUILD_H
#define VR_BUILD_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "vr_types.h"

typedef struct vr_device_tree {
    vr_node *nodes;            /* device, cudaMalloc'ed: ownership passes to the caller */
    uint8_t *types;
    uint64_t n_nodes, n_types, solid_voxels;
    int levels;
    float masks_ms;            /* CUDA-event time of the HBM-bound brick-mask kernel (reads the N^3 map once) */
    float total_ms;            /* whole build on the stream, including the host read of the level counts */
} vr_device_tree;

/* Builds the 64-tree of the dense map d_map[x + dim*(y + dim*z)] (device pointer; dim a power of two >= 4; solid =
 * value 5 or 6) on `stream`.  Emits exactly the arrays vr_native_from_dense (vr_octree.cpp) produces. */
cudaError_t vr_build_tree_device(const int8_t *d_map, int dim, cudaStream_t stream, vr_device_tree *out,
                                 unsigned long long *launches);

/* The same tree from a column description (device pointers, lo/hi[x + dim*y]: column (x, y) is solid for lo <= z <= hi,
 * every solid voxel has value `type`), without the N^3 volume and without any dense workspace: only the bricks that hold
 * a voxel are materialised (4096^3: 64 GiB dense).  dim a power of two, 4 .. 4096.  Emits exactly the arrays
 * vr_native_from_columns (vr_octree.cpp) produces. */
cudaError_t vr_build_tree_columns_device(const int32_t *d_lo, const int32_t *d_hi, int dim, uint8_t type, cudaStream_t stream,
                                         vr_device_tree *out, unsigned long long *launches);

/* Solid-subtree collapse of a tree just built by one of the two builders above (vr_types.h: VR_NODE_SOLID): the arrays
 * are replaced by the re-packed ones (same BFS order); a tree without a solid cube is left untouched.  Equals the host
 * version (vr_octree.cpp: vr_native_collapse_solid) array for array. */
cudaError_t vr_collapse_solid_device(vr_device_tree *t, cudaStream_t stream, unsigned long long *launches);

/* Top grid of the closed-form walk (vr_types.h: vr_frame_params::grid) from the 64-tree d_nodes: *grid_out is
 * cudaMalloc'ed (ownership passes to the caller).  cudaErrorInvalidValue when the tree is too shallow for a grid.
 * directed: the eight per-octant tables (8 << (3 * grid_bits) entries, vr_octree.cpp: vr_native_grid_directed). */
cudaError_t vr_build_grid_device(const vr_node *d_nodes, int levels, int dim, bool directed, cudaStream_t stream, uint32_t **grid_out,
                                 int *grid_shift, int *grid_bits, unsigned long long *launches);

#endif
