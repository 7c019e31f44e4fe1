/*
 * vr_octree.h -- host-side octree producers/consumers of the B200 caster (internal).
 *
 *  - reference-format child-descriptor buffers (the layout CLCaster::assign_octree uploads and
 *    kernels/ray_caster_kernel.cl:140-251 consumes; reference include/map/Octree.h:89-94):
 *    generator (replaces Octree::Generate, src/map/Octree.cpp:13-43,171-323) and point query
 *    (replaces get_oct_vox / Octree::GetVoxel);
 *  - the native 64-tree (vr_types.h: vr_node) the SVO kernel traverses: built from the dense char
 *    map or imported from a reference-format descriptor buffer.
 */
#ifndef VR_OCTREE_H
#define VR_OCTREE_H

#include <stdint.h>
#include <vector>

#include "vr_types.h"

struct vr_native_tree {
    std::vector<vr_node> nodes;        /* BFS order, root at 0, children of a node contiguous */
    std::vector<uint8_t> leaf_types;   /* voxel values of set leaf bits                       */
    int levels = 0;                    /* 64-tree levels; dimension covered = 4^levels        */
    int dim = 0;                       /* map edge (power of two)                             */
    uint64_t solid_voxels = 0;
};

/* Reference-format generator.  data: char map[x + N*(y + N*z)], any non-zero voxel is occupied
 * (src/map/Octree.cpp:198-201).  Only uniformly EMPTY subtrees collapse (:230-233).  Layout: root at
 * index 0 with relative pointer 1; a node's valid children are contiguous in ascending child order
 * and followed by their far-pointer slots; relative pointers are 15 bit, larger distances go
 * through a far pointer holding an absolute index (kernel:222-225).  Returns false on bad input. */
bool vr_ref_octree_generate(const int8_t *data, int dim, std::vector<uint64_t> &out, uint64_t *root_index);

/* get_oct_vox (kernel:140-251) on the host: frame-uniform, so the caster evaluates it once per
 * frame instead of once per pixel.  Returns found; writes the child-cell origin and `resolution`. */
int vr_ref_octree_query(const uint64_t *desc, uint64_t len, uint64_t root_index, int octdim, const int pos[3],
                        int sub_oct_pos[3], int *resolution);

/* Native tree from the dense map: a voxel is solid iff its value is 5 or 6 (kernel:575). */
bool vr_native_from_dense(const int8_t *map, int dim, vr_native_tree &out);

/* Native tree from a reference-format descriptor buffer (occupancy only; type 5 unless `types`,
 * a dense map of the same dimension, is given). */
bool vr_native_from_ref(const uint64_t *desc, uint64_t len, uint64_t root_index, int dim, const int8_t *types,
                        vr_native_tree &out);

/* Native tree straight from a column description, without materialising the N^3 volume (4096^3 = 64 GiB):
 * column (x, y) is solid for lo[x + dim*y] <= z <= hi[x + dim*y] (empty when lo > hi), every solid voxel has
 * value `type`.  Produces exactly the arrays vr_native_from_dense gives for the equivalent dense map. */
bool vr_native_from_columns(const int32_t *lo, const int32_t *hi, int dim, uint8_t type, vr_native_tree &out);

/* Solid-subtree collapse (vr_types.h: VR_NODE_SOLID), applied to a tree in BFS order by every producer above: bottom-up,
 * a leaf brick of 64 set voxels of one type and an inner node whose 64 children are solid nodes of one type become solid
 * nodes; the nodes and voxel types below a solid node are dropped and the arrays re-packed in the same BFS order.
 * Returns the number of nodes removed.  (Device version: vr_build.cu: vr_collapse_solid_device.) */
size_t vr_native_collapse_solid(vr_native_tree &t);

/* Top grid of the closed-form walk (vr_types.h: vr_frame_params::grid) from the 64-tree: host version (the caster builds
 * it on the device, vr_build.cu: vr_build_grid_device; this one serves the host emulation and checks that one).
 * Returns false when the tree is too shallow for a grid (a single level: maps up to 4^3). */
bool vr_native_grid(const vr_node *nodes, int levels, int dim, std::vector<uint32_t> &grid, int *grid_shift, int *grid_bits);
/* Directed top grids: eight tables, one per direction octant of the ray (see vr_octree.cpp); table o at o << (3 * bits). */
bool vr_native_grid_directed(const vr_node *nodes, int levels, int dim, std::vector<uint32_t> &grid, int *grid_shift, int *grid_bits);
#define VR_GRID_MAX_RADIUS 63

/* Point query on the native tree: voxel value (5/6) or 0, and the empty-cell shift if empty. */
int vr_native_query(const vr_native_tree &t, int x, int y, int z, int *cell_shift);

#endif
