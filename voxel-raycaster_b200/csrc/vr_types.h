/*
 * vr_types.h -- plain-old-data shared by the host runtime and the device kernels.
 *
 * Replaces the 16 positional OpenCL kernel arguments bound by CLCaster::validate
 * (reference src/CLCaster.cpp:186-202 <-> kernels/ray_caster_kernel.cl:256-273) with one
 * launch-constant parameter block.
 */
#ifndef VR_TYPES_H
#define VR_TYPES_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VR_HD __host__ __device__ __forceinline__
#else
#define VR_HD inline
#endif

/* 64-tree node, 16 bytes, read with one 128-bit load.
 * A node at shift s covers a (4<<s)^3 voxel cube split into 4x4x4 children of edge (1<<s).
 * Child slot ci = cx | cy<<2 | cz<<4.  Bit ci of `mask` = child is non-empty.
 *   s > 0 : children are nodes, stored contiguously in ascending set-bit order at
 *           nodes[child_base + popcount(mask & ((1<<ci)-1))].
 *   s == 0: children are voxels; voxel type at leaf_types[child_base + popcount(...)].
 * Two levels of the reference's 2^3 / 8-byte child descriptors (kernel:49-54, Octree.h:89-94)
 * collapse into one node: same bytes per level, half the dependent loads.
 * SOLID nodes (the reference collapses empty subtrees only, src/map/Octree.cpp:230-233; SURVEY 8f-1 asks for both): a
 * node whose whole cube is set voxels of ONE type is stored as a single node without children at whatever level it
 * sits: mask = ~0, child_base = VR_NODE_SOLID | type -- no nodes below it, no leaf_types entries.  A lookup that loads
 * such a node has found a set voxel of that type. */
#define VR_NODE_SOLID 0x80000000u
typedef struct vr_node {
    uint32_t mask_lo;
    uint32_t mask_hi;
    uint32_t child_base;
    uint32_t aux;          /* vr_node_planes(mask): which of the 4+4+4 axis-perpendicular slot planes are empty */
} vr_node;

/* aux word of a node: bit k = the 4x4 plane of slots x == k is empty, bit 4+k: y == k, bit 8+k: z == k.
 * A function of the mask alone, stored so that the traversal gets it with the node's 128-bit load. */
VR_HD uint32_t vr_node_planes(unsigned long long m) {
    uint32_t e = 0;
    for (int k = 0; k < 4; k++) {
        e |= ((m & (0x1111111111111111ull << k)) == 0ull ? 1u : 0u) << k;
        e |= ((m & (0x000F000F000F000Full << (4 * k))) == 0ull ? 1u : 0u) << (4 + k);
        e |= ((m & (0xFFFFull << (16 * k))) == 0ull ? 1u : 0u) << (8 + k);
    }
    return e;
}

/* Block edge (log2) of the top grid of the closed-form walk (vr_frame_params::grid) for a map of edge `dim` whose 64-tree
 * root has child shift root_shift: the finest one whose table stays within VR_GRID_MAX_BITS index bits (2^24 entries =
 * 64 MB: L2-resident on B200), never finer than a leaf brick (shift 2).  1024^3 -> bricks (256^3 entries), 4096^3 ->
 * 16^3 blocks. */
#ifndef VR_GRID_MAX_BITS
#define VR_GRID_MAX_BITS 24
#endif
VR_HD int vr_grid_shift_for(int root_shift, int dim) {
    int g = 2;
    while (g < root_shift && 3 * (31 - __builtin_clz((unsigned)(dim >> g))) > VR_GRID_MAX_BITS) g += 2;
    return g;
}

/* Directed top grids (vr_frame_params::grid_directed): edge cap of the empty cube stored per block and direction octant,
 * and the entry of an empty block from its cube edge e (blocks), the log2 edge `cell` of the aligned empty octree cell
 * around it and its mirrored block coordinates k (vr_octree.cpp: vr_native_grid_directed; vr_build.cu). */
#ifndef VR_GRID_MAX_CUBE
#define VR_GRID_MAX_CUBE 64u
#endif
VR_HD uint32_t vr_grid_directed_entry(uint32_t e, int cell, int g, int kx, int ky, int kz) {
    if (cell > g) {
        /* the aligned cell reaches the mirrored block (k | cm) on every axis, the cube k + e - 1 */
        const int cm = (1 << (cell - g)) - 1, c = (int)e - 1;
        const int rx = (kx | cm) - kx, ry = (ky | cm) - ky, rz = (kz | cm) - kz;
        if (rx >= c && ry >= c && rz >= c && rx + ry + rz > 3 * c) return (uint32_t)cell;
    }
    return (uint32_t)g | (((e - 1u) << g) << 8);
}

/* Per-pixel auxiliary record (32 bytes), layout-identical to the oracle's vro_aux. */
typedef struct vr_aux {
    int32_t hit[3];
    uint8_t face;
    uint8_t status;
    uint8_t flags;
    uint8_t hit_type;
    uint32_t steps_first;
    uint32_t steps_total;
    uint32_t node_fetches;   /* 16-byte node loads issued for this pixel (SVO kernels) */
    uint32_t lookups;        /* octree lookups (cell changes) for this pixel          */
} vr_aux;

enum {
    VR_ST_SKIP_PRIMARY = 0, VR_ST_OOB = 1, VR_ST_MAXDIST = 2, VR_ST_SHADOW_HIT = 3,
    VR_ST_SKIP_REDIRECT = 4, VR_ST_BOUNCES = 5
};
enum { VR_FL_LIT = 1, VR_FL_REFLECTED = 2, VR_FL_TIE = 4, VR_FL_ATLAS_CLAMP = 8, VR_FL_FRAC0 = 16 };

#define VR_MAX_LIGHTS 8          /* LightController's fixed slot count (ref src/LightController.cpp:3-11)  */
#define VR_MAX_LEVELS 8          /* 64-tree levels: dimension up to 4^8 = 65536 */

typedef struct vr_frame_params {
    /* viewport */
    int32_t width, height;
    /* multi-GPU screen-tile split: the frame is cut into row bands of band_rows rows; band b
     * belongs to this launch iff b % band_stride == band_first.  The launch renders local_rows
     * rows into a compact slab: local row ly -> frame row
     * ((ly / band_rows) * band_stride + band_first) * band_rows + ly % band_rows. */
    int32_t local_rows, band_rows, band_stride, band_first;
    /* multi-GPU 2-D tile interleave (vr_set_tiles): with tile_world > 1 the launch renders, IN PLACE in a full-size
     * frame (local memory or a peer GPU's frame mapped over NVLink), the CTA tiles (tx, ty) with
     * (tx + ty) % tile_world == tile_rank; bands are then off */
    int32_t tile_world, tile_rank;
    const float *ray_table;        /* float4 per pixel (kernel arg 3)                        */
    uint8_t *image;                /* RGBA8, row pitch width*4; local slab when banded       */
    vr_aux *aux;                   /* optional                                               */
    /* dense map (args 0,1) */
    const int8_t *map;
    int32_t dim[3];
    /* camera (args 4,5): trig = sinf/cosf of (inclination, azimuth), host evaluated          */
    float cam_pos[3];
    float trig[4];
    float bias[3];                 /* get_oct_vox start bias (kernel:353), host evaluated     */
    /* lights (arg 6): rgbi + position per slot.  The reference reads slot 0 only (kernel:660-670);
     * light_count > 1 (setting LIGHT_COUNT) enables the multi-light extension, see vr_next_light      */
    float light_rgbi[VR_MAX_LIGHTS][4];
    float light_pos[VR_MAX_LIGHTS][3];
    int32_t light_count;
    /* atlas (args 9-11) */
    unsigned long long atlas_tex;  /* cudaTextureObject_t                                     */
    const uint8_t *atlas;          /* same texels, linear RGBA8 (host emulation + fallback)   */
    int32_t atlas_dim[2];
    int32_t atlas_scale[2];        /* atlas_dim / tile_dim, integer (kernel:654)              */
    /* settings */
    int32_t max_distance;          /* kernel:326                                              */
    int32_t max_bounces;           /* kernel:357 `bounce_count < 2`; setting MAX_BOUNCES (TODO list, ref src/main.cpp:31-33) */
    /* native 64-tree */
    const vr_node *nodes;
    const uint8_t *leaf_types;
    int32_t root_shift;            /* child shift of the root node = 2*(levels-1)             */
    int32_t levels;
    /* top grid of the closed-form walk (walk = 2, vr_canon.h): the octree levels with child shift >= grid_shift flattened
     * into one dense table over the blocks of edge 1 << grid_shift, grid[bx + (by + bz * G) * G] with G = 1 << grid_bits.
     * Entry: bit 31 set = the block is not empty, bits 0-30 = index of the node (child shift grid_shift - 2) that covers
     * it; bit 31 clear = the block is empty and lies in an empty cell [p & ~m, (p | m) + ext] per axis with
     * m = (1 << (entry & 31)) - 1 and ext = entry >> 8: either the aligned empty octree cell around the block (ext = 0) or
     * the block grown by its empty radius in blocks (Chebyshev distance to the nearest non-empty block or to the map
     * boundary, minus one) -- whichever is wider. */
    const uint32_t *grid;
    int32_t grid_shift, grid_bits, grid_dim;       /* grid_dim = G */
    int32_t grid_directed;                         /* 1: eight tables, one per direction octant of the ray, table o at
                                                    * o << (3 * grid_bits) (vr_octree.cpp: vr_native_grid_directed) */
    int32_t cam_on_edge;                           /* host evaluated (vr_cam_on_edge): the camera sits on a voxel edge or corner */
} vr_frame_params;

/* A camera with two or three integer coordinates: intersection_t starts at the same value on those axes (kernel:317-323
 * with a zero fraction), so every primary ray makes a step along them at once (kernel:558), which counts once (kernel:714).
 * Returns 0: no such camera; 1: without a start bias -- the tie is the ray's first step, vr_canon.h corrects the step
 * count for it (vr_canon_first_step_tie); 2: with a get_oct_vox start bias (kernel:353) -- the bias shifts the axes apart,
 * the tie falls anywhere along the ray, and the closed-form walk hands such frames to its voxel-by-voxel form
 * (vr_canon_slow: every step observed). */
static inline int32_t vr_cam_on_edge(const float *cam_pos, const float *bias) {
    int n = 0;
    for (int i = 0; i < 3; i++) n += cam_pos[i] == floorf(cam_pos[i]) ? 1 : 0;
    if (n < 2) return 0;
    return (bias[0] != 0.0f || bias[1] != 0.0f || bias[2] != 0.0f) ? 2 : 1;
}

#endif
