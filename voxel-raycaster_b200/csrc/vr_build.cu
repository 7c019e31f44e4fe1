/*
 * vr_build.cu -- 64-tree builder on the GPU (sm_100a), from the dense char map already resident in HBM.
 *
 * Replaces, for the octree the traversal kernel reads, the reference's CPU builder Octree::Generate /
 * GenerationRecursion (src/map/Octree.cpp:13-43, 171-323: recursive, visits all N^3 voxels through
 * get1DIndexedVoxel, O(N^3) on one core, capped at 100 000 descriptors) -- SURVEY 8(f) rank 1.  The host
 * builder in vr_octree.cpp (vr_native_from_dense) stays as the definition of the layout; this builder emits
 * exactly the same arrays (tests/test_gpu_parity.py::test_gpu_builder_*).
 *
 * Data flow (everything stays in HBM; one small D2H read of the level counts):
 *   1. vr_brick_masks     N^3 bytes -> one 64-bit occupancy mask per 4^3 brick.  A lane reads 16 x 128 bits (4 bricks
 *                         x 16 voxel rows), a warp 16 x 512 contiguous bytes, and assembles its 4 masks with
 *                         byte-SIMD integer ops.  This is the HBM-bound kernel: 1 byte read per voxel.
 *   2. vr_reduce_masks    level l+1 -> level l: one warp per parent, two ballots over its 64 children.
 *   3. cub::DeviceScan    per level: exclusive sums of "mask != 0" (node index inside the level) and, on the leaf
 *                         level, of popcount(mask) (index of the brick's first voxel type).
 *   4. vr_emit_nodes      one thread per non-empty mask: vr_node {mask, child_base, plane bits} at its BFS slot.
 *   5. vr_emit_types      one thread per leaf brick: the voxel values of its set bits.
 * Masks are stored per level under a hierarchical key (6 bits per level: the child slot x | y<<2 | z<<4), so that
 * ascending key order IS the breadth-first order of the host builder and the children of a node are the 64
 * consecutive keys key*64 .. key*64+63.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "vr_build.h"

namespace {

__device__ __forceinline__ uint32_t solid4(uint32_t w) {
    /* 4 voxels per word -> 4 bits: value 5 or 6 (kernel:575).  Byte-SIMD in plain integer ops:
     *   t has a zero byte where the upper six bits of the voxel are 000001 (values 4..7),
     *   u has bit 0 of a byte set where the two low bits differ (01 or 10: 5 or 6 among 4..7),
     *   z = exact zero-byte test of t (no carries between bytes), bit 7 of every zero byte. */
    const uint32_t t = (w & 0xFCFCFCFCu) ^ 0x04040404u;
    const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
    const uint32_t u = (w ^ (w >> 1)) & 0x01010101u;
    const uint32_t v = (z >> 7) & u;                               /* bits 0, 8, 16, 24 */
    return ((v * 0x00204081u) >> 21) & 15u;                        /* gathered into one nibble (no colliding partial products) */
}

/* Hierarchical key of a leaf brick: one 6-bit child slot per level below the root, the root's slot on top.  `half`: the map
 * edge is 2 * 4^(levels-1), not 4^levels, so only the 2x2x2 root slots {0,1}^3 lie inside the map: that top digit is
 * stored as 3 bits (cx | cy << 1 | cz << 2 -- the same order as the slot numbers 0,1,4,5,16,17,20,21), which makes the key
 * space of every level exactly the cells of the map at that level instead of eight times as many. */
__device__ __forceinline__ uint32_t brick_key(uint32_t bx, uint32_t by, uint32_t bz, int digits, int half) {
    uint32_t key = 0;
    const int full = half ? digits - 1 : digits;
    for (int j = 0; j < full; j++)
        key |= (((bx >> (2 * j)) & 3u) | (((by >> (2 * j)) & 3u) << 2) | (((bz >> (2 * j)) & 3u) << 4)) << (6 * j);
    if (half && digits > 0)
        key |= (((bx >> (2 * full)) & 1u) | (((by >> (2 * full)) & 1u) << 1) | (((bz >> (2 * full)) & 1u) << 2)) << (6 * full);
    return key;
}

/* 1. one lane = 4 bricks adjacent in x (16 voxels = one 128-bit load per voxel row, 16 rows), one warp = 128 bricks:
 * every load instruction of a warp covers 512 contiguous bytes, 16 independent loads are in flight per lane, and
 * the 4 masks of a lane are 32 contiguous bytes of the key-ordered output (4 consecutive x slots of one node).
 * Requires dim >= 16; smaller maps take vr_brick_masks_small. */
__global__ void __launch_bounds__(128)
vr_brick_masks(const int8_t *__restrict__ map, int dim, int nb, int nquad, int digits, int half, unsigned long long *__restrict__ leaf) {
    const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = (unsigned long long)nb * nb * nquad;
    if (gid >= total) return;
    const unsigned qx = (unsigned)(gid % (unsigned)nquad);
    const unsigned row = (unsigned)(gid / (unsigned)nquad);
    const unsigned by = row % (unsigned)nb, bz = row / (unsigned)nb;
    const uint4 *rows = reinterpret_cast<const uint4 *>(map);
    const size_t qdim = (size_t)dim / 16;                      /* 128-bit words per voxel row */
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int z = 0; z < 4; z++)
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint4 w = __ldg(rows + qx + qdim * ((size_t)(4 * by + y) + (size_t)dim * (size_t)(4 * bz + z)));
            const int sh = 4 * y + 16 * (z & 1);
            if (z < 2) {
                lo[0] |= solid4(w.x) << sh; lo[1] |= solid4(w.y) << sh; lo[2] |= solid4(w.z) << sh; lo[3] |= solid4(w.w) << sh;
            } else {
                hi[0] |= solid4(w.x) << sh; hi[1] |= solid4(w.y) << sh; hi[2] |= solid4(w.z) << sh; hi[3] |= solid4(w.w) << sh;
            }
        }
    /* bricks 4*qx .. 4*qx+3 differ in the lowest key digit only: keys k, k+1, k+2, k+3 */
    uint4 *dst = reinterpret_cast<uint4 *>(leaf + brick_key(4 * qx, by, bz, digits, half));
    dst[0] = make_uint4(lo[0], hi[0], lo[1], hi[1]);
    dst[1] = make_uint4(lo[2], hi[2], lo[3], hi[3]);
}

/* maps narrower than 16 voxels: one thread per brick */
__global__ void vr_brick_masks_small(const int8_t *__restrict__ map, int dim, int nb, int digits, int half, unsigned long long *__restrict__ leaf) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (unsigned)(nb * nb * nb)) return;
    const unsigned bx = gid % (unsigned)nb, by = (gid / (unsigned)nb) % (unsigned)nb, bz = gid / (unsigned)(nb * nb);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(map);
    const size_t wdim = (size_t)dim / 4;
    unsigned long long m = 0ull;
    for (int z = 0; z < 4; z++)
        for (int y = 0; y < 4; y++)
            m |= (unsigned long long)solid4(words[bx + wdim * ((size_t)(4 * by + y) + (size_t)dim * (size_t)(4 * bz + z))]) << (4 * y + 16 * z);
    leaf[brick_key(bx, by, bz, digits, half)] = m;
}

/* 2. one warp per parent: bit ci = child ci has any voxel */
__global__ void __launch_bounds__(128)
vr_reduce_masks(const unsigned long long *__restrict__ child, unsigned long long *__restrict__ parent, unsigned nparent) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= nparent) return;
    const unsigned long long *c = child + (size_t)warp * 64;
    const unsigned lo = __ballot_sync(0xffffffffu, c[lane] != 0ull);
    const unsigned hi = __ballot_sync(0xffffffffu, c[lane + 32] != 0ull);
    if (lane == 0) parent[warp] = (unsigned long long)lo | ((unsigned long long)hi << 32);
}

struct NonZero {
    __host__ __device__ uint32_t operator()(unsigned long long m) const { return m != 0ull ? 1u : 0u; }
};
struct IsSolidFlag {
    __host__ __device__ uint32_t operator()(uint8_t v) const { return v != 0xFF ? 1u : 0u; }
};
struct PopCount64 {
    __host__ __device__ unsigned long long operator()(unsigned long long m) const {
#if defined(__CUDA_ARCH__)
        return (unsigned long long)__popcll(m);
#else
        return (unsigned long long)__builtin_popcountll(m);
#endif
    }
};
struct PopCount {
    __host__ __device__ uint32_t operator()(unsigned long long m) const {
#if defined(__CUDA_ARCH__)
        return (uint32_t)__popcll(m);
#else
        return (uint32_t)__builtin_popcountll(m);
#endif
    }
};
/* the root of a map of edge 2 * 4^(levels-1): its 8 children are the level-1 keys 0..7 (brick_key), slot cx | cy<<2 | cz<<4 */
__global__ void vr_reduce_root_half(const unsigned long long *__restrict__ child, unsigned long long *__restrict__ root) {
    unsigned long long m = 0ull;
    for (int c = 0; c < 8; c++)
        if (child[c] != 0ull) m |= 1ull << ((c & 1) | (((c >> 1) & 1) << 2) | (((c >> 2) & 1) << 4));
    root[0] = m;
}

/* totals[l] = non-empty masks of level l (l >= 1), totals[levels] = solid voxels */
__global__ void vr_level_totals(const unsigned long long *masks, const uint32_t *prefix, const uint32_t *vox_prefix,
                                const unsigned long long *level_off, int levels, uint32_t *totals) {
    const int l = threadIdx.x;
    if (l > levels) return;
    if (l == 0) { totals[0] = 1u; return; }
    const unsigned long long *level_keys = level_off + (VR_MAX_LEVELS + 1);          /* keys per level */
    if (l < levels) {
        const unsigned long long last = level_off[l] + level_keys[l] - 1;
        totals[l] = prefix[last] + (masks[last] != 0ull ? 1u : 0u);
    } else {
        const unsigned long long first = level_off[levels - 1], last = first + level_keys[levels - 1] - 1;
        totals[levels] = vox_prefix[last - first] + (uint32_t)__popcll(masks[last]);
    }
}

/* 4. nodes of one level */
__global__ void vr_emit_nodes(const unsigned long long *__restrict__ masks, const uint32_t *__restrict__ prefix,
                              const uint32_t *__restrict__ child_prefix, unsigned nkeys, uint32_t level_start,
                              uint32_t child_start, int leaf_level, vr_node *__restrict__ nodes) {
    const unsigned key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys) return;
    const unsigned long long m = masks[key];
    if (m == 0ull && nkeys != 1u) return;                       /* the root exists even when the map is empty */
    vr_node n;
    n.mask_lo = (uint32_t)m;
    n.mask_hi = (uint32_t)(m >> 32);
    /* inner level: index of the first child node = start of the next level + non-empty masks before key*64;
     * leaf level: index of the brick's first voxel type */
    n.child_base = leaf_level ? child_prefix[key] : child_start + child_prefix[(size_t)key * 64];
    n.aux = vr_node_planes(m);
    reinterpret_cast<uint4 *>(nodes)[level_start + (nkeys == 1u ? 0u : prefix[key])] = *reinterpret_cast<uint4 *>(&n);
}

/* 5. voxel values of the set bits of every leaf brick, in ascending bit order */
__global__ void vr_emit_types(const int8_t *__restrict__ map, int dim, int digits, int half, const unsigned long long *__restrict__ leaf,
                              const uint32_t *__restrict__ vox_prefix, unsigned nkeys, uint8_t *__restrict__ types) {
    const unsigned key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys) return;
    unsigned long long m = leaf[key];
    if (m == 0ull) return;
    unsigned bx = 0, by = 0, bz = 0;
    const int full = half ? digits - 1 : digits;
    for (int j = 0; j < full; j++) {
        const unsigned d = (key >> (6 * j)) & 63u;
        bx |= (d & 3u) << (2 * j);
        by |= ((d >> 2) & 3u) << (2 * j);
        bz |= (d >> 4) << (2 * j);
    }
    if (half && digits > 0) {
        const unsigned d = (key >> (6 * full)) & 7u;
        bx |= (d & 1u) << (2 * full);
        by |= ((d >> 1) & 1u) << (2 * full);
        bz |= (d >> 2) << (2 * full);
    }
    uint32_t at = vox_prefix[key];
    while (m) {
        const int ci = __ffsll((long long)m) - 1;
        m &= m - 1;
        const size_t x = 4 * bx + (ci & 3), y = 4 * by + ((ci >> 2) & 3), z = 4 * bz + (ci >> 4);
        types[at++] = (uint8_t)map[x + (size_t)dim * (y + (size_t)dim * z)];
    }
}

#define VRB(call)                                  \
    do {                                           \
        const cudaError_t e__ = (call);            \
        if (e__ != cudaSuccess) { cleanup(); return e__; } \
    } while (0)

}  // namespace

cudaError_t vr_build_tree_device(const int8_t *d_map, int dim, cudaStream_t stream, vr_device_tree *out,
                                 unsigned long long *launches) {
    if (!d_map || !out || dim < 4 || (dim & (dim - 1))) return cudaErrorInvalidValue;
    int L = 1;
    while ((1 << (2 * L)) < dim) L++;
    if (L > VR_MAX_LEVELS) return cudaErrorInvalidValue;
    const int nb = dim / 4, nquad = nb / 4, digits = L - 1;
    /* keys per level = cells of the map at that level: 64^l, or 8 * 64^(l-1) when the map edge is 2 * 4^(L-1) (only the
     * 2x2x2 root slots inside the map exist: brick_key) -- no workspace for the part of a wider root outside the map */
    const int half = (L > 1 && (1 << (2 * L)) != dim) ? 1 : 0;
    unsigned long long off[2 * (VR_MAX_LEVELS + 1)], *nk = off + (VR_MAX_LEVELS + 1);
    off[0] = 0;
    for (int l = 0; l < L; l++) {
        nk[l] = l == 0 ? 1ull : (half ? 8ull << (6 * (l - 1)) : 1ull << (6 * l));
        off[l + 1] = (off[l] + nk[l] + 3ull) & ~3ull;                                       /* 32-byte aligned */
    }
    const unsigned long long nkeys_all = off[L], nleaf = nk[L - 1];
    if (nleaf > (1ull << 31)) return cudaErrorInvalidValue;

    unsigned long long *masks = nullptr, *d_off = nullptr;
    uint32_t *prefix = nullptr, *vox_prefix = nullptr, *d_totals = nullptr;
    void *scan_tmp = nullptr;
    vr_node *nodes = nullptr;
    uint8_t *types = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    auto cleanup = [&]() {
        cudaFree(masks); cudaFree(d_off); cudaFree(prefix); cudaFree(vox_prefix); cudaFree(d_totals); cudaFree(scan_tmp);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (e2) cudaEventDestroy(e2);
    };
    auto fail_out = [&]() { cudaFree(nodes); cudaFree(types); };

    VRB(cudaMalloc(&masks, nkeys_all * sizeof(unsigned long long)));
    VRB(cudaMalloc(&prefix, nkeys_all * sizeof(uint32_t)));
    VRB(cudaMalloc(&vox_prefix, nleaf * sizeof(uint32_t)));
    VRB(cudaMalloc(&d_totals, (VR_MAX_LEVELS + 1) * sizeof(uint32_t)));
    VRB(cudaMalloc(&d_off, (2 * (VR_MAX_LEVELS + 1) + 1) * sizeof(unsigned long long)));     /* offsets, key counts, one sum */
    VRB(cudaMemcpyAsync(d_off, off, 2 * (VR_MAX_LEVELS + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    size_t tmp_bytes = 0;
    {
        auto it = thrust::make_transform_iterator((const unsigned long long *)masks, PopCount());
        VRB(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, vox_prefix, (int)nleaf, stream));
    }
    VRB(cudaMalloc(&scan_tmp, tmp_bytes));
    VRB(cudaEventCreate(&e0));
    VRB(cudaEventCreate(&e1));
    VRB(cudaEventCreate(&e2));

    VRB(cudaEventRecord(e0, stream));
    if (dim >= 16) {
        const unsigned long long lanes = (unsigned long long)nb * nb * nquad;
        vr_brick_masks<<<(unsigned)((lanes + 127) / 128), 128, 0, stream>>>(d_map, dim, nb, nquad, digits, half, masks + off[L - 1]);
    } else {
        vr_brick_masks_small<<<(nb * nb * nb + 63) / 64, 64, 0, stream>>>(d_map, dim, nb, digits, half, masks + off[L - 1]);
    }
    if (launches) ++*launches;
    VRB(cudaEventRecord(e1, stream));
    for (int l = L - 2; l >= 0; l--) {
        const unsigned nparent = (unsigned)nk[l];
        if (l == 0 && half) vr_reduce_root_half<<<1, 1, 0, stream>>>(masks + off[1], masks + off[0]);
        else vr_reduce_masks<<<(nparent * 32 + 127) / 128, 128, 0, stream>>>(masks + off[l + 1], masks + off[l], nparent);
        if (launches) ++*launches;
    }
    for (int l = 1; l < L; l++) {
        auto it = thrust::make_transform_iterator((const unsigned long long *)(masks + off[l]), NonZero());
        size_t need = tmp_bytes;
        VRB(cub::DeviceScan::ExclusiveSum(scan_tmp, need, it, prefix + off[l], (int)nk[l], stream));
    }
    {
        auto it = thrust::make_transform_iterator((const unsigned long long *)(masks + off[L - 1]), PopCount());
        size_t need = tmp_bytes;
        VRB(cub::DeviceScan::ExclusiveSum(scan_tmp, need, it, vox_prefix, (int)nleaf, stream));
    }
    {
        /* child_base / leaf_types positions are 32-bit: a map with 2^32 or more solid voxels cannot be represented */
        auto it = thrust::make_transform_iterator((const unsigned long long *)(masks + off[L - 1]), PopCount64());
        size_t need = 0;
        unsigned long long *d_sum = d_off + 2 * (VR_MAX_LEVELS + 1);
        VRB(cub::DeviceReduce::Sum(nullptr, need, it, d_sum, (int)nleaf, stream));
        if (need > tmp_bytes) { cudaFree(scan_tmp); scan_tmp = nullptr; tmp_bytes = need; VRB(cudaMalloc(&scan_tmp, tmp_bytes)); }
        need = tmp_bytes;
        VRB(cub::DeviceReduce::Sum(scan_tmp, need, it, d_sum, (int)nleaf, stream));
        unsigned long long total = 0;
        VRB(cudaMemcpyAsync(&total, d_sum, sizeof(total), cudaMemcpyDeviceToHost, stream));
        VRB(cudaStreamSynchronize(stream));
        if (total > 0xFFFFFFFFull) { cleanup(); return cudaErrorInvalidValue; }
    }
    vr_level_totals<<<1, 32, 0, stream>>>(masks, prefix, vox_prefix, d_off, L, d_totals);
    if (launches) ++*launches;
    uint32_t totals[VR_MAX_LEVELS + 1] = {0};
    VRB(cudaMemcpyAsync(totals, d_totals, (L + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VRB(cudaStreamSynchronize(stream));                          /* the output sizes are needed on the host */

    uint32_t start[VR_MAX_LEVELS + 1];
    start[0] = 0;
    for (int l = 0; l < L; l++) start[l + 1] = start[l] + totals[l];
    const uint64_t n_nodes = start[L], solid = totals[L], n_types = solid ? solid : 1;
    {
        cudaError_t e = cudaMalloc(&nodes, n_nodes * sizeof(vr_node));
        if (e == cudaSuccess) e = cudaMalloc(&types, n_types);
        if (e == cudaSuccess && !solid) e = cudaMemsetAsync(types, 0, 1, stream);
        if (e != cudaSuccess) { fail_out(); cleanup(); return e; }
    }
    for (int l = 0; l < L; l++) {
        const unsigned nkl = (unsigned)nk[l];
        const bool leaf = l == L - 1;
        vr_emit_nodes<<<(nkl + 255) / 256, 256, 0, stream>>>(masks + off[l], prefix + off[l], leaf ? vox_prefix : prefix + off[l + 1], nkl,
                                                             start[l], leaf ? 0u : start[l + 1], leaf ? 1 : 0, nodes);
        if (launches) ++*launches;
    }
    if (solid) {
        vr_emit_types<<<(unsigned)((nleaf + 255) / 256), 256, 0, stream>>>(d_map, dim, digits, half, masks + off[L - 1], vox_prefix,
                                                                           (unsigned)nleaf, types);
        if (launches) ++*launches;
    }
    cudaError_t e = cudaEventRecord(e2, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fail_out(); cleanup(); return e; }
    cudaEventElapsedTime(&out->masks_ms, e0, e1);
    cudaEventElapsedTime(&out->total_ms, e0, e2);
    out->nodes = nodes;
    out->types = types;
    out->n_nodes = n_nodes;
    out->n_types = n_types;
    out->solid_voxels = solid;
    out->levels = L;
    cleanup();
    return cudaSuccess;
}

/* ---- top grid of the closed-form walk (vr_types.h: vr_frame_params::grid), from the 64-tree in HBM ----------------
 * Same three steps as vr_native_grid (vr_octree.cpp), which the tests hold this against entry by entry. */
namespace {

/* one thread per block of edge 1 << g: descend to the slot that is the block */
__global__ void vr_grid_classify(const vr_node *nodes, int root_shift, int g, int G, int dim, uint32_t *grid, uint8_t *cell) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)G * G * G) return;
    const int bx = i % G, by = (i / G) % G, bz = i / (G * G);
    const int x = bx << g, y = by << g, z = bz << g;
    uint32_t idx = 0, entry = 0;
    uint8_t cs_out = 0;
    for (int s = root_shift;; s -= 2) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(nodes) + idx);
        if (v.z & VR_NODE_SOLID) { entry = 0x80000000u | idx; break; }                   /* the block lies in a solid node */
        const unsigned long long m = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
        const int ci = ((x >> s) & 3) | (((y >> s) & 3) << 2) | (((z >> s) & 3) << 4);
        if (!((m >> ci) & 1ull)) {
            int cs = s + ((((m >> (ci & 0x2A)) & 0x00330033ull) == 0ull) ? 1 : 0);
            while ((1 << cs) > dim) cs--;
            cs_out = (uint8_t)cs;
            break;
        }
        idx = v.z + (uint32_t)__popcll(m & ((1ull << ci) - 1ull));
        if (s == g) { entry = 0x80000000u | idx; break; }
    }
    grid[i] = entry;
    cell[i] = cs_out;
}

/* erosion step r: an empty block of radius r - 1 whose 26 neighbours all exist and have radius >= r - 1 gets radius r.
 * In place: the test is not affected by neighbours that were already raised to r in this launch. */
__global__ void vr_grid_erode(uint32_t *grid, int G, uint32_t r) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)G * G * G) return;
    const int bx = i % G, by = (i / G) % G, bz = i / (G * G);
    if (bx < 1 || by < 1 || bz < 1 || bx >= G - 1 || by >= G - 1 || bz >= G - 1) return;
    if (grid[i] != r - 1) return;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const uint32_t e = ((volatile uint32_t *)grid)[(unsigned)(bx + dx) + (unsigned)G * ((unsigned)(by + dy) + (unsigned)G * (unsigned)(bz + dz))];
                if (e < r - 1 || (e & 0x80000000u)) return;
            }
    grid[i] = r;
}

__global__ void vr_grid_finish(uint32_t *grid, const uint8_t *cell, int g, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t d = grid[i];
    if (d & 0x80000000u) return;
    grid[i] = (((2u * d + 1u) << g) > (1u << cell[i])) ? ((uint32_t)g | ((d << g) << 8)) : (uint32_t)cell[i];
}

/* ---- directed top grids (vr_octree.cpp: vr_native_grid_directed): eight tables, one per direction octant ----------
 * E[o][b] = edge (blocks, capped) of the largest empty cube with block b in its rear corner that extends along octant o's
 * direction of travel: the 3-D maximal-square recurrence E(b) = min(cap, 1 + min E(b + d)), d in {0,1}^3 \ 0 forward.
 * Every forward neighbour has a larger coordinate sum, so the recurrence is evaluated in ONE pass over the anti-diagonal
 * planes kx + ky + kz = 3(G-1) .. 0 of the mirrored coordinates, one launch per plane for all eight octants (766 small
 * launches at G = 256: ~4 ms; a relaxation "raise every block whose neighbours allow it, 63 times" read the neighbours of
 * every open-sky block in every launch and took 46 ms). */
__global__ void vr_cube_init(const uint32_t *__restrict__ base, unsigned n, uint8_t *__restrict__ E) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t v = (base[i] & 0x80000000u) ? 0 : 1;
#pragma unroll
    for (int o = 0; o < 8; o++) E[(size_t)o * n + i] = v;
}

/* one anti-diagonal plane kx + ky + kz = sum of the mirrored block coordinates, all eight octants: every forward
 * neighbour of a block on it lies on one of the three planes behind (sum + 1 .. sum + 3), which are final */
__global__ void vr_cube_plane(uint8_t *E, int bits, int sum, unsigned cap) {
    const unsigned G = 1u << bits, n = 1u << (3 * bits);
    const unsigned kx = blockIdx.x * blockDim.x + threadIdx.x, ky = blockIdx.y, o = blockIdx.z;
    const int kzs = sum - (int)kx - (int)ky;
    if (kx >= G || kzs < 0 || kzs >= (int)G) return;
    const unsigned kz = (unsigned)kzs;
    const unsigned bx = (o & 1u) ? G - 1u - kx : kx, by = (o & 2u) ? G - 1u - ky : ky, bz = (o & 4u) ? G - 1u - kz : kz;
    uint8_t *Eo = E + (size_t)o * n;
    const unsigned i = bx + (by << bits) + (bz << (2 * bits));
    if (Eo[i] == 0) return;                                              /* a block that holds voxels */
    unsigned mn = cap;
    if (kx == G - 1u || ky == G - 1u || kz == G - 1u) mn = 0;            /* a forward neighbour outside the map */
    else {
        const int sx = (o & 1u) ? -1 : 1, sy = (o & 2u) ? -1 : 1, sz = (o & 4u) ? -1 : 1;
#pragma unroll
        for (int d = 1; d < 8; d++) {
            const unsigned j = (unsigned)((int)bx + ((d & 1) ? sx : 0)) + (((unsigned)((int)by + ((d & 2) ? sy : 0))) << bits) +
                               (((unsigned)((int)bz + ((d & 4) ? sz : 0))) << (2 * bits));
            const unsigned e = Eo[j];
            mn = e < mn ? e : mn;
        }
    }
    Eo[i] = (uint8_t)(mn + 1u > cap ? cap : mn + 1u);
}

__global__ void vr_cube_finish(const uint32_t *__restrict__ base, const uint8_t *__restrict__ cell, const uint8_t *__restrict__ E,
                               int g, int bits, uint32_t *__restrict__ out) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = 1u << (3 * bits), G = 1u << bits;
    if (t >= 8u * n) return;
    const unsigned o = t >> (3 * bits), i = t & (n - 1u);
    const uint32_t b = base[i];
    if (b & 0x80000000u) { out[t] = b; return; }
    const unsigned bx = i & (G - 1u), by = (i >> bits) & (G - 1u), bz = i >> (2 * bits);
    const int kx = (o & 1u) ? (int)(G - 1u - bx) : (int)bx, ky = (o & 2u) ? (int)(G - 1u - by) : (int)by,
              kz = (o & 4u) ? (int)(G - 1u - bz) : (int)bz;
    out[t] = vr_grid_directed_entry(E[t], cell[i], g, kx, ky, kz);
}

}  // namespace

cudaError_t vr_build_grid_device(const vr_node *d_nodes, int levels, int dim, bool directed, cudaStream_t stream, uint32_t **grid_out,
                                 int *grid_shift, int *grid_bits, unsigned long long *launches) {
    const int root_shift = 2 * (levels - 1);
    if (root_shift < 2 || dim < 8) return cudaErrorInvalidValue;
    const int g = vr_grid_shift_for(root_shift, dim);
    const int G = dim >> g;
    int bits = 0;
    while ((1 << bits) < G) bits++;
    if (directed && (1 << bits) != G) return cudaErrorInvalidValue;
    const unsigned n = (unsigned)G * G * G;
    uint32_t *grid = nullptr, *tables = nullptr;
    uint8_t *cell = nullptr, *E = nullptr;
    cudaError_t e = cudaMalloc(&grid, (size_t)n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&cell, n);
    if (e == cudaSuccess && directed) e = cudaMalloc(&E, (size_t)8 * n);
    if (e == cudaSuccess && directed) e = cudaMalloc(&tables, (size_t)8 * n * sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(grid); cudaFree(cell); cudaFree(E); cudaFree(tables); return e; }
    const unsigned blocks = (n + 255) / 256;
    vr_grid_classify<<<blocks, 256, 0, stream>>>(d_nodes, root_shift, g, G, dim, grid, cell);
    if (directed) {
        const unsigned blocks8 = (unsigned)(((size_t)8 * n + 255) / 256);
        const unsigned cap = (unsigned)G < VR_GRID_MAX_CUBE ? (unsigned)G : VR_GRID_MAX_CUBE;
        vr_cube_init<<<blocks, 256, 0, stream>>>(grid, n, E);
        /* E(b) = min(cap, 1 + min over the 7 forward neighbours): one pass, plane by plane from the far corner */
        const dim3 pgrid(((unsigned)G + 63u) / 64u, (unsigned)G, 8u);
        for (int sum = 3 * (G - 1); sum >= 0; sum--) vr_cube_plane<<<pgrid, 64, 0, stream>>>(E, bits, sum, cap);
        vr_cube_finish<<<blocks8, 256, 0, stream>>>(grid, cell, E, g, bits, tables);
        if (launches) *launches += 3 + (unsigned long long)(3 * (G - 1) + 1);
    } else {
        const int rmax = G / 2 < 63 ? G / 2 : 63;                        /* a cube of radius r inside the grid needs G >= 2r + 1 */
        for (int r = 1; r <= rmax; r++) vr_grid_erode<<<blocks, 256, 0, stream>>>(grid, G, (uint32_t)r);
        vr_grid_finish<<<blocks, 256, 0, stream>>>(grid, cell, g, n);
        if (launches) *launches += 2 + (rmax > 0 ? rmax : 0);
    }
    e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(cell);
    cudaFree(E);
    if (directed) { cudaFree(grid); grid = tables; }
    if (e != cudaSuccess) { cudaFree(grid); return e; }
    *grid_out = grid;
    *grid_shift = g;
    *grid_bits = bits;
    return cudaSuccess;
}

/* ---- solid-subtree collapse on the device (vr_types.h: VR_NODE_SOLID; host version vr_octree.cpp:
 * vr_native_collapse_solid, which this reproduces array for array) --------------------------------------------------
 * Input: a complete tree in BFS order as the two builders emit it.  (1) bottom-up, per level: the type of a node's cube
 * if it is solid (leaf: 64 set voxels of one type; inner: 64 solid children of one type), else 0xFF; (2) top-down: a node
 * is kept unless its parent is solid or was dropped; (3) exclusive scans of "kept" and of the voxel types kept leaves
 * still need; (4) re-pack in the same order. */
namespace {

__global__ void vr_solid_leaves(const vr_node *__restrict__ nodes, const uint8_t *__restrict__ types, unsigned lo, unsigned hi, uint8_t *__restrict__ st) {
    const unsigned i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint4 v = reinterpret_cast<const uint4 *>(nodes)[i];
    uint8_t r = 0xFF;
    if ((v.x & v.y) == 0xFFFFFFFFu) {
        const uint8_t *ty = types + v.z;
        r = ty[0];
        for (int k = 1; k < 64; k++)
            if (ty[k] != r) { r = 0xFF; break; }
    }
    st[i] = r;
}

__global__ void vr_solid_inner(const vr_node *__restrict__ nodes, unsigned lo, unsigned hi, uint8_t *st) {
    const unsigned i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint4 v = reinterpret_cast<const uint4 *>(nodes)[i];
    uint8_t r = 0xFF;
    if ((v.x & v.y) == 0xFFFFFFFFu) {
        r = st[v.z];
        for (int k = 1; k < 64 && r != 0xFF; k++)
            if (st[v.z + k] != r) r = 0xFF;
    }
    st[i] = r;
}

__global__ void vr_solid_keep(const vr_node *__restrict__ nodes, const uint8_t *__restrict__ st, unsigned lo, unsigned hi, uint32_t *keep) {
    const unsigned i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint4 v = reinterpret_cast<const uint4 *>(nodes)[i];
    const uint32_t k = (keep[i] && st[i] == 0xFF) ? 1u : 0u;
    const int pc = __popc(v.x) + __popc(v.y);
    for (int j = 0; j < pc; j++) keep[v.z + j] = k;
}

/* voxel types a node still needs after the collapse: those of a kept leaf that is not solid */
__global__ void vr_solid_type_counts(const vr_node *__restrict__ nodes, const uint8_t *__restrict__ st, const uint32_t *__restrict__ keep, unsigned leaf_lo,
                                     unsigned n, uint32_t *__restrict__ cnt) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 0;
    if (i >= leaf_lo && keep[i] && st[i] == 0xFF) {
        const uint4 v = reinterpret_cast<const uint4 *>(nodes)[i];
        c = (uint32_t)(__popc(v.x) + __popc(v.y));
    }
    cnt[i] = c;
}

__global__ void vr_solid_emit(const vr_node *__restrict__ nodes, const uint8_t *__restrict__ types, const uint8_t *__restrict__ st,
                              const uint32_t *__restrict__ keep, const uint32_t *__restrict__ at, const uint32_t *__restrict__ tat, unsigned leaf_lo,
                              unsigned n, vr_node *__restrict__ out_nodes, uint8_t *__restrict__ out_types) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    uint4 v = reinterpret_cast<const uint4 *>(nodes)[i];
    if (st[i] != 0xFF) {
        v.z = VR_NODE_SOLID | (uint32_t)st[i];
    } else if (i >= leaf_lo) {
        const int pc = __popc(v.x) + __popc(v.y);
        const uint8_t *src = types + v.z;
        uint8_t *dst = out_types + tat[i];
        for (int j = 0; j < pc; j++) dst[j] = src[j];
        v.z = tat[i];
    } else if (v.x | v.y) {
        v.z = at[v.z];
    }
    reinterpret_cast<uint4 *>(out_nodes)[at[i]] = v;
}

}  // namespace

cudaError_t vr_collapse_solid_device(vr_device_tree *t, cudaStream_t stream, unsigned long long *launches) {
    if (!t || !t->nodes || t->levels < 1 || t->n_nodes < 1 || t->n_nodes >= (1ull << 31)) return cudaErrorInvalidValue;
    const int L = t->levels;
    const unsigned n = (unsigned)t->n_nodes;
    /* level l = nodes [start[l], start[l + 1]): the children of the first node of an inner level open the next one */
    unsigned start[VR_MAX_LEVELS + 1] = {0};
    start[1] = 1;
    for (int l = 1; l < L; l++) {
        if (l == L - 1) { start[l + 1] = n; break; }
        if (start[l] >= n) { start[l + 1] = n; continue; }
        vr_node first;
        cudaError_t e = cudaMemcpyAsync(&first, t->nodes + start[l], sizeof(vr_node), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        start[l + 1] = first.child_base;
        if (start[l + 1] < start[l] || start[l + 1] > n) return cudaErrorInvalidValue;
    }
    if (L == 1) start[1] = n;
    if (start[L] != n) return cudaErrorInvalidValue;
    const unsigned leaf_lo = start[L - 1];
    if (leaf_lo >= n) return cudaSuccess;                               /* no leaf level: an empty map (a root without children) */

    uint8_t *st = nullptr, *types = nullptr;
    uint32_t *keep = nullptr, *at = nullptr, *cnt = nullptr, *tat = nullptr, *d_solid = nullptr;
    vr_node *nodes = nullptr;
    void *tmp = nullptr;
    auto cleanup = [&]() { cudaFree(st); cudaFree(keep); cudaFree(at); cudaFree(cnt); cudaFree(tat); cudaFree(d_solid); cudaFree(tmp); };
    auto fail_out = [&]() { cudaFree(nodes); cudaFree(types); };
    VRB(cudaMalloc(&st, n));
    VRB(cudaMalloc(&keep, (size_t)(n + 1) * sizeof(uint32_t)));
    VRB(cudaMalloc(&at, (size_t)(n + 1) * sizeof(uint32_t)));
    VRB(cudaMalloc(&cnt, (size_t)(n + 1) * sizeof(uint32_t)));
    VRB(cudaMalloc(&tat, (size_t)(n + 1) * sizeof(uint32_t)));
    VRB(cudaMalloc(&d_solid, sizeof(uint32_t)));
    size_t tmp_bytes = 0, b = 0;
    VRB(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, keep, at, (int)(n + 1), stream));
    {
        auto it = thrust::make_transform_iterator((const uint8_t *)st, IsSolidFlag());
        VRB(cub::DeviceReduce::Sum(nullptr, b, it, d_solid, (int)n, stream));
        if (b > tmp_bytes) tmp_bytes = b;
    }
    VRB(cudaMalloc(&tmp, tmp_bytes));
    /* (1) */
    vr_solid_leaves<<<(n - leaf_lo + 255) / 256, 256, 0, stream>>>(t->nodes, t->types, leaf_lo, n, st);
    for (int l = L - 2; l >= 0; l--)
        if (start[l + 1] > start[l]) vr_solid_inner<<<(start[l + 1] - start[l] + 255) / 256, 256, 0, stream>>>(t->nodes, start[l], start[l + 1], st);
    uint32_t solid = 0;
    {
        auto it = thrust::make_transform_iterator((const uint8_t *)st, IsSolidFlag());
        b = tmp_bytes;
        VRB(cub::DeviceReduce::Sum(tmp, b, it, d_solid, (int)n, stream));
        VRB(cudaMemcpyAsync(&solid, d_solid, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        VRB(cudaStreamSynchronize(stream));
    }
    if (launches) *launches += (unsigned long long)L;
    if (!solid) { cleanup(); return cudaSuccess; }                       /* nothing to collapse: the tree stays as it is */
    /* (2) */
    VRB(cudaMemsetAsync(keep, 0, (size_t)(n + 1) * sizeof(uint32_t), stream));
    {
        const uint32_t one = 1;
        VRB(cudaMemcpyAsync(keep, &one, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    }
    for (int l = 0; l + 1 < L; l++)
        if (start[l + 1] > start[l]) vr_solid_keep<<<(start[l + 1] - start[l] + 255) / 256, 256, 0, stream>>>(t->nodes, st, start[l], start[l + 1], keep);
    /* (3) */
    b = tmp_bytes;
    VRB(cub::DeviceScan::ExclusiveSum(tmp, b, keep, at, (int)(n + 1), stream));
    VRB(cudaMemsetAsync(cnt + n, 0, sizeof(uint32_t), stream));
    vr_solid_type_counts<<<(n + 255) / 256, 256, 0, stream>>>(t->nodes, st, keep, leaf_lo, n, cnt);
    b = tmp_bytes;
    VRB(cub::DeviceScan::ExclusiveSum(tmp, b, cnt, tat, (int)(n + 1), stream));
    uint32_t kept = 0, ntypes = 0;
    VRB(cudaMemcpyAsync(&kept, at + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VRB(cudaMemcpyAsync(&ntypes, tat + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VRB(cudaStreamSynchronize(stream));
    /* (4) */
    {
        cudaError_t e = cudaMalloc(&nodes, (size_t)kept * sizeof(vr_node));
        if (e == cudaSuccess) e = cudaMalloc(&types, ntypes ? ntypes : 1);
        if (e == cudaSuccess && !ntypes) e = cudaMemsetAsync(types, 0, 1, stream);
        if (e != cudaSuccess) { fail_out(); cleanup(); return e; }
    }
    vr_solid_emit<<<(n + 255) / 256, 256, 0, stream>>>(t->nodes, t->types, st, keep, at, tat, leaf_lo, n, nodes, types);
    if (launches) *launches += (unsigned long long)L + 2;
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fail_out(); cleanup(); return e; }
    cudaFree(t->nodes);
    cudaFree(t->types);
    t->nodes = nodes;
    t->types = types;
    t->n_nodes = kept;
    t->n_types = ntypes ? ntypes : 1;
    cleanup();
    return cudaSuccess;
}

/* ---- 64-tree from a column description, built in HBM without the N^3 volume and without a dense workspace ----------
 * Column (x, y) is solid for lo[x + N y] <= z <= hi[x + N y] (vr_assign_columns; 4096^3 = 64 GiB dense cannot exist).
 * Only the bricks that can hold a voxel are ever materialised ("streaming SVO builder ... without a dense
 * array"):
 *   1. vr_col_count     per brick column (4x4 voxel columns): the range of brick layers its columns touch;
 *   2. exclusive scan   -> where the brick column's (key, mask) pairs go;
 *   3. vr_col_bricks    per brick column: the 64-bit occupancy mask of every touched brick + its hierarchical key;
 *   4. radix sort by key (empty masks get the key ~0 and fall off the end): ascending key = breadth-first order;
 *   5. per level, bottom up: reduce-by-key over key >> 6 -- OR of the child bits = the parent's mask, MIN of the child
 *      positions = where its children start -- then vr_pair_nodes writes the level's nodes.
 * Emits exactly the arrays vr_native_from_columns (vr_octree.cpp) produces. */
namespace {

struct KeyParent {
    __host__ __device__ uint32_t operator()(uint32_t k) const { return k >> 6; }
};
struct KeyBit {
    __host__ __device__ unsigned long long operator()(uint32_t k) const { return 1ull << (k & 63u); }
};
struct BitOr {
    __host__ __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a | b; }
};
struct MinU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; }
};

__device__ __forceinline__ void col_range(const int32_t *lo, const int32_t *hi, int dim, unsigned bx, unsigned by, int &zmin, int &zmax) {
    zmin = 0x7fffffff; zmax = -1;
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            const size_t c = (size_t)(4 * bx + x) + (size_t)dim * (size_t)(4 * by + y);
            int a = lo[c], b = hi[c];
            a = a < 0 ? 0 : a;
            b = b > dim - 1 ? dim - 1 : b;
            if (a > b) continue;
            zmin = a < zmin ? a : zmin;
            zmax = b > zmax ? b : zmax;
        }
}

__global__ void vr_col_count(const int32_t *__restrict__ lo, const int32_t *__restrict__ hi, int dim, int nb, uint32_t *__restrict__ count) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)(nb * nb)) return;
    int zmin, zmax;
    col_range(lo, hi, dim, i % (unsigned)nb, i / (unsigned)nb, zmin, zmax);
    count[i] = zmax < 0 ? 0u : (uint32_t)((zmax >> 2) - (zmin >> 2) + 1);
}

__global__ void vr_col_bricks(const int32_t *__restrict__ lo, const int32_t *__restrict__ hi, int dim, int nb, int digits,
                              const uint32_t *__restrict__ offset, uint32_t *__restrict__ keys, unsigned long long *__restrict__ masks) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)(nb * nb)) return;
    const unsigned bx = i % (unsigned)nb, by = i / (unsigned)nb;
    int zmin, zmax;
    col_range(lo, hi, dim, bx, by, zmin, zmax);
    if (zmax < 0) return;
    int a[16], b[16];
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            const size_t c = (size_t)(4 * bx + x) + (size_t)dim * (size_t)(4 * by + y);
            a[x + 4 * y] = lo[c] < 0 ? 0 : lo[c];
            b[x + 4 * y] = hi[c] > dim - 1 ? dim - 1 : hi[c];
        }
    uint32_t at = offset[i];
    for (int bz = zmin >> 2; bz <= (zmax >> 2); bz++, at++) {
        unsigned long long m = 0ull;
        for (int c = 0; c < 16; c++) {
            const int z0 = a[c] > 4 * bz ? a[c] : 4 * bz, z1 = b[c] < 4 * bz + 3 ? b[c] : 4 * bz + 3;
            for (int z = z0; z <= z1; z++) m |= 1ull << (c + 16 * (z - 4 * bz));
        }
        keys[at] = m ? brick_key(bx, by, (unsigned)bz, digits, 0) : 0xffffffffu;
        masks[at] = m;
    }
}

/* nodes of one level from its sorted (key, mask) arrays; first = per node the position of its first child in the next
 * level (inner levels) or of its first voxel type (leaf level) */
__global__ void vr_pair_nodes(const unsigned long long *__restrict__ masks, const uint32_t *__restrict__ first, unsigned n,
                              uint32_t level_start, uint32_t child_start, vr_node *__restrict__ nodes) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long m = masks[i];
    vr_node nd;
    nd.mask_lo = (uint32_t)m;
    nd.mask_hi = (uint32_t)(m >> 32);
    nd.child_base = child_start + first[i];
    nd.aux = vr_node_planes(m);
    reinterpret_cast<uint4 *>(nodes)[level_start + i] = *reinterpret_cast<uint4 *>(&nd);
}

}  // namespace

cudaError_t vr_build_tree_columns_device(const int32_t *d_lo, const int32_t *d_hi, int dim, uint8_t type, cudaStream_t stream,
                                         vr_device_tree *out, unsigned long long *launches) {
    if (!d_lo || !d_hi || !out || dim < 4 || (dim & (dim - 1))) return cudaErrorInvalidValue;
    int L = 1;
    while ((1 << (2 * L)) < dim) L++;
    if (L > 6) return cudaErrorInvalidValue;                     /* 32-bit keys: 6 bits per level below the root, up to 4096^3 */
    const int nb = dim / 4, digits = L - 1;
    const unsigned ncol = (unsigned)nb * (unsigned)nb;

    uint32_t *count = nullptr, *offset = nullptr, *keys[2] = {nullptr, nullptr}, *first[2] = {nullptr, nullptr}, *pkeys = nullptr, *d_num = nullptr;
    unsigned long long *masks[2] = {nullptr, nullptr}, *pmasks[VR_MAX_LEVELS] = {};
    uint32_t *pfirst[VR_MAX_LEVELS] = {}, *lkeys[VR_MAX_LEVELS] = {};
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    vr_node *nodes = nullptr;
    uint8_t *types = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(count); cudaFree(offset); cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(masks[0]); cudaFree(masks[1]);
        cudaFree(first[0]); cudaFree(first[1]); cudaFree(pkeys); cudaFree(d_num); cudaFree(tmp);
        for (int l = 0; l < VR_MAX_LEVELS; l++) { cudaFree(pmasks[l]); cudaFree(pfirst[l]); cudaFree(lkeys[l]); }
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    auto need_tmp = [&](size_t bytes) -> cudaError_t {
        if (bytes <= tmp_bytes) return cudaSuccess;
        cudaFree(tmp);
        tmp = nullptr;
        tmp_bytes = bytes;
        return cudaMalloc(&tmp, bytes);
    };
    VRB(cudaEventCreate(&e0));
    VRB(cudaEventCreate(&e1));
    VRB(cudaEventRecord(e0, stream));
    /* 1-2. brick layers per brick column and where they go */
    VRB(cudaMalloc(&count, (size_t)(ncol + 1) * sizeof(uint32_t)));
    VRB(cudaMalloc(&offset, (size_t)(ncol + 1) * sizeof(uint32_t)));
    VRB(cudaMemsetAsync(count + ncol, 0, sizeof(uint32_t), stream));
    vr_col_count<<<(ncol + 255) / 256, 256, 0, stream>>>(d_lo, d_hi, dim, nb, count);
    {
        size_t b = 0;
        VRB(cub::DeviceScan::ExclusiveSum(nullptr, b, count, offset, (int)(ncol + 1), stream));
        VRB(need_tmp(b));
        VRB(cub::DeviceScan::ExclusiveSum(tmp, b, count, offset, (int)(ncol + 1), stream));
    }
    uint32_t npairs = 0;
    VRB(cudaMemcpyAsync(&npairs, offset + ncol, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VRB(cudaStreamSynchronize(stream));
    if (launches) *launches += 1;
    const size_t cap = npairs ? npairs : 1;
    for (int i = 0; i < 2; i++) {
        VRB(cudaMalloc(&keys[i], cap * sizeof(uint32_t)));
        VRB(cudaMalloc(&masks[i], cap * sizeof(unsigned long long)));
        VRB(cudaMalloc(&first[i], cap * sizeof(uint32_t)));
    }
    VRB(cudaMalloc(&pkeys, cap * sizeof(uint32_t)));
    VRB(cudaMalloc(&d_num, sizeof(uint32_t)));
    uint32_t nleaf = 0;
    if (npairs) {
        /* 3-4. the bricks, then breadth-first (= ascending key) order; empty masks sort to the end */
        vr_col_bricks<<<(ncol + 127) / 128, 128, 0, stream>>>(d_lo, d_hi, dim, nb, digits, offset, keys[0], masks[0]);
        size_t b = 0;
        VRB(cub::DeviceRadixSort::SortPairs(nullptr, b, keys[0], keys[1], masks[0], masks[1], (int)npairs, 0, 32, stream));
        VRB(need_tmp(b));
        VRB(cub::DeviceRadixSort::SortPairs(tmp, b, keys[0], keys[1], masks[0], masks[1], (int)npairs, 0, 32, stream));
        auto nz = thrust::make_transform_iterator((const unsigned long long *)masks[1], NonZero());
        VRB(cub::DeviceReduce::Sum(nullptr, b, nz, d_num, (int)npairs, stream));
        VRB(need_tmp(b));
        VRB(cub::DeviceReduce::Sum(tmp, b, nz, d_num, (int)npairs, stream));
        VRB(cudaMemcpyAsync(&nleaf, d_num, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        VRB(cudaStreamSynchronize(stream));
        if (launches) *launches += 3;
    }
    /* 5. the levels, bottom up.  Level l (0 = root) has n[l] nodes; lkeys / pmasks / pfirst hold its keys, masks and the
     * position of each node's first child (leaf level: first voxel type) */
    uint32_t n[VR_MAX_LEVELS + 1] = {0};
    const int leaf = L - 1;
    n[leaf] = nleaf;
    unsigned long long solid = 0;
    if (nleaf) {
        VRB(cudaMalloc(&pmasks[leaf], (size_t)nleaf * sizeof(unsigned long long)));
        VRB(cudaMalloc(&pfirst[leaf], (size_t)nleaf * sizeof(uint32_t)));
        VRB(cudaMalloc(&lkeys[leaf], (size_t)nleaf * sizeof(uint32_t)));
        VRB(cudaMemcpyAsync(pmasks[leaf], masks[1], (size_t)nleaf * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
        VRB(cudaMemcpyAsync(lkeys[leaf], keys[1], (size_t)nleaf * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
        auto pc = thrust::make_transform_iterator((const unsigned long long *)pmasks[leaf], PopCount());
        size_t b = 0;
        VRB(cub::DeviceScan::ExclusiveSum(nullptr, b, pc, pfirst[leaf], (int)nleaf, stream));
        VRB(need_tmp(b));
        VRB(cub::DeviceScan::ExclusiveSum(tmp, b, pc, pfirst[leaf], (int)nleaf, stream));
        uint32_t last_first = 0;
        unsigned long long last_mask = 0;
        VRB(cudaMemcpyAsync(&last_first, pfirst[leaf] + nleaf - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        VRB(cudaMemcpyAsync(&last_mask, pmasks[leaf] + nleaf - 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        VRB(cudaStreamSynchronize(stream));
        solid = (unsigned long long)last_first + (unsigned long long)__builtin_popcountll(last_mask);
        for (int l = leaf - 1; l >= 0; l--) {
            const uint32_t nc = n[l + 1];
            auto parent = thrust::make_transform_iterator((const uint32_t *)lkeys[l + 1], KeyParent());
            auto bit = thrust::make_transform_iterator((const uint32_t *)lkeys[l + 1], KeyBit());
            thrust::counting_iterator<uint32_t> pos(0u);
            VRB(cudaMalloc(&pmasks[l], (size_t)nc * sizeof(unsigned long long)));
            VRB(cudaMalloc(&pfirst[l], (size_t)nc * sizeof(uint32_t)));
            VRB(cudaMalloc(&lkeys[l], (size_t)nc * sizeof(uint32_t)));
            size_t b1 = 0, b2 = 0;
            VRB(cub::DeviceReduce::ReduceByKey(nullptr, b1, parent, lkeys[l], bit, pmasks[l], d_num, BitOr(), (int)nc, stream));
            VRB(cub::DeviceReduce::ReduceByKey(nullptr, b2, parent, pkeys, pos, pfirst[l], d_num, MinU32(), (int)nc, stream));
            VRB(need_tmp(b1 > b2 ? b1 : b2));
            VRB(cub::DeviceReduce::ReduceByKey(tmp, b1, parent, lkeys[l], bit, pmasks[l], d_num, BitOr(), (int)nc, stream));
            VRB(cub::DeviceReduce::ReduceByKey(tmp, b2, parent, pkeys, pos, pfirst[l], d_num, MinU32(), (int)nc, stream));
            VRB(cudaMemcpyAsync(&n[l], d_num, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            VRB(cudaStreamSynchronize(stream));
            if (launches) *launches += 2;
        }
    }
    /* output arrays */
    uint32_t start[VR_MAX_LEVELS + 1];
    const bool empty = nleaf == 0;
    if (empty) for (int l = 0; l < L; l++) n[l] = l == 0 ? 1u : 0u;
    start[0] = 0;
    for (int l = 0; l < L; l++) start[l + 1] = start[l] + n[l];
    const uint64_t n_nodes = start[L], n_types = solid ? solid : 1;
    {
        cudaError_t e = cudaMalloc(&nodes, n_nodes * sizeof(vr_node));
        if (e == cudaSuccess) e = cudaMalloc(&types, n_types);
        if (e == cudaSuccess) e = cudaMemsetAsync(types, solid ? (int)type : 0, n_types, stream);
        if (e == cudaSuccess && empty) e = cudaMemsetAsync(nodes, 0, sizeof(vr_node), stream);     /* an empty root (planes: all empty) */
        if (e != cudaSuccess) { cudaFree(nodes); cudaFree(types); cleanup(); return e; }
    }
    if (empty) {
        vr_node root = {0u, 0u, L > 1 ? 1u : 0u, vr_node_planes(0ull)};      /* child_base as the host builder leaves it */
        cudaError_t e = cudaMemcpyAsync(nodes, &root, sizeof(root), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { cudaFree(nodes); cudaFree(types); cleanup(); return e; }
    } else {
        for (int l = 0; l < L; l++) {
            vr_pair_nodes<<<(n[l] + 255) / 256, 256, 0, stream>>>(pmasks[l], pfirst[l], n[l], start[l], l == leaf ? 0u : start[l + 1], nodes);
            if (launches) ++*launches;
        }
    }
    cudaError_t e = cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(nodes); cudaFree(types); cleanup(); return e; }
    out->masks_ms = 0.f;
    cudaEventElapsedTime(&out->total_ms, e0, e1);
    out->nodes = nodes;
    out->types = types;
    out->n_nodes = n_nodes;
    out->n_types = n_types;
    out->solid_voxels = solid;
    out->levels = L;
    cleanup();
    return cudaSuccess;
}
