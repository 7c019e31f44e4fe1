/*
 * vr_kernels.cu -- sm_100a kernels of the B200 voxel ray caster.
 *
 * Replaces clEnqueueNDRangeKernel("raycaster", global=(W,H), local=NULL)
 * (reference src/CLCaster.cpp:946-987) and the OpenCL kernel it runs
 * (kernels/ray_caster_kernel.cl:256).  The per-pixel logic lives in vr_trace.h.
 *
 * Pixel -> thread mapping: a CTA of 128 threads covers a 32x4 pixel tile, each warp an 8x4 block
 * of it, so that the 32 rays of a warp form a compact bundle (coherent DDA trip counts and octree
 * paths) while ray-table loads (16 B/pixel) and RGBA8 stores (4 B/pixel) still fill whole
 * 128-byte / 32-byte sectors.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "vr_kernels.h"
#include "vr_trace.h"
#include "vr_canon.h"

namespace {

constexpr int kTileW = 32;
#ifndef VR_TILE_H
#define VR_TILE_H 4         /* rows of a CTA tile: CTA = 32 x VR_TILE_H pixels, one warp per 8x4 block (4 warps) */
#endif
constexpr int kTileH = VR_TILE_H;
constexpr int kThreads = kTileW * kTileH;
#ifndef VR_SVO_MIN_CTAS
#define VR_SVO_MIN_CTAS 8      /* CTAs per SM the register allocation of vr_svo_kernel targets: 8 x 4 warps, 64 registers
                                * (measured against 3 x 8 warps / 80 registers and 7 x 4 / 72: DESIGN.md section 5) */
#endif

/* thread -> pixel inside the CTA tile: warp w = (w&3, w>>2) of 8x4 blocks, lane = (l&7, l>>3) */
__device__ __forceinline__ void tile_xy(int tid, int &lx, int &ly) {
    const int warp = tid >> 5, lane = tid & 31;
    lx = ((warp & 3) << 3) | (lane & 7);
    ly = ((warp >> 2) << 2) | (lane >> 3);
}

/* local slab row -> frame row (multi-GPU row-band interleave, see vr_types.h) */
__device__ __forceinline__ int frame_row(const vr_frame_params &P, int ly) {
    if (P.band_stride == 1) return ly + P.band_first * P.band_rows;      /* one rank: no integer division per thread */
    const int lb = ly / P.band_rows;
    return (lb * P.band_stride + P.band_first) * P.band_rows + (ly - lb * P.band_rows);
}

/* CTA -> pixel of this thread.  Returns false when the thread has no pixel.  Band mode (default): the grid covers the
 * rank's compact slab, `local` addresses the slab, y is the frame row.  Tile mode (P.tile_world > 1): the grid covers
 * ceil(tiles_x / world) x tiles_y CTAs, CTA (k, ty) is tile tx = ((rank - ty) mod world) + k * world of tile row ty, and
 * `local` addresses the full frame: the pixel is written where it belongs, possibly in a peer GPU's memory. */
__device__ __forceinline__ bool cta_pixel(const vr_frame_params &P, int &x, int &y, size_t &local) {
    int lx, ly;
    tile_xy(threadIdx.x, lx, ly);
    if (P.tile_world > 1) {
        int first = (P.tile_rank - (int)blockIdx.y) % P.tile_world;
        first += first < 0 ? P.tile_world : 0;
        x = (first + (int)blockIdx.x * P.tile_world) * kTileW + lx;
        y = (int)blockIdx.y * kTileH + ly;
        if (x >= P.width || y >= P.height) return false;
        local = (size_t)x + (size_t)P.width * (size_t)y;
        return true;
    }
    x = blockIdx.x * kTileW + lx;
    const int row = blockIdx.y * kTileH + ly;
    if (x >= P.width || row >= P.local_rows) return false;
    y = frame_row(P, row);
    if (y >= P.height) return false;
    local = (size_t)x + (size_t)P.width * (size_t)row;
    return true;
}

struct SmemStack {
    uint32_t *base;   /* &stack[0][tid]; level stride = kThreads */
    __device__ __forceinline__ void set(int level, uint32_t v) { base[level * kThreads] = v; }
    __device__ __forceinline__ uint32_t get(int level) const { return base[level * kThreads]; }
};

/* the same stack addressed through a 32-bit shared-memory address held in a register (the generic pointer above is
 * re-derived from %tid and the CTA's shared window at every access) */
struct SmemStackA {
    uint32_t addr;
    __device__ __forceinline__ void set(int level, uint32_t v) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr + (uint32_t)level * (kThreads * 4)), "r"(v));
    }
    __device__ __forceinline__ uint32_t get(int level) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr + (uint32_t)level * (kThreads * 4)));
        return v;
    }
};
#ifndef VR_CANON_MIN_CTAS
#define VR_CANON_MIN_CTAS 8
#endif

/* MULTI = the multi-light extension (LIGHT_COUNT > 1): a separate instantiation, so that the reference-parity
 * kernels do not carry its per-ray state */
template <bool AUX, bool MULTI>
__global__ void __launch_bounds__(kThreads)
vr_dense_kernel(const __grid_constant__ vr_frame_params P) {
    int x, y;
    size_t local;
    if (!cta_pixel(P, x, y, local)) return;
    uint32_t rgba;
    vr_aux a;
    const bool write = vr_trace_dense<AUX, MULTI>(P, x, y, &rgba, &a);
    if (write) __stcs(reinterpret_cast<uint32_t *>(P.image) + local, rgba);       /* written once, not read by this kernel: streaming */
    if (AUX) reinterpret_cast<uint4 *>(P.aux)[2 * local] = *reinterpret_cast<uint4 *>(&a),
             reinterpret_cast<uint4 *>(P.aux)[2 * local + 1] = *(reinterpret_cast<uint4 *>(&a) + 1);
}

template <bool AUX, int WALK, bool MULTI>
__global__ void __launch_bounds__(kThreads, WALK == 2 ? VR_CANON_MIN_CTAS : VR_SVO_MIN_CTAS)
vr_svo_kernel(const __grid_constant__ vr_frame_params P) {
    __shared__ uint32_t stack[VR_MAX_LEVELS * kThreads];
    int x, y;
    size_t local;
    if (!cta_pixel(P, x, y, local)) return;
    uint32_t rgba;
    vr_aux a;
    bool write;
    if constexpr (WALK == 2) {
        SmemStackA stk{(uint32_t)__cvta_generic_to_shared(stack + threadIdx.x)};
        VR_PIN(stk.addr);
        write = vr_trace_svo_canon<AUX, MULTI>(P, x, y, &rgba, &a, stk);
    } else {
        SmemStack stk{stack + threadIdx.x};
        write = vr_trace_svo<AUX, WALK, MULTI>(P, x, y, &rgba, &a, stk);
    }
    if (write) __stcs(reinterpret_cast<uint32_t *>(P.image) + local, rgba);       /* written once, not read by this kernel: streaming */
    if (AUX) reinterpret_cast<uint4 *>(P.aux)[2 * local] = *reinterpret_cast<uint4 *>(&a),
             reinterpret_cast<uint4 *>(P.aux)[2 * local + 1] = *(reinterpret_cast<uint4 *>(&a) + 1);
}

/* Persistent-warp variant (Aila/Laine style): the grid is sized to the machine (ctas_per_sm * #SM CTAs), every
 * warp pulls pixels from a global counter.  When at least `refill_min` lanes of a warp have finished their
 * pixel, ONE lane reserves that many new pixels with a single atomicAdd and the reservation is handed out to
 * the idle lanes with __ballot_sync / __shfl_sync / popc-prefix, so lanes never idle behind the slowest ray of
 * their tile.  Pixel order: index -> 8x4 tile -> 32x8 super tile (same as the static kernel), so one refill
 * batch is a compact screen block.  WALK = the in-cell walk (0 merged, 1 per-axis), as in vr_svo_kernel. */
template <bool AUX, int WALK>
__global__ void __launch_bounds__(kThreads)
vr_svo_persistent_kernel(const __grid_constant__ vr_frame_params P, unsigned int *counter, int refill_min) {
    __shared__ uint32_t stack[VR_MAX_LEVELS * kThreads];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned stiles_x = (unsigned)(P.width + kTileW - 1) / kTileW;
    const unsigned stiles_y = (unsigned)(P.local_rows + kTileH - 1) / kTileH;
    const unsigned total = stiles_x * stiles_y * (unsigned)kThreads;
    vr_svo_ray<SmemStack> q;
    q.stk = SmemStack{stack + threadIdx.x};
    vr_aux a;
    bool active = false, more = true;
    size_t local = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu || (more && __popc(idle) >= refill_min)) {
            if (!more) break;                                   /* every lane idle and the counter is exhausted */
            const int nidle = __popc(idle);
            const int leader = __ffs(idle) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(counter, (unsigned)nidle);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + (unsigned)nidle >= total) more = false;
            if (!active) {
                const unsigned idx = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (idx < total) {
                    const unsigned st = idx / (unsigned)kThreads, in = idx % (unsigned)kThreads;
                    int lx, ly;
                    tile_xy((int)in, lx, ly);
                    const int x = (int)(st % stiles_x) * kTileW + lx;
                    const int row = (int)(st / stiles_x) * kTileH + ly;
                    if (x < P.width && row < P.local_rows) {
                        const int y = frame_row(P, row);
                        if (y < P.height) {
                            local = (size_t)x + (size_t)P.width * (size_t)row;
                            active = vr_svo_begin<AUX>(P, x, y, q, &a);
                            if (AUX && !active) {
                                reinterpret_cast<uint4 *>(P.aux)[2 * local] = *reinterpret_cast<uint4 *>(&a);
                                reinterpret_cast<uint4 *>(P.aux)[2 * local + 1] = *(reinterpret_cast<uint4 *>(&a) + 1);
                            }
                        }
                    }
                }
            }
            if (idle == 0xffffffffu && !more && !__any_sync(0xffffffffu, active)) break;
        }
        if (active) {
            const int rc = vr_svo_round<AUX, WALK, false>(P, q, &a);
            if (rc != VR_CELL_CONTINUE) {
                if (rc != VR_CELL_NO_WRITE) reinterpret_cast<uint32_t *>(P.image)[local] = vr_svo_finish<AUX>(q, rc, &a);
                if (AUX) {
                    reinterpret_cast<uint4 *>(P.aux)[2 * local] = *reinterpret_cast<uint4 *>(&a);
                    reinterpret_cast<uint4 *>(P.aux)[2 * local + 1] = *(reinterpret_cast<uint4 *>(&a) + 1);
                }
                active = false;
            }
        }
    }
}

__global__ void vr_fill_kernel(uint32_t *dst, size_t n, uint32_t value) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = value;
}

}  // namespace

cudaError_t vr_launch_raycast(const vr_frame_params &P, int use_svo, int with_aux, cudaStream_t stream,
                              unsigned long long *launches, const vr_launch_options *opt) {
    if (P.width <= 0 || P.local_rows <= 0) return cudaSuccess;
    const int tiles_x = (P.width + kTileW - 1) / kTileW;
    const dim3 grid(P.tile_world > 1 ? (tiles_x + P.tile_world - 1) / P.tile_world : tiles_x,
                    ((P.tile_world > 1 ? P.height : P.local_rows) + kTileH - 1) / kTileH);
    const dim3 block(kThreads);
    const bool multi = P.light_count > 1;
    const bool aux = with_aux != 0;
    if (use_svo && opt && opt->persistent && !multi && P.tile_world <= 1) {   /* (no multi-light / tile-interleave build) */
        cudaError_t e = cudaMemsetAsync(opt->counter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return e;
        unsigned ctas = (unsigned)(opt->num_sms * opt->ctas_per_sm);
        if (ctas > grid.x * grid.y) ctas = grid.x * grid.y;
        /* literal walks only (0 merged, 1 per-axis): the closed-form walk keeps 30 of 32 lanes busy with static tiles
         * (profiles/r2_persistent_l2_evidence.txt), so it has no persistent variant and renders as walk 1 here */
        void (*k)(vr_frame_params, unsigned int *, int) =
            opt->walk == 0 ? (aux ? vr_svo_persistent_kernel<true, 0> : vr_svo_persistent_kernel<false, 0>)
                           : (aux ? vr_svo_persistent_kernel<true, 1> : vr_svo_persistent_kernel<false, 1>);
        k<<<ctas, block, 0, stream>>>(P, opt->counter, opt->refill_min);
    } else if (use_svo) {
        const int walk = opt ? opt->walk : 0;
        void (*k)(vr_frame_params) =
            walk == 2 ? (multi ? (aux ? vr_svo_kernel<true, 2, true> : vr_svo_kernel<false, 2, true>)
                               : (aux ? vr_svo_kernel<true, 2, false> : vr_svo_kernel<false, 2, false>))
            : walk == 1 ? (multi ? (aux ? vr_svo_kernel<true, 1, true> : vr_svo_kernel<false, 1, true>)
                                 : (aux ? vr_svo_kernel<true, 1, false> : vr_svo_kernel<false, 1, false>))
                        : (multi ? (aux ? vr_svo_kernel<true, 0, true> : vr_svo_kernel<false, 0, true>)
                                 : (aux ? vr_svo_kernel<true, 0, false> : vr_svo_kernel<false, 0, false>));
        k<<<grid, block, 0, stream>>>(P);
    } else {
        void (*k)(vr_frame_params) = multi ? (aux ? vr_dense_kernel<true, true> : vr_dense_kernel<false, true>)
                                           : (aux ? vr_dense_kernel<true, false> : vr_dense_kernel<false, false>);
        k<<<grid, block, 0, stream>>>(P);
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t vr_launch_fill(uint32_t *dst, size_t n, uint32_t value, cudaStream_t stream,
                           unsigned long long *launches) {
    if (!n) return cudaSuccess;
    vr_fill_kernel<<<148 * 4, 256, 0, stream>>>(dst, n, value);
    if (launches) ++*launches;
    return cudaGetLastError();
}
