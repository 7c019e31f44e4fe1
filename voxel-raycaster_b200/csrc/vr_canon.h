/*
 * vr_canon.h -- octree traversal with closed-form crossing times ("canonical t", option walk = 2).
 *
 * The reference advances intersection_t by one rounded addition per DDA step (kernels/ray_caster_kernel.cl:559), so
 * the k-th crossing time of an axis is the result of k roundings.  SURVEY.md Appendix E proposes the closed form
 *      t_axis(k) = fma(k, delta_t_axis, t0_axis)          (t0 = intersection_t when the ray (re)starts, one rounding)
 * which makes every crossing time a pure function of the integer crossing count: a traversal that skips a whole empty
 * octree cell obtains exactly the values a voxel-by-voxel walk would, in O(1) per cell -- no add chains, no loops
 * whose trip count differs between the lanes of a warp.  The voxel sequence, the hit voxel and the face are those of
 * the reference unless two crossing times are closer to each other than the rounding the reference accumulates
 * (rays through a voxel edge: BASELINE.json's "degenerate" rays); colours differ by float noise in the hit UV.
 * Measured against the reference walk on the bench frame (profiles/canonical_t.py): RGBA8
 * max-abs-diff <= 1 on 99.98 % of the pixels, first hit / face identical on all but 1.8e-5 of them.
 *
 * Everything else is the reference's: ray setup (kernel:276-354), hit / UV / atlas / view_light / redirect
 * (kernel:575-711, vr_hit_block), tie rule `<=` of kernel:558 at every cell exit, distance_traveled / max_distance
 * accounting (kernel:357, 667, 714), fog and the RGBA8 store (kernel:716-721).
 *
 * State per ray: crossing counts are kept as MIRRORED voxel coordinates p = (voxel_step < 0) ? N-1-voxel : voxel, so
 * that every axis steps by +1, the crossing count of an axis is p - p0, the last voxel of an aligned power-of-two cell
 * is p | (size-1), and "left the map" is one unsigned compare of px|py|pz.  The octree is addressed with the child
 * slot of p XOR the slot bits of the mirror mask.
 * Integer <-> float conversions run on the FMA pipe through the 1.5 * 2^23 trick (the conversion instructions are
 * quarter rate): as_float(i + 0x4B400000) - 12582912.0f == (float)i for |i| < 2^22.
 */
#ifndef VR_CANON_H
#define VR_CANON_H

#include "vr_trace.h"

#define VR_MAGIC_F 12582912.0f
#define VR_MAGIC_I 0x4B400000

/* keeps a per-ray constant in its register: without this the compiler re-derives such values from their sources inside
 * the traversal loop (a dozen ALU instructions per lookup) to stay below the register limit */
#if defined(__CUDA_ARCH__) && !defined(VR_CANON_NO_PIN)
#define VR_PIN(x) asm volatile("" : "+r"(x))
#else
#define VR_PIN(x) ((void)0)
#endif

/* a point every lane of the warp passes once per turn of the traversal loop: without it the compiler threads the
 * early exits of the lookup / the walk straight back to the loop head, the lanes of a warp then run the loop body
 * in separate groups and never reconverge (measured: 20 instead of 29 active lanes per instruction) */
#if defined(__CUDA_ARCH__)
#define VR_JOIN(x) asm volatile("" : "+r"(x))
#else
#define VR_JOIN(x) ((void)0)
#endif

/* marks a branch as one that must stay a branch (rare path): keeps ptxas from turning it into a chain of selects
 * that every turn of the loop would pay for */
#if defined(__CUDA_ARCH__)
#define VR_RARE() asm volatile("" ::: "memory")
#else
#define VR_RARE() ((void)0)
#endif

/* analysis hook (profiles/canon_stats.py builds the host emulation with it): one call per turn of the cell loop */
#ifndef VR_CANON_STAT
#define VR_CANON_STAT(brick, m, ext, steps, shadow) ((void)0)
#endif

VR_HD int vr_hibit(int x) {          /* index of the highest set bit of x != 0 (31 for negative x) */
#if defined(__CUDA_ARCH__)
    return 31 - __clz(x);
#else
    return 31 - __builtin_clz((unsigned)x);
#endif
}

template <class Stack>
struct vr_cray {
    RayState r;               /* the reference's private variables; r.t / r.voxel / r.fm only materialised at a hit */
    float t0x, t0y, t0z;      /* intersection_t when the ray (re)started */
    float ix, iy, iz;         /* |ray_dir| ~ 1 / delta_t: crossing-count estimates only, never a value */
    int px, py, pz;           /* mirrored voxel */
    int bx, by, bz;           /* p0 - VR_MAGIC_I (p0 = mirrored voxel at the (re)start) */
    uint32_t flip;            /* child-slot XOR of the mirrored axes (levels below the grid: 3 / 0xC / 0x30 per axis) */
    uint32_t gflip;           /* grid-index XOR of the mirrored axes */
    vr_node_regs node;        /* octree cursor below the grid: current node and its child shift */
    int s;
    bool first_hit_done;
    Stack stk;                /* node indices of the current block's subtree, slot = child shift / 2 */
};

/* (re)start of a ray: from the reference state (voxel, voxel_step, intersection_t, ray_dir) to the mirrored form */
template <class Stack>
VR_HD void vr_canon_enter(const vr_frame_params &P, vr_cray<Stack> &q) {
    const RayState &r = q.r;
    const int N = P.dim[0];
    const bool nx = r.step.x < 0, ny = r.step.y < 0, nz = r.step.z < 0;
    q.px = nx ? N - 1 - r.voxel.x : r.voxel.x;
    q.py = ny ? N - 1 - r.voxel.y : r.voxel.y;
    q.pz = nz ? N - 1 - r.voxel.z : r.voxel.z;
    q.bx = q.px - VR_MAGIC_I; q.by = q.py - VR_MAGIC_I; q.bz = q.pz - VR_MAGIC_I;
    q.t0x = r.t.x; q.t0y = r.t.y; q.t0z = r.t.z;
    q.ix = fabsf(r.ray_dir.x); q.iy = fabsf(r.ray_dir.y); q.iz = fabsf(r.ray_dir.z);
    q.flip = (nx ? 3u : 0u) | (ny ? 0xCu : 0u) | (nz ? 0x30u : 0u);
    const uint32_t gm = (1u << P.grid_bits) - 1u;
    q.gflip = (nx ? gm : 0u) | (ny ? gm << P.grid_bits : 0u) | (nz ? gm << (2 * P.grid_bits) : 0u);
    if (P.grid_directed) q.gflip |= ((nx ? 1u : 0u) | (ny ? 2u : 0u) | (nz ? 4u : 0u)) << (3 * P.grid_bits);   /* this octant's table */
    VR_PIN(q.bx); VR_PIN(q.by); VR_PIN(q.bz); VR_PIN(q.flip); VR_PIN(q.gflip);
}

/* the reference's voxel / intersection_t / face_mask of the step just made, whose time was T: the axes of that step
 * are those whose latest crossing has exactly that time (kernel:558: every axis whose time equals the minimum steps) */
template <class Stack>
VR_HD void vr_canon_materialize(const vr_frame_params &P, vr_cray<Stack> &q, float T) {
    RayState &r = q.r;
    const int N = P.dim[0];
    r.step = {(q.flip & 1u) ? -1 : 1, (q.flip & 4u) ? -1 : 1, (q.flip & 16u) ? -1 : 1};   /* (not kept live across the walk) */
    r.voxel = {r.step.x < 0 ? N - 1 - q.px : q.px, r.step.y < 0 ? N - 1 - q.py : q.py, r.step.z < 0 ? N - 1 - q.pz : q.pz};
    const float kx = VR_SUB(vr_bits2f(q.px - q.bx), VR_MAGIC_F), ky = VR_SUB(vr_bits2f(q.py - q.by), VR_MAGIC_F),
                kz = VR_SUB(vr_bits2f(q.pz - q.bz), VR_MAGIC_F);
    r.t.x = VR_FMA(kx, r.delta.x, q.t0x);
    r.t.y = VR_FMA(ky, r.delta.y, q.t0y);
    r.t.z = VR_FMA(kz, r.delta.z, q.t0z);
    r.fm = ((kx >= 1.0f && VR_FMA(VR_SUB(kx, 1.0f), r.delta.x, q.t0x) == T) ? 1 : 0) |
           ((ky >= 1.0f && VR_FMA(VR_SUB(ky, 1.0f), r.delta.y, q.t0y) == T) ? 2 : 0) |
           ((kz >= 1.0f && VR_FMA(VR_SUB(kz, 1.0f), r.delta.z, q.t0z) == T) ? 4 : 0);
}

/* The empty cell a walk crosses: the cube [p & ~m, (p | m) + ext] per axis (mirrored coordinates) -- an aligned
 * power-of-two cell of the octree (ext = 0) or a block of the top grid grown by its empty radius -- or a leaf brick. */
struct vr_ccell {
    int m, ext;
    bool brick;
};

/* Looks the (in-map) voxel p up.  `xr` has a bit set wherever p differs from the voxel of the previous lookup.
 * Returns the voxel value if the voxel is set; otherwise 0 and the empty cell around p.
 * Two stages: the top grid (vr_types.h) answers for whole blocks in one 4-byte load, without stack or descent; only
 * inside a non-empty block the 64-tree below it is descended, one 16-byte node per two octree levels. */
template <bool AUX, class Stack>
VR_HD int vr_canon_lookup(const vr_frame_params &P, vr_cray<Stack> &q, int xr, vr_ccell &c, vr_aux *a) {
    const int g = P.grid_shift;
    if (AUX) a->lookups++;
    if ((xr >> g) != 0) {                                                 /* another block */
        const int G = P.grid_dim;
        const uint32_t key = (uint32_t)((q.px >> g) + ((q.py >> g) + (q.pz >> g) * G) * G) ^ q.gflip;
#if defined(__CUDA_ARCH__)
        const uint32_t e = __ldg(P.grid + key);
#else
        const uint32_t e = P.grid[key];
#endif
        if (!(e & 0x80000000u)) {
            c.m = (1 << (e & 31u)) - 1;
            c.ext = (int)(e >> 8);
            c.brick = false;
            return 0;
        }
        const uint32_t idx = e & 0x7fffffffu;
        q.s = g - 2;
        q.stk.set(q.s >> 1, idx);
        q.node = vr_load_node(P, idx);
        if (AUX) a->node_fetches++;
    } else if ((xr >> (q.s + 2)) != 0) {                                  /* same block: pop to the lowest ancestor containing p */
        q.s = vr_hibit(xr) & ~1;
        q.node = vr_load_node(P, q.stk.get(q.s >> 1));
        if (AUX) a->node_fetches++;
    }
    const uint32_t cf = q.flip;
    for (;;) {
        if (q.node.base & VR_NODE_SOLID) return (int)(int8_t)(q.node.base & 0xffu);      /* a collapsed solid subtree: p is a set voxel */
        const int s = q.s;
        const uint32_t ci = ((uint32_t)((q.px >> s) & 3) | ((uint32_t)((q.py >> s) & 3) << 2) | ((uint32_t)((q.pz >> s) & 3) << 4)) ^ cf;
        if (!((uint32_t)(q.node.mask >> ci) & 1u)) {
            /* the 2x2x2 octant of slots around an empty slot is empty as a whole <=> the cell is twice as wide (the odd
             * levels of the reference's 2^3 octree; slots ci&0x2A + {0,1,4,5,16,17,20,21}) */
            const bool wide = ((uint32_t)(q.node.mask >> (ci & 0x2Au)) & 0x00330033u) == 0u;
            c.brick = s == 0;
            c.m = (1 << (s + (wide ? 1 : 0))) - 1;
            c.ext = 0;
            return 0;
        }
        const uint32_t below = (uint32_t)VR_POPC64(q.node.mask & ((1ull << ci) - 1ull));
        if (s == 0) return (int)(int8_t)P.leaf_types[q.node.base + below];
        const uint32_t child = q.node.base + below;
        q.s = s - 2;
        q.stk.set(q.s >> 1, child);
        q.node = vr_load_node(P, child);
        if (AUX) a->node_fetches++;
    }
}

/* Walks the empty cell `c` the mirrored voxel p lies in, up to and including the step that leaves it.
 * T = min over the axes of the time of the crossing that leaves the cell; every axis then makes all its crossings with
 * time <= T (kernel:558: an axis steps when its time is <= the others', ties step together).
 * Step count: the walk does not count steps, it moves the voxel.  Between multi-axis steps every step changes exactly one
 * coordinate by one, so distance_traveled = dbase + px + py + pz with a per-segment constant dbase; a step along k axes
 * at once lowers dbase by k - 1 (done here for the last step of the cell; ties strictly inside a cell are not observed:
 * see DESIGN.md, tie rays).  Returns the new px + py + pz; T = time of the last step; tie = that step moved along more
 * than one axis; xr = changed voxel bits.  `biased`: the frame has a get_oct_vox start bias (kernel:353). */
template <class Stack>
VR_HD int vr_canon_walk(vr_cray<Stack> &q, const vr_ccell &c, bool biased, float &T, int &dbase, bool &tie, int &xr) {
    const RayState &r = q.r;
    const int ox = (q.px | c.m) + c.ext, oy = (q.py | c.m) + c.ext, oz = (q.pz | c.m) + c.ext;   /* last voxel of the cell per axis */
    const float Tx = VR_FMA(VR_SUB(vr_bits2f(ox - q.bx), VR_MAGIC_F), r.delta.x, q.t0x);
    const float Ty = VR_FMA(VR_SUB(vr_bits2f(oy - q.by), VR_MAGIC_F), r.delta.y, q.t0y);
    const float Tz = VR_FMA(VR_SUB(vr_bits2f(oz - q.bz), VR_MAGIC_F), r.delta.z, q.t0z);
    T = vr_min3(Tx, Ty, Tz);
    /* per axis: k = the last crossing with time <= T.  The estimate RN((T - t0) / delta) is k or k + 1 (its error is
     * below 1e-2 crossings: |ray_dir| * delta_t = 1 +- 2^-24, at most 2^16 crossings), one evaluation decides. */
    const float mx = VR_ADD(VR_MUL(VR_SUB(T, q.t0x), q.ix), VR_MAGIC_F);
    const float my = VR_ADD(VR_MUL(VR_SUB(T, q.t0y), q.iy), VR_MAGIC_F);
    const float mz = VR_ADD(VR_MUL(VR_SUB(T, q.t0z), q.iz), VR_MAGIC_F);
    int nx = vr_f2bits(mx) + q.bx + 1, ny = vr_f2bits(my) + q.by + 1, nz = vr_f2bits(mz) + q.bz + 1;
    if (VR_FMA(VR_SUB(mx, VR_MAGIC_F), r.delta.x, q.t0x) > T) nx -= 1;
    if (VR_FMA(VR_SUB(my, VR_MAGIC_F), r.delta.y, q.t0y) > T) ny -= 1;
    if (VR_FMA(VR_SUB(mz, VR_MAGIC_F), r.delta.z, q.t0z) > T) nz -= 1;
    if (biased) {
        VR_RARE();
        /* an axis whose first crossing lies beyond T makes none.  With 0 <= t0 <= delta_t and T >= 0 the estimate above is
         * never below that; a negative start bias (a camera in a collapsed empty octree cell, kernel:353) makes T negative
         * and the linear estimate of such an axis point far below zero */
        nx = nx > q.px ? nx : q.px; ny = ny > q.py ? ny : q.py; nz = nz > q.pz ? nz : q.pz;
    }
    const float together = VR_ADD(VR_ADD(Tx == T ? 1.0f : 0.0f, Ty == T ? 1.0f : 0.0f), Tz == T ? 1.0f : 0.0f);
    tie = together > 1.5f;
    if (tie) {                                                            /* the axes of the last step moved together */
        VR_RARE();
        dbase -= (int)together - 1;
    }
    xr = (nx ^ q.px) | (ny ^ q.py) | (nz ^ q.pz);
    q.px = nx; q.py = ny; q.pz = nz;
    return nx + ny + nz;
}

/* The first step of a ray is made at m0 = min(t0) by every axis whose t0 equals it: c0 axes, ONE step (kernel:558, 714).
 * vr_canon_walk counts steps per axis and sees a multi-axis step only where it looks -- at the step that leaves the cell,
 * and there only the axes that leave it.  For generic rays a tie at the very first step is as rare as any other; with the
 * camera on a voxel edge EVERY primary ray starts with one, and counting it per axis would end rays up to two steps early
 * (kernel:357, 667).  Returns what has to be taken off dbase for the first walk of the segment over cell `c`: the c0 - 1
 * surplus steps, less what the walk will take off itself when the first step is the step that leaves the cell (m0 == T) and
 * `together` >= 2 of the axes leave with it.  A brick walk observes every step: nothing to correct. */
template <class Stack>
VR_HD int vr_canon_first_step_tie(const vr_cray<Stack> &q, vr_ccell c, bool bricks_walked) {
    const RayState &r = q.r;
    const float m0 = vr_min3(q.t0x, q.t0y, q.t0z);
    const int c0 = (q.t0x == m0 ? 1 : 0) + (q.t0y == m0 ? 1 : 0) + (q.t0z == m0 ? 1 : 0);
    if (c0 < 2) return 0;
    if (c.brick) {
        if (bricks_walked) return 0;
        c = {0, 0, false};                     /* as in the cell loop: a brick is walked voxel by voxel when the ray is about to end */
    }
    const int ox = (q.px | c.m) + c.ext, oy = (q.py | c.m) + c.ext, oz = (q.pz | c.m) + c.ext;
    const float Tx = VR_FMA(VR_SUB(vr_bits2f(ox - q.bx), VR_MAGIC_F), r.delta.x, q.t0x);
    const float Ty = VR_FMA(VR_SUB(vr_bits2f(oy - q.by), VR_MAGIC_F), r.delta.y, q.t0y);
    const float Tz = VR_FMA(VR_SUB(vr_bits2f(oz - q.bz), VR_MAGIC_F), r.delta.z, q.t0z);
    const float T = vr_min3(Tx, Ty, Tz);
    const int together = (Tx == T ? 1 : 0) + (Ty == T ? 1 : 0) + (Tz == T ? 1 : 0);
    return c0 - ((m0 < T || together < 2) ? 1 : together);
}

/* Leaf brick (4^3 voxels, occupancy = mask): the step of kernel:558-560 with closed-form times, followed by a bit test
 * of the voxel entered.  Stops when a step leaves the brick (returns false) or lands on a set voxel (returns true,
 * `bit` = its slot).  The caller guarantees that max_distance cannot be reached inside (a brick holds <= 10 steps).
 * Every step is observed here, so multi-axis steps are exact (n counts them once). */
template <class Stack>
VR_HD bool vr_canon_brick(vr_cray<Stack> &q, unsigned long long mask, int &n, float &T, int &xr, int &bit) {
    const RayState &r = q.r;
    const int lx = q.px & 3, ly = q.py & 3, lz = q.pz & 3;
    float kx = VR_SUB(vr_bits2f(q.px - q.bx), VR_MAGIC_F), ky = VR_SUB(vr_bits2f(q.py - q.by), VR_MAGIC_F),
          kz = VR_SUB(vr_bits2f(q.pz - q.bz), VR_MAGIC_F);
    float tx = VR_FMA(kx, r.delta.x, q.t0x), ty = VR_FMA(ky, r.delta.y, q.t0y), tz = VR_FMA(kz, r.delta.z, q.t0z);
    float rx = (float)(4 - lx), ry = (float)(4 - ly), rz = (float)(4 - lz), steps = 0.0f;
    float bitf = (float)(lx | (ly << 2) | (lz << 4));
    const int cf = (int)q.flip;
    float ex, ey, ez;
    bool hit;
    float mn;
    for (;;) {
        mn = vr_min3(tx, ty, tz);
        ex = (tx == mn) ? 1.0f : 0.0f;
        ey = (ty == mn) ? 1.0f : 0.0f;
        ez = (tz == mn) ? 1.0f : 0.0f;
        kx = VR_ADD(kx, ex); ky = VR_ADD(ky, ey); kz = VR_ADD(kz, ez);
        tx = VR_FMA(kx, r.delta.x, q.t0x); ty = VR_FMA(ky, r.delta.y, q.t0y); tz = VR_FMA(kz, r.delta.z, q.t0z);
        rx = VR_SUB(rx, ex); ry = VR_SUB(ry, ey); rz = VR_SUB(rz, ez);
        bitf = VR_FMA_EXACT(ex, 1.0f, VR_FMA_EXACT(ey, 4.0f, VR_FMA_EXACT(ez, 16.0f, bitf)));
        steps = VR_ADD(steps, 1.0f);
        if (VR_MUL(VR_MUL(rx, ry), rz) == 0.0f) { hit = false; break; }
        bit = (int)bitf ^ cf;
        if ((uint32_t)(mask >> bit) & 1u) { hit = true; break; }
    }
    const int nx = q.px + (4 - lx) - (int)rx, ny = q.py + (4 - ly) - (int)ry, nz = q.pz + (4 - lz) - (int)rz;
    xr = (nx ^ q.px) | (ny ^ q.py) | (nz ^ q.pz);
    q.px = nx; q.py = ny; q.pz = nz;
    T = mn;
    n = (int)steps;
    return hit;
}

/* point query from the root (slow path only) */
VR_HD int vr_tree_voxel(const vr_frame_params &P, int x, int y, int z) {
    uint32_t idx = 0;
    for (int s = P.root_shift;; s -= 2) {
        const vr_node_regs nd = vr_load_node(P, idx);
        if (nd.base & VR_NODE_SOLID) return (int)(int8_t)(nd.base & 0xffu);
        const int ci = ((x >> s) & 3) | (((y >> s) & 3) << 2) | (((z >> s) & 3) << 4);
        if (!((nd.mask >> ci) & 1ull)) return 0;
        const uint32_t rank = (uint32_t)VR_POPC64(nd.mask & ((1ull << ci) - 1ull));
        if (s == 0) return (int)(int8_t)P.leaf_types[nd.base + rank];
        idx = nd.base + rank;
    }
}

/* Rays whose float state is not finite (delta_t = inf from a denormal direction component): the estimates above
 * need finite numbers, so such a ray is finished voxel by voxel -- kernel:558-560 with the same closed form, applied
 * to the axes that step (an axis that never steps keeps its t0, infinite or not).  Returns the terminal status or
 * VR_CELL_NO_WRITE. */
template <bool AUX, bool MULTI>
VR_HD int vr_canon_slow(const vr_frame_params &P, RayState &r, vr_aux *a, bool &first_hit_done) {
    const int N = P.dim[0];
    vf3 t0 = r.t;
    float kx = 0.0f, ky = 0.0f, kz = 0.0f;
    for (;;) {
        if (!(r.dist < r.max_distance && r.bounce < P.max_bounces)) {
            if (MULTI && r.bounce < P.max_bounces && vr_more_lights(P, r)) {
                if (!vr_next_light(P, r)) return VR_ST_SKIP_REDIRECT;
                r.dist++;
                t0 = r.t; kx = ky = kz = 0.0f;
                continue;
            }
            return r.bounce >= P.max_bounces ? VR_ST_BOUNCES : VR_ST_MAXDIST;
        }
        const int mx = (r.t.x <= vr_min(r.t.y, r.t.z)) ? 1 : 0;
        const int my = (r.t.y <= vr_min(r.t.z, r.t.x)) ? 1 : 0;
        const int mz = (r.t.z <= vr_min(r.t.x, r.t.y)) ? 1 : 0;
        r.fm = mx | (my << 1) | (mz << 2);
        if (mx) { kx += 1.0f; r.t.x = VR_FMA(kx, r.delta.x, t0.x); }
        if (my) { ky += 1.0f; r.t.y = VR_FMA(ky, r.delta.y, t0.y); }
        if (mz) { kz += 1.0f; r.t.z = VR_FMA(kz, r.delta.z, t0.z); }
        r.voxel.x += r.step.x * mx; r.voxel.y += r.step.y * my; r.voxel.z += r.step.z * mz;
        if (AUX && (r.fm & (r.fm - 1))) a->flags |= VR_FL_TIE;
        if ((unsigned)r.voxel.x >= (unsigned)N || (unsigned)r.voxel.y >= (unsigned)N || (unsigned)r.voxel.z >= (unsigned)N) {
            if (MULTI && vr_more_lights(P, r)) {
                if (!vr_next_light(P, r)) return VR_ST_SKIP_REDIRECT;
                r.dist++;
                t0 = r.t; kx = ky = kz = 0.0f;
                continue;
            }
            vr_out_of_bounds(r);
            return VR_ST_OOB;
        }
        const int voxel_data = vr_tree_voxel(P, r.voxel.x, r.voxel.y, r.voxel.z);
        if (voxel_data == 5 || voxel_data == 6) {
            const int st = vr_hit_block<AUX, MULTI>(P, r, voxel_data, a, first_hit_done);
            if (st >= 0) return st;
            t0 = r.t; kx = ky = kz = 0.0f;
        }
        r.dist++;
    }
}

/* Whole pixel.  Returns true if the pixel must be written (packed colour in *rgba_out).
 * distance_traveled bookkeeping: r.dist is the reference's counter at the top of its loop (kernel:357).  A walk of n
 * steps is n loop iterations; the last of them loads the voxel entered and, on a hit, runs the hit block with the
 * counter one below the value it has at the top of the next iteration (kernel:714 increments it afterwards).  Inside
 * the cell loop the counter is not kept: it is dbase + px + py + pz (vr_canon_walk). */
template <bool AUX, bool MULTI, class Stack>
VR_HD bool vr_trace_svo_canon(const vr_frame_params &P, int x, int y, uint32_t *rgba_out, vr_aux *a, Stack &stk) {
    vr_cray<Stack> q;
    q.stk = stk;
    RayState &r = q.r;
    if (AUX) vr_aux_init(a, P);
    if (!vr_ray_setup(P, x, y, r)) {
        if (AUX) a->status = VR_ST_SKIP_PRIMARY;
        return false;
    }
    const int N = P.dim[0];
    const bool biased = P.bias[0] != 0.0f || P.bias[1] != 0.0f || P.bias[2] != 0.0f;      /* frame-uniform */
    /* frame-uniform too: a camera on a voxel edge or corner (two or three integer coordinates).  intersection_t then starts
     * at the same value on those axes, and the first step of EVERY primary ray moves along them at once (kernel:558) */
    const bool on_edge = P.cam_on_edge != 0;
    const bool voxelwise = P.cam_on_edge == 2;      /* ... inside a collapsed empty octree cell: see vr_cam_on_edge (vr_types.h) */
    q.first_hit_done = false;
    q.s = 0;
    int status = VR_ST_MAXDIST;
    bool slow = false;
    int xr = -1;                           /* changed voxel bits since the last lookup; -1 = everything */
    for (;;) {                             /* one turn per ray segment: primary ray, shadow ray(s), reflections */
        if (!vr_ray_finite(r) || voxelwise) { slow = true; break; }
        vr_canon_enter(P, q);
        /* the cell around the voxel the segment starts in; that voxel itself is never tested (the reference steps before
         * it loads, kernel:555-570) and may even lie outside the map */
        vr_ccell c = {0, 0, false};
        if ((unsigned)(q.px | q.py | q.pz) < (unsigned)N) {
            if (vr_canon_lookup<AUX>(P, q, xr, c, a) != 0) c = {0, 0, false};
        }
        int voxel_data = 0;
        float T = 0.0f;                    /* time of the last step */
        enum { EV_HIT, EV_UNBLOCKED_MAXDIST, EV_UNBLOCKED_OOB } ev = EV_UNBLOCKED_MAXDIST;
        if (!(r.bounce < P.max_bounces)) { status = VR_ST_BOUNCES; break; }            /* kernel:357; bounce_count only changes at a hit */
        if (r.dist < r.max_distance) {                                     /* kernel:357 */
            int dbase = r.dist - (q.px + q.py + q.pz);                     /* distance_traveled = dbase + px + py + pz */
            if (on_edge) {
                VR_RARE();
                if (!r.shadow && r.bounce == 0) dbase -= vr_canon_first_step_tie(q, c, r.max_distance - r.dist > 12);
            }
            for (;;) {                     /* one turn per empty cell */
                int sum, bit = 0;
                bool known = false;        /* the brick walk already knows that the voxel entered is set */
                bool tie = false;
                bool as_brick = false;
                if (c.brick) {
                    /* a brick holds at most 10 steps: walked as such unless the ray is about to end (then voxel by voxel) */
                    as_brick = r.max_distance - (dbase + q.px + q.py + q.pz) > 12;
                    if (!as_brick) c = {0, 0, false};
                }
                if (as_brick) {
                    const int before = q.px + q.py + q.pz;
                    int n;
                    known = vr_canon_brick(q, q.node.mask, n, T, xr, bit);
                    sum = q.px + q.py + q.pz;
                    VR_CANON_STAT(1, 3, 0, n, r.shadow);
                    if (sum - before != n) {                               /* multi-axis steps inside the brick */
                        VR_RARE();
                        dbase -= sum - before - n;
                        tie = true;
                    }
                } else {
                    VR_CANON_STAT(0, c.m, c.ext, -(q.px + q.py + q.pz), r.shadow);
                    sum = vr_canon_walk(q, c, biased, T, dbase, tie, xr);
                    VR_CANON_STAT(2, c.m, c.ext, sum, r.shadow);
                }
                VR_JOIN(sum);
                r.dist = dbase + sum;
                if (r.dist > r.max_distance) {                             /* kernel:357 ended the loop inside this cell */
                    r.dist = r.max_distance;
                    ev = EV_UNBLOCKED_MAXDIST;
                    break;
                }
                if (AUX && tie) a->flags |= VR_FL_TIE;
                if (known) {                                               /* the step landed on a set voxel of the brick */
                    voxel_data = (int)(int8_t)P.leaf_types[q.node.base + (uint32_t)VR_POPC64(q.node.mask & ((1ull << bit) - 1ull))];
                } else {
                    if ((unsigned)(q.px | q.py | q.pz) >= (unsigned)N) { ev = EV_UNBLOCKED_OOB; break; }   /* kernel:563 */
                    voxel_data = vr_canon_lookup<AUX>(P, q, xr, c, a);
                }
                VR_JOIN(voxel_data);
                if (voxel_data == 5 || voxel_data == 6) { ev = EV_HIT; break; }
                if (voxel_data != 0) c = {0, 0, false};
            }
        }
        /* ---- the segment ended */
        bool relight = false;              /* multi-light extension: this light is not blocked, on to the next one */
        if (ev == EV_HIT) {
            r.dist -= 1;                   /* the counter inside the iteration of the hit */
            if (r.shadow) {                /* kernel:706-710; nothing else of the hit block is live */
                r.color.w = MULTI ? VR_ADD(r.alpha_before, 0.1f) : 0.1f;
                if (!(MULTI && vr_more_lights(P, r))) { status = VR_ST_SHADOW_HIT; break; }
                relight = true;
            } else {
                vr_canon_materialize(P, q, T);
                const vi3 hv = r.voxel;
                const int st = vr_hit_block<AUX, MULTI>(P, r, voxel_data, a, q.first_hit_done);
                if (st >= 0) { status = st; break; }
                /* redirected: the ray restarts in the voxel it came from */
                xr = (hv.x ^ r.voxel.x) | (hv.y ^ r.voxel.y) | (hv.z ^ r.voxel.z);
                r.dist += 1;               /* kernel:714 ends the iteration of the hit */
            }
        } else if (ev == EV_UNBLOCKED_MAXDIST) {
            if (!(MULTI && r.bounce < P.max_bounces && vr_more_lights(P, r))) { status = VR_ST_MAXDIST; break; }
            relight = true;
        } else {
            if (!(MULTI && vr_more_lights(P, r))) {
                r.dist -= 1;
                r.fm = 0;                  /* the voxel is dead: only the colour of kernel:565 is needed */
                vr_out_of_bounds(r);
                status = VR_ST_OOB;
                break;
            }
            relight = true;
        }
        if (MULTI && relight) {
            if (!vr_next_light(P, r)) { status = VR_ST_SKIP_REDIRECT; break; }
            r.dist += 1;
            xr = -1;
        }
    }
    if (slow) status = vr_canon_slow<AUX, MULTI>(P, r, a, q.first_hit_done);
    if (status == VR_ST_SKIP_REDIRECT) {
        if (AUX) { a->status = (uint8_t)VR_ST_SKIP_REDIRECT; a->steps_total = (uint32_t)r.dist; }
        return false;
    }
    if (AUX) { a->status = (uint8_t)status; a->steps_total = (uint32_t)r.dist; }
    *rgba_out = vr_epilogue(r);
    return true;
}

#endif
