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
    uint32_t flip;            /* child-slot XOR of the mirrored axes: bits 0-5 below the root, bits 8-13 at the root */
    vr_node_regs node;        /* octree cursor: current node, its child shift, its depth */
    int s, level;
    bool first_hit_done;
    Stack stk;
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
    const uint32_t rb = (uint32_t)(N - 1) >> P.root_shift;               /* slot bits of N-1 at the root level: 1 or 3 */
    q.flip = (nx ? 3u | (rb << 8) : 0u) | (ny ? 0xCu | (rb << 10) : 0u) | (nz ? 0x30u | (rb << 12) : 0u);
}

/* the reference's voxel / intersection_t / face_mask of the step just made */
template <class Stack>
VR_HD void vr_canon_materialize(const vr_frame_params &P, vr_cray<Stack> &q, int fm) {
    RayState &r = q.r;
    const int N = P.dim[0];
    r.voxel = {r.step.x < 0 ? N - 1 - q.px : q.px, r.step.y < 0 ? N - 1 - q.py : q.py, r.step.z < 0 ? N - 1 - q.pz : q.pz};
    r.t.x = VR_FMA(VR_SUB(vr_bits2f(q.px - q.bx), VR_MAGIC_F), r.delta.x, q.t0x);
    r.t.y = VR_FMA(VR_SUB(vr_bits2f(q.py - q.by), VR_MAGIC_F), r.delta.y, q.t0y);
    r.t.z = VR_FMA(VR_SUB(vr_bits2f(q.pz - q.bz), VR_MAGIC_F), r.delta.z, q.t0z);
    r.fm = fm;
}

template <class Stack>
VR_HD void vr_canon_cursor_reset(const vr_frame_params &P, vr_cray<Stack> &q) {
    q.s = P.root_shift;
    q.level = 0;
    q.node = vr_load_node(P, 0);
}

/* Looks the (in-map) voxel p up.  `xr` has a bit set wherever p differs from the voxel of the previous lookup.
 * Returns the voxel value if the voxel is set; otherwise 0 and the empty cell around p: edge 1 << cs, or the leaf
 * brick (4^3 voxels, occupancy = q.node.mask) when `brick`. */
template <bool AUX, class Stack>
VR_HD int vr_canon_lookup(const vr_frame_params &P, vr_cray<Stack> &q, int xr, int &cs, bool &brick, vr_aux *a) {
    if (AUX) a->lookups++;
    if ((xr >> (q.s + 2)) != 0) {                                         /* pop to the lowest ancestor containing p */
        int s2 = vr_hibit(xr) & ~1;
        s2 = s2 < P.root_shift ? s2 : P.root_shift;
        q.level -= (s2 - q.s) >> 1;
        q.s = s2;
        q.node = vr_load_node(P, q.stk.get(q.level));
        if (AUX) a->node_fetches++;
    }
    for (;;) {
        const int s = q.s;
        const uint32_t cf = (s == P.root_shift ? q.flip >> 8 : q.flip) & 63u;
        const int ci = (int)(((uint32_t)((q.px >> s) & 3) | ((uint32_t)((q.py >> s) & 3) << 2) | ((uint32_t)((q.pz >> s) & 3) << 4)) ^ cf);
        if (!((q.node.mask >> ci) & 1ull)) {
            brick = s == 0;
            /* the 2x2x2 octant of slots around an empty slot is empty as a whole <=> the cell is twice as wide (the odd
             * levels of the reference's 2^3 octree; slots ci&0x2A + {0,1,4,5,16,17,20,21}) */
            const bool wide = ((q.node.mask >> (ci & 0x2A)) & 0x00330033ull) == 0ull;
            cs = brick ? 2 : s + (wide ? 1 : 0);
            return 0;
        }
        const uint32_t below = (uint32_t)VR_POPC64(q.node.mask & ((1ull << ci) - 1ull));
        if (s == 0) return (int)(int8_t)P.leaf_types[q.node.base + below];
        const uint32_t child = q.node.base + below;
        q.level++;
        q.s = s - 2;
        q.stk.set(q.level, child);
        q.node = vr_load_node(P, child);
        if (AUX) a->node_fetches++;
    }
}

/* Walks the empty cell of edge m+1 (aligned, power of two) the mirrored voxel p lies in, up to and including the step
 * that leaves it.  T = min over the axes of the time of the crossing that leaves the cell; every axis then makes all
 * its crossings with time <= T (kernel:558: an axis steps when its time is <= the others', ties step together).
 * Returns the number of steps (multi-axis steps inside the cell are counted per axis: see DESIGN.md, tie rays);
 * fm = axes of the last step; xr = changed voxel bits. */
template <class Stack>
VR_HD int vr_canon_walk(vr_cray<Stack> &q, int m, int &fm, int &xr) {
    const RayState &r = q.r;
    const int ox = q.px | m, oy = q.py | m, oz = q.pz | m;               /* last voxel of the cell along each axis */
    const float Tx = VR_FMA(VR_SUB(vr_bits2f(ox - q.bx), VR_MAGIC_F), r.delta.x, q.t0x);
    const float Ty = VR_FMA(VR_SUB(vr_bits2f(oy - q.by), VR_MAGIC_F), r.delta.y, q.t0y);
    const float Tz = VR_FMA(VR_SUB(vr_bits2f(oz - q.bz), VR_MAGIC_F), r.delta.z, q.t0z);
    const float T = vr_min3(Tx, Ty, Tz);
    /* per axis: k = the last crossing with time <= T.  The estimate RN((T - t0) / delta) is k or k + 1 (its error is
     * below 1e-3 crossings: |ray_dir| * delta_t = 1 +- 2^-24, at most 2^22 crossings), one evaluation decides. */
    const float mx = VR_ADD(VR_MUL(VR_SUB(T, q.t0x), q.ix), VR_MAGIC_F);
    const float my = VR_ADD(VR_MUL(VR_SUB(T, q.t0y), q.iy), VR_MAGIC_F);
    const float mz = VR_ADD(VR_MUL(VR_SUB(T, q.t0z), q.iz), VR_MAGIC_F);
    int nx = vr_f2bits(mx) + q.bx + 1, ny = vr_f2bits(my) + q.by + 1, nz = vr_f2bits(mz) + q.bz + 1;
    if (VR_FMA(VR_SUB(mx, VR_MAGIC_F), r.delta.x, q.t0x) > T) nx -= 1;
    if (VR_FMA(VR_SUB(my, VR_MAGIC_F), r.delta.y, q.t0y) > T) ny -= 1;
    if (VR_FMA(VR_SUB(mz, VR_MAGIC_F), r.delta.z, q.t0z) > T) nz -= 1;
    /* an axis whose next crossing lies beyond T makes none: the linear estimate may point far below that when T is
     * much smaller than the axis' first crossing time (negative get_oct_vox bias, kernel:353) */
    nx = nx > q.px ? nx : q.px; ny = ny > q.py ? ny : q.py; nz = nz > q.pz ? nz : q.pz;
    fm = (Tx == T ? 1 : 0) | (Ty == T ? 2 : 0) | (Tz == T ? 4 : 0);
    int n = (nx - q.px) + (ny - q.py) + (nz - q.pz);
    if (fm & (fm - 1)) n -= (fm == 7) ? 2 : 1;                            /* the axes of the last step moved together */
    xr = (nx ^ q.px) | (ny ^ q.py) | (nz ^ q.pz);
    q.px = nx; q.py = ny; q.pz = nz;
    return n;
}

/* Leaf brick (4^3 voxels, occupancy = mask): the step of kernel:558-560 with closed-form times, followed by a bit test
 * of the voxel entered.  Stops when a step leaves the brick (returns false) or lands on a set voxel (returns true,
 * `bit` = its slot).  The caller guarantees that max_distance cannot be reached inside (a brick holds <= 10 steps).
 * Every step is observed here, so multi-axis steps are exact (n counts them once). */
template <class Stack>
VR_HD bool vr_canon_brick(vr_cray<Stack> &q, unsigned long long mask, int &n, int &fm, int &xr, int &bit) {
    const RayState &r = q.r;
    const int lx = q.px & 3, ly = q.py & 3, lz = q.pz & 3;
    float kx = VR_SUB(vr_bits2f(q.px - q.bx), VR_MAGIC_F), ky = VR_SUB(vr_bits2f(q.py - q.by), VR_MAGIC_F),
          kz = VR_SUB(vr_bits2f(q.pz - q.bz), VR_MAGIC_F);
    float tx = VR_FMA(kx, r.delta.x, q.t0x), ty = VR_FMA(ky, r.delta.y, q.t0y), tz = VR_FMA(kz, r.delta.z, q.t0z);
    float rx = (float)(4 - lx), ry = (float)(4 - ly), rz = (float)(4 - lz), steps = 0.0f;
    float bitf = (float)(lx | (ly << 2) | (lz << 4));
    const int cf = (int)(q.flip & 63u);     /* (a leaf is the root only in a 4^3 map, where both flips coincide) */
    float ex, ey, ez;
    bool hit;
    for (;;) {
        const float mn = vr_min3(tx, ty, tz);
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
        if ((mask >> bit) & 1ull) { hit = true; break; }
    }
    const int nx = q.px + (4 - lx) - (int)rx, ny = q.py + (4 - ly) - (int)ry, nz = q.pz + (4 - lz) - (int)rz;
    xr = (nx ^ q.px) | (ny ^ q.py) | (nz ^ q.pz);
    q.px = nx; q.py = ny; q.pz = nz;
    fm = (ex != 0.0f ? 1 : 0) | (ey != 0.0f ? 2 : 0) | (ez != 0.0f ? 4 : 0);
    n = (int)steps;
    return hit;
}

/* point query from the root (slow path only) */
VR_HD int vr_tree_voxel(const vr_frame_params &P, int x, int y, int z) {
    uint32_t idx = 0;
    for (int s = P.root_shift;; s -= 2) {
        const vr_node_regs nd = vr_load_node(P, idx);
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
        if (!(r.dist < r.max_distance && r.bounce < 2)) {
            if (MULTI && r.bounce < 2 && vr_more_lights(P, r)) {
                if (!vr_next_light(P, r)) return VR_ST_SKIP_REDIRECT;
                r.dist++;
                t0 = r.t; kx = ky = kz = 0.0f;
                continue;
            }
            return r.bounce >= 2 ? VR_ST_BOUNCES : VR_ST_MAXDIST;
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

/* Whole pixel.  Returns true if the pixel must be written (packed colour in *rgba_out). */
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
    q.first_hit_done = false;
    int status = VR_ST_MAXDIST;
    bool slow = !vr_ray_finite(r);
    if (!slow) {
        vr_canon_enter(P, q);
        vr_canon_cursor_reset(P, q);
        q.stk.set(0, 0u);
        if (AUX) a->node_fetches = 1;
        /* the voxel a ray (re)starts in is never tested (the reference steps before it loads, kernel:555-570) and
         * its entry is not a step: distance_traveled starts one below, the common increment below brings it to 0 */
        r.dist = -1;
        bool fresh = true;
        int xr = 0, fm = 0, bit = 0;
        bool known = false;                /* the brick walk already knows that the voxel entered is set */
        for (;;) {
            /* ---- (1) what does voxel p hold? */
            int cs = 0, voxel_data = 0;
            bool brick = false;
            bool relight = false;          /* multi-light extension: this light is not blocked, on to the next one */
            if (known) {
                voxel_data = (int)(int8_t)P.leaf_types[q.node.base + (uint32_t)VR_POPC64(q.node.mask & ((1ull << bit) - 1ull))];
            } else if ((unsigned)(q.px | q.py | q.pz) < (unsigned)N) {
                voxel_data = vr_canon_lookup<AUX>(P, q, xr, cs, brick, a);
            } else {
                /* a ray may start outside the map (and a redirect may restart there): that voxel is a cell of its own */
                vr_canon_cursor_reset(P, q);
            }
            known = false;
            /* ---- (2) hit handling (kernel:575-711) */
            if ((voxel_data == 5 || voxel_data == 6) && !fresh) {
                if (r.shadow) {                                          /* kernel:706-710; nothing else of the block is live */
                    r.color.w = MULTI ? VR_ADD(r.alpha_before, 0.1f) : 0.1f;
                    if (!(MULTI && vr_more_lights(P, r))) { status = VR_ST_SHADOW_HIT; break; }
                    relight = true;
                } else {
                    vr_canon_materialize(P, q, fm);
                    const vi3 hv = r.voxel;
                    const int st = vr_hit_block<AUX, MULTI>(P, r, voxel_data, a, q.first_hit_done);
                    if (st >= 0) { status = st; break; }
                    /* redirected: the ray restarts in the voxel it came from; the increment of kernel:714 that ends
                     * the hit iteration is the common one below */
                    xr = (hv.x ^ r.voxel.x) | (hv.y ^ r.voxel.y) | (hv.z ^ r.voxel.z);
                    if (!vr_ray_finite(r)) { slow = true; break; }
                    vr_canon_enter(P, q);
                    fresh = true;
                    continue;
                }
            } else {
                if (voxel_data == 5 || voxel_data == 6) { cs = 0; brick = false; }   /* a ray that starts inside a set voxel */
                fresh = false;
                r.dist++;                                                 /* kernel:714 */
                /* ---- (3) kernel:357, then the walk through the empty cell around p */
                if (!(r.dist < r.max_distance && r.bounce < 2)) {
                    if (!(MULTI && r.bounce < 2 && vr_more_lights(P, r))) { status = r.bounce >= 2 ? VR_ST_BOUNCES : VR_ST_MAXDIST; break; }
                    relight = true;
                } else {
                    const int nmax = r.max_distance - r.dist;
                    int n;
                    bool tie = false;
                    if (brick && nmax > 12) {
                        const int before = q.px + q.py + q.pz;
                        known = vr_canon_brick(q, q.node.mask, n, fm, xr, bit);
                        tie = (q.px + q.py + q.pz - before) != n;
                    } else {
                        n = vr_canon_walk(q, brick ? 0 : (1 << cs) - 1, fm, xr);
                    }
                    if (AUX && ((fm & (fm - 1)) || tie)) a->flags |= VR_FL_TIE;
                    if (n > nmax) {                                       /* max_distance is reached inside the cell */
                        r.dist = r.max_distance;
                        if (!(MULTI && r.bounce < 2 && vr_more_lights(P, r))) { status = VR_ST_MAXDIST; break; }
                        relight = true;
                    } else {
                        r.dist += n - 1;
                        if (!known && (unsigned)(q.px | q.py | q.pz) >= (unsigned)N) {     /* kernel:563 */
                            if (!(MULTI && vr_more_lights(P, r))) {
                                r.fm = 0;                                 /* the voxel is dead: only the colour of kernel:565 */
                                vr_out_of_bounds(r);
                                status = VR_ST_OOB;
                                break;
                            }
                            relight = true;
                        }
                    }
                }
            }
            if (MULTI && relight) {
                if (!vr_next_light(P, r)) { status = VR_ST_SKIP_REDIRECT; break; }
                if (!vr_ray_finite(r)) { slow = true; break; }
                vr_canon_enter(P, q);
                vr_canon_cursor_reset(P, q);
                fresh = true;
                known = false;
                xr = 0;
            }
        }
        if (slow) r.dist++;                /* the increment that ends the iteration of the redirect */
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
