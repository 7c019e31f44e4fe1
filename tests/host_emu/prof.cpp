/*
 * prof.cpp -- TEST / ANALYSIS INFRASTRUCTURE: runs the device core (vr_trace.h, host build with VR_PROFILE) the
 * way a warp runs it -- the 32 pixels of an 8x4 block in lockstep, one vr_svo_cell call per lane per round --
 * and reports where the cells, steps and (modelled) issue slots of a frame go.  It lets the traversal be
 * redesigned on the CPU box; the numbers that count are still the ncu ones (profiles/).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#define VR_PROFILE 1
#include "../../voxel-raycaster_b200/csrc/vr_trace.h"
#include "../../voxel-raycaster_b200/csrc/vr_octree.h"

thread_local vr_prof_rec vr_prof;

struct LocalStack {
    uint32_t v[VR_MAX_LEVELS];
    void set(int l, uint32_t x) { v[l] = x; }
    uint32_t get(int l) const { return v[l]; }
};

enum { P_NONE = 0, P_BRICK = 1, P_AXES = 2, P_MERGED = 3, P_MERGED_FB = 4, NPATH = 5 };
enum { NH = 12 };

struct Totals {
    double cells[NPATH], steps[NPATH], hist[NPATH][NH], hist_steps[NPATH][NH];
    double lookups, pops, loads, hits, replays, chain, fix, rounds, lane_rounds, rays, adds, jumps;
    double model_warp, model_lane;       /* modelled warp issue slots (SIMT: max per path) and lane slots (sum) */
    double comp_warp[8];
    double grp_rounds[8], grp_adds[8], grp_model[8], grp_rays[8];   /* by (tile row mod 8): screen-row periodicity of the cost */
    double mixed_rounds, alt_warp, alt_iters_small, alt_iters_axes;  /* phase-batched schedule (see emu_profile) */
};

static int bucket(int n) { int b = 0; while ((1 << b) < n && b < NH - 1) b++; return b; }

/* crude issue-slot model of one vr_svo_cell call, split in components that diverge from each other */
struct Cost { float brick, axes, merged, lookup, hit; };
static Cost cost_of(const vr_prof_rec &p, const float *k) {
    Cost c = {0, 0, 0, 0, 0};
    if (p.path == P_BRICK) c.brick = k[0] + k[1] * p.n;
    if (p.path == P_AXES) c.axes = k[2] + k[3] * p.adds + k[12] * p.jumps + k[4] * p.fix;
    if (p.path == P_MERGED) c.merged = k[5] + k[6] * p.n;
    if (p.path == P_MERGED_FB) { c.axes = k[2] + k[3] * p.adds + k[12] * p.jumps + k[4] * p.fix; c.merged = k[5] + k[6] * p.n; }
    if (p.lookup) c.lookup = k[7] + k[8] * p.pops + k[9] * p.loads;
    if (p.hit) c.hit = k[10] + (p.replay ? k[6] * p.n : 0);
    return c;
}

extern "C" int emu_profile(int width, int height, const float *ray_table, const int32_t *col_lo, const int32_t *col_hi, int n,
                           const float *cam_pos, const float *cam_dir, const float *lights, const uint8_t *atlas, int atlas_w,
                           int atlas_h, int tile_w, int tile_h, int max_distance, int warp_stride, const float *model,
                           double *out, int nout) {
    vr_native_tree tree;
    if (!vr_native_from_columns(col_lo, col_hi, n, 5, tree)) return -1;
    vr_frame_params P;
    memset(&P, 0, sizeof(P));
    P.width = width; P.height = height;
    P.local_rows = height; P.band_rows = 1; P.band_stride = 1; P.band_first = 0;
    P.ray_table = ray_table;
    P.dim[0] = P.dim[1] = P.dim[2] = n;
    for (int i = 0; i < 3; i++) { P.cam_pos[i] = cam_pos[i]; P.bias[i] = 0.0f; P.light_pos[0][i] = lights[4 + i]; }
    P.cam_on_edge = vr_cam_on_edge(P.cam_pos, P.bias);
    for (int i = 0; i < 4; i++) P.light_rgbi[0][i] = lights[i];
    P.light_count = 1;
    P.trig[0] = sinf(cam_dir[0]); P.trig[1] = cosf(cam_dir[0]); P.trig[2] = sinf(cam_dir[1]); P.trig[3] = cosf(cam_dir[1]);
    P.atlas = atlas; P.atlas_dim[0] = atlas_w; P.atlas_dim[1] = atlas_h;
    P.atlas_scale[0] = atlas_w / tile_w; P.atlas_scale[1] = atlas_h / tile_h;
    P.max_distance = max_distance;
    P.max_bounces = 2;
    P.nodes = tree.nodes.data(); P.leaf_types = tree.leaf_types.data();
    P.levels = tree.levels; P.root_shift = 2 * (tree.levels - 1);

    const int wx = width / 8, wy = height / 4;
    Totals T;
    memset(&T, 0, sizeof(T));
#pragma omp parallel
    {
        Totals L;
        memset(&L, 0, sizeof(L));
#pragma omp for schedule(dynamic, 4)
        for (int w = 0; w < wx * wy; w++) {
            const int bx = w % wx, by = w / wx;
            if ((bx + 3 * by) % warp_stride) continue;
            vr_svo_ray<LocalStack> q[32];
            bool active[32];
            vr_aux a;
            int nact = 0;
            for (int l = 0; l < 32; l++) {
                active[l] = vr_svo_begin<false>(P, bx * 8 + (l & 7), by * 4 + (l >> 3), q[l], &a);
                nact += active[l];
                L.rays += active[l];
            }
            std::vector<Cost> ev[32];           /* per lane: the cost record of every vr_svo_cell call, in order */
            std::vector<int> evpath[32];
            while (nact) {
                Cost mx = {0, 0, 0, 0, 0};
                for (int l = 0; l < 32; l++) {
                    if (!active[l]) continue;
                    memset(&vr_prof, 0, sizeof(vr_prof));
                    const bool was_shadow = q[l].r.shadow;
                    const int rc = vr_svo_round<false, 1, false>(P, q[l], &a);
                    if (!was_shadow && q[l].r.shadow) L.rays += 1;
                    const vr_prof_rec &p = vr_prof;
                    const int b = bucket(p.n);
                    L.cells[p.path] += 1; L.steps[p.path] += p.n; L.hist[p.path][b] += 1; L.hist_steps[p.path][b] += p.n;
                    L.lookups += p.lookup; L.pops += p.pops; L.loads += p.loads; L.hits += p.hit; L.replays += p.replay;
                    L.chain += p.chain_a + p.chain_b; L.fix += p.fix; L.adds += p.adds; L.jumps += p.jumps;
                    L.lane_rounds += 1;
                    L.grp_adds[by & 7] += p.adds + p.fix + 35 * p.jumps;
                    const Cost c = cost_of(p, model);
                    ev[l].push_back(c);
                    evpath[l].push_back(p.path);
                    L.model_lane += c.brick + c.axes + c.merged + c.lookup + c.hit + model[11];
                    mx.brick = fmaxf(mx.brick, c.brick); mx.axes = fmaxf(mx.axes, c.axes); mx.merged = fmaxf(mx.merged, c.merged);
                    mx.lookup = fmaxf(mx.lookup, c.lookup); mx.hit = fmaxf(mx.hit, c.hit);
                    if (rc != VR_CELL_CONTINUE) { active[l] = false; nact--; }
                }
                L.rounds += 1;
                if (mx.axes > 0 && (mx.brick > 0 || mx.merged > 0)) L.mixed_rounds += 1;
                L.grp_rounds[by & 7] += 1;
                L.grp_model[by & 7] += mx.brick + mx.axes + mx.merged + mx.lookup + mx.hit + model[11];
                L.model_warp += mx.brick + mx.axes + mx.merged + mx.lookup + mx.hit + model[11];
                L.comp_warp[0] += mx.brick; L.comp_warp[1] += mx.axes; L.comp_warp[2] += mx.merged; L.comp_warp[3] += mx.lookup;
                L.comp_warp[4] += mx.hit; L.comp_warp[5] += model[11];
            }
            /* alternative schedule: lanes whose next cell is small (brick / merged walk) iterate together until none is
             * left in a small cell, then the lanes whose next cell takes the per-axis walk do one cell together, ... */
            size_t at[32] = {0};
            for (;;) {
                bool any_small = false, any_axes = false;
                for (int l = 0; l < 32; l++) {
                    if (at[l] >= ev[l].size()) continue;
                    const int pth = evpath[l][at[l]];
                    if (pth == P_AXES || pth == P_MERGED_FB) any_axes = true; else any_small = true;
                }
                if (!any_small && !any_axes) break;
                Cost mx = {0, 0, 0, 0, 0};
                for (int l = 0; l < 32; l++) {
                    if (at[l] >= ev[l].size()) continue;
                    const int pth = evpath[l][at[l]];
                    const bool is_axes = (pth == P_AXES || pth == P_MERGED_FB);
                    if (any_small ? is_axes : !is_axes) continue;
                    const Cost &c = ev[l][at[l]++];
                    mx.brick = fmaxf(mx.brick, c.brick); mx.axes = fmaxf(mx.axes, c.axes); mx.merged = fmaxf(mx.merged, c.merged);
                    mx.lookup = fmaxf(mx.lookup, c.lookup); mx.hit = fmaxf(mx.hit, c.hit);
                }
                L.alt_warp += mx.brick + mx.axes + mx.merged + mx.lookup + mx.hit + model[11];
                if (any_small) L.alt_iters_small += 1; else L.alt_iters_axes += 1;
            }
        }
#pragma omp critical
        {
            double *t = (double *)&T, *l = (double *)&L;
            for (size_t i = 0; i < sizeof(Totals) / sizeof(double); i++) t[i] += l[i];
        }
    }
    const int need = (int)(sizeof(Totals) / sizeof(double));
    if (nout < need) return -2;
    memcpy(out, &T, sizeof(T));
    return need;
}
