/*
 * emu.cpp -- TEST INFRASTRUCTURE: compiles the device core (voxel-raycaster_b200/csrc/vr_trace.h) for
 * the host so its control flow can be single-stepped against the oracle without a GPU.  The product
 * never loads this file; it exists because a gpurun round-trip takes minutes.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <vector>

#include "../../voxel-raycaster_b200/csrc/vr_trace.h"
#include "../../voxel-raycaster_b200/csrc/vr_canon.h"
#include "../../voxel-raycaster_b200/csrc/vr_octree.h"

struct LocalStack {
    uint32_t v[VR_MAX_LEVELS];
    void set(int l, uint32_t x) { v[l] = x; }
    uint32_t get(int l) const { return v[l]; }
};

/* solid-subtree collapse (vr_octree.cpp: vr_native_collapse_solid) applied to every tree built below: 0 / 1 */
static int g_collapse = 1;
extern "C" void emu_set_collapse(int on) { g_collapse = on; }

extern "C" int emu_raycast(int width, int height, const float *ray_table, const int8_t *map, int n,
                           const float *cam_pos, const float *cam_dir, const int32_t *bias, const float *lights,
                           const uint8_t *atlas, int atlas_w, int atlas_h, int tile_w, int tile_h, int max_distance,
                           int use_svo, int shadow_lights, uint8_t *rgba, vr_aux *aux) {
    vr_native_tree tree;
    vr_frame_params P;
    memset(&P, 0, sizeof(P));
    P.width = width; P.height = height;
    P.local_rows = height; P.band_rows = 1; P.band_stride = 1; P.band_first = 0;
    P.ray_table = ray_table;
    P.map = map;
    P.dim[0] = P.dim[1] = P.dim[2] = n;
    for (int i = 0; i < 3; i++) { P.cam_pos[i] = cam_pos[i]; P.bias[i] = (float)bias[i]; }
    P.cam_on_edge = vr_cam_on_edge(P.cam_pos, P.bias);
    P.light_count = shadow_lights < 1 ? 1 : (shadow_lights > VR_MAX_LIGHTS ? VR_MAX_LIGHTS : shadow_lights);
    for (int l = 0; l < P.light_count; l++) {
        for (int i = 0; i < 4; i++) P.light_rgbi[l][i] = lights[10 * l + i];
        for (int i = 0; i < 3; i++) P.light_pos[l][i] = lights[10 * l + 4 + i];
    }
    P.trig[0] = sinf(cam_dir[0]); P.trig[1] = cosf(cam_dir[0]); P.trig[2] = sinf(cam_dir[1]); P.trig[3] = cosf(cam_dir[1]);
    P.atlas = atlas; P.atlas_dim[0] = atlas_w; P.atlas_dim[1] = atlas_h;
    P.atlas_scale[0] = atlas_w / tile_w; P.atlas_scale[1] = atlas_h / tile_h;
    P.max_distance = max_distance;
    P.max_bounces = 2;
    if (use_svo) {
        if (!vr_native_from_dense(map, n, tree)) return -1;
        if (g_collapse) vr_native_collapse_solid(tree);
        P.nodes = tree.nodes.data(); P.leaf_types = tree.leaf_types.data();
        P.levels = tree.levels; P.root_shift = 2 * (tree.levels - 1);
    }
    std::vector<uint32_t> grid;
    if (use_svo >= 3) {                      /* 3: closed-form walk over the undirected top grid, 4: over the directed ones */
        int gs = 0, gb = 0;
        if (!(use_svo == 4 ? vr_native_grid_directed : vr_native_grid)(tree.nodes.data(), tree.levels, n, grid, &gs, &gb)) return -3;
        P.grid = grid.data(); P.grid_shift = gs; P.grid_bits = gb; P.grid_dim = 1 << gb;
        P.grid_directed = use_svo == 4;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            uint32_t px;
            vr_aux a;
            bool w;
            LocalStack s;
            if (P.light_count > 1) {
                if (use_svo >= 3) w = vr_trace_svo_canon<true, true>(P, x, y, &px, &a, s);
                else if (use_svo == 2) w = vr_trace_svo<true, 1, true>(P, x, y, &px, &a, s);
                else if (use_svo) w = vr_trace_svo<true, 0, true>(P, x, y, &px, &a, s);
                else w = vr_trace_dense<true, true>(P, x, y, &px, &a);
            } else {
                if (use_svo >= 3) w = vr_trace_svo_canon<true, false>(P, x, y, &px, &a, s);
                else if (use_svo == 2) w = vr_trace_svo<true, 1, false>(P, x, y, &px, &a, s);
                else if (use_svo) w = vr_trace_svo<true, 0, false>(P, x, y, &px, &a, s);
                else w = vr_trace_dense<true, false>(P, x, y, &px, &a);
            }
            const size_t i = (size_t)x + (size_t)width * y;
            if (w) memcpy(rgba + 4 * i, &px, 4);
            if (aux) aux[i] = a;
        }
    return 0;
}

/* 64-tree builders exposed for CPU tests: both write up to cap_nodes 16-byte nodes / cap_types bytes and return
 * the node count (types count in *ntypes, levels in *levels); -1 on failure. */
static long emu_export(const vr_native_tree &t, void *nodes, long cap_nodes, uint8_t *types, long cap_types, long *ntypes, int *levels) {
    if ((long)t.nodes.size() > cap_nodes || (long)t.leaf_types.size() > cap_types) return -2;
    memcpy(nodes, t.nodes.data(), t.nodes.size() * sizeof(vr_node));
    memcpy(types, t.leaf_types.data(), t.leaf_types.size());
    *ntypes = (long)t.leaf_types.size();
    *levels = t.levels;
    return (long)t.nodes.size();
}

extern "C" long emu_tree_from_dense(const int8_t *map, int n, void *nodes, long cap_nodes, uint8_t *types, long cap_types, long *ntypes, int *levels) {
    vr_native_tree t;
    if (!vr_native_from_dense(map, n, t)) return -1;
    if (g_collapse) vr_native_collapse_solid(t);
    return emu_export(t, nodes, cap_nodes, types, cap_types, ntypes, levels);
}

extern "C" long emu_tree_from_columns(const int32_t *lo, const int32_t *hi, int n, int type, void *nodes, long cap_nodes, uint8_t *types, long cap_types, long *ntypes, int *levels) {
    vr_native_tree t;
    if (!vr_native_from_columns(lo, hi, n, (uint8_t)type, t)) return -1;
    if (g_collapse) vr_native_collapse_solid(t);
    return emu_export(t, nodes, cap_nodes, types, cap_types, ntypes, levels);
}

/* the host version of the closed-form walk's top grid, from 64-tree arrays (nodes as 16-byte records) */
extern "C" long emu_grid_from_tree(const void *nodes, int levels, int dim, uint32_t *out, long cap, int *shift, int *bits) {
    std::vector<uint32_t> g;
    if (!vr_native_grid((const vr_node *)nodes, levels, dim, g, shift, bits)) return -1;
    if ((long)g.size() > cap) return -2;
    memcpy(out, g.data(), g.size() * sizeof(uint32_t));
    return (long)g.size();
}

/* vr_div_const (vr_trace.h) against the IEEE division it replaces, over the whole range of its operands: texel / 255 for
 * texel = 0..255 and steps / 700 for steps = 0..2^24.  Returns the number of differing quotients. */
extern "C" long emu_div_const_mismatches() {
    long bad = 0;
    for (int i = 0; i <= 255; i++) {
        volatile float x = (float)i;
        const float q = VR_DIV_255((float)i), ref = x / 255.0f;
        bad += memcmp(&q, &ref, 4) != 0;
    }
    for (int i = 0; i <= (1 << 24); i++) {
        volatile float x = (float)i;
        const float q = VR_DIV_700((float)i), ref = x / 700.0f;
        bad += memcmp(&q, &ref, 4) != 0;
    }
    return bad;
}

/* the directed top grids (vr_octree.cpp: vr_native_grid_directed): eight tables, octant 0 first */
extern "C" long emu_grid_directed_from_tree(const void *nodes, int levels, int dim, uint32_t *out, long cap, int *shift, int *bits) {
    std::vector<uint32_t> g;
    if (!vr_native_grid_directed((const vr_node *)nodes, levels, dim, g, shift, bits)) return -1;
    if ((long)g.size() > cap) return -2;
    memcpy(out, g.data(), g.size() * sizeof(uint32_t));
    return (long)g.size();
}

/* vr_add_chain (binade jumps) against the literal chain of additions it stands for */
extern "C" void emu_add_chain(const float *t, const float *d, const int32_t *n, int count, float *jumped, float *literal) {
    for (int i = 0; i < count; i++) {
        jumped[i] = vr_add_chain(t[i], d[i], n[i]);
        volatile float v = t[i];
        for (int k = 0; k < n[i]; k++) v = v + d[i];
        literal[i] = v;
    }
}

/* importer of reference-format descriptor buffers (what vr_assign_octree alone leads to): occupancy only, type 5 */
extern "C" long emu_tree_from_ref(const uint64_t *desc, uint64_t len, uint64_t root, int dim, void *nodes, long cap_nodes, uint8_t *types,
                                  long cap_types, long *ntypes, int *levels) {
    vr_native_tree t;
    if (!vr_native_from_ref(desc, len, root, dim, nullptr, t)) return -1;
    if (g_collapse) vr_native_collapse_solid(t);
    return emu_export(t, nodes, cap_nodes, types, cap_types, ntypes, levels);
}
