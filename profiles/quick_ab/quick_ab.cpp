/*
 * quick_ab.cpp -- a GPU check that fits into the last seconds of a round's GPU budget: no Python, no torch import.
 *
 *   1. A/B on one box, one process: the C3 bench frame (1024^3 shell terrain from column tables, 3840x2160, bench camera,
 *      one light, library defaults) rendered by the CURRENT libvrcaster.so and by the build of the commit before the
 *      session's kernel change (build/quick/libvrcaster_prev.so), frames alternating; per library the minimum and median of
 *      vr_stats.last_kernel_ms (CUDA events around the ray kernel), the frame's FNV-1a hash and bench.py's checksum
 *      (sum of frame[::64, ::64]).
 *   2. The three camera-on-a-voxel-edge scenes of tests/test_gpu_zz_edge_camera.py against their Oracle-B frames
 *      (inputs + expected RGBA8 written by make_data.py), current library, both kinds of top grid.
 *
 *   g++ -O2 -std=c++17 -fopenmp profiles/quick_ab/quick_ab.cpp -o build/quick/quick_ab -ldl
 *   ./build/quick/quick_ab            (from the repo root; prints progressively, flushes after every line)
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/vr_caster.h"

struct Lib {
    void *h = nullptr;
    decltype(&vr_init) init;
    decltype(&vr_destroy) destroy;
    decltype(&vr_last_error) last_error;
    decltype(&vr_add_to_settings_buffer) add_setting;
    decltype(&vr_assign_columns) assign_columns;
    decltype(&vr_assign_map) assign_map;
    decltype(&vr_assign_octree) assign_octree;
    decltype(&vr_octree_generate) octree_generate;
    decltype(&vr_assign_camera) assign_camera;
    decltype(&vr_create_viewport) create_viewport;
    decltype(&vr_assign_lights) assign_lights;
    decltype(&vr_create_texture_atlas) create_atlas;
    decltype(&vr_validate) validate;
    decltype(&vr_compute) compute;
    decltype(&vr_read_framebuffer) read_fb;
    decltype(&vr_get_stats) get_stats;
    decltype(&vr_set_option) set_option;
};

static double now_s() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static double T0;
#define SAY(...) do { printf("[%6.3f s] ", now_s() - T0); printf(__VA_ARGS__); printf("\n"); fflush(stdout); } while (0)

static bool load(Lib &L, const char *path) {
    L.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!L.h) { SAY("dlopen %s: %s", path, dlerror()); return false; }
#define SYM(field, name) *(void **)&L.field = dlsym(L.h, name); if (!L.field) { SAY("missing %s in %s", name, path); return false; }
    SYM(init, "vr_init") SYM(destroy, "vr_destroy") SYM(last_error, "vr_last_error") SYM(add_setting, "vr_add_to_settings_buffer")
    SYM(assign_columns, "vr_assign_columns") SYM(assign_map, "vr_assign_map") SYM(assign_octree, "vr_assign_octree")
    SYM(octree_generate, "vr_octree_generate") SYM(assign_camera, "vr_assign_camera") SYM(create_viewport, "vr_create_viewport")
    SYM(assign_lights, "vr_assign_lights") SYM(create_atlas, "vr_create_texture_atlas") SYM(validate, "vr_validate")
    SYM(compute, "vr_compute") SYM(read_fb, "vr_read_framebuffer") SYM(get_stats, "vr_get_stats") SYM(set_option, "vr_set_option")
#undef SYM
    return true;
}

struct Scene {
    int32_t n = 0, w = 0, h = 0, lights_n = 1, max_distance = 0;
    float pos[3], dir[2], lights[80];
    std::vector<int32_t> lo, hi;       /* column tables (c3) */
    std::vector<int8_t> vol;           /* dense map (corner scenes) */
    std::vector<uint8_t> expect;       /* Oracle-B frame (corner scenes) */
};

static bool read_scene(const char *path, bool columns, Scene &s) {
    FILE *f = fopen(path, "rb");
    if (!f) { SAY("cannot open %s", path); return false; }
    int32_t hd[5];
    bool ok = fread(hd, 4, 5, f) == 5 && fread(s.pos, 4, 3, f) == 3 && fread(s.dir, 4, 2, f) == 2 && fread(s.lights, 4, 80, f) == 80;
    s.n = hd[0]; s.w = hd[1]; s.h = hd[2]; s.lights_n = hd[3]; s.max_distance = hd[4];
    if (ok && columns) {
        s.lo.resize((size_t)s.n * s.n); s.hi.resize((size_t)s.n * s.n);
        ok = fread(s.lo.data(), 4, s.lo.size(), f) == s.lo.size() && fread(s.hi.data(), 4, s.hi.size(), f) == s.hi.size();
    } else if (ok) {
        s.vol.resize((size_t)s.n * s.n * s.n); s.expect.resize((size_t)s.w * s.h * 4);
        ok = fread(s.vol.data(), 1, s.vol.size(), f) == s.vol.size() && fread(s.expect.data(), 1, s.expect.size(), f) == s.expect.size();
    }
    fclose(f);
    if (!ok) SAY("short read of %s", path);
    return ok;
}

static std::vector<uint8_t> g_atlas;

/* the call order of CUDACaster.load_scene (voxel-raycaster_b200/caster.py) */
static vr_ctx *setup(Lib &L, Scene &s, const char *tag) {
    vr_ctx *c = nullptr;
    if (!L.init(&c, 0, 0)) { SAY("%s: vr_init failed", tag); return nullptr; }
    bool ok = L.add_setting(c, "octree_dimensions", "OCTDIM", s.n) && L.add_setting(c, "using_octree", "OCTENABLED", 0) &&
              L.add_setting(c, "max_distance", "MAX_DISTANCE", s.max_distance);
    if (ok && s.lights_n > 1) ok = L.add_setting(c, "light_count", "LIGHT_COUNT", s.lights_n) != 0;
    if (ok && !s.lo.empty()) {
        ok = L.assign_columns(c, s.lo.data(), s.hi.data(), s.n, 5) != 0;
    } else if (ok) {
        uint64_t entries = 0, root = 0;
        ok = L.octree_generate(s.vol.data(), s.n, nullptr, &entries, &root) != 0;
        std::vector<uint64_t> desc(entries);
        ok = ok && L.octree_generate(s.vol.data(), s.n, desc.data(), &entries, &root) && L.assign_octree(c, desc.data(), nullptr, nullptr, entries, root) &&
             L.assign_map(c, s.vol.data(), s.n, s.n, s.n);
    }
    ok = ok && L.assign_camera(c, s.dir, s.pos) && L.create_viewport(c, s.w, s.h, 0.625f * 90.0f, 90.0f) && L.assign_lights(c, s.lights, 8) &&
         L.create_atlas(c, g_atlas.data(), 256, 256, 16, 16) && L.validate(c);
    if (!ok) { SAY("%s: set-up failed: %s", tag, L.last_error(c)); L.destroy(c); return nullptr; }
    return c;
}

static uint64_t fnv(const std::vector<uint8_t> &v) {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : v) { h ^= b; h *= 1099511628211ull; }
    return h;
}

int main() {
    T0 = now_s();
    Lib cur, prev;
    if (!load(cur, "voxel-raycaster_b200/libvrcaster.so")) return 2;
    const bool have_prev = load(prev, "build/quick/libvrcaster_prev.so");
    {
        FILE *f = fopen("build/quick/atlas.bin", "rb");
        g_atlas.resize(256 * 256 * 4);
        if (!f || fread(g_atlas.data(), 1, g_atlas.size(), f) != g_atlas.size()) { SAY("atlas.bin missing"); return 2; }
        fclose(f);
    }
    int rc = 0;
    /* ---- 1. C3 A/B */
    Scene c3;
    if (read_scene("build/quick/c3.bin", true, c3)) {
        SAY("c3 inputs read");
        vr_ctx *a = setup(cur, c3, "current"), *b = have_prev ? setup(prev, c3, "previous") : nullptr;
        SAY("contexts ready (current %s, previous %s)", a ? "ok" : "FAILED", b ? "ok" : "absent");
        if (a) {
            const int warm = 10, frames = 60;
            std::vector<float> ta, tb;
            vr_stats st;
            bool ok = true;
            for (int i = 0; ok && i < warm + frames; i++) {
                ok = cur.compute(a) != 0;
                if (ok && i >= warm) { cur.get_stats(a, &st); ta.push_back(st.last_kernel_ms); }
                if (ok && b) {
                    ok = prev.compute(b) != 0;
                    if (ok && i >= warm) { prev.get_stats(b, &st); tb.push_back(st.last_kernel_ms); }
                }
            }
            if (!ok) { SAY("compute failed: %s", cur.last_error(a)); rc = 1; }
            auto report = [&](const char *tag, Lib &L, vr_ctx *c, std::vector<float> &t) {
                if (!c || t.empty()) return;
                std::sort(t.begin(), t.end());
                std::vector<uint8_t> fb((size_t)c3.w * c3.h * 4);
                L.read_fb(c, fb.data(), fb.size());
                long long sum = 0;
                for (int y = 0; y < c3.h; y += 64)
                    for (int x = 0; x < c3.w; x += 64)
                        for (int k = 0; k < 4; k++) sum += fb[4 * ((size_t)x + (size_t)c3.w * y) + k];
                L.get_stats(c, &st);
                SAY("C3 %-8s kernel ms/frame: min %.4f  median %.4f  p90 %.4f  (%zu frames)  frame fnv %016llx  checksum %lld  nodes %llu",
                    tag, t.front(), t[t.size() / 2], t[t.size() * 9 / 10], t.size(), (unsigned long long)fnv(fb), sum, (unsigned long long)st.native_nodes);
            };
            report("current", cur, a, ta);
            report("previous", prev, b, tb);
        } else rc = 1;
        if (a) cur.destroy(a);
        if (b) prev.destroy(b);
    }
    /* ---- 2. camera on a voxel edge / corner: current library == Oracle-B */
    for (const char *name : {"corner", "edge", "biased"}) {
        Scene s;
        if (!read_scene((std::string("build/quick/") + name + ".bin").c_str(), false, s)) { rc = 1; continue; }
        vr_ctx *c = setup(cur, s, name);
        if (!c) { rc = 1; continue; }
        for (int directed = 1; directed >= 0; directed--) {
            std::vector<uint8_t> fb(s.expect.size());
            if (!cur.set_option(c, "directed_grid", directed) || !cur.compute(c) || !cur.read_fb(c, fb.data(), fb.size())) {
                SAY("%s: compute failed: %s", name, cur.last_error(c)); rc = 1; continue;
            }
            size_t bad = 0;
            for (size_t i = 0; i < fb.size(); i += 4) bad += memcmp(&fb[i], &s.expect[i], 4) != 0;
            vr_stats st;
            cur.get_stats(c, &st);
            SAY("%-7s %d^3 %dx%d lights %d max_distance %d bias (%d,%d,%d) directed_grid %d: %zu of %zu pixels differ from Oracle-B%s", name, s.n, s.w, s.h,
                s.lights_n, s.max_distance, st.bias[0], st.bias[1], st.bias[2], directed, bad, fb.size() / 4, bad ? "  <-- MISMATCH" : "");
            if (bad) rc = 1;
        }
        cur.destroy(c);
    }
    SAY("done rc=%d", rc);
    return rc;
}
