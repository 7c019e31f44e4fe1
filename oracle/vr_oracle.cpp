/*
 * vr_oracle.cpp -- CPU restatement of the reference ray caster.  TEST INFRASTRUCTURE ONLY.
 * See vr_oracle.h for scope, pinning status ("parity unpinned") and the pinned built-ins.
 *
 * "kernel:N" = /root/reference/kernels/ray_caster_kernel.cl line N
 * "host:N"   = /root/reference/src/CLCaster.cpp line N
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 */
#include "vr_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct i3 { int x, y, z; };

inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 to_f3(i3 v) { return {(float)v.x, (float)v.y, (float)v.z}; }

/* pinned built-ins (vr_oracle.h) */
inline float cl_dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float cl_length(f3 a) { return sqrtf(cl_dot(a, a)); }
inline f3 cl_normalize(f3 a) { float l = cl_length(a); return {a.x / l, a.y / l, a.z / l}; }
inline float cl_max(float x, float y) { return (x < y) ? y : x; }
inline float cl_min(float x, float y) { return (y < x) ? y : x; }
inline int cl_sign_step(float d) { return (d > 0.0f) - (d < 0.0f); }      /* kernel:298,675 */
inline int cl_convert_int(float v) {                                      /* convert_int2, rtz */
    if (!(v > -2147483648.0f && v < 2147483648.0f)) return 0;             /* NaN/out of range: pinned 0 */
    return (int)v;
}
inline uint8_t unorm8(float c) {                                          /* write_imagef, kernel:717 */
    float v = c * 255.0f;
    if (!(v > 0.0f)) return 0;                                            /* also NaN -> 0 */
    if (v > 255.0f) v = 255.0f;
    return (uint8_t)nearbyintf(v);                                        /* round half to even */
}

/* ---- view_light, kernel:78-99 ---------------------------------------------------------------- */
inline f4 view_light(f4 in_color, f3 light, f4 light_color, f3 view, i3 mask) {
    if (light.x == 0.0f && light.y == 0.0f && light.z == 0.0f)            /* kernel:80 */
        return {0.0f, 0.0f, 0.0f, 0.0f};
    float d = cl_length(light) * 0.01f;                                   /* kernel:83 */
    d *= d;
    f3 nmask = cl_normalize(to_f3(mask));
    f3 nlight = cl_normalize(light);
    float diffuse = cl_max(cl_dot(nmask, nlight), 0.1f);                  /* kernel:86 */
    float specular = 0.0f;
    if (diffuse > 0.0f) {                                                 /* kernel:89 */
        f3 halfway = cl_normalize(nlight + cl_normalize(view));           /* kernel:92 */
        float spec_tmp = cl_max(cl_dot(nmask, halfway), 0.0f);            /* kernel:93 */
        specular = spec_tmp;                                              /* pow(x, 1.0f), kernel:94 */
    }
    /* kernel:97: in_color += diffuse * light_color + specular * light_color / d */
    f4 o;
    o.x = in_color.x + (diffuse * light_color.x + (specular * light_color.x) / d);
    o.y = in_color.y + (diffuse * light_color.y + (specular * light_color.y) / d);
    o.z = in_color.z + (diffuse * light_color.z + (specular * light_color.z) / d);
    o.w = in_color.w + (diffuse * light_color.w + (specular * light_color.w) / d);
    return o;
}

/* ---- get_oct_vox, kernel:140-251 ------------------------------------------------------------- */
struct TraversalState {
    i3 sub_oct_pos;
    int parent_stack_position;
    uint64_t parent_stack[32];          /* kernel:119 uses 8; widened */
    uint64_t parent_stack_index[32];
    int scale;
    uint8_t idx_stack[32];
    uint64_t current_descriptor;
    uint64_t current_descriptor_index;
    i3 oct_pos;
    int resolution;
    int found;
};

const uint8_t mask_8[8] = {0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80};          /* kernel:37 */
const uint8_t count_mask_8[8] = {0x1, 0x3, 0x7, 0xF, 0x1F, 0x3F, 0x7F, 0xFF};    /* kernel:43 */
const uint64_t child_pointer_mask = 0x7fff;                                       /* kernel:49 */
const uint64_t far_bit_mask = 0x8000;                                             /* kernel:50 */

TraversalState get_oct_vox(i3 position, const uint64_t *buf, int64_t root_index, int64_t octdim) {
    TraversalState ts;
    memset(&ts, 0, sizeof(ts));
    ts.current_descriptor_index = (uint64_t)root_index;                   /* kernel:150 */
    ts.current_descriptor = buf[ts.current_descriptor_index];
    ts.scale = 0;
    ts.parent_stack_position = 0;
    ts.found = 0;
    ts.parent_stack[0] = ts.current_descriptor;
    ts.parent_stack_index[0] = ts.current_descriptor_index;
    int dimension = (int)octdim;                                          /* kernel:162 */
    ts.resolution = dimension / 2;
    ts.oct_pos = {0, 0, 0};
    ts.sub_oct_pos = ts.oct_pos;

    while (dimension > 1) {                                               /* kernel:176 */
        ts.oct_pos = ts.sub_oct_pos;
        int half = dimension / 2;
        bool gx = position.x >= half + ts.oct_pos.x;                      /* kernel:181-183 */
        bool gy = position.y >= half + ts.oct_pos.y;
        bool gz = position.z >= half + ts.oct_pos.z;
        ts.idx_stack[ts.scale] = (uint8_t)((gx ? 1 : 0) | (gy ? 2 : 0) | (gz ? 4 : 0));
        ts.sub_oct_pos.x += gx ? half : 0;                                /* kernel:191 */
        ts.sub_oct_pos.y += gy ? half : 0;
        ts.sub_oct_pos.z += gz ? half : 0;
        int mask_index = ts.idx_stack[ts.scale];

        if ((ts.current_descriptor >> 16) & mask_8[mask_index]) {         /* valid, kernel:196 */
            if ((ts.current_descriptor >> 24) & mask_8[mask_index]) {     /* leaf, kernel:199 */
                ts.found = 1;
                return ts;                                                /* kernel:205 */
            }
            ts.scale++;
            ts.parent_stack_position++;
            dimension /= 2;
            ts.resolution /= 2;
            int count = __builtin_popcount((uint8_t)(ts.current_descriptor >> 16) &
                                           count_mask_8[mask_index]) - 1;         /* kernel:218 */
            if (far_bit_mask & buf[ts.current_descriptor_index]) {        /* kernel:222 */
                uint64_t far_pointer_index =
                    ts.current_descriptor_index + (ts.current_descriptor & child_pointer_mask);
                ts.current_descriptor_index = buf[far_pointer_index] + (uint64_t)count;
            } else {                                                      /* kernel:229 */
                ts.current_descriptor_index =
                    ts.current_descriptor_index + (ts.current_descriptor & child_pointer_mask) + (uint64_t)count;
            }
            ts.current_descriptor = buf[ts.current_descriptor_index];     /* kernel:233 */
            ts.parent_stack[ts.parent_stack_position] = ts.current_descriptor;
            ts.parent_stack_index[ts.parent_stack_position] = ts.current_descriptor_index;
        } else {
            ts.found = 0;                                                 /* kernel:245 */
            return ts;
        }
    }
    ts.found = 1;                                                         /* kernel:249 */
    return ts;
}

/* ---- canonical descent bookkeeping for the D_svo byte model ------------
 * Not part of the reference: it counts how many child descriptors a stack-based traversal of the
 * reference's own octree must fetch when the ray moves into a new cell. */
struct CellTrack {
    int path_len = 0;        /* descriptors on the root..cell path (root included)               */
    i3 origin = {0, 0, 0};   /* cell origin                                                      */
    int size = 0;            /* cell edge in voxels; 0 = none                                    */
    i3 last = {0, 0, 0};
};

inline void svo_lookup(const vro_scene *s, i3 v, CellTrack &c, vro_counters &k) {
    if (c.size && v.x >= c.origin.x && v.x < c.origin.x + c.size && v.y >= c.origin.y &&
        v.y < c.origin.y + c.size && v.z >= c.origin.z && v.z < c.origin.z + c.size) {
        c.last = v;
        return;                                                           /* same cell: no fetch */
    }
    /* full descent to find the new cell */
    int dimension = (int)s->octdim;
    i3 o = {0, 0, 0};
    int len = 1;
    if (!s->oct_desc) {
        /* the same path read off a 4^3-per-node occupancy tree (vro_scene::tree64): every node is two 2^3 levels */
        const uint32_t *t = s->tree64;
        uint32_t idx = 0;
        for (int sh = 2 * (s->tree64_levels - 1);; sh -= 2) {
            const uint64_t m = (uint64_t)t[4 * idx] | ((uint64_t)t[4 * idx + 1] << 32);
            const int cx = (v.x >> sh) & 3, cy = (v.y >> sh) & 3, cz = (v.z >> sh) & 3;
            const int ci = cx | (cy << 2) | (cz << 4);
            /* upper 2^3 level: the octant of 2x2x2 slots around the slot, edge 2 << sh */
            dimension = 2 << sh;
            o = {(v.x >> (sh + 1)) << (sh + 1), (v.y >> (sh + 1)) << (sh + 1), (v.z >> (sh + 1)) << (sh + 1)};
            if (((m >> (ci & 0x2A)) & 0x00330033ull) == 0ull) break;      /* !valid: an empty child */
            len++;                                                        /* the octant's descriptor */
            /* lower 2^3 level: the slot, edge 1 << sh */
            dimension = 1 << sh;
            o = {(v.x >> sh) << sh, (v.y >> sh) << sh, (v.z >> sh) << sh};
            if (!((m >> ci) & 1ull) || sh == 0) break;                    /* !valid, or half == 1 */
            len++;
            idx = t[4 * idx + 2] + (uint32_t)__builtin_popcountll(m & ((1ull << ci) - 1ull));
        }
    } else {
    const uint64_t *buf = s->oct_desc;
    uint64_t idx = (uint64_t)s->oct_root_index;
    uint64_t cd = buf[idx];
    for (;;) {
        int half = dimension / 2;
        bool gx = v.x >= half + o.x, gy = v.y >= half + o.y, gz = v.z >= half + o.z;
        int ci = (gx ? 1 : 0) | (gy ? 2 : 0) | (gz ? 4 : 0);
        o.x += gx ? half : 0; o.y += gy ? half : 0; o.z += gz ? half : 0;
        bool valid = (cd >> 16) & mask_8[ci];
        bool leaf = (cd >> 24) & mask_8[ci];
        if (!valid || leaf || half == 1) { dimension = half; break; }
        int count = __builtin_popcount((uint8_t)(cd >> 16) & count_mask_8[ci]) - 1;
        if (far_bit_mask & cd) idx = buf[idx + (cd & child_pointer_mask)] + (uint64_t)count;
        else idx = idx + (cd & child_pointer_mask) + (uint64_t)count;
        cd = buf[idx];
        dimension = half;
        len++;
    }
    }
    int shared = 0;   /* path nodes (depth 0..shared-1) that contain both the old and new voxel */
    if (c.size) {
        int x = (c.last.x ^ v.x) | (c.last.y ^ v.y) | (c.last.z ^ v.z);
        int node = (int)s->octdim;
        while (shared < c.path_len && shared < len && (x / node) == 0) { shared++; node /= 2; }
    }
    k.svo_desc_fetches += (uint64_t)(len - shared);
    k.svo_cell_changes += 1;
    c.path_len = len; c.origin = o; c.size = dimension; c.last = v;
}

/* ---- atlas fetch, kernel:652-656 / 684-688 ---------------------------------------------------- */
inline f3 atlas_fetch(const vro_scene *s, float u, float v, int tile_x, int tile_y, bool &clamped) {
    int sx = s->atlas_dim[0] / s->tile_dim[0];                            /* *atlas_dim / *tile_dim */
    int sy = s->atlas_dim[1] / s->tile_dim[1];
    int px = cl_convert_int(u * (float)sx) + cl_convert_int((float)tile_x * (float)sx);
    int py = cl_convert_int(v * (float)sy) + cl_convert_int((float)tile_y * (float)sy);
    /* sampler-less read_imagef with out-of-range coordinates is undefined: pinned clamp-to-edge */
    int cx = std::min(std::max(px, 0), s->atlas_dim[0] - 1);
    int cy = std::min(std::max(py, 0), s->atlas_dim[1] - 1);
    if (cx != px || cy != py) clamped = true;
    const uint8_t *t = s->atlas + 4 * ((size_t)cx + (size_t)s->atlas_dim[0] * (size_t)cy);
    return {(float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f};
}

/* ---- raycaster, kernel:256-724, one pixel ------------------------------------------------------ */
template <bool COUNT>
void cast_pixel(const vro_scene *s, int px, int py, i3 bias, uint8_t *rgba, vro_aux *aux,
                vro_counters &k, bool count_svo) {
    vro_aux a;
    memset(&a, 0, sizeof(a));
    a.hit[0] = a.hit[1] = a.hit[2] = -1;
    bool first_hit_done = false;
    auto finish = [&](uint8_t status, int dist) {
        a.status = status;
        a.steps_total = (uint32_t)dist;
        if (aux) *aux = a;
    };

    const float *rt = s->ray_table + 4 * ((size_t)px + (size_t)s->width * (size_t)py);   /* kernel:277 */
    f3 ray_dir = {rt[0], rt[1], rt[2]};
    const float sp = s->trig[0], cp = s->trig[1], sy = s->trig[2], cy = s->trig[3];
    ray_dir = {ray_dir.z * sp + ray_dir.x * cp, ray_dir.y, ray_dir.z * cp - ray_dir.x * sp};    /* kernel:280 */
    ray_dir = {ray_dir.x * cy - ray_dir.y * sy, ray_dir.x * sy + ray_dir.y * cy, ray_dir.z};    /* kernel:287 */

    const f3 cam = {s->cam_pos[0], s->cam_pos[1], s->cam_pos[2]};
    f3 fl = {floorf(cam.x), floorf(cam.y), floorf(cam.z)};
    if (cam.x == fl.x || cam.y == fl.y || cam.z == fl.z) a.flags |= VRO_FL_FRAC0;

    if (ray_dir.x == 0.0f || ray_dir.y == 0.0f || ray_dir.z == 0.0f) {    /* kernel:293 */
        finish(VRO_ST_SKIP_PRIMARY, 0);
        return;
    }
    if (COUNT) k.primary_rays++;

    i3 voxel_step = {cl_sign_step(ray_dir.x), cl_sign_step(ray_dir.y), cl_sign_step(ray_dir.z)};  /* kernel:298 */
    i3 voxel = {(int)fl.x, (int)fl.y, (int)fl.z};                         /* convert_int3_rtn, kernel:302 */
    f3 delta_t = {fabsf(1.0f / ray_dir.x), fabsf(1.0f / ray_dir.y), fabsf(1.0f / ray_dir.z)};      /* kernel:307 */
    f3 offset = {delta_t.x * (cam.x - fl.x), delta_t.y * (cam.y - fl.y), delta_t.z * (cam.z - fl.z)}; /* :313 */
    f3 t = {offset.x * -(float)voxel_step.x, offset.y * -(float)voxel_step.y, offset.z * -(float)voxel_step.z}; /* :317 */
    /* kernel:323: t += delta_t * -1 * convert_float3(isless(t, 0))   (isless -> -1 for true) */
    t.x += (delta_t.x * -1.0f) * ((t.x < 0.0f) ? -1.0f : 0.0f);
    t.y += (delta_t.y * -1.0f) * ((t.y < 0.0f) ? -1.0f : 0.0f);
    t.z += (delta_t.z * -1.0f) * ((t.z < 0.0f) ? -1.0f : 0.0f);

    int distance_traveled = 0;                                            /* kernel:325-337 */
    int max_distance = s->max_distance;
    unsigned bounce_count = 0;
    const unsigned max_bounces = s->max_bounces > 0 ? (unsigned)s->max_bounces : 2u;   /* kernel:357 hard-codes 2; 0 = that */
    i3 face_mask = {0, 0, 0};
    int voxel_data = 0;
    f3 face_position = {0, 0, 0};
    f4 voxel_color = {0, 0, 0, 0};
    float tfx = 0.0f, tfy = 0.0f;        /* tile_face_position */
    f3 sign = {0, 0, 0};
    f4 color = {0, 0, 0, 0};             /* color_accumulator */
    float fog_distance = 0.0f;
    bool shadow_ray = false;

    /* kernel:342-354: get_oct_vox(camera voxel) biases intersection_t; uniform per frame, so the
     * caller evaluates it once and passes ((sub_oct_pos - voxel) * resolution) / 2 as `bias`. */
    t.x += (float)bias.x; t.y += (float)bias.y; t.z += (float)bias.z;

    /* PARITY-CHAIN LINK "Oracle-B" (SURVEY Appendix E), not part of the reference: the crossing times of an axis in
     * closed form, t(k) = fma(k, delta_t, t0) with k the crossings made since the ray (re)started and t0 the value
     * intersection_t had then.  With s->canonical_t != 0 the walk below USES them instead of the accumulated sums of
     * kernel:559 (what the octree kernel's walk = 2 computes); with canonical_t == 0 (the reference) they are only
     * evaluated next to the real ones to flag VRO_FL_NEAR: some step of this ray would have been taken along other
     * axes under the closed form, i.e. two crossing times are closer than the rounding the additions accumulate --
     * the ray passes within float noise of a voxel edge (BASELINE.json: "degenerate" rays). */
    const bool canon = s->canonical_t != 0;
    f3 t0 = t, tc = t;
    float kx = 0.0f, ky = 0.0f, kz = 0.0f;
    auto restart_canon = [&]() { t0 = t; tc = t; kx = ky = kz = 0.0f; };
    auto blocked_at = [&](const i3 &v) -> bool {      /* set (5 / 6) or outside the map */
        if (v.x < 0 || v.y < 0 || v.z < 0 || v.x >= s->map_dim[0] || v.y >= s->map_dim[1] || v.z >= s->map_dim[2]) return true;
        int d;
        if (s->map) d = (int)s->map[(size_t)v.x + (size_t)s->map_dim[0] * ((size_t)v.y + (size_t)s->map_dim[2] * (size_t)v.z)];
        else { const size_t c = (size_t)v.x + (size_t)s->map_dim[0] * (size_t)v.y; d = (v.z >= s->col_lo[c] && v.z <= s->col_hi[c]) ? 5 : 0; }
        return d == 5 || d == 6;
    };

    CellTrack cell;
    if (COUNT && count_svo) svo_lookup(s, voxel, cell, k);                /* the camera-voxel descent */

    /* The reference reads light 0 only (kernel:660-670).  EXTENSION (`shadow_lights` > 1): the
     * lights are taken one after the other from the same hit point.  Light i adds its view_light term to the
     * colour accumulated so far and casts its own shadow ray, with the reference's own budget (max_distance =
     * steps at the hit + distance to that light); a blocked light leaves 0.1 instead of its term in alpha.
     * With shadow_lights == 1 every statement below is the reference's. */
    const int n_lights = s->shadow_lights < 1 ? 1 : (s->shadow_lights > s->light_count ? s->light_count : s->shadow_lights);
    int light_i = 0;
    f3 L = {s->lights[4], s->lights[5], s->lights[6]};
    f4 Lc = {s->lights[0], s->lights[1], s->lights[2], s->lights[3]};
    i3 hit_voxel = {0, 0, 0}, hit_normal = {0, 0, 0}, hit_empty = {0, 0, 0};
    f3 hit_point = {0, 0, 0};
    int hit_steps = 0;
    float alpha_before = 0.0f;
    const int X = s->map_dim[0], Y = s->map_dim[1], Z = s->map_dim[2];

    /* extension: start the shadow ray of the next light from the stored hit; false = pixel skipped (kernel:671) */
    auto next_light = [&]() -> bool {
        light_i++;
        const float *lp = s->lights + 10 * light_i;
        L = {lp[4], lp[5], lp[6]};
        Lc = {lp[0], lp[1], lp[2], lp[3]};
        alpha_before = color.w;
        color = view_light(color, hit_point - L, Lc, hit_point - cam, hit_normal);
        distance_traveled = hit_steps;                                    /* a skipped pixel reports the counter of the hit's iteration, as at the first light (kernel:671) */
        max_distance = (int)((float)hit_steps + cl_length(to_f3(hit_voxel) - L));
        ray_dir = cl_normalize(L - hit_point);
        if (ray_dir.x == 0.0f || ray_dir.y == 0.0f || ray_dir.z == 0.0f) return false;
        distance_traveled = hit_steps + 1;                                /* as after the first redirect (kernel:714) */
        if (COUNT) k.shadow_rays++;
        voxel = hit_empty;
        voxel_step = {cl_sign_step(ray_dir.x), cl_sign_step(ray_dir.y), cl_sign_step(ray_dir.z)};
        delta_t = {fabsf(1.0f / ray_dir.x), fabsf(1.0f / ray_dir.y), fabsf(1.0f / ray_dir.z)};
        t.x = (delta_t.x * (hit_point.x - floorf(hit_point.x))) * (float)voxel_step.x;
        t.y = (delta_t.y * (hit_point.y - floorf(hit_point.y))) * (float)voxel_step.y;
        t.z = (delta_t.z * (hit_point.z - floorf(hit_point.z))) * (float)voxel_step.z;
        t.x += delta_t.x * ((t.x < 0.0f) ? 1.0f : -0.0f);
        t.y += delta_t.y * ((t.y < 0.0f) ? 1.0f : -0.0f);
        t.z += delta_t.z * ((t.z < 0.0f) ? 1.0f : -0.0f);
        restart_canon();
        if (COUNT && count_svo) { cell = CellTrack(); svo_lookup(s, voxel, cell, k); }
        return true;
    };

    uint8_t status = VRO_ST_MAXDIST;
    for (;;) {
        if (!(distance_traveled < max_distance && bounce_count < max_bounces)) {    /* kernel:357 */
            if (shadow_ray && light_i + 1 < n_lights && bounce_count < max_bounces) {   /* extension: this light is not blocked */
                if (!next_light()) { finish(VRO_ST_SKIP_REDIRECT, distance_traveled); return; }
                continue;
            }
            break;
        }
        /* dense branch, kernel:555-570 */
        face_mask.x = (t.x <= cl_min(t.y, t.z)) ? 1 : 0;                  /* kernel:558 */
        face_mask.y = (t.y <= cl_min(t.z, t.x)) ? 1 : 0;
        face_mask.z = (t.z <= cl_min(t.x, t.y)) ? 1 : 0;
        if (face_mask.x + face_mask.y + face_mask.z > 1) a.flags |= VRO_FL_TIE;
        if (!canon) {
            const int cx = (tc.x <= cl_min(tc.y, tc.z)) ? 1 : 0, cy = (tc.y <= cl_min(tc.z, tc.x)) ? 1 : 0,
                      cz = (tc.z <= cl_min(tc.x, tc.y)) ? 1 : 0;
            if (cx != face_mask.x || cy != face_mask.y || cz != face_mask.z) {
                /* the two evaluation orders step differently here.  That is only visible if one of the two voxels
                 * entered is set or outside the map; between empty voxels the two paths meet again a step later */
                const i3 va = {voxel.x + voxel_step.x * face_mask.x, voxel.y + voxel_step.y * face_mask.y, voxel.z + voxel_step.z * face_mask.z};
                const i3 vb = {voxel.x + voxel_step.x * cx, voxel.y + voxel_step.y * cy, voxel.z + voxel_step.z * cz};
                a.flags |= (blocked_at(va) || blocked_at(vb)) ? VRO_FL_NEAR : VRO_FL_NEAR_AIR;
            }
        }
        t.x += delta_t.x * (float)face_mask.x;                            /* kernel:559 */
        t.y += delta_t.y * (float)face_mask.y;
        t.z += delta_t.z * (float)face_mask.z;
        if (face_mask.x) { kx += 1.0f; tc.x = fmaf(kx, delta_t.x, t0.x); }
        if (face_mask.y) { ky += 1.0f; tc.y = fmaf(ky, delta_t.y, t0.y); }
        if (face_mask.z) { kz += 1.0f; tc.z = fmaf(kz, delta_t.z, t0.z); }
        if (canon) t = tc;
        voxel.x += voxel_step.x * face_mask.x;                            /* kernel:560 */
        voxel.y += voxel_step.y * face_mask.y;
        voxel.z += voxel_step.z * face_mask.z;

        if (voxel.x >= X || voxel.y >= Y || voxel.z >= Z || voxel.x < 0 || voxel.y < 0 || voxel.z < 0) { /* :563 */
            if (shadow_ray && light_i + 1 < n_lights) {                   /* extension: left the map unblocked */
                if (!next_light()) { finish(VRO_ST_SKIP_REDIRECT, distance_traveled); return; }
                continue;
            }
            voxel.x -= voxel_step.x * face_mask.x;
            voxel.y -= voxel_step.y * face_mask.y;
            voxel.z -= voxel_step.z * face_mask.z;
            float m = 1.0f - cl_max((float)distance_traveled / 700.0f, 0.0f);
            color = {0.0f + (voxel_color.x - 0.0f) * m, 0.0f + (voxel_color.y - 0.0f) * m,     /* mix, :565 */
                     0.0f + (voxel_color.z - 0.0f) * m, 0.0f + (voxel_color.w - 0.0f) * m};
            color.w *= 4.0f;                                              /* kernel:566 */
            status = VRO_ST_OOB;
            break;
        }
        if (s->map) {
            voxel_data = (int)s->map[(size_t)voxel.x + (size_t)X * ((size_t)voxel.y + (size_t)Z * (size_t)voxel.z)]; /* :569 */
        } else {                                                          /* column-table map, vr_oracle.h */
            const size_t c = (size_t)voxel.x + (size_t)X * (size_t)voxel.y;
            voxel_data = (voxel.z >= s->col_lo[c] && voxel.z <= s->col_hi[c]) ? 5 : 0;
        }
        if (COUNT) {
            k.dda_steps++;
            if (count_svo) svo_lookup(s, voxel, cell, k);
        }

        if (voxel_data == 5 || voxel_data == 6) {                         /* kernel:575 */
            if (!first_hit_done) {
                first_hit_done = true;
                a.hit[0] = voxel.x; a.hit[1] = voxel.y; a.hit[2] = voxel.z;
                a.face = (uint8_t)(face_mask.x | (face_mask.y << 1) | (face_mask.z << 2) |
                                   ((voxel_step.x < 0) << 3) | ((voxel_step.y < 0) << 4) | ((voxel_step.z < 0) << 5));
                a.hit_type = (uint8_t)voxel_data;
                a.steps_first = (uint32_t)distance_traveled;
            }
            face_position = {0, 0, 0};
            tfx = tfy = 0.0f;
            sign = {1.0f, 1.0f, 1.0f};                                    /* kernel:582 */
            if (face_mask.x == 1) {                                       /* kernel:586 */
                sign.x *= -1.0f;
                float z_percent = (t.z - (t.x - delta_t.x)) / delta_t.z;
                float y_percent = (t.y - (t.x - delta_t.x)) / delta_t.y;
                face_position = {1.00001f, y_percent, z_percent};
                tfx = face_position.y; tfy = face_position.z;
            } else if (face_mask.y == 1) {                                /* kernel:601 */
                sign.y *= -1.0f;
                float x_percent = (t.x - (t.y - delta_t.y)) / delta_t.x;
                float z_percent = (t.z - (t.y - delta_t.y)) / delta_t.z;
                face_position = {x_percent, 1.00001f, z_percent};
                tfx = face_position.x; tfy = face_position.z;
            } else if (face_mask.z == 1) {                                /* kernel:610 */
                sign.z *= -1.0f;
                float x_percent = (t.x - (t.z - delta_t.z)) / delta_t.x;
                float y_percent = (t.y - (t.z - delta_t.z)) / delta_t.y;
                face_position = {x_percent, y_percent, 1.00001f};
                tfx = face_position.x; tfy = face_position.y;
            }
            /* quadrant fix-ups, kernel:626-643 */
            if (ray_dir.x > 0.0f) face_position.x = -face_position.x + 1.0f;
            if (ray_dir.x < 0.0f) tfx = -tfx + 1.0f;
            if (ray_dir.y > 0.0f) {
                face_position.y = -face_position.y + 1.0f;
            } else {
                tfx = 1.0f - tfx;                                         /* kernel:632 (1.0 literal) */
                if (face_mask.z == 1) {
                    tfx = 1.0f - tfx;
                    tfy = 1.0f - tfy;
                }
            }
            if (ray_dir.z > 0.0f) face_position.z = -face_position.z + 1.0f;
            if (ray_dir.z < 0.0f) tfy = -tfy + 1.0f;

            if (voxel_data == 5 && !shadow_ray) {                         /* kernel:649 */
                shadow_ray = true;
                a.flags |= VRO_FL_LIT;
                bool clamped = false;
                f3 tex = atlas_fetch(s, tfx, tfy, 5, 0, clamped);         /* kernel:652-656 */
                if (clamped) a.flags |= VRO_FL_ATLAS_CLAMP;
                if (COUNT) k.texel_fetches++;
                voxel_color.x += tex.x / 2.0f;
                voxel_color.y += tex.y / 2.0f;
                voxel_color.z += tex.z / 2.0f;

                f3 hit_pos = to_f3(voxel) + face_position;
                hit_point = hit_pos;                                      /* extension state (unused with one light) */
                hit_voxel = voxel;
                hit_steps = distance_traveled;
                hit_normal = {face_mask.x * voxel_step.x, face_mask.y * voxel_step.y, face_mask.z * voxel_step.z};
                hit_empty = {voxel.x - voxel_step.x * face_mask.x, voxel.y - voxel_step.y * face_mask.y,
                             voxel.z - voxel_step.z * face_mask.z};
                color = view_light(voxel_color, hit_pos - L, Lc, hit_pos - cam,      /* kernel:658 */
                                   {face_mask.x * voxel_step.x, face_mask.y * voxel_step.y, face_mask.z * voxel_step.z});
                fog_distance = (float)distance_traveled;                  /* kernel:666 */
                max_distance = (int)((float)distance_traveled + cl_length(to_f3(voxel) - L));   /* kernel:667 */

                ray_dir = cl_normalize(L - hit_pos);                      /* kernel:670 */
                if (ray_dir.x == 0.0f || ray_dir.y == 0.0f || ray_dir.z == 0.0f) {
                    finish(VRO_ST_SKIP_REDIRECT, distance_traveled);
                    return;
                }
                if (COUNT) k.shadow_rays++;
                voxel.x -= voxel_step.x * face_mask.x;                    /* kernel:674 */
                voxel.y -= voxel_step.y * face_mask.y;
                voxel.z -= voxel_step.z * face_mask.z;
                voxel_step = {cl_sign_step(ray_dir.x), cl_sign_step(ray_dir.y), cl_sign_step(ray_dir.z)};
                delta_t = {fabsf(1.0f / ray_dir.x), fabsf(1.0f / ray_dir.y), fabsf(1.0f / ray_dir.z)};
                t.x = (delta_t.x * (hit_pos.x - floorf(hit_pos.x))) * (float)voxel_step.x;   /* kernel:678 */
                t.y = (delta_t.y * (hit_pos.y - floorf(hit_pos.y))) * (float)voxel_step.y;
                t.z = (delta_t.z * (hit_pos.z - floorf(hit_pos.z))) * (float)voxel_step.z;
                t.x += delta_t.x * ((t.x < 0.0f) ? 1.0f : -0.0f);         /* kernel:679 */
                t.y += delta_t.y * ((t.y < 0.0f) ? 1.0f : -0.0f);
                t.z += delta_t.z * ((t.z < 0.0f) ? 1.0f : -0.0f);
                restart_canon();
            } else if (voxel_data == 6 && !shadow_ray) {                  /* kernel:682 */
                a.flags |= VRO_FL_REFLECTED;
                bool clamped = false;
                f3 tex = atlas_fetch(s, tfx, tfy, 3, 4, clamped);         /* kernel:684-688 */
                if (clamped) a.flags |= VRO_FL_ATLAS_CLAMP;
                if (COUNT) k.texel_fetches++;
                voxel_color.x += tex.x / 4.0f;
                voxel_color.y += tex.y / 4.0f;
                voxel_color.z += tex.z / 4.0f;

                f3 hit_pos = to_f3(voxel) + face_position;
                ray_dir = {ray_dir.x * sign.x, ray_dir.y * sign.y, ray_dir.z * sign.z};   /* kernel:693 */
                if (ray_dir.x == 0.0f || ray_dir.y == 0.0f || ray_dir.z == 0.0f) {
                    finish(VRO_ST_SKIP_REDIRECT, distance_traveled);
                    return;
                }
                if (COUNT) k.reflect_rays++;
                voxel.x -= voxel_step.x * face_mask.x;                    /* kernel:697 */
                voxel.y -= voxel_step.y * face_mask.y;
                voxel.z -= voxel_step.z * face_mask.z;
                /* kernel:698: (-1,-1,-1) * (ray_dir > 0) - (ray_dir < 0)  ==> +1 for any non-zero d */
                voxel_step = {(-1 * ((ray_dir.x > 0.0f) ? -1 : 0)) - ((ray_dir.x < 0.0f) ? -1 : 0),
                              (-1 * ((ray_dir.y > 0.0f) ? -1 : 0)) - ((ray_dir.y < 0.0f) ? -1 : 0),
                              (-1 * ((ray_dir.z > 0.0f) ? -1 : 0)) - ((ray_dir.z < 0.0f) ? -1 : 0)};
                delta_t = {fabsf(1.0f / ray_dir.x), fabsf(1.0f / ray_dir.y), fabsf(1.0f / ray_dir.z)};
                t.x = (delta_t.x * (hit_pos.x - floorf(hit_pos.x))) * (float)voxel_step.x;   /* kernel:701 */
                t.y = (delta_t.y * (hit_pos.y - floorf(hit_pos.y))) * (float)voxel_step.y;
                t.z = (delta_t.z * (hit_pos.z - floorf(hit_pos.z))) * (float)voxel_step.z;
                t.x += delta_t.x * ((t.x < 0.0f) ? 1.0f : -0.0f);         /* kernel:702 */
                t.y += delta_t.y * ((t.y < 0.0f) ? 1.0f : -0.0f);
                t.z += delta_t.z * ((t.z < 0.0f) ? 1.0f : -0.0f);
                restart_canon();
                bounce_count += 1;                                        /* kernel:704 */
            } else {                                                      /* kernel:707 */
                color.w = alpha_before + 0.1f;                            /* one light: 0 + 0.1 = kernel:708 */
                if (light_i + 1 < n_lights) {                             /* extension */
                    if (!next_light()) { finish(VRO_ST_SKIP_REDIRECT, distance_traveled); return; }
                    continue;
                }
                status = VRO_ST_SHADOW_HIT;
                break;
            }
        }
        distance_traveled++;                                              /* kernel:714 */
    }
    if (status == VRO_ST_MAXDIST && bounce_count >= max_bounces) status = VRO_ST_BOUNCES;

    float m = 1.0f - cl_max(fog_distance / 700.0f, 0.0f);                 /* kernel:716 */
    color = {0.0f + (color.x - 0.0f) * m, 0.0f + (color.y - 0.0f) * m, 0.0f + (color.z - 0.0f) * m,
             0.0f + (color.w - 0.0f) * m};
    uint8_t *o = rgba + 4 * ((size_t)px + (size_t)s->width * (size_t)py); /* kernel:717 */
    o[0] = unorm8(color.x); o[1] = unorm8(color.y); o[2] = unorm8(color.z); o[3] = unorm8(color.w);
    if (COUNT) { k.pixels_written++; if (a.flags & VRO_FL_TIE) k.tie_pixels++; }
    finish(status, distance_traveled);
}

}  // namespace

extern "C" {

int vro_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* host:233-299 */
void vro_make_ray_table(int width, int height, float *out) {
    memset(out, 0, sizeof(float) * 4 * (size_t)width * (size_t)height);
    const double s157 = sin(1.57), c157 = cos(1.57);
    for (int y = -height / 2; y < height / 2; y++) {
        for (int x = -width / 2; x < width / 2; x++) {
            float rx = -800.0f, ry = (float)x, rz = (float)y;             /* host:249 */
            float nx = (float)((double)rz * s157 + (double)rx * c157);    /* host:252-256 */
            float ny = ry;
            float nz = (float)((double)rz * c157 - (double)rx * s157);
            float mult = sqrtf(nx * nx + ny * ny + nz * nz);              /* Normalize, util.hpp:64 */
            size_t index = (size_t)(x + width / 2) + (size_t)width * (size_t)(y + height / 2);   /* host:264 */
            out[4 * index + 0] = nx / mult;
            out[4 * index + 1] = ny / mult;
            out[4 * index + 2] = nz / mult;
            out[4 * index + 3] = 0.0f;
        }
    }
}

void vro_get_oct_vox(const uint64_t *desc, int64_t root_index, int64_t octdim, const int32_t pos[3],
                     int32_t *found, int32_t sub_oct_pos[3], int32_t *resolution, int32_t *scale) {
    TraversalState ts = get_oct_vox({pos[0], pos[1], pos[2]}, desc, root_index, octdim);
    *found = ts.found;
    sub_oct_pos[0] = ts.sub_oct_pos.x; sub_oct_pos[1] = ts.sub_oct_pos.y; sub_oct_pos[2] = ts.sub_oct_pos.z;
    *resolution = ts.resolution;
    *scale = ts.scale;
}

int vro_raycast(const vro_scene *s, int y0, int y1, int row_stride, uint8_t *rgba, vro_aux *aux,
                vro_counters *counters, int count_svo, int num_threads) {
    if (!s || !s->ray_table || (!s->map && !(s->col_lo && s->col_hi)) || !s->lights || !s->atlas || !rgba) return -1;
    if (count_svo && !s->oct_desc && !(s->tree64 && s->tree64_levels >= 1 && (1ll << (2 * s->tree64_levels)) == s->octdim)) return -2;
    y0 = std::max(y0, 0);
    y1 = std::min(y1, (int)s->height);
    if (row_stride < 1) row_stride = 1;

    /* kernel:342-354, hoisted: uniform per frame */
    i3 bias = {0, 0, 0};
    if (s->oct_desc) {
        i3 v = {(int)floorf(s->cam_pos[0]), (int)floorf(s->cam_pos[1]), (int)floorf(s->cam_pos[2])};
        TraversalState ts = get_oct_vox(v, s->oct_desc, s->oct_root_index, s->octdim);
        bias.x = ((ts.sub_oct_pos.x - v.x) * ts.resolution) / 2;          /* kernel:354, int3 arithmetic */
        bias.y = ((ts.sub_oct_pos.y - v.y) * ts.resolution) / 2;
        bias.z = ((ts.sub_oct_pos.z - v.z) * ts.resolution) / 2;
    }

    vro_counters total;
    memset(&total, 0, sizeof(total));
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_num_procs();
#else
    num_threads = 1;
#endif
    const bool counting = counters != nullptr;
#pragma omp parallel num_threads(num_threads)
    {
        vro_counters k;
        memset(&k, 0, sizeof(k));
#pragma omp for schedule(dynamic, 1)
        for (int yi = 0; yi < (y1 - y0 + row_stride - 1) / row_stride; yi++) {
            const int y = y0 + yi * row_stride;
            for (int x = 0; x < s->width; x++) {
                vro_aux *a = aux ? aux + ((size_t)x + (size_t)s->width * (size_t)y) : nullptr;
                if (counting) cast_pixel<true>(s, x, y, bias, rgba, a, k, count_svo != 0);
                else cast_pixel<false>(s, x, y, bias, rgba, a, k, false);
            }
        }
#pragma omp critical
        {
            total.pixels_written += k.pixels_written; total.primary_rays += k.primary_rays;
            total.shadow_rays += k.shadow_rays; total.reflect_rays += k.reflect_rays;
            total.dda_steps += k.dda_steps; total.texel_fetches += k.texel_fetches;
            total.svo_desc_fetches += k.svo_desc_fetches; total.svo_cell_changes += k.svo_cell_changes;
            total.tie_pixels += k.tie_pixels;
        }
    }
    total.pixels = (uint64_t)((y1 - y0 + row_stride - 1) / row_stride) * (uint64_t)s->width;
    if (counters) *counters = total;
    return 0;
}

/* ---- Octree::Generate / GenerationRecursion, src/map/Octree.cpp:13-43,171-323 ------------------ */
namespace {
struct Gen {
    const int8_t *data;
    int dim;
    uint64_t *buf;
    uint64_t size;
    int64_t pos;                  /* descriptor_buffer_position */
    int page_header_counter;      /* include/map/Octree.h:55 */
    bool overflow;

    inline bool is_leaf(uint64_t d) const {                               /* util.hpp:219 */
        const uint64_t vm = 0xFF0000, lm = 0xFF000000;
        if (((d & vm) == vm) || ((d & vm) == 0)) return (d & lm) == lm;
        return false;
    }
    inline void put(int64_t at, uint64_t v) {
        if (at < 0 || (uint64_t)at >= size) { overflow = true; return; }
        buf[at] = v;
    }
    /* returns (descriptor, absolute position of its first child) */
    std::pair<uint64_t, uint64_t> rec(int px, int py, int pz, int scale) {
        const int vx[8] = {px, px + scale, px, px + scale, px, px + scale, px, px + scale};   /* :176-185 */
        const int vy[8] = {py, py, py + scale, py + scale, py, py, py + scale, py + scale};
        const int vz[8] = {pz, pz, pz, pz, pz + scale, pz + scale, pz + scale, pz + scale};
        uint64_t d = 0;
        if (scale == 1) {                                                 /* :195-209 */
            for (int i = 0; i < 8; i++)
                if (data[(size_t)vx[i] + (size_t)dim * ((size_t)vy[i] + (size_t)dim * (size_t)vz[i])])
                    d |= (uint64_t)1 << (i + 16);
            d |= 0xFF000000;
            return {d, 0};
        }
        std::pair<uint64_t, uint64_t> kids[8];
        int nk = 0;
        for (int i = 0; i < 8; i++) {                                     /* :218-241 */
            auto c = rec(vx[i], vy[i], vz[i], scale / 2);
            if (is_leaf(c.first) && (c.first & 0xFF0000) == 0) d |= (uint64_t)1 << (i + 16 + 8);
            else { d |= (uint64_t)1 << (i + 16); kids[nk++] = c; }
        }
        int worst = nk * 2;                                               /* :247 */
        if (page_header_counter - worst <= 0) {                           /* :251-262 */
            pos -= page_header_counter;
            page_header_counter = 0x8000;
            put(pos, ~(uint64_t)0);
            pos--;
        }
        int64_t far_block = pos;                                          /* :266 */
        for (int i = nk - 1; i >= 0; i--) {                               /* :270-286 */
            int64_t rel = (int64_t)kids[i].second - (pos - worst);
            if (rel > 0x8000) { put(pos, kids[i].second); pos--; page_header_counter--; }
        }
        for (int i = nk - 1; i >= 0; i--) {                               /* :289-315 */
            int64_t rel = (int64_t)kids[i].second - pos;
            uint64_t desc = kids[i].first;
            if (rel > 0x8000) {
                desc |= 0x8000;
                desc |= (uint64_t)(far_block - pos);
                far_block--;
            } else if (rel > 0) {
                desc |= (uint64_t)rel;
            }
            put(pos, desc);
            pos--;
            page_header_counter--;
        }
        return {d, (uint64_t)(pos + 1)};                                  /* :319 */
    }
};
}  // namespace

int64_t vro_octree_generate(const int8_t *data, int dim, uint64_t *buffer, uint64_t buffer_size, uint64_t *used) {
    memset(buffer, 0, sizeof(uint64_t) * buffer_size);
    Gen g{data, dim, buffer, buffer_size, (int64_t)buffer_size - 1, 0x8000, false};
    auto root = g.rec(0, 0, 0, dim / 2);                                  /* :19 */
    root.first |= 1;                                                      /* :27 */
    g.put(g.pos, root.first);
    int64_t root_index = g.pos;
    g.pos--;
    if (used) *used = (uint64_t)((int64_t)buffer_size - 1 - g.pos);
    return g.overflow ? -1 : root_index;
}

}  // extern "C"
