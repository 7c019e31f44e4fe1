/*
 * vr_trace.h -- per-pixel ray casting core of the B200 caster (device code; also compilable for the
 * host so tests/host_emu can single-step the same logic without a GPU -- the product never runs it
 * on the CPU).
 *
 * What it computes is the reference kernel `raycaster` (kernels/ray_caster_kernel.cl:256-724):
 * ray fetch + camera rotation, Amanatides-Woo DDA, hit UV, atlas fetch, Blinn-Phong `view_light`,
 * shadow / reflection redirect inside the same loop, fog, RGBA8 store.  How it does it is new:
 *   - all frame-uniform work (sin/cos, the get_oct_vox camera-cell bias) is hoisted to the host;
 *   - the dense variant walks the char map exactly like kernel:555-570;
 *   - the SVO variant replaces the per-step `map[]` gather by a lookup in a 16-byte-node 64-tree
 *     whose empty cells are cached, so that steps inside a known-empty cell touch no memory, while
 *     the float state (intersection_t) is advanced by exactly the same additions as the reference
 *     -- every pixel, hit voxel, face and step count is bit-identical to the dense walk.
 *
 * Float discipline: every arithmetic op that feeds a compare, an index or the colour goes through
 * VR_ADD/VR_MUL/... (round-to-nearest intrinsics on the device, never FMA-contracted) so the result
 * does not depend on -fmad; division and sqrt are IEEE (__fdiv_rn/__fsqrt_rn).
 */
#ifndef VR_TRACE_H
#define VR_TRACE_H

#include "vr_types.h"
#include <math.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define VR_ADD(a, b) __fadd_rn((a), (b))
#define VR_SUB(a, b) __fsub_rn((a), (b))
#define VR_MUL(a, b) __fmul_rn((a), (b))
#define VR_DIV(a, b) __fdiv_rn((a), (b))
#define VR_SQRT(a) __fsqrt_rn((a))
#define VR_POPC64(x) __popcll((x))
#else
#define VR_ADD(a, b) ((a) + (b))
#define VR_SUB(a, b) ((a) - (b))
#define VR_MUL(a, b) ((a) * (b))
#define VR_DIV(a, b) ((a) / (b))
#define VR_SQRT(a) sqrtf((a))
#define VR_POPC64(x) __builtin_popcountll((x))
#endif

/* profiling hooks of the host emulation (tests/host_emu/prof.cpp); nothing on the device */
#if defined(VR_PROFILE) && !defined(__CUDA_ARCH__)
struct vr_prof_rec { int path, n, chain_a, chain_b, chain_c, fix, pops, loads, lookup, hit, replay, adds, jumps; };
extern thread_local vr_prof_rec vr_prof;
#define VR_PROF(field, v) (vr_prof.field = (v))
#define VR_PROF_ADD(field, v) (vr_prof.field += (v))
#else
#define VR_PROF(field, v) ((void)0)
#define VR_PROF_ADD(field, v) ((void)0)
#endif

struct vf3 { float x, y, z; };
struct vf4 { float x, y, z, w; };
struct vi3 { int x, y, z; };

VR_HD vf3 vr_add3(vf3 a, vf3 b) { return {VR_ADD(a.x, b.x), VR_ADD(a.y, b.y), VR_ADD(a.z, b.z)}; }
VR_HD vf3 vr_sub3(vf3 a, vf3 b) { return {VR_SUB(a.x, b.x), VR_SUB(a.y, b.y), VR_SUB(a.z, b.z)}; }
VR_HD vf3 vr_i2f3(vi3 a) { return {(float)a.x, (float)a.y, (float)a.z}; }
VR_HD float vr_dot(vf3 a, vf3 b) { return VR_ADD(VR_ADD(VR_MUL(a.x, b.x), VR_MUL(a.y, b.y)), VR_MUL(a.z, b.z)); }
VR_HD float vr_length(vf3 a) { return VR_SQRT(vr_dot(a, a)); }
VR_HD vf3 vr_normalize(vf3 a) { float l = vr_length(a); return {VR_DIV(a.x, l), VR_DIV(a.y, l), VR_DIV(a.z, l)}; }
/* x / c for a compile-time constant c, with r = RN(1 / c): q0 = x r, q = fma(fma(-c, q0, x), r, q0).  The correctly
 * rounded quotient -- the reference's IEEE division, bit for bit -- for the operands it is used on: texel / 255 with
 * texel = 0..255 and steps / 700 with steps = 0..2^24 (tests/test_emu_canonical.py checks both ranges exhaustively);
 * three FMA-pipe instructions instead of the ten of a general division.  -DVR_HIT_LITERAL restores the divisions. */
#if defined(__CUDA_ARCH__)
#define VR_FMA_RN(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define VR_FMA_RN(a, b, c) fmaf((a), (b), (c))
#endif
VR_HD float vr_div_const(float x, float c, float r) {
#if defined(VR_HIT_LITERAL)
    (void)r;
    return VR_DIV(x, c);
#else
    const float q0 = VR_MUL(x, r);
    return VR_FMA_RN(VR_FMA_RN(-c, q0, x), r, q0);
#endif
}
#define VR_DIV_255(x) vr_div_const((x), 255.0f, 1.0f / 255.0f)
#define VR_DIV_700(x) vr_div_const((x), 700.0f, 1.0f / 700.0f)
/* x / 2 and x / 4 are exact scalings */
#if defined(VR_HIT_LITERAL)
#define VR_HALF(x) VR_DIV((x), 2.0f)
#define VR_QUARTER(x) VR_DIV((x), 4.0f)
#else
#define VR_HALF(x) VR_MUL((x), 0.5f)
#define VR_QUARTER(x) VR_MUL((x), 0.25f)
#endif
VR_HD float vr_max(float x, float y) { return (x < y) ? y : x; }      /* OpenCL max(), NaN -> x */
VR_HD float vr_min(float x, float y) { return (y < x) ? y : x; }
VR_HD int vr_sign(float d) { return (d > 0.0f) - (d < 0.0f); }
VR_HD bool vr_any_zero(vf3 d) { return d.x == 0.0f || d.y == 0.0f || d.z == 0.0f; }
VR_HD int vr_f2i_rz(float v) {                                        /* convert_int, pinned */
    if (!(v > -2147483648.0f && v < 2147483648.0f)) return 0;
    return (int)v;
}
VR_HD uint32_t vr_unorm8(float c) {                                   /* write_imagef, kernel:717 */
    float v = VR_MUL(c, 255.0f);
    if (!(v > 0.0f)) return 0u;
    if (v > 255.0f) v = 255.0f;
#if defined(__CUDA_ARCH__)
    return (uint32_t)__float2int_rn(v);
#else
    return (uint32_t)nearbyintf(v);
#endif
}

/* Everything the reference keeps in private variables across loop iterations (kernel:298-337). */
struct RayState {
    vf3 ray_dir;
    vi3 step;            /* voxel_step */
    vi3 voxel;
    vf3 delta;           /* delta_t */
    vf3 t;               /* intersection_t */
    int dist;            /* distance_traveled */
    int max_distance;
    int bounce;          /* bounce_count */
    int fm;              /* face_mask, bit0 x, bit1 y, bit2 z */
    vf3 voxel_color;     /* .w stays 0 (kernel:690 subtracts 0) */
    vf4 color;           /* color_accumulator */
    float fog_distance;
    bool shadow;         /* shadow_ray */
    /* multi-light extension only (dead in the single-light kernels): the hit the shadow rays start from */
    int light;           /* light whose shadow ray is being traced */
    vf3 hit_point;
    vi3 hit_voxel, hit_empty, hit_normal;
    float alpha_before;  /* colour alpha before the current light was added */
};

/* kernel:276-337 + 353.  Returns false when the pixel is skipped (kernel:293). */
VR_HD bool vr_ray_setup(const vr_frame_params &P, int x, int y, RayState &r) {
#if defined(__CUDA_ARCH__)
    /* read once per frame: streaming (evict-first), so that the 133 MB table does not push the octree / top grid out of L2 */
    const float4 rt = __ldcs(reinterpret_cast<const float4 *>(P.ray_table) + ((size_t)x + (size_t)P.width * (size_t)y));
    vf3 d = {rt.x, rt.y, rt.z};
#else
    const float *rt = P.ray_table + 4 * ((size_t)x + (size_t)P.width * (size_t)y);
    vf3 d = {rt[0], rt[1], rt[2]};
#endif
    const float sp = P.trig[0], cp = P.trig[1], sy = P.trig[2], cy = P.trig[3];
    d = {VR_ADD(VR_MUL(d.z, sp), VR_MUL(d.x, cp)), d.y, VR_SUB(VR_MUL(d.z, cp), VR_MUL(d.x, sp))};   /* pitch */
    d = {VR_SUB(VR_MUL(d.x, cy), VR_MUL(d.y, sy)), VR_ADD(VR_MUL(d.x, sy), VR_MUL(d.y, cy)), d.z};   /* yaw */
    r.ray_dir = d;
    if (vr_any_zero(d)) return false;

    r.step = {vr_sign(d.x), vr_sign(d.y), vr_sign(d.z)};
    const vf3 cam = {P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]};
    const vf3 fl = {floorf(cam.x), floorf(cam.y), floorf(cam.z)};
    r.voxel = {(int)fl.x, (int)fl.y, (int)fl.z};
    r.delta = {fabsf(VR_DIV(1.0f, d.x)), fabsf(VR_DIV(1.0f, d.y)), fabsf(VR_DIV(1.0f, d.z))};
    vf3 off = {VR_MUL(r.delta.x, VR_SUB(cam.x, fl.x)), VR_MUL(r.delta.y, VR_SUB(cam.y, fl.y)),
               VR_MUL(r.delta.z, VR_SUB(cam.z, fl.z))};
    /* t = off * -step; where t < 0: t += delta (kernel:317-323) */
    vf3 t = {VR_MUL(off.x, -(float)r.step.x), VR_MUL(off.y, -(float)r.step.y), VR_MUL(off.z, -(float)r.step.z)};
    if (t.x < 0.0f) t.x = VR_ADD(t.x, r.delta.x);
    if (t.y < 0.0f) t.y = VR_ADD(t.y, r.delta.y);
    if (t.z < 0.0f) t.z = VR_ADD(t.z, r.delta.z);
    /* get_oct_vox start bias (kernel:353-354), frame-uniform, host evaluated */
    r.t = {VR_ADD(t.x, P.bias[0]), VR_ADD(t.y, P.bias[1]), VR_ADD(t.z, P.bias[2])};

    r.dist = 0;
    r.max_distance = P.max_distance;
    r.bounce = 0;
    r.fm = 0;
    r.voxel_color = {0.0f, 0.0f, 0.0f};
    r.color = {0.0f, 0.0f, 0.0f, 0.0f};
    r.fog_distance = 0.0f;
    r.shadow = false;
    r.light = 0;
    r.alpha_before = 0.0f;
    return true;
}

/* kernel:558-560: one DDA step; ties step every tied axis. */
VR_HD void vr_dda_step(RayState &r) {
    const int mx = (r.t.x <= vr_min(r.t.y, r.t.z)) ? 1 : 0;
    const int my = (r.t.y <= vr_min(r.t.z, r.t.x)) ? 1 : 0;
    const int mz = (r.t.z <= vr_min(r.t.x, r.t.y)) ? 1 : 0;
    r.fm = mx | (my << 1) | (mz << 2);
    r.t.x = VR_ADD(r.t.x, VR_MUL(r.delta.x, (float)mx));
    r.t.y = VR_ADD(r.t.y, VR_MUL(r.delta.y, (float)my));
    r.t.z = VR_ADD(r.t.z, VR_MUL(r.delta.z, (float)mz));
    r.voxel.x += r.step.x * mx;
    r.voxel.y += r.step.y * my;
    r.voxel.z += r.step.z * mz;
}

/* kernel:563-568: the ray left the map. */
VR_HD void vr_out_of_bounds(RayState &r) {
    r.voxel.x -= r.step.x * (r.fm & 1);
    r.voxel.y -= r.step.y * ((r.fm >> 1) & 1);
    r.voxel.z -= r.step.z * ((r.fm >> 2) & 1);
    const float m = VR_SUB(1.0f, vr_max(VR_DIV_700((float)r.dist), 0.0f));
    r.color = {VR_MUL(r.voxel_color.x, m), VR_MUL(r.voxel_color.y, m), VR_MUL(r.voxel_color.z, m), VR_MUL(0.0f, m)};
    r.color.w = VR_MUL(r.color.w, 4.0f);
}

/* kernel:652-656 / 684-688 */
VR_HD vf3 vr_atlas_fetch(const vr_frame_params &P, float u, float v, int tile_x, int tile_y, bool &clamped) {
    const int px = vr_f2i_rz(VR_MUL(u, (float)P.atlas_scale[0])) + vr_f2i_rz(VR_MUL((float)tile_x, (float)P.atlas_scale[0]));
    const int py = vr_f2i_rz(VR_MUL(v, (float)P.atlas_scale[1])) + vr_f2i_rz(VR_MUL((float)tile_y, (float)P.atlas_scale[1]));
    const int cx = px < 0 ? 0 : (px > P.atlas_dim[0] - 1 ? P.atlas_dim[0] - 1 : px);
    const int cy = py < 0 ? 0 : (py > P.atlas_dim[1] - 1 ? P.atlas_dim[1] - 1 : py);
    clamped = (cx != px) || (cy != py);
#if defined(__CUDA_ARCH__)
    const uchar4 c = tex2D<uchar4>((cudaTextureObject_t)P.atlas_tex, (float)cx + 0.5f, (float)cy + 0.5f);
    return {VR_DIV_255((float)c.x), VR_DIV_255((float)c.y), VR_DIV_255((float)c.z)};
#else
    const uint8_t *c = P.atlas + 4 * ((size_t)cx + (size_t)P.atlas_dim[0] * (size_t)cy);
    return {VR_DIV_255((float)c[0]), VR_DIV_255((float)c[1]), VR_DIV_255((float)c[2])};
#endif
}

/* kernel:78-99.  `nlight_out` (optional) receives normalize(light) when it was computed (*nlight_ok): the shadow ray's
 * direction normalize(-light) (kernel:669) is its exact negation -- a - b = -(b - a), (-x)(-x) = x x and (-x) / l = -(x / l)
 * hold bit for bit in IEEE arithmetic -- so the caller needs no second square root and three more divisions.
 * The face normal `mask` is a signed unit axis vector on every ray but the exact-tie ones: normalize() of it is then the
 * vector itself (1 / sqrt(1) = 1, 0 / 1 = +0). */
VR_HD vf4 vr_view_light(vf3 in_color, float in_w, vf3 light, const float *rgbi, vf3 view, vi3 mask, vf3 *nlight_out = nullptr,
                        bool *nlight_ok = nullptr) {
    if (nlight_ok) *nlight_ok = false;
    if (light.x == 0.0f && light.y == 0.0f && light.z == 0.0f) return {0.0f, 0.0f, 0.0f, 0.0f};
    const float ll = vr_length(light);
    float d = VR_MUL(ll, 0.01f);
    d = VR_MUL(d, d);
#if defined(VR_HIT_LITERAL)
    const vf3 nmask = vr_normalize(vr_i2f3(mask));
#else
    vf3 nmask = vr_i2f3(mask);
    if (mask.x * mask.x + mask.y * mask.y + mask.z * mask.z != 1) nmask = vr_normalize(nmask);
#endif
    const vf3 nlight = {VR_DIV(light.x, ll), VR_DIV(light.y, ll), VR_DIV(light.z, ll)};
    if (nlight_out) { *nlight_out = nlight; *nlight_ok = true; }
    const float diffuse = vr_max(vr_dot(nmask, nlight), 0.1f);
    float specular = 0.0f;
    if (diffuse > 0.0f) {
        const vf3 halfway = vr_normalize(vr_add3(nlight, vr_normalize(view)));
        specular = vr_max(vr_dot(nmask, halfway), 0.0f);               /* pow(x, 1) */
    }
    vf4 o;
    o.x = VR_ADD(in_color.x, VR_ADD(VR_MUL(diffuse, rgbi[0]), VR_DIV(VR_MUL(specular, rgbi[0]), d)));
    o.y = VR_ADD(in_color.y, VR_ADD(VR_MUL(diffuse, rgbi[1]), VR_DIV(VR_MUL(specular, rgbi[1]), d)));
    o.z = VR_ADD(in_color.z, VR_ADD(VR_MUL(diffuse, rgbi[2]), VR_DIV(VR_MUL(specular, rgbi[2]), d)));
    o.w = VR_ADD(in_w, VR_ADD(VR_MUL(diffuse, rgbi[3]), VR_DIV(VR_MUL(specular, rgbi[3]), d)));
    return o;
}

/* normalize(L - hit_pos) (kernel:669) from view_light's normalize(hit_pos - L) */
VR_HD vf3 vr_shadow_dir(vf3 L, vf3 hit_pos, vf3 nlight, bool nlight_ok) {
#if !defined(VR_HIT_LITERAL)
    if (nlight_ok) return {-nlight.x, -nlight.y, -nlight.z};
#endif
    return vr_normalize(vr_sub3(L, hit_pos));
}

/* kernel:674-679 / 697-702: restart the DDA from hit_pos along r.ray_dir with the given step. */
VR_HD void vr_restart_dda(RayState &r, vf3 hit_pos, vi3 new_step) {
    r.voxel.x -= r.step.x * (r.fm & 1);
    r.voxel.y -= r.step.y * ((r.fm >> 1) & 1);
    r.voxel.z -= r.step.z * ((r.fm >> 2) & 1);
    r.step = new_step;
    r.delta = {fabsf(VR_DIV(1.0f, r.ray_dir.x)), fabsf(VR_DIV(1.0f, r.ray_dir.y)), fabsf(VR_DIV(1.0f, r.ray_dir.z))};
    vf3 t = {VR_MUL(VR_MUL(r.delta.x, VR_SUB(hit_pos.x, floorf(hit_pos.x))), (float)r.step.x),
             VR_MUL(VR_MUL(r.delta.y, VR_SUB(hit_pos.y, floorf(hit_pos.y))), (float)r.step.y),
             VR_MUL(VR_MUL(r.delta.z, VR_SUB(hit_pos.z, floorf(hit_pos.z))), (float)r.step.z)};
    if (t.x < 0.0f) t.x = VR_ADD(t.x, r.delta.x);
    if (t.y < 0.0f) t.y = VR_ADD(t.y, r.delta.y);
    if (t.z < 0.0f) t.z = VR_ADD(t.z, r.delta.z);
    r.t = t;
}

/* kernel:575-711, entered when the voxel just stepped into holds 5 or 6.
 * Returns -1 to continue the loop (ray was redirected), or the terminal VR_ST_* code. */
/* EXTENSION beyond the reference (setting LIGHT_COUNT > 1): the lights are taken one after the other from
 * the same hit point.  Light i adds its view_light term to the colour accumulated so far and casts its own shadow
 * ray with the reference's budget (max_distance = steps at the hit + distance to that light); a blocked light leaves
 * 0.1 instead of its term in alpha.  Restated in the oracle (vr_oracle.cpp: next_light).  Returns false when the pixel
 * is skipped because the direction to the light has a zero component (kernel:671). */
VR_HD bool vr_more_lights(const vr_frame_params &P, const RayState &r) { return r.shadow && r.light + 1 < P.light_count; }

VR_HD bool vr_next_light(const vr_frame_params &P, RayState &r) {
    r.light++;
    const vf3 L = {P.light_pos[r.light][0], P.light_pos[r.light][1], P.light_pos[r.light][2]};
    const vf3 cam = {P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]};
    r.alpha_before = r.color.w;
    vf3 nlight = {0.0f, 0.0f, 0.0f};
    bool nlight_ok;
    r.color = vr_view_light({r.color.x, r.color.y, r.color.z}, r.color.w, vr_sub3(r.hit_point, L), P.light_rgbi[r.light],
                            vr_sub3(r.hit_point, cam), r.hit_normal, &nlight, &nlight_ok);
    const int hit_steps = (int)r.fog_distance;
    r.dist = hit_steps;
    r.max_distance = (int)VR_ADD((float)hit_steps, vr_length(vr_sub3(vr_i2f3(r.hit_voxel), L)));
    r.ray_dir = vr_shadow_dir(L, r.hit_point, nlight, nlight_ok);
    if (vr_any_zero(r.ray_dir)) return false;
    r.voxel = r.hit_empty;
    r.fm = 0;                                                    /* vr_restart_dda steps back by step * fm: already done */
    vr_restart_dda(r, r.hit_point, {vr_sign(r.ray_dir.x), vr_sign(r.ray_dir.y), vr_sign(r.ray_dir.z)});
    return true;
}

template <bool AUX, bool MULTI>
VR_HD int vr_hit_block(const vr_frame_params &P, RayState &r, int voxel_data, vr_aux *a, bool &first_hit_done) {
    if (AUX && !first_hit_done) {
        first_hit_done = true;
        a->hit[0] = r.voxel.x; a->hit[1] = r.voxel.y; a->hit[2] = r.voxel.z;
        a->face = (uint8_t)(r.fm | ((r.step.x < 0) << 3) | ((r.step.y < 0) << 4) | ((r.step.z < 0) << 5));
        a->hit_type = (uint8_t)voxel_data;
        a->steps_first = (uint32_t)r.dist;
    }
    vf3 fp = {0.0f, 0.0f, 0.0f};     /* face_position */
    float tu = 0.0f, tv = 0.0f;       /* tile_face_position */
    vf3 sgn = {1.0f, 1.0f, 1.0f};
    if (r.fm & 1) {                   /* kernel:586 */
        sgn.x = -1.0f;
        const float tc = VR_SUB(r.t.x, r.delta.x);
        const float zp = VR_DIV(VR_SUB(r.t.z, tc), r.delta.z);
        const float yp = VR_DIV(VR_SUB(r.t.y, tc), r.delta.y);
        fp = {1.00001f, yp, zp}; tu = yp; tv = zp;
    } else if (r.fm & 2) {            /* kernel:601 */
        sgn.y = -1.0f;
        const float tc = VR_SUB(r.t.y, r.delta.y);
        const float xp = VR_DIV(VR_SUB(r.t.x, tc), r.delta.x);
        const float zp = VR_DIV(VR_SUB(r.t.z, tc), r.delta.z);
        fp = {xp, 1.00001f, zp}; tu = xp; tv = zp;
    } else if (r.fm & 4) {            /* kernel:610 */
        sgn.z = -1.0f;
        const float tc = VR_SUB(r.t.z, r.delta.z);
        const float xp = VR_DIV(VR_SUB(r.t.x, tc), r.delta.x);
        const float yp = VR_DIV(VR_SUB(r.t.y, tc), r.delta.y);
        fp = {xp, yp, 1.00001f}; tu = xp; tv = yp;
    }
    /* kernel:626-643 */
    if (r.ray_dir.x > 0.0f) fp.x = VR_ADD(-fp.x, 1.0f);
    if (r.ray_dir.x < 0.0f) tu = VR_ADD(-tu, 1.0f);
    if (r.ray_dir.y > 0.0f) {
        fp.y = VR_ADD(-fp.y, 1.0f);
    } else {
        tu = VR_SUB(1.0f, tu);
        if (r.fm & 4) { tu = VR_SUB(1.0f, tu); tv = VR_SUB(1.0f, tv); }
    }
    if (r.ray_dir.z > 0.0f) fp.z = VR_ADD(-fp.z, 1.0f);
    if (r.ray_dir.z < 0.0f) tv = VR_ADD(-tv, 1.0f);

    const vf3 hit_pos = vr_add3(vr_i2f3(r.voxel), fp);
    const vf3 L = {P.light_pos[0][0], P.light_pos[0][1], P.light_pos[0][2]};

    if (voxel_data == 5 && !r.shadow) {                                  /* kernel:649 */
        r.shadow = true;
        bool clamped;
        const vf3 tex = vr_atlas_fetch(P, tu, tv, 5, 0, clamped);
        if (AUX) a->flags |= VR_FL_LIT | (clamped ? VR_FL_ATLAS_CLAMP : 0);
        r.voxel_color.x = VR_ADD(r.voxel_color.x, VR_HALF(tex.x));
        r.voxel_color.y = VR_ADD(r.voxel_color.y, VR_HALF(tex.y));
        r.voxel_color.z = VR_ADD(r.voxel_color.z, VR_HALF(tex.z));
        const vf3 cam = {P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]};
        const vi3 nrm = {(r.fm & 1) * r.step.x, ((r.fm >> 1) & 1) * r.step.y, ((r.fm >> 2) & 1) * r.step.z};
        if (MULTI) {
            r.hit_point = hit_pos;
            r.hit_voxel = r.voxel;
            r.hit_normal = nrm;
            r.hit_empty = {r.voxel.x - nrm.x, r.voxel.y - nrm.y, r.voxel.z - nrm.z};
        }
        vf3 nlight = {0.0f, 0.0f, 0.0f};
        bool nlight_ok;
        r.color = vr_view_light(r.voxel_color, 0.0f, vr_sub3(hit_pos, L), P.light_rgbi[0], vr_sub3(hit_pos, cam), nrm, &nlight, &nlight_ok);
        r.fog_distance = (float)r.dist;
        r.max_distance = (int)VR_ADD((float)r.dist, vr_length(vr_sub3(vr_i2f3(r.voxel), L)));   /* kernel:667 */
        r.ray_dir = vr_shadow_dir(L, hit_pos, nlight, nlight_ok);
        if (vr_any_zero(r.ray_dir)) return VR_ST_SKIP_REDIRECT;          /* kernel:671 */
        vr_restart_dda(r, hit_pos, {vr_sign(r.ray_dir.x), vr_sign(r.ray_dir.y), vr_sign(r.ray_dir.z)});
        return -1;
    }
    if (voxel_data == 6 && !r.shadow) {                                  /* kernel:682 */
        bool clamped;
        const vf3 tex = vr_atlas_fetch(P, tu, tv, 3, 4, clamped);
        if (AUX) a->flags |= VR_FL_REFLECTED | (clamped ? VR_FL_ATLAS_CLAMP : 0);
        r.voxel_color.x = VR_ADD(r.voxel_color.x, VR_QUARTER(tex.x));
        r.voxel_color.y = VR_ADD(r.voxel_color.y, VR_QUARTER(tex.y));
        r.voxel_color.z = VR_ADD(r.voxel_color.z, VR_QUARTER(tex.z));
        r.ray_dir = {VR_MUL(r.ray_dir.x, sgn.x), VR_MUL(r.ray_dir.y, sgn.y), VR_MUL(r.ray_dir.z, sgn.z)};
        if (vr_any_zero(r.ray_dir)) return VR_ST_SKIP_REDIRECT;          /* kernel:694 */
        vr_restart_dda(r, hit_pos, {1, 1, 1});                           /* kernel:698 precedence: always +1 */
        r.bounce += 1;
        return -1;
    }
    r.color.w = MULTI ? VR_ADD(r.alpha_before, 0.1f) : 0.1f;             /* kernel:708 */
    if (MULTI && vr_more_lights(P, r)) return vr_next_light(P, r) ? -1 : VR_ST_SKIP_REDIRECT;
    return VR_ST_SHADOW_HIT;
}

/* kernel:716-721.  Packs RGBA8 little-endian (R in the low byte). */
VR_HD uint32_t vr_epilogue(const RayState &r) {
    const float m = VR_SUB(1.0f, vr_max(VR_DIV_700(r.fog_distance), 0.0f));
    return vr_unorm8(VR_MUL(r.color.x, m)) | (vr_unorm8(VR_MUL(r.color.y, m)) << 8) |
           (vr_unorm8(VR_MUL(r.color.z, m)) << 16) | (vr_unorm8(VR_MUL(r.color.w, m)) << 24);
}

VR_HD void vr_aux_init(vr_aux *a, const vr_frame_params &P) {
    a->hit[0] = a->hit[1] = a->hit[2] = -1;
    a->face = 0; a->status = 0; a->flags = 0; a->hit_type = 0;
    a->steps_first = 0; a->steps_total = 0; a->node_fetches = 0; a->lookups = 0;
    if (P.cam_pos[0] == floorf(P.cam_pos[0]) || P.cam_pos[1] == floorf(P.cam_pos[1]) ||
        P.cam_pos[2] == floorf(P.cam_pos[2]))
        a->flags |= VR_FL_FRAC0;
}

/* ---------------------------------------------------------------------------------------------
 * Dense variant: kernel:357 loop with the `else` branch (kernel:555-570).
 * Returns true if the pixel must be written (packed colour in *rgba_out).
 * ------------------------------------------------------------------------------------------- */
template <bool AUX, bool MULTI>
VR_HD bool vr_trace_dense(const vr_frame_params &P, int x, int y, uint32_t *rgba_out, vr_aux *a) {
    RayState r;
    if (AUX) vr_aux_init(a, P);
    if (!vr_ray_setup(P, x, y, r)) {
        if (AUX) a->status = VR_ST_SKIP_PRIMARY;
        return false;
    }
    const int X = P.dim[0], Y = P.dim[1], Z = P.dim[2];
    bool first_hit_done = false;
    int status = VR_ST_MAXDIST;
    for (;;) {
        if (!(r.dist < r.max_distance && r.bounce < P.max_bounces)) {
            /* multi-light extension: this light is not blocked, on to the next one */
            if (MULTI && r.bounce < P.max_bounces && vr_more_lights(P, r)) {
                if (!vr_next_light(P, r)) { status = VR_ST_SKIP_REDIRECT; break; }
                r.dist++;
                continue;
            }
            break;
        }
        vr_dda_step(r);
        if (AUX && (r.fm & (r.fm - 1))) a->flags |= VR_FL_TIE;
        if (r.voxel.x >= X || r.voxel.y >= Y || r.voxel.z >= Z || r.voxel.x < 0 || r.voxel.y < 0 || r.voxel.z < 0) {
            if (MULTI && vr_more_lights(P, r)) {
                if (!vr_next_light(P, r)) { status = VR_ST_SKIP_REDIRECT; break; }
                r.dist++;
                continue;
            }
            vr_out_of_bounds(r);
            status = VR_ST_OOB;
            break;
        }
        const int voxel_data =
            (int)P.map[(size_t)r.voxel.x + (size_t)X * ((size_t)r.voxel.y + (size_t)Z * (size_t)r.voxel.z)];
        if (voxel_data == 5 || voxel_data == 6) {
            const int st = vr_hit_block<AUX, MULTI>(P, r, voxel_data, a, first_hit_done);
            if (st >= 0) { status = st; break; }
        }
        r.dist++;
    }
    if (status == VR_ST_SKIP_REDIRECT) {
        if (AUX) { a->status = (uint8_t)status; a->steps_total = (uint32_t)r.dist; }
        return false;
    }
    if (status == VR_ST_MAXDIST && r.bounce >= P.max_bounces) status = VR_ST_BOUNCES;
    if (AUX) { a->status = (uint8_t)status; a->steps_total = (uint32_t)r.dist; }
    *rgba_out = vr_epilogue(r);
    return true;
}

/* ---------------------------------------------------------------------------------------------
 * SVO variant.  Same loop, but `map[voxel]` is answered by the 64-tree:
 *   - the empty cell (edge 1<<cs at origin co) found by the last lookup is cached; a step that
 *     stays inside it needs no lookup, no bounds test and no memory access;
 *   - a lookup resumes from the lowest stacked ancestor that still contains the new voxel
 *     (XOR of the coordinates gives the level), then descends one 16-byte node per two octree
 *     levels until it meets an empty slot (new cached cell) or a set voxel bit (hit).
 * Stack: node indices per level, provided by the caller (shared memory on the device).
 * Requires a cubic power-of-two map so that cells never straddle the map boundary.
 * ------------------------------------------------------------------------------------------- */
struct vr_node_regs { unsigned long long mask; uint32_t base; uint32_t planes; };

VR_HD vr_node_regs vr_load_node(const vr_frame_params &P, uint32_t idx) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(P.nodes) + idx);          /* one 128-bit load */
    return {(unsigned long long)v.x | ((unsigned long long)v.y << 32), v.z, v.w};
#else
    const vr_node &n = P.nodes[idx];
    return {(unsigned long long)n.mask_lo | ((unsigned long long)n.mask_hi << 32), n.child_base, n.aux};
#endif
}

/* ---- in-cell walk ---------------------------------------------------------------------------------
 * Inside a known-empty cell the reference loop (kernel:558-560) reduces to: find the smallest
 * intersection_t, add delta_t on every axis that attains it (ties step several axes), count the step.
 * Per axis the state is (t, k): the next crossing time and the number of crossings still needed to leave
 * the cell along that axis; `rem` counts the steps max_distance still allows.  The walk stops when some
 * k reaches 0 (that step left the cell) or rem reaches 0.
 * Written so that almost all work lands on the FMA pipe (2x the ALU pipe's rate on sm_100):
 *   m  = min3(tx,ty,tz)                         1 FMNMX3          (ALU)
 *   ma = (ta == m) ? 1.0f : 0.0f                3 FSET.BF         (ALU)   exact tie semantics of kernel:558
 *   ta = fma(da, ma, ta);  ka -= ma             3 FFMA + 3 FADD   (FMA)   da*ma is exact => same as ta + da*ma
 *   stop <=> kx*ky*kz*rem == 0                  1 FADD + 3 FMUL   (FMA) + 1 FSETP (ALU)
 * No integer work, no memory access. */
struct vr_walk_state {
    float tx, ty, tz, kx, ky, kz, rem;
};

VR_HD float vr_min3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float m;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
    return m;
#else
    return fminf(fminf(a, b), c);
#endif
}

#if defined(__CUDA_ARCH__)
#define VR_FMA_EXACT(a, b, c) __fmaf_rn((a), (b), (c))      /* only used where a*b is exact */
#define VR_FMA(a, b, c) __fmaf_rn((a), (b), (c))            /* a true fused multiply-add on both sides */
#else
#define VR_FMA_EXACT(a, b, c) ((c) + (a) * (b))
#define VR_FMA(a, b, c) fmaf((a), (b), (c))
#endif

/* one step; returns kx*ky*kz*rem (zero <=> stop) */
VR_HD float vr_walk_step(vr_walk_state &w, const vf3 &d) {
    const float m = vr_min3(w.tx, w.ty, w.tz);
    const float mx = (w.tx == m) ? 1.0f : 0.0f;
    const float my = (w.ty == m) ? 1.0f : 0.0f;
    const float mz = (w.tz == m) ? 1.0f : 0.0f;
    w.tx = VR_FMA_EXACT(d.x, mx, w.tx);
    w.ty = VR_FMA_EXACT(d.y, my, w.ty);
    w.tz = VR_FMA_EXACT(d.z, mz, w.tz);
    w.kx = VR_SUB(w.kx, mx);
    w.ky = VR_SUB(w.ky, my);
    w.kz = VR_SUB(w.kz, mz);
    w.rem = VR_SUB(w.rem, 1.0f);
    return VR_MUL(VR_MUL(w.kx, w.ky), VR_MUL(w.kz, w.rem));
}

/* crossings of one axis needed to leave the cell [origin, origin+size) from `voxel`; clamped to 512
 * (a wider cell is then simply left early and found again by the lookup) */
VR_HD int vr_exit_count(int step, int voxel, int origin, int size, int cap) {
    const int r = step > 0 ? origin + size - voxel : voxel - origin + 1;
    return r > cap ? cap : r;
}

/* The literal walk of one cell (kernel:558-560 step by step, full face mask): used when the float state
 * is not finite (the packed walk assumes ordered compares) and to recover the face mask of a hit that
 * follows a cell in which a multi-axis (tie) step occurred.  Returns true if the last step left the cell. */
template <bool AUX>
VR_HD bool vr_walk_literal(RayState &r, vi3 co, vi3 ce, vr_aux *a) {
    for (;;) {
        vr_dda_step(r);
        if (AUX && (r.fm & (r.fm - 1))) a->flags |= VR_FL_TIE;
        if ((unsigned)(r.voxel.x - co.x) >= (unsigned)ce.x || (unsigned)(r.voxel.y - co.y) >= (unsigned)ce.y ||
            (unsigned)(r.voxel.z - co.z) >= (unsigned)ce.z) return true;
        r.dist++;
        if (!(r.dist < r.max_distance)) return false;
    }
}

VR_HD bool vr_ray_finite(const RayState &r) {
    const float s = ((r.t.x + r.t.y) + r.t.z) + ((r.delta.x + r.delta.y) + r.delta.z);
    return s - s == 0.0f;             /* false for inf / NaN */
}

/* ---- per-axis walk (option "walk" = 1) ------------------------------------------------------------
 * The three intersection_t sequences are independent: axis a's crossing times are t_a, t_a+d_a, (t_a+d_a)+d_a, ...
 * whatever the other axes do (kernel:559 adds delta_t only on the axes that step).  So a cell can be walked
 * one AXIS at a time instead of one STEP at a time:
 *   1. predict the exit axis A from the closed forms t + (r-1)*d (a prediction only, never used as a value);
 *   2. run A's r_A - 1 additions as a bare, unrolled FADD chain: T = the exact time of the step that leaves
 *      the cell;
 *   3. for each of the other two axes estimate the number of crossings before T by a division, run that many
 *      additions minus a safety margin as a bare chain, and finish with a compare-per-step loop (2-3 rounds);
 *      a crossing exactly at T joins the exit step (multi-axis step, kernel:558 tie rule).
 * About 1.3 issue slots per step instead of 16 for the merged walk, and the float state / crossing counts are
 * EXACTLY those of the merged walk (same additions, same order per axis).  What is not observed is a tie
 * between two crossings strictly inside the cell (the reference then moves diagonally and counts ONE step):
 * distance_traveled is then one too large per such tie.  Those rays are "degenerate" in BASELINE.json's sense
 * (they pass exactly through a voxel edge); the oracle flags them (VRO_FL_TIE) and the tests compare them with
 * the north_star tolerance instead of bit-exactly.
 * Returns VR_AXES_DONE, VR_AXES_FALLBACK (exit axis mispredicted: use the merged walk) or VR_AXES_MAXDIST
 * (max_distance is reached inside the cell: the ray ends here and nothing of its state is needed any more). */
enum { VR_AXES_FALLBACK = 0, VR_AXES_DONE = 1, VR_AXES_MAXDIST = 2 };
#ifndef VR_AXES_MIN_COUNT
#define VR_AXES_MIN_COUNT 12     /* rx+ry+rz above which a cell is walked per axis (measured) */
#endif

VR_HD int vr_f2bits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(v);
#else
    int b; memcpy(&b, &v, 4); return b;
#endif
}
VR_HD float vr_bits2f(int b) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(b);
#else
    float v; memcpy(&v, &b, 4); return v;
#endif
}

/* t after n additions of d, each rounded to nearest-even like kernel:559 does them one by one -- but in O(number of
 * binades crossed) instead of O(n).  While t stays inside one binade [2^e, 2^(e+1)) every value is a multiple of
 * u = ulp = 2^(e-23), so RN(t + d) = t + dr with the SAME dr = "d rounded to the grid u" for every such addition;
 * j of them are t + j*dr, which one FMA evaluates exactly (the result is a grid point below 2^(e+1)).  Only the
 * addition that leaves the binade (rounded on the coarser grid) and the case where d falls exactly half way between
 * two grid points (ties-to-even: the increment then depends on the parity of t) are done literally.
 * dr is measured, not derived: two literal additions t1 = t+d, t2 = t1+d, dr = t2 - t1 (exact, same binade).
 * All of this holds for t > 0 and d > 0 (t grows); negative t is walked literally until it turns positive. */
#ifndef VR_JUMP_MIN
#define VR_JUMP_MIN 64           /* chains at least this long use the jumps (one jump costs about 35 issue slots); 0 = never */
#endif
#ifndef VR_CHAIN_UNROLL
#define VR_CHAIN_UNROLL 8
#endif
VR_HD float vr_add_chain(float t, float d, int n) {
    while (VR_JUMP_MIN > 0 && n >= VR_JUMP_MIN) {
        const float t1 = VR_ADD(t, d);
        const float t2 = VR_ADD(t1, d);
        t = t2;
        n -= 2;
        VR_PROF_ADD(jumps, 1);
        const int e1 = vr_f2bits(t1) & (int)0xff800000;      /* sign + exponent: negative for t1 < 0 */
        /* literal while t is negative (the get_oct_vox start bias, kernel:353, can make it so: |t| then SHRINKS towards
         * finer grids), when the two additions left the binade, and where u/2 or 2^(e+1) would not be normal numbers */
        if ((vr_f2bits(t2) & 0x7f800000) != e1 || e1 < (30 << 23) || e1 >= (250 << 23)) continue;
        const float dr = VR_SUB(t2, t1);
        if (fabsf(VR_SUB(dr, d)) == vr_bits2f(e1 - (24 << 23))) break;   /* d is half way between grid points: literal */
        if (dr == 0.0f) return t;                                         /* d < u/2: t no longer moves */
        const float top = vr_bits2f(e1 + (1 << 23));
#if defined(__CUDA_ARCH__)
        const float q = __fdividef(VR_SUB(top, t2), dr);
#else
        const float q = VR_SUB(top, t2) / dr;
#endif
        int k = (q < 1.0e6f ? (int)q : 1000000) - 2;                      /* t2 + k*dr stays below top (q is off by << 1) */
        k = k < n ? k : n;
        if (k > 0) {
            t = VR_FMA((float)k, dr, t2);
            n -= k;
        }
    }
    VR_PROF_ADD(adds, n);
#if VR_CHAIN_UNROLL >= 8
    for (; n >= 8; n -= 8) t = VR_ADD(VR_ADD(VR_ADD(VR_ADD(VR_ADD(VR_ADD(VR_ADD(VR_ADD(t, d), d), d), d), d), d), d), d);
#endif
    for (; n >= 4; n -= 4) t = VR_ADD(VR_ADD(VR_ADD(VR_ADD(t, d), d), d), d);
    for (; n > 0; --n) t = VR_ADD(t, d);
    return t;
}

#ifndef VR_COUNT_MARGIN
#define VR_COUNT_MARGIN 1
#endif
/* crossings of one non-exit axis strictly before T; t ends at the first crossing time >= T */
VR_HD int vr_count_before(float &t, float d, float inv_d, float T) {
    int c = (int)(VR_MUL(VR_SUB(T, t), inv_d)) - VR_COUNT_MARGIN;   /* estimate minus margin: never overshoots (the estimate is off by < 0.5) */
    c = c < 0 ? 0 : c;
    t = vr_add_chain(t, d, c);
    VR_PROF_ADD(chain_b, c);
    while (t < T) { t = VR_ADD(t, d); c++; VR_PROF_ADD(fix, 1); }
    return c;
}

VR_HD int vr_walk_axes(RayState &r, int rx, int ry, int rz, int nmax, int &ax, int &ay, int &az, int &n, bool &exit_tie) {
    const float ex = VR_FMA_EXACT((float)(rx - 1), r.delta.x, r.t.x);     /* predictions; rounding is irrelevant */
    const float ey = VR_FMA_EXACT((float)(ry - 1), r.delta.y, r.t.y);
    const float ez = VR_FMA_EXACT((float)(rz - 1), r.delta.z, r.t.z);
    const int sel = (ex <= ey && ex <= ez) ? 0 : (ey <= ez ? 1 : 2);
    /* roles: A = predicted exit axis, B and C the next two in cyclic order */
    float tA = sel == 0 ? r.t.x : (sel == 1 ? r.t.y : r.t.z), dA = sel == 0 ? r.delta.x : (sel == 1 ? r.delta.y : r.delta.z);
    float tB = sel == 0 ? r.t.y : (sel == 1 ? r.t.z : r.t.x), dB = sel == 0 ? r.delta.y : (sel == 1 ? r.delta.z : r.delta.x);
    float tC = sel == 0 ? r.t.z : (sel == 1 ? r.t.x : r.t.y), dC = sel == 0 ? r.delta.z : (sel == 1 ? r.delta.x : r.delta.y);
    const float iB = fabsf(sel == 0 ? r.ray_dir.y : (sel == 1 ? r.ray_dir.z : r.ray_dir.x));   /* ~ 1 / dB */
    const float iC = fabsf(sel == 0 ? r.ray_dir.z : (sel == 1 ? r.ray_dir.x : r.ray_dir.y));
    const int rA = sel == 0 ? rx : (sel == 1 ? ry : rz);
    const int rB = sel == 0 ? ry : (sel == 1 ? rz : rx);
    const int rC = sel == 0 ? rz : (sel == 1 ? rx : ry);
    tA = vr_add_chain(tA, dA, rA - 1);
    VR_PROF(chain_a, rA - 1);
    const float T = tA;                                                   /* time of the step that leaves the cell */
    tA = VR_ADD(tA, dA);
    int tieB = 0, tieC = 0;
    int cB = vr_count_before(tB, dB, iB, T);
    if (tB == T && cB < rB) { tB = VR_ADD(tB, dB); cB++; tieB = 1; }
    if (cB > rB || (cB == rB && !tieB)) return VR_AXES_FALLBACK;          /* B left the cell before A: mispredicted */
    int cC = vr_count_before(tC, dC, iC, T);
    if (tC == T && cC < rC) { tC = VR_ADD(tC, dC); cC++; tieC = 1; }
    if (cC > rC || (cC == rC && !tieC)) return VR_AXES_FALLBACK;
    n = rA + cB + cC - tieB - tieC;
    if (n > nmax) return VR_AXES_MAXDIST;
    r.t.x = sel == 0 ? tA : (sel == 1 ? tC : tB);
    r.t.y = sel == 0 ? tB : (sel == 1 ? tA : tC);
    r.t.z = sel == 0 ? tC : (sel == 1 ? tB : tA);
    ax = sel == 0 ? rA : (sel == 1 ? cC : cB);
    ay = sel == 0 ? cB : (sel == 1 ? rA : cC);
    az = sel == 0 ? cC : (sel == 1 ? cB : rA);
    const int bA = 1 << sel, bB = sel == 2 ? 1 : (bA << 1), bC = sel == 0 ? 4 : (sel == 1 ? 1 : 2);
    r.fm = bA | (tieB ? bB : 0) | (tieC ? bC : 0);
    exit_tie = (tieB | tieC) != 0;
    return VR_AXES_DONE;
}

/* ---- leaf-brick walk --------------------------------------------------------------------------------
 * Inside a leaf node (a brick of 4^3 voxels whose occupancy is the node's 64-bit mask, held in registers) the
 * merged step is followed by a bit test of the voxel just entered, so that the empty voxels of the brick cost no
 * lookup at all: rays that skim a surface visit one brick after the other instead of one 1^3 / 2^3 cell after the
 * other.  Stops when a step leaves the brick (VR_BRICK_EXIT) or lands on a solid voxel (VR_BRICK_HIT, `bit` = its
 * slot).  The caller guarantees that max_distance cannot be reached inside the brick (a brick holds at most 10
 * steps).  All masks of the last step are known here, so the face mask is exact even with ties. */
enum { VR_BRICK_EXIT = 0, VR_BRICK_HIT = 1 };

VR_HD int vr_walk_brick(RayState &r, unsigned long long mask, vi3 co, int &n, int &bit, bool &tie) {
    const int lx = r.voxel.x - co.x, ly = r.voxel.y - co.y, lz = r.voxel.z - co.z;
    const int rx = r.step.x > 0 ? 4 - lx : lx + 1, ry = r.step.y > 0 ? 4 - ly : ly + 1, rz = r.step.z > 0 ? 4 - lz : lz + 1;
    float tx = r.t.x, ty = r.t.y, tz = r.t.z;
    float kx = (float)rx, ky = (float)ry, kz = (float)rz, steps = 0.0f;
    float bitf = (float)(lx | (ly << 2) | (lz << 4));
    const float sx = (float)r.step.x, sy = (float)(4 * r.step.y), sz = (float)(16 * r.step.z);
    float mx, my, mz;
    int res;
    for (;;) {
        const float m = vr_min3(tx, ty, tz);
        mx = (tx == m) ? 1.0f : 0.0f;
        my = (ty == m) ? 1.0f : 0.0f;
        mz = (tz == m) ? 1.0f : 0.0f;
        tx = VR_FMA_EXACT(r.delta.x, mx, tx);
        ty = VR_FMA_EXACT(r.delta.y, my, ty);
        tz = VR_FMA_EXACT(r.delta.z, mz, tz);
        kx = VR_SUB(kx, mx);
        ky = VR_SUB(ky, my);
        kz = VR_SUB(kz, mz);
        bitf = VR_FMA_EXACT(mx, sx, VR_FMA_EXACT(my, sy, VR_FMA_EXACT(mz, sz, bitf)));   /* small integers: exact */
        steps = VR_ADD(steps, 1.0f);
        if (VR_MUL(VR_MUL(kx, ky), kz) == 0.0f) { res = VR_BRICK_EXIT; break; }
        if ((mask >> (int)bitf) & 1ull) { res = VR_BRICK_HIT; break; }
    }
    const int ax = rx - (int)kx, ay = ry - (int)ky, az = rz - (int)kz;
    r.t = {tx, ty, tz};
    r.voxel.x += r.step.x * ax;
    r.voxel.y += r.step.y * ay;
    r.voxel.z += r.step.z * az;
    r.fm = (mx != 0.0f ? 1 : 0) | (my != 0.0f ? 2 : 0) | (mz != 0.0f ? 4 : 0);
    n = (int)steps;
    bit = (int)bitf;
    tie = (ax + ay + az) != n;
    return res;
}

/* The empty box handed to the in-cell walk when the descent meets the empty slot ci of a node (slots of edge 1<<s).
 * The walk works in any empty box, so the box is grown inside the node as far as the 64-bit mask proves it empty:
 *   - if the whole 4x4 plane of slots through ci perpendicular to some axis is empty, the box is the run of
 *     consecutive empty planes around it (heightfield-like scenes: everything above the surface inside the node);
 *   - else, if the 2x2x2 octant of slots around ci is empty, that octant (the odd levels of the reference's 2^3
 *     octree, recovered from the 4^3 mask: slots ci&0x2A + {0,1,4,5,16,17,20,21});
 *   - else the slot itself. */
#ifndef VR_PLANE_BOXES
#define VR_PLANE_BOXES 0         /* measured on B200: 12 % fewer lookups at C3 but no faster (2.305 vs 2.301 ms), slower at C2 */
#endif
VR_HD void vr_empty_box(unsigned long long m, uint32_t planes, int ci, int s, vi3 v, int N, vi3 &co, vi3 &ce) {
    if (VR_PLANE_BOXES) {
        const int cx = ci & 3, cy = (ci >> 2) & 3, cz = ci >> 4;
        const int axis = ((planes >> (8 + cz)) & 1u) ? 2 : (((planes >> (4 + cy)) & 1u) ? 1 : (((planes >> cx) & 1u) ? 0 : -1));
        if (axis >= 0) {
            const uint32_t e = (planes >> (4 * axis)) & 15u;             /* the 4 planes along `axis`; bit c is set */
            const int c = axis == 2 ? cz : (axis == 1 ? cy : cx);
#if defined(__CUDA_ARCH__)
            const int up = __ffs((int)~(e >> (c + 1))) - 1;               /* empty planes right above / below plane c */
            const int dn = __clz((int)~((e << (31 - c)) << 1));
#else
            const int up = __builtin_ffs((int)~(e >> (c + 1))) - 1;
            const int dn = __builtin_clz(~((e << (31 - c)) << 1));
#endif
            const int ns = s + 2;                                        /* the node's own edge is 1 << ns */
            const int full = 4 << s, o = (c - dn) << s, len = (up + dn + 1) << s;
            co = {(v.x >> ns) << ns, (v.y >> ns) << ns, (v.z >> ns) << ns};
            ce = {full, full, full};
            if (axis == 0) { co.x += o; ce.x = len; }
            if (axis == 1) { co.y += o; ce.y = len; }
            if (axis == 2) { co.z += o; ce.z = len; }
            /* the root may be wider than the map (N not a power of 4): the box must end at the map boundary, where the
             * walk has to stop for the bounds test */
            ce = {ce.x < N - co.x ? ce.x : N - co.x, ce.y < N - co.y ? ce.y : N - co.y, ce.z < N - co.z ? ce.z : N - co.z};
            return;
        }
    }
    const int cs = s + ((((m >> (ci & 0x2A)) & 0x00330033ull) == 0ull) ? 1 : 0);
    co = {(v.x >> cs) << cs, (v.y >> cs) << cs, (v.z >> cs) << cs};
    ce = {1 << cs, 1 << cs, 1 << cs};
}

/* Per-ray traversal state of the SVO variant: everything that lives across cells. */
template <class Stack>
struct vr_svo_ray {
    RayState r;
    vr_node_regs node;        /* current node of the descent */
    int s;                    /* its child shift */
    int level;
    vi3 nv;                   /* a voxel inside the current node */
    vi3 co, ce;               /* cached empty cell: the box of ce voxels at origin co */
    bool brick;               /* the cached cell is the leaf brick of `node` (4^3 voxels, occupancy = node.mask) */
    bool finite;
    bool first_hit_done;
    Stack stk;
};

enum { VR_CELL_CONTINUE = -1, VR_CELL_NO_WRITE = -2, VR_CELL_NEXT_LIGHT = -3 };   /* otherwise: terminal VR_ST_* status */

/* kernel:276-354 for one pixel.  Returns false when the pixel is skipped (kernel:293). */
template <bool AUX, class Stack>
VR_HD bool vr_svo_begin(const vr_frame_params &P, int x, int y, vr_svo_ray<Stack> &q, vr_aux *a) {
    if (AUX) vr_aux_init(a, P);
    if (!vr_ray_setup(P, x, y, q.r)) {
        if (AUX) a->status = VR_ST_SKIP_PRIMARY;
        return false;
    }
    q.s = P.root_shift;
    q.level = 0;
    q.node = vr_load_node(P, 0);
    q.stk.set(0, 0u);
    q.nv = {0, 0, 0};
    if (AUX) a->node_fetches = 1;
    /* the first "cell" is the camera voxel itself: the reference steps before it loads (kernel:555-570),
     * so that voxel is never tested */
    q.co = q.r.voxel;
    q.ce = {1, 1, 1};
    q.brick = false;
    q.finite = vr_ray_finite(q.r);
    q.first_hit_done = false;
    return true;
}

/* One iteration of the cell loop: walk the cached cell, then look the new voxel up.
 * Returns VR_CELL_CONTINUE, VR_CELL_NO_WRITE (pixel skipped after a redirect, kernel:671/694) or the
 * terminal status. */
template <bool AUX, int WALK, bool MULTI, class Stack>
VR_HD int vr_svo_cell(const vr_frame_params &P, vr_svo_ray<Stack> &q, vr_aux *a) {
    RayState &r = q.r;
    if (!(r.dist < r.max_distance && r.bounce < P.max_bounces)) return r.bounce >= P.max_bounces ? VR_ST_BOUNCES : VR_ST_MAXDIST;
    const int N = P.dim[0];
    /* ---- (1) walk inside the cached cell: no memory access, no bounds test */
    bool tie_cell = false;
    const vf3 t0 = r.t;                                                  /* state at cell entry, for a replay */
    int n = 0, ax = 0, ay = 0, az = 0;
    int voxel_data = 0;
    bool known = false;                      /* the brick walk already knows what the voxel just entered holds */
    const int nmax = r.max_distance - r.dist;
    /* a brick is walked as such unless the ray is about to end (then voxel by voxel, like any 1^3 cell) */
    const bool as_brick = q.brick && q.finite && nmax > 12;
    const vi3 wce = q.brick && !as_brick ? vi3{1, 1, 1} : q.ce;
    const vi3 wco = q.brick && !as_brick ? r.voxel : q.co;
    if (as_brick) {
        int bit;
        bool tie;
        const int res = vr_walk_brick(r, q.node.mask, q.co, n, bit, tie);
        VR_PROF(path, 1); VR_PROF(n, n);
        if (AUX && tie) a->flags |= VR_FL_TIE;
        r.dist += n - 1;
        if (res == VR_BRICK_HIT) {
            voxel_data = (int)(int8_t)P.leaf_types[q.node.base + (uint32_t)VR_POPC64(q.node.mask & ((1ull << bit) - 1ull))];
            known = true;
        }
    } else if (q.finite) {
        /* no axis can cross more often than max_distance leaves steps: a ray that ends inside the cell (most shadow
         * rays end in mid-air, at the light's distance) must not pay for the walk to the far side of a wide cell */
        const int rcap = nmax < 511 ? nmax + 1 : 512;
        const int rx = vr_exit_count(r.step.x, r.voxel.x, wco.x, wce.x, rcap);
        const int ry = vr_exit_count(r.step.y, r.voxel.y, wco.y, wce.y, rcap);
        const int rz = vr_exit_count(r.step.z, r.voxel.z, wco.z, wce.z, rcap);
        bool exit_tie = false;
        /* short walks (cells of a few voxels next to surfaces) are cheaper step by step: 16 slots per step against
         * ~200 of fixed cost for the per-axis machinery */
        const int axes = (WALK == 1 && rx + ry + rz > VR_AXES_MIN_COUNT)
                             ? vr_walk_axes(r, rx, ry, rz, nmax, ax, ay, az, n, exit_tie) : VR_AXES_FALLBACK;
        if (axes == VR_AXES_MAXDIST) {
            r.dist = r.max_distance;                                     /* kernel:357 ends the loop; t / voxel are dead */
            return VR_ST_MAXDIST;
        }
        if (axes == VR_AXES_DONE) {
            VR_PROF(path, 2); VR_PROF(n, n);
            /* per-axis walk: float state, crossing counts and face mask are exact; see vr_walk_axes */
            r.voxel.x += r.step.x * ax;
            r.voxel.y += r.step.y * ay;
            r.voxel.z += r.step.z * az;
            if (AUX && exit_tie) a->flags |= VR_FL_TIE;
            r.dist += n - 1;
        } else {
            vr_walk_state w = {r.t.x, r.t.y, r.t.z, (float)rx, (float)ry, (float)rz, (float)nmax};
            while (vr_walk_step(w, r.delta) != 0.0f) {}
            r.t = {w.tx, w.ty, w.tz};
            const float kx = w.kx, ky = w.ky, kz = w.kz;
            ax = rx - (int)kx; ay = ry - (int)ky; az = rz - (int)kz;     /* crossings done per axis */
            n = nmax - (int)w.rem;                                       /* steps done */
            VR_PROF(path, (WALK == 1 && rx + ry + rz > VR_AXES_MIN_COUNT) ? 4 : 3); VR_PROF(n, n);
            r.voxel.x += r.step.x * ax;
            r.voxel.y += r.step.y * ay;
            r.voxel.z += r.step.z * az;
            /* without ties exactly one axis moved per step, and the axis whose k hit 0 is the face crossed by the
             * last step; with a tie somewhere in the cell the mask is recovered by a replay if it is needed */
            tie_cell = (ax + ay + az) != n;
            r.fm = (kx == 0.0f ? 1 : 0) | (ky == 0.0f ? 2 : 0) | (kz == 0.0f ? 4 : 0);
            if (AUX && tie_cell) a->flags |= VR_FL_TIE;
            if (r.fm == 0) { r.dist += n; return VR_ST_MAXDIST; }        /* max_distance reached inside the cell */
            r.dist += n - 1;
        }
    } else if (!vr_walk_literal<AUX>(r, wco, wce, a)) {
        return VR_ST_MAXDIST;
    }
    /* ---- (2) the last step left the cell: bounds test, octree lookup, hit handling */
    if (!known) {
        if ((unsigned)r.voxel.x >= (unsigned)N || (unsigned)r.voxel.y >= (unsigned)N || (unsigned)r.voxel.z >= (unsigned)N) {
            if (MULTI && vr_more_lights(P, r)) return VR_CELL_NEXT_LIGHT;   /* left the map unblocked: next light */
            vr_out_of_bounds(r);
            return VR_ST_OOB;
        }
        if (AUX) a->lookups++;
        VR_PROF(lookup, 1);
        /* pop to the lowest ancestor containing the voxel */
        const int nx = (r.voxel.x ^ q.nv.x) | (r.voxel.y ^ q.nv.y) | (r.voxel.z ^ q.nv.z);
        if ((nx >> (q.s + 2)) != 0) {
            do { q.s += 2; q.level--; VR_PROF_ADD(pops, 1); } while ((nx >> (q.s + 2)) != 0);
            q.node = vr_load_node(P, q.stk.get(q.level));
            if (AUX) a->node_fetches++;
        }
        q.nv = r.voxel;
        for (;;) {
            if (q.node.base & VR_NODE_SOLID) {                           /* a collapsed solid subtree: a set voxel of its type */
                voxel_data = (int)(int8_t)(q.node.base & 0xffu);
                break;
            }
            const int s = q.s;
            const int ci = ((r.voxel.x >> s) & 3) | (((r.voxel.y >> s) & 3) << 2) | (((r.voxel.z >> s) & 3) << 4);
            if (!((q.node.mask >> ci) & 1ull)) {                         /* empty slot: cache the cell */
                if (s == 0) {
                    /* an empty voxel of a leaf brick: the whole brick becomes the cell, walked with bit tests */
                    q.brick = true;
                    q.ce = {4, 4, 4};
                    q.co = {r.voxel.x & ~3, r.voxel.y & ~3, r.voxel.z & ~3};
                    break;
                }
                /* if the whole 2x2x2 octant of slots around it is empty the cell is twice as wide -- the odd levels
                 * of the reference's 2^3 octree, recovered from the 4^3 mask (slots ci&0x2A + {0,1,4,5,16,17,20,21}) */
                q.brick = false;
                vr_empty_box(q.node.mask, q.node.planes, ci, s, r.voxel, N, q.co, q.ce);
                break;
            }
            const uint32_t rank = (uint32_t)VR_POPC64(q.node.mask & ((1ull << ci) - 1ull));
            if (s == 0) {                                                /* a set voxel bit */
                voxel_data = (int)(int8_t)P.leaf_types[q.node.base + rank];
                break;
            }
            const uint32_t child = q.node.base + rank;
            q.level++;
            q.s -= 2;
            q.stk.set(q.level, child);
            q.node = vr_load_node(P, child);
            VR_PROF_ADD(loads, 1);
            if (AUX) a->node_fetches++;
        }
    }
    if (voxel_data == 5 || voxel_data == 6) {
        if (tie_cell) {
            /* a multi-axis step happened in the cell just left: the hit's face mask may have more than the exit
             * axis set.  Replay the cell literally from its entry state (same t, voxel, dist; exact fm). */
            r.voxel.x -= r.step.x * ax;
            r.voxel.y -= r.step.y * ay;
            r.voxel.z -= r.step.z * az;
            r.dist -= n - 1;
            r.t = t0;
            vr_walk_literal<false>(r, wco, wce, a);
            VR_PROF(replay, 1);
        }
        VR_PROF(hit, 1);
        const int st = vr_hit_block<AUX, MULTI>(P, r, voxel_data, a, q.first_hit_done);
        if (st == VR_ST_SKIP_REDIRECT) {
            if (AUX) { a->status = (uint8_t)st; a->steps_total = (uint32_t)r.dist; }
            return VR_CELL_NO_WRITE;
        }
        if (st >= 0) return st;
        /* the ray was redirected and restarts from the voxel it came from, which is known to be empty: that voxel
         * is the cell (the previous cell may have been a brick whose node is no longer the current one) */
        q.finite = vr_ray_finite(r);
        q.brick = false;
        q.ce = {1, 1, 1};
        q.co = r.voxel;
    }
    r.dist++;
    return VR_CELL_CONTINUE;
}

/* kernel:716-721 + aux bookkeeping for a finished ray */
template <bool AUX, class Stack>
VR_HD uint32_t vr_svo_finish(vr_svo_ray<Stack> &q, int status, vr_aux *a) {
    if (AUX) { a->status = (uint8_t)status; a->steps_total = (uint32_t)q.r.dist; }
    return vr_epilogue(q.r);
}

/* One vr_svo_cell plus the multi-light extension's turn-over: a shadow ray that ends unblocked (max_distance or
 * out of the map) hands over to the next light, restarting from the stored hit. */
template <bool AUX, int WALK, bool MULTI, class Stack>
VR_HD int vr_svo_round(const vr_frame_params &P, vr_svo_ray<Stack> &q, vr_aux *a) {
    int rc = vr_svo_cell<AUX, WALK, MULTI>(P, q, a);
    if (MULTI && (rc == VR_CELL_NEXT_LIGHT || (rc == VR_ST_MAXDIST && q.r.bounce < P.max_bounces && vr_more_lights(P, q.r)))) {
        if (!vr_next_light(P, q.r)) {
            if (AUX) { a->status = (uint8_t)VR_ST_SKIP_REDIRECT; a->steps_total = (uint32_t)q.r.dist; }
            return VR_CELL_NO_WRITE;
        }
        q.r.dist++;
        q.finite = vr_ray_finite(q.r);
        q.brick = false;
        q.ce = {1, 1, 1};
        q.co = q.r.voxel;
        rc = VR_CELL_CONTINUE;
    }
    return rc;
}

/* whole pixel, static pixel->thread mapping */
template <bool AUX, int WALK, bool MULTI, class Stack>
VR_HD bool vr_trace_svo(const vr_frame_params &P, int x, int y, uint32_t *rgba_out, vr_aux *a, Stack &stk) {
    vr_svo_ray<Stack> q;
    q.stk = stk;
    if (!vr_svo_begin<AUX>(P, x, y, q, a)) return false;
    int rc;
    while ((rc = vr_svo_round<AUX, WALK, MULTI>(P, q, a)) == VR_CELL_CONTINUE) {}
    if (rc == VR_CELL_NO_WRITE) return false;
    *rgba_out = vr_svo_finish<AUX>(q, rc, a);
    return true;
}

#endif
