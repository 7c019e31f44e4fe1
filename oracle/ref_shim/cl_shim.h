/*
 * cl_shim.h -- TEST INFRASTRUCTURE.  Just enough of OpenCL C 1.2, written in C++17, for g++ to compile the
 * reference's kernels/ray_caster_kernel.cl FROM WHERE IT LIES under /root/reference and run it on the CPU
 * (oracle/Makefile target `ref`, output oracle/_ref/libref_kernel*.so).  Nothing here restates the kernel: it only
 * supplies what an OpenCL runtime would -- vector types with the swizzles the kernel uses, operators with OpenCL's
 * semantics (vector comparisons give -1 / 0, select() tests the MSB, ...), the built-ins and the image accessors.
 *
 * The built-ins whose results OpenCL leaves to the implementation are given the same IEEE binary32 definitions the
 * oracle pins (oracle/vr_oracle.h): normalize(v) = v / sqrtf(dot(v,v)) with dot summed x,y,z in order,
 * fast_length = sqrtf(dot), max/min = the spec's comparison form, read_imagef = UNORM8 / 255.0f with clamp-to-edge,
 * write_imagef = saturate + round-to-nearest-even.  Everything else -- control flow, arithmetic order, constants --
 * is the reference's own source text.
 *
 * The one thing C++ cannot parse is the vector literal `(float3)(a, b, c)` (in C++ a cast of a comma expression);
 * the build recipe rewrites the token sequence `(typeN)(` to `typeN(` with sed on the fly.  Plain `(a, b, c)` without
 * a cast -- which the kernel also contains -- IS a comma expression in OpenCL C too and is left alone.
 */
#ifndef VR_CL_SHIM_H
#define VR_CL_SHIM_H

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <sys/types.h>      /* uint, ulong */
#include <type_traits>

typedef unsigned char uchar;

/* address-space and access qualifiers */
#define __kernel
#define __global
#define global
#define __constant
#define constant
#define __read_only
#define __write_only

/* ---- swizzle proxies: members of a union with the components, so `v.yzx` is plain member access ----------- */
template <class V, class T, int A, int B, int C>
struct vr_swz3 {
    T c[4];
    operator V() const { return V(c[A], c[B], c[C]); }
    vr_swz3 &operator=(const V &o) { T a = o.x, b = o.y, d = o.z; c[A] = a; c[B] = b; c[C] = d; return *this; }
    vr_swz3 &operator+=(const V &o) { return *this = V(*this) + o; }
    vr_swz3 &operator-=(const V &o) { return *this = V(*this) - o; }
    vr_swz3 &operator*=(const V &o) { return *this = V(*this) * o; }
};
template <class V, class T, int A, int B>
struct vr_swz2 {
    T c[4];
    operator V() const { return V(c[A], c[B]); }
};

struct int2;
struct int3;
struct float2;
struct float4;

struct float2 {
    float x, y;
    float2() {}
    float2(float s) : x(s), y(s) {}
    float2(float a, float b) : x(a), y(b) {}
};

struct alignas(16) float3 {
    union {
        struct { float x, y, z, pad_; };
        vr_swz3<float3, float, 0, 1, 2> xyz;
        vr_swz3<float3, float, 1, 2, 0> yzx;
        vr_swz3<float3, float, 2, 0, 1> zxy;
        vr_swz2<float2, float, 0, 1> xy;
        vr_swz2<float2, float, 0, 2> xz;
        vr_swz2<float2, float, 1, 2> yz;
    };
    float3() {}
    float3(float s) : x(s), y(s), z(s), pad_(0) {}
    float3(float a, float b, float c) : x(a), y(b), z(c), pad_(0) {}
    float3(const float3 &o) : x(o.x), y(o.y), z(o.z), pad_(0) {}
    float3 &operator=(const float3 &o) { x = o.x; y = o.y; z = o.z; return *this; }
};

struct alignas(16) float4 {
    union {
        struct { float x, y, z, w; };
        vr_swz3<float3, float, 0, 1, 2> xyz;
    };
    float4() {}
    float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(const float4 &o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    float4 &operator=(const float4 &o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    explicit operator float3() const { return float3(x, y, z); }
};

struct int2 {
    int x, y;
    int2() {}
    int2(int s) : x(s), y(s) {}
    int2(int a, int b) : x(a), y(b) {}
};

struct alignas(16) int3 {
    union {
        struct { int x, y, z, pad_; };
        vr_swz3<int3, int, 0, 1, 2> xyz;
    };
    int3() {}
    int3(int s) : x(s), y(s), z(s), pad_(0) {}
    int3(int a, int b, int c) : x(a), y(b), z(c), pad_(0) {}
    int3(const int3 &o) : x(o.x), y(o.y), z(o.z), pad_(0) {}
    int3 &operator=(const int3 &o) { x = o.x; y = o.y; z = o.z; return *this; }
};

struct int4 {
    int x, y, z, w;
    int4() {}
    int4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
};

template <class T>
struct vr_small3 {
    T x, y, z, pad_;
    vr_small3() {}
    vr_small3(int s) : x((T)s), y((T)s), z((T)s), pad_(0) {}
    vr_small3(int a, int b, int c) : x((T)a), y((T)b), z((T)c), pad_(0) {}
};
typedef vr_small3<uchar> uchar3;
typedef vr_small3<signed char> char3;
struct uint3 { uint x, y, z, pad_; };

/* ---- float3 / float4 / float2 arithmetic (component-wise, scalar operands widened) -------------------------- */
#define VR_F3_OP(op)                                                                                                    \
    inline float3 operator op(const float3 &a, const float3 &b) { return float3(a.x op b.x, a.y op b.y, a.z op b.z); }   \
    inline float3 operator op(const float3 &a, float b) { return float3(a.x op b, a.y op b, a.z op b); }                 \
    inline float3 operator op(float a, const float3 &b) { return float3(a op b.x, a op b.y, a op b.z); }                 \
    inline float3 &operator op##=(float3 &a, const float3 &b) { a = a op b; return a; }                                  \
    inline float3 &operator op##=(float3 &a, float b) { a = a op b; return a; }
VR_F3_OP(+) VR_F3_OP(-) VR_F3_OP(*) VR_F3_OP(/)
inline float3 operator-(const float3 &a) { return float3(-a.x, -a.y, -a.z); }

#define VR_F4_OP(op)                                                                                                              \
    inline float4 operator op(const float4 &a, const float4 &b) { return float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    inline float4 operator op(const float4 &a, float b) { return float4(a.x op b, a.y op b, a.z op b, a.w op b); }                 \
    inline float4 operator op(float a, const float4 &b) { return float4(a op b.x, a op b.y, a op b.z, a op b.w); }                 \
    inline float4 &operator op##=(float4 &a, const float4 &b) { a = a op b; return a; }
VR_F4_OP(+) VR_F4_OP(-) VR_F4_OP(*) VR_F4_OP(/)

#define VR_F2_OP(op)                                                                                    \
    inline float2 operator op(const float2 &a, const float2 &b) { return float2(a.x op b.x, a.y op b.y); } \
    inline float2 operator op(const float2 &a, float b) { return float2(a.x op b, a.y op b); }
VR_F2_OP(+) VR_F2_OP(-) VR_F2_OP(*) VR_F2_OP(/)

/* ---- int3 / int2 arithmetic --------------------------------------------------------------------------------- */
#define VR_I3_OP(op)                                                                                              \
    inline int3 operator op(const int3 &a, const int3 &b) { return int3(a.x op b.x, a.y op b.y, a.z op b.z); }     \
    inline int3 operator op(const int3 &a, int b) { return int3(a.x op b, a.y op b, a.z op b); }                   \
    inline int3 operator op(int a, const int3 &b) { return int3(a op b.x, a op b.y, a op b.z); }                   \
    inline int3 &operator op##=(int3 &a, const int3 &b) { a = a op b; return a; }
VR_I3_OP(+) VR_I3_OP(-) VR_I3_OP(*) VR_I3_OP(/)
inline int3 operator-(const int3 &a) { return int3(-a.x, -a.y, -a.z); }
inline int2 operator+(const int2 &a, const int2 &b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator/(const int2 &a, const int2 &b) { return int2(a.x / b.x, a.y / b.y); }

/* ---- relational operators: vectors give -1 (true) / 0 (false) per component (OpenCL C 6.3.d/e) -------------- */
#define VR_REL(op)                                                                                                                     \
    inline int3 operator op(const float3 &a, const float3 &b) { return int3(-(int)(a.x op b.x), -(int)(a.y op b.y), -(int)(a.z op b.z)); } \
    inline int3 operator op(const float3 &a, float b) { return int3(-(int)(a.x op b), -(int)(a.y op b), -(int)(a.z op b)); }               \
    inline int3 operator op(const int3 &a, const int3 &b) { return int3(-(int)(a.x op b.x), -(int)(a.y op b.y), -(int)(a.z op b.z)); }     \
    inline int3 operator op(const int3 &a, int b) { return int3(-(int)(a.x op b), -(int)(a.y op b), -(int)(a.z op b)); }
VR_REL(==) VR_REL(!=) VR_REL(<) VR_REL(>) VR_REL(<=) VR_REL(>=)
inline int3 isless(const float3 &a, const float3 &b) { return a < b; }
inline int3 isless(const float3 &a, float b) { return a < b; }
inline int any(const int3 &v) { return (v.x < 0) || (v.y < 0) || (v.z < 0); }       /* MSB of any component */
inline int all(const int3 &v) { return (v.x < 0) && (v.y < 0) && (v.z < 0); }

/* ---- select: vectors take b where the MSB of c is set; scalars where c != 0 (OpenCL C 6.12.6) ---------------- */
template <class T, class C>
inline vr_small3<T> select(const vr_small3<T> &a, const vr_small3<T> &b, const vr_small3<C> &c) {
    vr_small3<T> r;
    r.x = (c.x & 0x80) ? b.x : a.x; r.y = (c.y & 0x80) ? b.y : a.y; r.z = (c.z & 0x80) ? b.z : a.z; r.pad_ = 0;
    return r;
}
inline int3 select(const int3 &a, const int3 &b, const int3 &c) { return int3(c.x < 0 ? b.x : a.x, c.y < 0 ? b.y : a.y, c.z < 0 ? b.z : a.z); }
template <class A, class B, class C, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C>::value>::type>
inline typename std::common_type<A, B>::type select(A a, B b, C c) { return c ? b : a; }

/* ---- conversions ------------------------------------------------------------------------------------------- */
inline int vr_cvt_int(float v) { return (v > -2147483648.0f && v < 2147483648.0f) ? (int)v : 0; }   /* rtz; out of range pinned to 0 like the oracle */
inline float3 convert_float3(const int3 &v) { return float3((float)v.x, (float)v.y, (float)v.z); }
inline float2 convert_float2(const int2 &v) { return float2((float)v.x, (float)v.y); }
inline int3 convert_int3(const int3 &v) { return v; }
inline int3 convert_int3(const float3 &v) { return int3(vr_cvt_int(v.x), vr_cvt_int(v.y), vr_cvt_int(v.z)); }
inline int3 convert_int3_rtn(const float3 &v) { return int3((int)floorf(v.x), (int)floorf(v.y), (int)floorf(v.z)); }
inline int2 convert_int2(const float2 &v) { return int2(vr_cvt_int(v.x), vr_cvt_int(v.y)); }
inline char3 convert_char3(const int3 &v) { return char3(v.x, v.y, v.z); }
inline uchar3 convert_uchar3(const int3 &v) { return uchar3(v.x, v.y, v.z); }
inline uint3 convert_uint3(const float3 &v) { uint3 r = {(uint)v.x, (uint)v.y, (uint)v.z, 0}; return r; }

/* ---- math built-ins (definitions pinned as in oracle/vr_oracle.h) ------------------------------------------ */
inline float max(float x, float y) { return (x < y) ? y : x; }
inline float min(float x, float y) { return (y < x) ? y : x; }
inline int max(int x, int y) { return (x < y) ? y : x; }
inline int min(int x, int y) { return (y < x) ? y : x; }
inline float3 min(const float3 &a, const float3 &b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(const float3 &a, const float3 &b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float3 fabs(const float3 &v) { return float3(fabsf(v.x), fabsf(v.y), fabsf(v.z)); }
inline float3 floor(const float3 &v) { return float3(floorf(v.x), floorf(v.y), floorf(v.z)); }
inline int3 abs(const int3 &v) { return int3(v.x < 0 ? -v.x : v.x, v.y < 0 ? -v.y : v.y, v.z < 0 ? -v.z : v.z); }
inline float dot(const float3 &a, const float3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float fast_length(const float3 &v) { return sqrtf(dot(v, v)); }
inline float fast_distance(const float3 &a, const float3 &b) { return fast_length(a - b); }
inline float3 normalize(const float3 &v) { const float l = sqrtf(dot(v, v)); return float3(v.x / l, v.y / l, v.z / l); }
inline float4 mix(const float4 &x, const float4 &y, float a) { return x + (y - x) * a; }
inline float4 mix(const float4 &x, float y, float a) { return x + (float4(y) - x) * a; }
inline int popcount(int v) { return __builtin_popcount((unsigned)v); }

/* ---- images: RGBA8, sampler-less integer coordinates -------------------------------------------------------- */
struct vr_image { int width, height; uint8_t *data; uint8_t *written; };
typedef vr_image *image2d_t;
inline float4 read_imagef(image2d_t im, const int2 &p) {
    const int cx = p.x < 0 ? 0 : (p.x > im->width - 1 ? im->width - 1 : p.x);
    const int cy = p.y < 0 ? 0 : (p.y > im->height - 1 ? im->height - 1 : p.y);
    const uint8_t *t = im->data + 4 * ((size_t)cx + (size_t)im->width * (size_t)cy);
    return float4((float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f, (float)t[3] / 255.0f);
}
inline uint8_t vr_unorm8(float c) {
    float v = c * 255.0f;
    if (!(v > 0.0f)) return 0;
    if (v > 255.0f) v = 255.0f;
    return (uint8_t)nearbyintf(v);
}
inline void write_imagef(image2d_t im, const int2 &p, const float4 &c) {
    if (p.x < 0 || p.y < 0 || p.x >= im->width || p.y >= im->height) return;
    uint8_t *o = im->data + 4 * ((size_t)p.x + (size_t)im->width * (size_t)p.y);
    o[0] = vr_unorm8(c.x); o[1] = vr_unorm8(c.y); o[2] = vr_unorm8(c.z); o[3] = vr_unorm8(c.w);
    if (im->written) im->written[(size_t)p.x + (size_t)im->width * (size_t)p.y] = 1;
}

/* ---- work-item functions ------------------------------------------------------------------------------------ */
extern thread_local int vr_global_id[2];
inline int get_global_id(int d) { return vr_global_id[d]; }

#endif
