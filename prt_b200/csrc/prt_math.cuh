// prt_math.cuh -- pinned float32 arithmetic ("PRT-ARITH v1") for the sm_100a kernels.
//
// Everything that decides a ray/triangle hit, builds a sample direction or advances a bounce uses the
// explicitly rounded intrinsics below, so nvcc can neither contract nor reassociate it and the results
// are bit-identical to the CPU oracle's statement of the same formulas.  The file also compiles as plain
// C++ (-ffp-contract=off) for the host-side BVH self-check used by the CPU test-suite; that harness is
// test tooling, never part of libprt_b200.so.
//
// Follows reference src/raytracing/raytracing.cpp:101-107 (frame), :130-160 (sampling), util.h:6 (PI).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PRT_HD __host__ __device__ __forceinline__
#else
#define PRT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PRT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define PRT_MUL(a, b) __fmul_rn((a), (b))
#define PRT_ADD(a, b) __fadd_rn((a), (b))
#define PRT_SUB(a, b) __fsub_rn((a), (b))
#define PRT_DIV(a, b) __fdiv_rn((a), (b))
#define PRT_SQRT(a) __fsqrt_rn((a))
#define PRT_F2U(f) __float_as_uint(f)
#define PRT_U2F(u) __uint_as_float(u)
#else
#define PRT_FMA(a, b, c) fmaf((a), (b), (c))
#define PRT_MUL(a, b) ((a) * (b))
#define PRT_ADD(a, b) ((a) + (b))
#define PRT_SUB(a, b) ((a) - (b))
#define PRT_DIV(a, b) ((a) / (b))
#define PRT_SQRT(a) sqrtf((a))
static inline uint32_t prt_f2u_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float prt_u2f_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define PRT_F2U(f) prt_f2u_host(f)
#define PRT_U2F(u) prt_u2f_host(u)
#endif

namespace prt {

constexpr float kPiF = 3.14159265359f;  // reference util.h:6

struct f3 { float x, y, z; };
PRT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
PRT_HD f3 sub3(f3 a, f3 b) { return mk3(PRT_SUB(a.x, b.x), PRT_SUB(a.y, b.y), PRT_SUB(a.z, b.z)); }
PRT_HD f3 scale3(f3 a, float s) { return mk3(PRT_MUL(a.x, s), PRT_MUL(a.y, s), PRT_MUL(a.z, s)); }
PRT_HD f3 cross3(f3 a, f3 b) {
    return mk3(PRT_FMA(a.y, b.z, -PRT_MUL(a.z, b.y)), PRT_FMA(a.z, b.x, -PRT_MUL(a.x, b.z)), PRT_FMA(a.x, b.y, -PRT_MUL(a.y, b.x)));
}
PRT_HD float dot3(f3 a, f3 b) { return PRT_FMA(a.z, b.z, PRT_FMA(a.y, b.y, PRT_MUL(a.x, b.x))); }
PRT_HD f3 normalize3(f3 a) { return scale3(a, PRT_DIV(1.0f, PRT_SQRT(dot3(a, a)))); }
PRT_HD f3 madd3(f3 a, float s, f3 b) { return mk3(PRT_FMA(s, b.x, a.x), PRT_FMA(s, b.y, a.y), PRT_FMA(s, b.z, a.z)); }

// sin/cos(2*pi*v), v in [0,1]: exact quadrant reduction + fixed Taylor/Horner polynomials.
PRT_HD void sincos2pi(float v, float &so, float &co) {
    float x4 = PRT_MUL(v, 4.0f);
    float kf = rintf(x4);
    float r = PRT_SUB(x4, kf);
    float a = PRT_MUL(r, 1.57079632679489661923f);
    float a2 = PRT_MUL(a, a);
    float sp = PRT_FMA(a2, 2.7557319224e-6f, -1.9841269841e-4f);
    sp = PRT_FMA(a2, sp, 8.3333333333e-3f);
    sp = PRT_FMA(a2, sp, -1.6666666667e-1f);
    sp = PRT_FMA(a2, sp, 1.0f);
    float s = PRT_MUL(a, sp);
    float cp = PRT_FMA(a2, -2.7557319224e-7f, 2.4801587302e-5f);
    cp = PRT_FMA(a2, cp, -1.3888888889e-3f);
    cp = PRT_FMA(a2, cp, 4.1666666667e-2f);
    cp = PRT_FMA(a2, cp, -0.5f);
    float c = PRT_FMA(a2, cp, 1.0f);
    int k = ((int)kf) & 3;
    so = k == 0 ? s : (k == 1 ? c : (k == 2 ? -s : -c));
    co = k == 0 ? c : (k == 1 ? -s : (k == 2 ? -c : s));
}

// raytracing.cpp:130-146
PRT_HD f3 cosine_local(float u, float v) {
    float r = PRT_SQRT(u), s, c;
    sincos2pi(v, s, c);
    float x = PRT_MUL(r, c), y = PRT_MUL(r, s);
    float z = PRT_SQRT(fmaxf(0.0f, PRT_FMA(-y, y, PRT_FMA(-x, x, 1.0f))));
    return mk3(x, y, z);
}

struct Frame { f3 right, up, n; };
// raytracing.cpp:101-107 (fabsf: SURVEY section 7 notes the reference's unqualified abs)
PRT_HD Frame make_frame(f3 N) {
    Frame f;
    f3 up0 = fabsf(N.z) < 0.99f ? mk3(0.f, 0.f, 1.f) : mk3(1.f, 0.f, 0.f);
    f.right = normalize3(cross3(up0, N));
    f.up = cross3(N, f.right);
    f.n = N;
    return f;
}
// Packed variants (sm_100a, fma.rn.f32x2 / mul.rn.f32x2): component-wise the SAME IEEE operations in the same order, so the results
// are bit-identical to the scalar forms -- only the number of issue slots changes.  The host build keeps the scalar forms.
// Measured on the headline bake (profiles/r2_packed_math_ab.jsonl): 44.87 -> 44.50 ms, with the FMNMX clamp of rcp_box 44.39 ms.
#ifndef PRT_PACKED_MATH
#define PRT_PACKED_MATH 1
#endif
struct fpair { float x, y; };
#if defined(__CUDA_ARCH__) && PRT_PACKED_MATH
__device__ __forceinline__ f3 to_world(const Frame &f, f3 l) {
    float2 t = __fmul2_rn(make_float2(f.right.x, f.right.y), make_float2(l.x, l.x));
    t = __ffma2_rn(make_float2(f.up.x, f.up.y), make_float2(l.y, l.y), t);
    t = __ffma2_rn(make_float2(f.n.x, f.n.y), make_float2(l.z, l.z), t);
    return mk3(t.x, t.y, PRT_FMA(f.n.z, l.z, PRT_FMA(f.up.z, l.y, PRT_MUL(f.right.z, l.x))));
}
// (dot3(a, c), dot3(b, c))
__device__ __forceinline__ fpair dot3_pair(f3 a, f3 b, f3 c) {
    float2 t = __fmul2_rn(make_float2(a.x, b.x), make_float2(c.x, c.x));
    t = __ffma2_rn(make_float2(a.y, b.y), make_float2(c.y, c.y), t);
    t = __ffma2_rn(make_float2(a.z, b.z), make_float2(c.z, c.z), t);
    fpair r; r.x = t.x; r.y = t.y; return r;
}
#else
PRT_HD f3 to_world(const Frame &f, f3 l) {
    return mk3(PRT_FMA(f.n.x, l.z, PRT_FMA(f.up.x, l.y, PRT_MUL(f.right.x, l.x))),
               PRT_FMA(f.n.y, l.z, PRT_FMA(f.up.y, l.y, PRT_MUL(f.right.y, l.x))),
               PRT_FMA(f.n.z, l.z, PRT_FMA(f.up.z, l.y, PRT_MUL(f.right.z, l.x))));
}
PRT_HD fpair dot3_pair(f3 a, f3 b, f3 c) { fpair r; r.x = dot3(a, c); r.y = dot3(b, c); return r; }
#endif

// Philox4x32-10 (Random123 constants); integer only.
PRT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
constexpr uint32_t kPhiloxKey1 = 0x50525421u;
PRT_HD float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
PRT_HD void rand2(uint32_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t stream, float &u, float &v) {
    uint32_t o[4];
    philox4x32_10(a, b, c, stream, seed, kPhiloxKey1, o);
    u = u01(o[0]); v = u01(o[1]);
}

// Real SH basis, bands 0..ORDER-1, sh-space argument (x,y,z) = world (z,x,y) (raytracing.cpp:226).
// l<=2: reference src/sh/SH_function.h:7-41 (no Condon-Shortley sign); sgn = -1 reproduces the
// google/spherical-harmonics sign on odd |m|.
template <int ORDER>
PRT_HD void sh_eval(float x, float y, float z, float sgn, float *out) {
    out[0] = 0.282095f;
    if (ORDER >= 2) {
        out[1] = sgn * 0.488603f * y; out[2] = 0.488603f * z; out[3] = sgn * 0.488603f * x;
    }
    if (ORDER >= 3) {
        float x2 = x * x, y2 = y * y, z2 = z * z;
        out[4] = 1.092548f * x * y; out[5] = sgn * 1.092548f * y * z; out[6] = 0.315392f * (3.0f * z2 - 1.0f);
        out[7] = sgn * 1.092548f * x * z; out[8] = 0.546274f * (x2 - y2);
        if (ORDER >= 4) {
            out[9] = sgn * 0.590044f * y * (3.0f * x2 - y2); out[10] = 2.890611f * x * y * z;
            out[11] = sgn * 0.457046f * y * (5.0f * z2 - 1.0f); out[12] = 0.373176f * z * (5.0f * z2 - 3.0f);
            out[13] = sgn * 0.457046f * x * (5.0f * z2 - 1.0f); out[14] = 1.445306f * z * (x2 - y2);
            out[15] = sgn * 0.590044f * x * (x2 - 3.0f * y2);
        }
        if (ORDER >= 5) {
            out[16] = 2.503343f * x * y * (x2 - y2); out[17] = sgn * 1.770131f * y * z * (3.0f * x2 - y2);
            out[18] = 0.946175f * x * y * (7.0f * z2 - 1.0f); out[19] = sgn * 0.669047f * y * z * (7.0f * z2 - 3.0f);
            out[20] = 0.105786f * (35.0f * z2 * z2 - 30.0f * z2 + 3.0f); out[21] = sgn * 0.669047f * x * z * (7.0f * z2 - 3.0f);
            out[22] = 0.473087f * (x2 - y2) * (7.0f * z2 - 1.0f); out[23] = sgn * 1.770131f * x * z * (x2 - 3.0f * y2);
            out[24] = 0.625836f * (x2 * (x2 - 3.0f * y2) - y2 * (3.0f * x2 - y2));
        }
    }
}

}  // namespace prt
