/*
 * oracle/arith.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product library).
 *
 * Pinned float32 arithmetic ("PRT-ARITH v1", stated in DESIGN.md section 3) for the
 * parts of the path whose results must agree bit-for-bit between this CPU oracle and the
 * sm_100a kernels: sampling (reference src/raytracing/raytracing.cpp:101-107,130-160),
 * the ray/triangle test that stands in for Embree's rtcIntersect1/rtcOccluded1
 * (raytracing.cpp:167,200,254; light_probe.cpp:119,128) and the bounce step
 * (raytracing.cpp:263-275).
 *
 * Rules: every operation is an IEEE-754 binary32 round-to-nearest op; fmaf() is a single
 * rounding; nothing else may be contracted (compile with -ffp-contract=off).
 * Division and sqrt are correctly rounded.
 */
#ifndef PRT_ORACLE_ARITH_H
#define PRT_ORACLE_ARITH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

/* reference util.h:6 : constexpr float PI = 3.14159265359 */
#define PRT_PI_F 3.14159265359f

static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }

/* cross(a,b).x = fma(a.y, b.z, -(a.z*b.y)) etc. */
static inline v3 v3_cross(v3 a, v3 b) {
    v3 r;
    r.x = fmaf(a.y, b.z, -(a.z * b.y));
    r.y = fmaf(a.z, b.x, -(a.x * b.z));
    r.z = fmaf(a.x, b.y, -(a.y * b.x));
    return r;
}
/* dot(a,b) = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x)) */
static inline float v3_dot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

/* glm::normalize = v * inversesqrt(dot(v,v)); pinned as v * (1/sqrt(dot)) */
static inline v3 v3_normalize(v3 a) {
    float inv = 1.0f / sqrtf(v3_dot(a, a));
    return v3_scale(a, inv);
}
/* a + s*b, one fma per component */
static inline v3 v3_madd(v3 a, float s, v3 b) {
    return v3_make(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z));
}

/*
 * sin/cos of 2*pi*v for v in [0,1], pinned polynomial (libm and CUDA sinf/cosf differ in
 * the last ulp, so neither may be used on a bit-exact path).  Quadrant reduction is exact:
 * k = rint(4v), r = 4v - k in [-1/2,1/2], a = r*(pi/2); Taylor polynomials in Horner/fma form.
 * Replaces std::cos/std::sin(2*PI*v) of raytracing.cpp:131-133 (deviation <= 6e-7 in angle,
 * far below the reference's own float-PI rounding of the angle).
 */
static inline void prt_sincos2pi(float v, float *s_out, float *c_out) {
    float x4 = v * 4.0f;
    float kf = rintf(x4);
    float r = x4 - kf;
    float a = r * 1.57079632679489661923f;
    float a2 = a * a;
    float sp = fmaf(a2, 2.7557319224e-6f, -1.9841269841e-4f);
    sp = fmaf(a2, sp, 8.3333333333e-3f);
    sp = fmaf(a2, sp, -1.6666666667e-1f);
    sp = fmaf(a2, sp, 1.0f);
    float s = a * sp;
    float cp = fmaf(a2, -2.7557319224e-7f, 2.4801587302e-5f);
    cp = fmaf(a2, cp, -1.3888888889e-3f);
    cp = fmaf(a2, cp, 4.1666666667e-2f);
    cp = fmaf(a2, cp, -0.5f);
    float c = fmaf(a2, cp, 1.0f);
    int k = ((int)kf) & 3;
    float so, co;
    switch (k) {
    case 0: so = s; co = c; break;
    case 1: so = c; co = -s; break;
    case 2: so = -s; co = -c; break;
    default: so = -c; co = s; break;
    }
    *s_out = so; *c_out = co;
}

/* raytracing.cpp:130-146: disk sample -> cosine-weighted local direction (+z up) */
static inline v3 prt_cosine_local(float u, float v) {
    float r = sqrtf(u);
    float s, c;
    prt_sincos2pi(v, &s, &c);
    float x = r * c, y = r * s;
    float z = sqrtf(fmaxf(0.0f, fmaf(-y, y, fmaf(-x, x, 1.0f))));
    return v3_make(x, y, z);
}

/* raytracing.cpp:101-107 / 154-157: tangent frame about N (fabsf) */
typedef struct { v3 right, up, n; } frame3;
static inline frame3 prt_frame(v3 N) {
    frame3 f;
    v3 up0 = fabsf(N.z) < 0.99f ? v3_make(0.f, 0.f, 1.f) : v3_make(1.f, 0.f, 0.f);
    f.right = v3_normalize(v3_cross(up0, N));
    f.up = v3_cross(N, f.right);
    f.n = N;
    return f;
}
/* glm mat3{right,up,N} * l */
static inline v3 prt_to_world(const frame3 *f, v3 l) {
    v3 r;
    r.x = fmaf(f->n.x, l.z, fmaf(f->up.x, l.y, f->right.x * l.x));
    r.y = fmaf(f->n.y, l.z, fmaf(f->up.y, l.y, f->right.y * l.x));
    r.z = fmaf(f->n.z, l.z, fmaf(f->up.z, l.y, f->right.z * l.x));
    return r;
}

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/*
 * Pinned ray/triangle test (Moeller-Trumbore, sign-folded like Embree's so that no division
 * is needed for the decision).  Triangle is stored as v0, e1=v1-v0, e2=v2-v0.
 * Returns 1 when the ray (O + t*D, tnear < t <= tfar) hits; *t_out = Ts/|det| (IEEE division).
 * Both faces hit (Embree default: no back-face culling); det == 0 never hits.
 */
static inline int prt_tri_test(v3 O, v3 D, float tnear, float tfar, v3 v0, v3 e1, v3 e2, float *t_out) {
    v3 tv = v3_sub(O, v0);
    v3 pv = v3_cross(D, e2);
    float det = v3_dot(e1, pv);
    float U = v3_dot(tv, pv);
    v3 qv = v3_cross(tv, e1);
    float V = v3_dot(D, qv);
    float T = v3_dot(e2, qv);
    uint32_t sgn = f2u(det) & 0x80000000u;
    float ad = fabsf(det);
    float Us = u2f(f2u(U) ^ sgn), Vs = u2f(f2u(V) ^ sgn), Ts = u2f(f2u(T) ^ sgn);
    if (!(ad > 0.0f)) return 0;
    if (!(Us >= 0.0f) || !(Vs >= 0.0f) || !(Us + Vs <= ad)) return 0;
    if (!(Ts > tnear * ad) || !(Ts <= tfar * ad)) return 0;
    if (t_out) *t_out = Ts / ad;
    return 1;
}

#endif
