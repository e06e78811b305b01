// horizon_math.cuh -- per-item bounds of the per-origin horizon map (horizon.cu, entry_list.cuh): for one triangle or one box,
// a conservative upper bound of sin(elevation above the tangent plane) and the azimuth bins it can occupy.  Pure functions of
// their arguments, compiled for the device by the kernels and as plain C++ by tests/hostcheck, where the CPU test-suite checks
// them against brute-force sampling (tests/test_hostcheck.py) -- the warp-collective builder around them is in entry_list.cuh.
#pragma once
#include "prt_math.cuh"

#if defined(__CUDA_ARCH__)
#define PRT_FDIVIDEF(a, b) __fdividef((a), (b))
#define PRT_RSQRTF(a) rsqrtf((a))
#else
#define PRT_FDIVIDEF(a, b) ((a) / (b))
#define PRT_RSQRTF(a) (1.0f / sqrtf((a)))
#endif
#if defined(__CUDACC__)
#define PRT_HZ_NOINLINE static __host__ __device__ __noinline__
#else
#define PRT_HZ_NOINLINE static inline
#endif

namespace prt {

// azimuth bins of the horizon map: one bin per lane of the building warp.  The host bins the sample directions with the same
// constant (abi.cu, ensure_samples).
constexpr int kHzBins = 32;
constexpr float kHzPerUnit = (float)(kHzBins / 4);   // bins per unit of pseudo-angle
constexpr int kHzLevels = 5;                          // published map: level k = minima over 2^k consecutive bins (circular)
constexpr int kHzWords = kHzLevels * kHzBins;

struct HzItem { int b0, b1; float v; };          // bins b0..b1 (unwrapped, b1-b0 <= kHzBins-1), value; v <= 0: empty

PRT_HD float hz_pang(float x, float y) {
    const float p = PRT_FDIVIDEF(y, fabsf(x) + fabsf(y));
    return x < 0.f ? 2.f - p : (y < 0.f ? 4.f + p : p);
}
PRT_HD HzItem hz_item(float lo, float hi, bool all, float sinh) {
    HzItem it; it.b0 = 0; it.b1 = 0; it.v = 0.f;
    if (!(sinh > 0.0f)) return it;
    it.v = fminf(sinh + 2e-4f, 2.0f);
    it.b1 = kHzBins - 1;
    if (all || !(hi - lo < 3.9f)) return it;
    const int b0 = (int)floorf((lo - 0.02f) * kHzPerUnit), b1 = (int)floorf((hi + 0.02f) * kHzPerUnit);
    if (b1 - b0 >= kHzBins - 1) return it;
    it.b0 = b0; it.b1 = b1;
    return it;
}
// maximum of (v.z / |v|) over the segment a + t (b - a), t in [0,1]  (local frame: z = height above the tangent plane)
PRT_HD float hz_edge_max(const f3 a, const f3 b) {
    const float la2 = a.x * a.x + a.y * a.y + a.z * a.z, lb2 = b.x * b.x + b.y * b.y + b.z * b.z;
    float m = fmaxf(la2 > 0.f ? a.z * PRT_RSQRTF(la2) : 1.f, lb2 > 0.f ? b.z * PRT_RSQRTF(lb2) : 1.f);
    const f3 e = mk3(b.x - a.x, b.y - a.y, b.z - a.z);
    const float ae = a.x * e.x + a.y * e.y + a.z * e.z, ee = e.x * e.x + e.y * e.y + e.z * e.z;
    const float den = e.z * ae - a.z * ee;
    if (fabsf(den) > 0.f) {
        const float t = PRT_FDIVIDEF(a.z * ae - e.z * la2, den);
        if (t > 0.f && t < 1.f) {
            const f3 v = mk3(a.x + t * e.x, a.y + t * e.y, a.z + t * e.z);
            const float lv2 = v.x * v.x + v.y * v.y + v.z * v.z;
            m = fmaxf(m, lv2 > 0.f ? v.z * PRT_RSQRTF(lv2) : 1.f);
        }
    }
    return m;
}

PRT_HD HzItem hz_triangle(const f3 q0, const f3 q1, const f3 q2) {
    const float zmax = fmaxf(q0.z, fmaxf(q1.z, q2.z));
    const float scale = fmaxf(fmaxf(fabsf(q0.x) + fabsf(q0.y) + fabsf(q0.z), fabsf(q1.x) + fabsf(q1.y) + fabsf(q1.z)), fabsf(q2.x) + fabsf(q2.y) + fabsf(q2.z));
    if (zmax < -1e-5f * scale) return hz_item(0.f, 0.f, false, 0.f);       // wholly below the tangent plane
    float sinh = fmaxf(hz_edge_max(q0, q1), fmaxf(hz_edge_max(q1, q2), hz_edge_max(q2, q0)));
    // vertical axis through the triangle: (0,0,1) . (qi x qj) all of one sign (with tolerance)
    const float c01 = q0.x * q1.y - q0.y * q1.x, c12 = q1.x * q2.y - q1.y * q2.x, c20 = q2.x * q0.y - q2.y * q0.x;
    const float tol = 1e-6f * scale * scale;
    const bool surround = (c01 >= -tol && c12 >= -tol && c20 >= -tol) || (c01 <= tol && c12 <= tol && c20 <= tol);
    // azimuth range: minimal arc containing the pseudo-angles of the vertices that are not (numerically) on the vertical axis
    float ph[3]; int np = 0;
    const float rmin = 1e-4f * scale;
    if (fabsf(q0.x) + fabsf(q0.y) > rmin) ph[np++] = hz_pang(q0.x, q0.y);
    if (fabsf(q1.x) + fabsf(q1.y) > rmin) ph[np++] = hz_pang(q1.x, q1.y);
    if (fabsf(q2.x) + fabsf(q2.y) > rmin) ph[np++] = hz_pang(q2.x, q2.y);
    bool all = false;
    float lo = 0.f, hi = 0.f;
    if (np == 3 && surround) { all = true; if (zmax > 0.f) sinh = 1.0f; }
    else if (np == 0) all = true;
    else if (np == 1) { lo = hi = ph[0]; }
    else if (np == 2) { lo = fminf(ph[0], ph[1]); hi = fmaxf(ph[0], ph[1]); if (hi - lo > 2.f) { const float t = lo; lo = hi; hi = t + 4.f; } }
    else {
        float a = ph[0], b = ph[1], c = ph[2], t;
        if (a > b) { t = a; a = b; b = t; } if (b > c) { t = b; b = c; c = t; } if (a > b) { t = a; a = b; b = t; }
        const float g0 = b - a, g1 = c - b, g2 = a + 4.f - c;
        if (g2 >= g0 && g2 >= g1) { lo = a; hi = c; }
        else if (g0 >= g1) { lo = b; hi = a + 4.f; }
        else { lo = c; hi = b + 4.f; }
        if (hi - lo > 2.f) all = true;
    }
    return hz_item(lo, hi, all, sinh);
}

// exact bound of a box (centre c relative to the origin, half extents e): 12 edge maxima, or 1 if the vertical ray pierces it
PRT_HZ_NOINLINE HzItem hz_box(const f3 c, const f3 e, const Frame &fr) {
    if (!(e.x < 1e30f)) return hz_item(0.f, 0.f, true, 1.0f);               // overflow candidate: unbounded
    {
        float t0 = 0.f, t1 = 3.0e38f;
        const float cc[3] = {c.x, c.y, c.z}, ee[3] = {e.x, e.y, e.z}, nn3[3] = {fr.n.x, fr.n.y, fr.n.z};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float inv = 1.0f / (fabsf(nn3[a]) < 1e-12f ? copysignf(1e-12f, nn3[a]) : nn3[a]);
            const float ta = (cc[a] - ee[a]) * inv, tb = (cc[a] + ee[a]) * inv;
            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
        }
        if (t0 <= t1 * 1.0001f + 1e-6f) return hz_item(0.f, 0.f, true, 1.0f);
    }
    const f3 cl = mk3(c.x * fr.right.x + c.y * fr.right.y + c.z * fr.right.z, c.x * fr.up.x + c.y * fr.up.y + c.z * fr.up.z, c.x * fr.n.x + c.y * fr.n.y + c.z * fr.n.z);
    const f3 hx = mk3(e.x * fr.right.x, e.x * fr.up.x, e.x * fr.n.x), hy = mk3(e.y * fr.right.y, e.y * fr.up.y, e.y * fr.n.y), hzv = mk3(e.z * fr.right.z, e.z * fr.up.z, e.z * fr.n.z);
    f3 q[8];
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        const float sx = (i & 1) ? 1.f : -1.f, sy = (i & 2) ? 1.f : -1.f, sz = (i & 4) ? 1.f : -1.f;
        q[i] = mk3(cl.x + sx * hx.x + sy * hy.x + sz * hzv.x, cl.y + sx * hx.y + sy * hy.y + sz * hzv.y, cl.z + sx * hx.z + sy * hy.z + sz * hzv.z);
    }
    float sinh = -1.f;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        if (!(i & 1)) sinh = fmaxf(sinh, hz_edge_max(q[i], q[i | 1]));
        if (!(i & 2)) sinh = fmaxf(sinh, hz_edge_max(q[i], q[i | 2]));
        if (!(i & 4)) sinh = fmaxf(sinh, hz_edge_max(q[i], q[i | 4]));
    }
    const float rc = fabsf(cl.x) + fabsf(cl.y), sc = e.x + e.y + e.z + fabsf(c.x) + fabsf(c.y) + fabsf(c.z);
    bool all = !(rc > 1e-4f * sc);
    float lo = 0.f, hi = 0.f, pc = 0.f;
    if (!all) {
        pc = hz_pang(cl.x, cl.y);
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            if (!(fabsf(q[i].x) + fabsf(q[i].y) > 1e-4f * sc)) { all = true; continue; }
            float dl = hz_pang(q[i].x, q[i].y) - pc;
            if (dl > 2.f) dl -= 4.f;
            if (dl < -2.f) dl += 4.f;
            lo = fminf(lo, dl); hi = fmaxf(hi, dl);
        }
        if (hi - lo >= 1.98f) all = true;                                  // the box surrounds the vertical axis
    }
    if (pc + lo < 0.f) pc += 4.f;
    return hz_item(pc + lo, pc + hi, all, sinh);
}

// cone around the bounding sphere of a box (r^2 = |e|^2, d^2 = |c|^2 > r^2): cheap, slightly loose
PRT_HD HzItem hz_sphere(const f3 c, const float r2, const float d2, const Frame &fr) {
    const float id = PRT_RSQRTF(d2), sina = fminf(sqrtf(r2) * id * 1.0001f + 1e-6f, 1.0f), cosa = sqrtf(fmaxf(0.f, 1.f - sina * sina));
    const float ax = c.x * fr.right.x + c.y * fr.right.y + c.z * fr.right.z, ay = c.x * fr.up.x + c.y * fr.up.y + c.z * fr.up.z;
    const float sinb = fminf(1.f, fmaxf(-1.f, (c.x * fr.n.x + c.y * fr.n.y + c.z * fr.n.z) * id)), cosb = sqrtf(fmaxf(0.f, 1.f - sinb * sinb));
    float sinh = sinb * cosa + cosb * sina;
    if (sinb >= cosa) sinh = 1.0f;                   // elevation + half angle >= 90 degrees
    if (sina >= 0.98f * cosb || !(fabsf(ax) + fabsf(ay) > 0.f)) return hz_item(0.f, 0.f, true, sinh);
    // asin(x) <= x (1 + 0.58 x^2) on [0,1]; d(pseudo-angle) <= d(angle), so the true half width bounds the pseudo one
    const float x = PRT_FDIVIDEF(sina, cosb), dphi = x * (1.f + 0.58f * x * x) + 1e-3f;
    float p = hz_pang(ax, ay);
    if (p - dphi < 0.f) p += 4.f;
    return hz_item(p - dphi, p + dphi, false, sinh);
}

// Cheap bound of a box (centre c relative to the origin, half extents e; r2 = |e|^2, d2 = |c|^2): the cone around its bounding
// sphere (needs the origin outside the sphere), clamped by z_max / d_min over the box -- the largest height above the tangent
// plane over the smallest distance from the origin (d_min > 0: the origin is outside the box).
PRT_HD HzItem hz_cheap_box(const f3 c, const f3 e, const float r2, const float d2, const Frame &fr) {
    HzItem cb = d2 > 1.05f * r2 ? hz_sphere(c, r2, d2, fr) : hz_item(0.f, 0.f, true, 1.0f);
    if (cb.v > 0.f) {
        const float zt = c.x * fr.n.x + c.y * fr.n.y + c.z * fr.n.z + fabsf(fr.n.x) * e.x + fabsf(fr.n.y) * e.y + fabsf(fr.n.z) * e.z;
        const float mx = fmaxf(fabsf(c.x) - e.x, 0.f), my_ = fmaxf(fabsf(c.y) - e.y, 0.f), mz = fmaxf(fabsf(c.z) - e.z, 0.f);
        const float dm2 = mx * mx + my_ * my_ + mz * mz;
        if (dm2 > 0.f) cb.v = fminf(cb.v, fmaxf(zt, 0.f) * PRT_RSQRTF(dm2) * 1.0001f + (1e-6f + 2e-4f));
    }
    return cb;
}

// Bound of a box CUT BY THE ORIENTED SLAB of its node (bvh8.h, Slab32): the geometry lies in {x : |x - c| <= e} and between the planes
// L0 <= m . x <= U0 (x relative to the origin: L0 = d0 - m . O, U0 = d1 - m . O).  The height above the tangent plane n . x = n . c + n . y,
// |y| <= e, L <= m . y <= U (L = L0 - m . c, U = U0 - m . c), is a linear programme with one two-sided constraint; by weak duality
//     n . y = (n - lambda m) . y + lambda (m . y) <= sum_i |n_i - lambda m_i| e_i + (lambda >= 0 ? lambda U : lambda L)      for EVERY lambda,
// so any choice of lambda gives a valid bound (lambda = 0 is the plain box).  The dual function is convex and piecewise linear with
// breakpoints n_i / m_i, hence its minimum -- the exact optimum of the programme -- is attained at one of them or at 0; n . m (the
// projection) is tried as well: it is the minimiser for a cube.  The distance bound uses the slab too (the origin may lie outside it).
// Returns the sin(elevation) bound in the form of hz_cheap_box (same margins); the caller keeps the azimuth range of the box.
PRT_HD float hz_slab_value(const f3 c, const f3 e, const f3 n, const f3 m, const float L0, const float U0) {
    const float mc = m.x * c.x + m.y * c.y + m.z * c.z;
    // rounding of L0, U0 (a difference of two dot products of scene-scale numbers) and of mc is covered by the padding of d0 / d1
    const float L = L0 - mc, U = U0 - mc;
    float best = fabsf(n.x) * e.x + fabsf(n.y) * e.y + fabsf(n.z) * e.z;
    const float nm = n.x * m.x + n.y * m.y + n.z * m.z;
    const float lam[4] = {nm, fabsf(m.x) > 1e-6f ? PRT_FDIVIDEF(n.x, m.x) : nm, fabsf(m.y) > 1e-6f ? PRT_FDIVIDEF(n.y, m.y) : nm,
                          fabsf(m.z) > 1e-6f ? PRT_FDIVIDEF(n.z, m.z) : nm};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float l = lam[k];
        const float f = fabsf(n.x - l * m.x) * e.x + fabsf(n.y - l * m.y) * e.y + fabsf(n.z - l * m.z) * e.z + l * (l >= 0.f ? U : L);
        best = fminf(best, f);
    }
    const float zt = n.x * c.x + n.y * c.y + n.z * c.z + best;
    const float mx = fmaxf(fabsf(c.x) - e.x, 0.f), my_ = fmaxf(fabsf(c.y) - e.y, 0.f), mz = fmaxf(fabsf(c.z) - e.z, 0.f);
    const float gap = fmaxf(fmaxf(L0, -U0), 0.f);                  // distance from the origin to the slab (|m| = 1 up to rounding)
    const float dm2 = fmaxf(mx * mx + my_ * my_ + mz * mz, gap * gap * 0.999f);
    if (!(dm2 > 0.f)) return 2.0f;
    return fmaxf(zt, 0.f) * PRT_RSQRTF(dm2) * 1.0001f + (1e-6f + 2e-4f);
}

}  // namespace prt
