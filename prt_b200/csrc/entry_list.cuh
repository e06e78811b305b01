// entry_list.cuh -- per-origin entry lists (device only), shared by the bake kernels.
#pragma once
#include "traverse.cuh"
#include <cuda_runtime.h>

namespace prt {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- per-origin entry list ---------------------------------------------------------------------------------------
// All S rays of a vertex share one origin, so every one of them would repeat the same descent through the chain of
// nodes whose boxes contain that origin.  Once per vertex the warp flattens that chain: it expands (breadth of 8
// lanes = 8 children) every node containing the origin and records each remaining child -- subtree root or leaf --
// that is not wholly below the tangent plane as a *candidate* box in shared memory (centre/half-extent relative to
// the origin).  Each ray then tests the candidate list in lockstep (no divergence, no node decode) and traverses only
// the subtrees whose candidate box it hits.  Decisions are still made by the pinned triangle test, so results are
// unchanged; the list only removes work.
constexpr int kMaxCand = 96;
struct EntryList {
    float4 ca[kMaxCand];       // centre - origin (xyz), half extent x
    float4 cb[kMaxCand];       // half extent y, z, group x, group y (see Trav::start_group)
    uint32_t queue[kMaxCand];
};

// sphere_expand: also expand children whose *bounding sphere* contains the origin (needed by the horizon map, whose cone bound
// for a subtree candidate requires the origin to be outside that sphere)
__device__ __forceinline__ int build_entry_list(const Node8 *nodes, const f3 O, const f3 N, EntryList &W, const int lane, const bool sphere_expand = false) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int n = 0, qn = 1;
    if (lane == 0) W.queue[0] = 0u;
    __syncwarp();
    while (qn > 0) {
        const uint32_t x = W.queue[qn - 1];
        qn--;
        __syncwarp();
        if (n + qn + 8 > kMaxCand) {
            // out of room: keep the node itself as an always-hit candidate (graceful fallback to plain traversal)
            if (lane == 0) {
                W.ca[n] = make_float4(0.f, 0.f, 0.f, INFINITY);
                W.cb[n] = make_float4(INFINITY, INFINITY, __uint_as_float(x), __uint_as_float(0x80000000u));
            }
            n++;
            continue;
        }
        const char *np = reinterpret_cast<const char *>(nodes + x);
        const u4 n0 = ld16(np), n1 = ld16(np + 16), n2 = ld16(np + 32), n3 = ld16(np + 48), n4 = ld16(np + 64);
        bool keep = false, expand = false;
        float cx = 0.f, cy = 0.f, cz = 0.f, ex = 0.f, ey = 0.f, ez = 0.f;
        uint32_t gx = 0u, gy = 0u;
        if (lane < 8) {
            const int h = lane >> 2, sh = 8 * (lane & 3);
            const uint32_t meta = ((h ? n1.w : n1.z) >> sh) & 0xFFu;
            if (meta) {
                const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
                            sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
                const float lox = __uint_as_float(n0.x) + (float)(((h ? n2.y : n2.x) >> sh) & 0xFFu) * sx;
                const float loy = __uint_as_float(n0.y) + (float)(((h ? n2.w : n2.z) >> sh) & 0xFFu) * sy;
                const float loz = __uint_as_float(n0.z) + (float)(((h ? n3.y : n3.x) >> sh) & 0xFFu) * sz;
                const float hix = __uint_as_float(n0.x) + (float)(((h ? n3.w : n3.z) >> sh) & 0xFFu) * sx;
                const float hiy = __uint_as_float(n0.y) + (float)(((h ? n4.y : n4.x) >> sh) & 0xFFu) * sy;
                const float hiz = __uint_as_float(n0.z) + (float)(((h ? n4.w : n4.z) >> sh) & 0xFFu) * sz;
                cx = 0.5f * (lox + hix) - O.x; cy = 0.5f * (loy + hiy) - O.y; cz = 0.5f * (loz + hiz) - O.z;
                ex = 0.5f * (hix - lox); ey = 0.5f * (hiy - loy); ez = 0.5f * (hiz - loz);
                // absorb the rounding of the centre/half-extent form (boxes are already padded by the builder)
                ex += 4e-7f * (fabsf(cx) + ex); ey += 4e-7f * (fabsf(cy) + ey); ez += 4e-7f * (fabsf(cz) + ez);
                const float top = N.x * cx + N.y * cy + N.z * cz + fabsf(N.x) * ex + fabsf(N.y) * ey + fabsf(N.z) * ez;
                const float far = fmaxf(fabsf(cx) + ex, fmaxf(fabsf(cy) + ey, fabsf(cz) + ez));
                keep = !(top < -1e-5f * far);   // something of the box lies above the tangent plane
                const bool inside = sphere_expand ? (cx * cx + cy * cy + cz * cz <= 1.1f * (ex * ex + ey * ey + ez * ez))
                                                  : (fabsf(cx) <= ex && fabsf(cy) <= ey && fabsf(cz) <= ez);
                const bool inner = (n0.w >> (24 + lane)) & 1u;
                if (inner) {
                    gx = n1.x + __popc((n0.w >> 24) & lt_mask);
                    gy = 0x80000000u;
                    expand = keep && inside;
                } else {
                    gx = n1.y + (meta & 31u);
                    gy = meta >> 5;          // unary count: 1, 3, 7 = one bit per triangle
                }
            }
        }
        const unsigned keep_b = __ballot_sync(kFull, keep), exp_b = __ballot_sync(kFull, expand), cand_b = keep_b & ~exp_b;
        if (keep && !expand) {
            const int slot = n + __popc(cand_b & lt_mask);
            W.ca[slot] = make_float4(cx, cy, cz, ex);
            W.cb[slot] = make_float4(ey, ez, __uint_as_float(gx), __uint_as_float(gy));
        }
        if (expand) W.queue[qn + __popc(exp_b & lt_mask)] = gx;
        n += __popc(cand_b);
        qn += __popc(exp_b);
        __syncwarp();
    }
    return n;
}

// Bit k of the result words = candidate k is a leaf (triangle group) rather than a subtree root.
__device__ __forceinline__ void entry_list_leaf_mask(const EntryList &W, const int n, const int lane, uint32_t lm[3]) {
#pragma unroll
    for (int w = 0; w < 3; w++) {
        const int k = 32 * w + lane;
        const bool leaf = k < n && __float_as_uint(W.cb[k].w) <= 0x00FFFFFFu;
        lm[w] = __ballot_sync(kFull, leaf);
    }
}

// approximate reciprocal for box tests only (never used by the pinned triangle test)
__device__ __forceinline__ float rcp_box(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}

// exclusive prefix sum over the warp of two 16-bit counters packed in one word; total returned through `total`
__device__ __forceinline__ uint32_t warp_excl_scan_packed(uint32_t v, const int lane, uint32_t &total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(kFull, incl, 31);
    return incl - v;
}

// ---- per-origin horizon map -------------------------------------------------------------------------------------------
// kHzBins azimuth bins (local frame of the vertex, azimuth = atan2(y, x)); hz[b] = a conservative upper bound of
// sin(elevation above the tangent plane) of ALL geometry seen from the origin in that azimuth range.  A ray whose local z
// exceeds hz[bin] cannot hit anything and is visible without any traversal.  Every piece of geometry is covered by exactly one
// entry-list candidate (culled ones lie wholly below the tangent plane), so bounding every candidate bounds the scene:
//   leaf candidate     exact maximum elevation of each of its triangles: maximum over the three edge arcs (end points and the
//                      interior critical point, which is the root of a LINEAR equation) or 1 if the zenith pierces the triangle;
//                      azimuth range = minimal arc containing the (non-degenerate) vertex azimuths
//   subtree candidate  cone around the bounding sphere of its box (the builder expands any child whose sphere contains the origin)
// Margins (2e-4 in sin-elevation, 0.02 rad in azimuth) absorb float rounding; rays inside the margin simply take the full path.
constexpr int kHzBins = 32;
constexpr float kHzTwoPi = 6.283185307179586f;

__device__ __forceinline__ void hz_update(uint32_t *hz, float phi_lo, float phi_hi, bool all, float sinh) {
    if (!(sinh > 0.0f)) return;
    const uint32_t v = __float_as_uint(fminf(sinh + 2e-4f, 2.0f));
    if (all || !(phi_hi - phi_lo < kHzTwoPi - 0.1f)) {
        for (int b = 0; b < kHzBins; b++) atomicMax(&hz[b], v);
        return;
    }
    // bins are [b, b+1) * 2pi/kHzBins - pi; walk from the bin of phi_lo - margin to the bin of phi_hi + margin (with wrap)
    const float scale = (float)kHzBins / kHzTwoPi;
    const int b0 = (int)floorf((phi_lo - 0.02f + 3.14159265358979f) * scale), b1 = (int)floorf((phi_hi + 0.02f + 3.14159265358979f) * scale);
    for (int b = b0; b <= b1; b++) atomicMax(&hz[((b % kHzBins) + kHzBins) % kHzBins], v);
}

// maximum of (v.z / |v|) over the segment a + t (b - a), t in [0,1]  (local frame: z = height above the tangent plane)
__device__ __forceinline__ float hz_edge_max(const f3 a, const f3 b) {
    const float la = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z), lb = sqrtf(b.x * b.x + b.y * b.y + b.z * b.z);
    float m = fmaxf(la > 0.f ? a.z / la : 1.f, lb > 0.f ? b.z / lb : 1.f);
    const f3 e = mk3(b.x - a.x, b.y - a.y, b.z - a.z);
    const float ae = a.x * e.x + a.y * e.y + a.z * e.z, ee = e.x * e.x + e.y * e.y + e.z * e.z;
    const float den = e.z * ae - a.z * ee;
    if (fabsf(den) > 0.f) {
        const float t = (a.z * ae - e.z * la * la) / den;
        if (t > 0.f && t < 1.f) {
            const f3 v = mk3(a.x + t * e.x, a.y + t * e.y, a.z + t * e.z);
            const float lv = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
            m = fmaxf(m, lv > 0.f ? v.z / lv : 1.f);
        }
    }
    return m;
}

__device__ __forceinline__ void hz_triangle(uint32_t *hz, const f3 q0, const f3 q1, const f3 q2) {
    const float zmax = fmaxf(q0.z, fmaxf(q1.z, q2.z));
    const float scale = fmaxf(fmaxf(fabsf(q0.x) + fabsf(q0.y) + fabsf(q0.z), fabsf(q1.x) + fabsf(q1.y) + fabsf(q1.z)), fabsf(q2.x) + fabsf(q2.y) + fabsf(q2.z));
    if (zmax < -1e-5f * scale) return;                       // wholly below the tangent plane
    float sinh = fmaxf(hz_edge_max(q0, q1), fmaxf(hz_edge_max(q1, q2), hz_edge_max(q2, q0)));
    // zenith inside the triangle's cone: (0,0,1) . (qi x qj) all of one sign (with tolerance) and the plane is above the origin
    const float c01 = q0.x * q1.y - q0.y * q1.x, c12 = q1.x * q2.y - q1.y * q2.x, c20 = q2.x * q0.y - q2.y * q0.x;
    const float tol = 1e-6f * scale * scale;
    const bool surround = (c01 >= -tol && c12 >= -tol && c20 >= -tol) || (c01 <= tol && c12 <= tol && c20 <= tol);
    // azimuth range: minimal arc containing the azimuths of the vertices that are not (numerically) on the vertical axis
    float ph[3]; int np = 0;
    const float rmin = 1e-4f * scale;
    if (fabsf(q0.x) + fabsf(q0.y) > rmin) ph[np++] = atan2f(q0.y, q0.x);
    if (fabsf(q1.x) + fabsf(q1.y) > rmin) ph[np++] = atan2f(q1.y, q1.x);
    if (fabsf(q2.x) + fabsf(q2.y) > rmin) ph[np++] = atan2f(q2.y, q2.x);
    bool all = false;
    float lo = 0.f, hi = 0.f;
    if (np == 3 && surround) { all = true; if (zmax > 0.f) sinh = 1.0f; }        // the vertical axis passes through the triangle
    else if (np == 0) all = true;
    else if (np == 1) { lo = hi = ph[0]; }
    else {
        // sort, then drop the largest gap
        if (np == 2) { lo = fminf(ph[0], ph[1]); hi = fmaxf(ph[0], ph[1]); if (hi - lo > 3.14159265358979f) { const float t = lo; lo = hi; hi = t + kHzTwoPi; } }
        else {
            float a = ph[0], b = ph[1], c = ph[2], t;
            if (a > b) { t = a; a = b; b = t; } if (b > c) { t = b; b = c; c = t; } if (a > b) { t = a; a = b; b = t; }
            const float g0 = b - a, g1 = c - b, g2 = a + kHzTwoPi - c;
            if (g2 >= g0 && g2 >= g1) { lo = a; hi = c; }
            else if (g0 >= g1) { lo = b; hi = a + kHzTwoPi; }
            else { lo = c; hi = b + kHzTwoPi; }
            if (hi - lo > 3.14159265358979f) all = true;      // cannot happen for a planar triangle that does not surround the axis; be safe
        }
    }
    hz_update(hz, lo, hi, all, sinh);
}

__device__ __forceinline__ void build_horizon(const EntryList &W, const int n_cand, const Tri48 *tris, const f3 O, const Frame &fr, uint32_t *hz, const int lane) {
    hz[lane] = 0u;                                           // kHzBins == 32
    __syncwarp();
    for (int k = lane; k < n_cand; k += 32) {
        const float4 ca = W.ca[k], cb = W.cb[k];
        const uint32_t gx = __float_as_uint(cb.z), gy = __float_as_uint(cb.w);
        if (gy > 0x00FFFFFFu) {
            // subtree candidate: exact bound of its box.  For a convex polyhedron that the vertical axis does not pierce the
            // maximum elevation lies on an edge (elevation is quasi-concave on every face plane), so 12 edge maxima suffice.
            if (!(ca.w < 1e30f)) { hz_update(hz, 0.f, 0.f, true, 1.0f); continue; }      // overflow candidate: unbounded
            {
                // vertical ray O + t n, t >= 0, against the box (slab test)
                float t0 = 0.f, t1 = 3.0e38f;
                const float cc[3] = {ca.x, ca.y, ca.z}, ee[3] = {ca.w, cb.x, cb.y}, nn3[3] = {fr.n.x, fr.n.y, fr.n.z};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float inv = 1.0f / (fabsf(nn3[a]) < 1e-12f ? copysignf(1e-12f, nn3[a]) : nn3[a]);
                    const float ta = (cc[a] - ee[a]) * inv, tb = (cc[a] + ee[a]) * inv;
                    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                }
                if (t0 <= t1 * 1.0001f + 1e-6f) { hz_update(hz, 0.f, 0.f, true, 1.0f); continue; }
            }
            const f3 cl = mk3(ca.x * fr.right.x + ca.y * fr.right.y + ca.z * fr.right.z, ca.x * fr.up.x + ca.y * fr.up.y + ca.z * fr.up.z,
                              ca.x * fr.n.x + ca.y * fr.n.y + ca.z * fr.n.z);
            const f3 hx = mk3(ca.w * fr.right.x, ca.w * fr.up.x, ca.w * fr.n.x), hy = mk3(cb.x * fr.right.y, cb.x * fr.up.y, cb.x * fr.n.y),
                     hzv = mk3(cb.y * fr.right.z, cb.y * fr.up.z, cb.y * fr.n.z);
            f3 q[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float sx = (i & 1) ? 1.f : -1.f, sy = (i & 2) ? 1.f : -1.f, sz = (i & 4) ? 1.f : -1.f;
                q[i] = mk3(cl.x + sx * hx.x + sy * hy.x + sz * hzv.x, cl.y + sx * hx.y + sy * hy.y + sz * hzv.y, cl.z + sx * hx.z + sy * hy.z + sz * hzv.z);
            }
            float sinh = -1.f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (!(i & 1)) sinh = fmaxf(sinh, hz_edge_max(q[i], q[i | 1]));
                if (!(i & 2)) sinh = fmaxf(sinh, hz_edge_max(q[i], q[i | 2]));
                if (!(i & 4)) sinh = fmaxf(sinh, hz_edge_max(q[i], q[i | 4]));
            }
            if (!(sinh > 0.f)) continue;
            // azimuth range relative to the centre's azimuth; a span >= pi means the box surrounds the vertical axis
            const float rc = fabsf(cl.x) + fabsf(cl.y), sc = fabsf(ca.w) + fabsf(cb.x) + fabsf(cb.y) + fabsf(ca.x) + fabsf(ca.y) + fabsf(ca.z);
            bool all = !(rc > 1e-4f * sc);
            float lo = 0.f, hi = 0.f;
            const float phic = atan2f(cl.y, cl.x);
            if (!all) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (!(fabsf(q[i].x) + fabsf(q[i].y) > 1e-4f * sc)) { all = true; continue; }
                    float dl = atan2f(q[i].y, q[i].x) - phic;
                    if (dl > 3.14159265358979f) dl -= kHzTwoPi;
                    if (dl < -3.14159265358979f) dl += kHzTwoPi;
                    lo = fminf(lo, dl); hi = fmaxf(hi, dl);
                }
                if (hi - lo >= 3.14159265358979f - 0.02f) all = true;
            }
            hz_update(hz, phic + lo, phic + hi, all, sinh);
        } else {
            for (uint32_t bits = gy, j = 0; bits; bits >>= 1, j++) {
                const char *tp = reinterpret_cast<const char *>(tris + gx + j);
                const u4 a4 = ld16(tp), b4 = ld16(tp + 16), c4 = ld16(tp + 32);
                const f3 p0 = mk3(PRT_U2F(a4.x) - O.x, PRT_U2F(a4.y) - O.y, PRT_U2F(a4.z) - O.z);
                const f3 p1 = mk3(p0.x + PRT_U2F(b4.x), p0.y + PRT_U2F(b4.y), p0.z + PRT_U2F(b4.z));
                const f3 p2 = mk3(p0.x + PRT_U2F(c4.x), p0.y + PRT_U2F(c4.y), p0.z + PRT_U2F(c4.z));
                const f3 q0 = mk3(p0.x * fr.right.x + p0.y * fr.right.y + p0.z * fr.right.z, p0.x * fr.up.x + p0.y * fr.up.y + p0.z * fr.up.z, p0.x * fr.n.x + p0.y * fr.n.y + p0.z * fr.n.z);
                const f3 q1 = mk3(p1.x * fr.right.x + p1.y * fr.right.y + p1.z * fr.right.z, p1.x * fr.up.x + p1.y * fr.up.y + p1.z * fr.up.z, p1.x * fr.n.x + p1.y * fr.n.y + p1.z * fr.n.z);
                const f3 q2 = mk3(p2.x * fr.right.x + p2.y * fr.right.y + p2.z * fr.right.z, p2.x * fr.up.x + p2.y * fr.up.y + p2.z * fr.up.z, p2.x * fr.n.x + p2.y * fr.n.y + p2.z * fr.n.z);
                hz_triangle(hz, q0, q1, q2);
            }
        }
    }
    __syncwarp();
}

// Tests the lane's ray (origin = list origin, interval [0, inf)) against all candidate boxes; one bit per candidate.
__device__ __forceinline__ void scan_entry_list(const EntryList &W, const int n, const float idx, const float idy, const float idz, uint32_t m[3]) {
    const float aix = fabsf(idx), aiy = fabsf(idy), aiz = fabsf(idz);
#pragma unroll
    for (int w = 0; w < 3; w++) {
        uint32_t bits = 0u, bit = 1u;
        const int k1 = min(n, 32 * (w + 1));
#pragma unroll 4
        for (int k = 32 * w; k < k1; k++) {
            const float4 a = W.ca[k], b = W.cb[k];
            const float tx = a.x * idx, ty = a.y * idy, tz = a.z * idz;
            const float tmin = fmaxf(fmaxf(fmaf(-a.w, aix, tx), fmaf(-b.x, aiy, ty)), fmaf(-b.y, aiz, tz));
            const float tmax = fminf(fminf(fmaf(a.w, aix, tx), fmaf(b.x, aiy, ty)), fmaf(b.y, aiz, tz));
            if (tmin <= tmax && tmax >= 0.0f) bits |= bit;
            bit += bit;
        }
        m[w] = bits;
    }
}

}  // namespace prt
