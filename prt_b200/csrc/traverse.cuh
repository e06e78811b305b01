// traverse.cuh -- per-ray traversal of the 8-wide compressed BVH (bvh8.h) with the pinned triangle test.
//
// Stands in for rtcOccluded1 / rtcIntersect1 (reference src/raytracing/light_probe.cpp:119,128;
// raytracing.cpp:167,200,254).  One ray per thread; the traversal state is resumable so a persistent warp
// can break out, refill idle lanes with fresh rays and continue (bake.cu).  Node fetch = 5 x 16-byte loads,
// triangle fetch = 3 x 16-byte loads through the read-only path.
//
// Hit rule: Moeller-Trumbore on (v0,e1,e2), sign-folded, tnear < t <= tfar, both faces,
// det == 0 never hits; closest hit = min (t, prim) over all candidates valid in the *initial* interval, so the
// answer does not depend on traversal order or on the tree.
#pragma once
#include "bvh8.h"
#include "prt_math.cuh"

namespace prt {

struct u4 { uint32_t x, y, z, w; };
struct u2 { uint32_t x, y; };

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ u4 ld16(const void *p) {
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    u4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
#define PRT_CLZ(x) __clz((int)(x))
#define PRT_POPC(x) __popc(x)
#define PRT_FFS(x) __ffs((int)(x))
#define PRT_ACTIVE_LANES() __popc(__activemask())
#else
inline u4 ld16(const void *p) { u4 r; memcpy(&r, p, 16); return r; }
#define PRT_CLZ(x) ((x) ? __builtin_clz(x) : 32)
#define PRT_POPC(x) __builtin_popcount(x)
#define PRT_FFS(x) __builtin_ffs((int)(x))
#define PRT_ACTIVE_LANES() 32
#endif

PRT_HD float safe_rcp(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    return 1.0f / d;
}

// ---- quantised child boxes of one node against one ray (the wavefront kernels: one lane = one (ray, node) item) -----------------------
// Bit s of the result = slot s hit in (RANGE ? [tnear, tfar] : [0, inf)).
// These kernels are bound by instruction issue with the XU pipe (I2F, MUFU.RCP; quarter rate) their busiest unit (62 % in the traversal
// pass) and the ALU pipe at 49 %.  PRT_NODE_PRMT = n moves the byte -> float conversion of n of the six plane groups (far x, far y, far z,
// near x, near y, near z) from I2F to a byte permute on the ALU pipe: byte q placed in mantissa bits 8..15 of 2^15 is the float
// 32768 + q, and the constant is folded into the plane offset, t = (32768 + q) * s + (a - 32768 * s) -- still one FFMA per plane.  The
// offset is rounded once more, at 2^-9 of a quantisation step; it is pushed OUTWARD by 2^-8 of a step (one more FFMA per axis and
// node), so the test stays conservative whatever the node's size -- box tests only ever have to be conservative.
// Measured on the headline bake (profiles/r2_node_test_ab.jsonl, step ms): n = 0: 51.1, 3: 49.8, 6: 50.4 -- balancing the two pipes is
// what pays: far planes through the ALU, near planes on the XU.  PRT_NODE_FFMA2 pairs the two planes of an axis in one packed
// fma.rn.f32x2 (sm_100a; near plane from I2F with offset a, far plane from the byte permute with offset f): 24 FFMA2 instead of 48 FFMA per
// node test, 24 instructions net after the moves that build the operand pairs.  With all six planes on I2F (first half of round 2) it
// gained nothing -- the XU pipe was what bound; with the 3 / 3 split it is -1.8 % on the headline step (45.66 -> 44.85 ms) and -1.7 % on
// the folded mesh (profiles/r2_ffma2_ab.jsonl).  The host build (tests/hostcheck) evaluates the same expressions with scalar fmaf.
#ifndef PRT_NODE_FFMA2
#define PRT_NODE_FFMA2 1
#endif
#ifndef PRT_NODE_PRMT
#define PRT_NODE_PRMT 3
#endif
#define PRT_Q2F_I2F(w, j) ((float)(((w) >> (8 * (j))) & 0xFFu))
#ifndef PRT_NODE_PRMT_REG
#define PRT_NODE_PRMT_REG 1         // 0: the exponent word is an immediate again (A/B)
#endif
#if defined(__CUDACC__) && PRT_NODE_PRMT_REG
static __constant__ uint32_t prt_q2f_exp = 0x47000000u;
#endif
#if defined(__CUDA_ARCH__)
// The exponent word lives in constant memory so that it is an operand the compiler has to keep in a REGISTER: PRMT takes one immediate,
// and with the word as the immediate every selector (four per plane group) needed a MOV of its own -- 24 of the 204 instructions of a
// node test.
#define PRT_Q2F_PRMT(w, j) __uint_as_float(__byte_perm((w), q2f_exp, 0x7604u | ((j) << 4)))
#else
#define PRT_Q2F_PRMT(w, j) (32768.0f + PRT_Q2F_I2F(w, j))
#endif
// plane group g (0..5 = far x, far y, far z, near x, near y, near z): conversion by byte permute (offset b) or by I2F (offset a)
#define PRT_PLANE(g, w, j, s, a, b) ((g) < PRT_NODE_PRMT ? fmaf(PRT_Q2F_PRMT(w, j), s, b) : fmaf(PRT_Q2F_I2F(w, j), s, a))
// DOP: a fourth slab axis (bvh8.h Dop32: p0 = scaled mean normal M and offset D0, p1 = the children's quantised extents along it; d = the ray
// direction): s(t) = M . o - D0 + t (M . d) must meet [qlo, qhi] of a child inside the ray's interval in the child's box -- the same
// arithmetic as the three box axes with "scale" 1 / (M . d), folded into the interval with one more min (the max comes for free: the
// clamp at 0 becomes a three-input max).  Box tests only ever have to be conservative; the extents carry a step of padding either side.
template <bool RANGE, bool DOP = false>
PRT_HD uint32_t node_slots_hit_t(const u4 n0, const u4 n2, const u4 n3, const u4 n4, const f3 o, const float idx, const float idy,
                                 const float idz, const float tnear, const float tfar, const u4 p0 = u4{0u, 0u, 0u, 0u},
                                 const u4 p1 = u4{0u, 0u, 0u, 0u}, const f3 d = f3{0.f, 0.f, 0.f}) {
    const float sx = PRT_U2F((n0.w & 0xFFu) << 23) * idx;
    const float sy = PRT_U2F(((n0.w >> 8) & 0xFFu) << 23) * idy;
    const float sz = PRT_U2F(((n0.w >> 16) & 0xFFu) << 23) * idz;
    const float ax = (PRT_U2F(n0.x) - o.x) * idx;
    const float ay = (PRT_U2F(n0.y) - o.y) * idy;
    const float az = (PRT_U2F(n0.z) - o.z) * idz;
    // offsets of the permuted planes: far planes pushed out (+), near planes pushed out (-) by 2^-8 of a quantisation step (dead code
    // for the groups that stay on I2F)
    const float gx = fabsf(sx) * 0.00390625f, gy = fabsf(sy) * 0.00390625f, gz = fabsf(sz) * 0.00390625f;
    const float fx = fmaf(-32768.0f, sx, ax) + gx, fy = fmaf(-32768.0f, sy, ay) + gy, fz = fmaf(-32768.0f, sz, az) + gz;
    const float mx = fmaf(-32768.0f, sx, ax) - gx, my = fmaf(-32768.0f, sy, ay) - gy, mz = fmaf(-32768.0f, sz, az) - gz;
    const bool nx = idx < 0.f, ny = idy < 0.f, nz = idz < 0.f;
    // fourth axis: s = so + t sd; t = (q - so) / sd = q * ss + as
    float ss = 0.f, as = 0.f, fs = 0.f;
    bool ns = false;
    if (DOP) {
        const float so = fmaf(PRT_U2F(p0.z), o.z, fmaf(PRT_U2F(p0.y), o.y, PRT_U2F(p0.x) * o.x)) - PRT_U2F(p0.w);
        float sd = fmaf(PRT_U2F(p0.z), d.z, fmaf(PRT_U2F(p0.y), d.y, PRT_U2F(p0.x) * d.x));
        sd = copysignf(fmaxf(fabsf(sd), 1e-18f), sd);
#if defined(__CUDA_ARCH__)
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ss) : "f"(sd));       // box tests only (2 ulp; the extents carry a step of padding)
#else
        ss = 1.0f / sd;
#endif
        as = -so * ss;
        fs = fmaf(-32768.0f, ss, as) + fabsf(ss) * 0.00390625f;
        ns = ss < 0.f;
    }
#if defined(__CUDA_ARCH__) && PRT_NODE_PRMT_REG
    const uint32_t q2f_exp = prt_q2f_exp;
#elif defined(__CUDA_ARCH__)
    const uint32_t q2f_exp = 0x47000000u;
#endif
    uint32_t hits = 0u;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t qlx = h ? n2.y : n2.x, qly = h ? n2.w : n2.z, qlz = h ? n3.y : n3.x;
        const uint32_t qhx = h ? n3.w : n3.z, qhy = h ? n4.y : n4.x, qhz = h ? n4.w : n4.z;
        const uint32_t nearx = nx ? qhx : qlx, farx = nx ? qlx : qhx;
        const uint32_t neary = ny ? qhy : qly, fary = ny ? qly : qhy;
        const uint32_t nearz = nz ? qhz : qlz, farz = nz ? qlz : qhz;
        const uint32_t qls = h ? p1.y : p1.x, qhs = h ? p1.w : p1.z;
        const uint32_t nears = ns ? qhs : qls, fars = ns ? qls : qhs;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float t0s = 0.0f, t1s = 0.0f;
            if (DOP) {
#if defined(__CUDA_ARCH__) && PRT_NODE_FFMA2 && PRT_NODE_PRMT == 3
                const float2 ts = __ffma2_rn(make_float2(PRT_Q2F_I2F(nears, j), PRT_Q2F_PRMT(fars, j)), make_float2(ss, ss), make_float2(as, fs));
                t0s = ts.x; t1s = ts.y;
#else
                t0s = fmaf(PRT_Q2F_I2F(nears, j), ss, as); t1s = fmaf(PRT_Q2F_PRMT(fars, j), ss, fs);     // the same formulas, scalar
#endif
            }
#if defined(__CUDA_ARCH__) && PRT_NODE_FFMA2 && PRT_NODE_PRMT == 3
            // near plane (I2F, offset a) and far plane (byte permute, offset f) of an axis in one packed FFMA2
            const float2 tx = __ffma2_rn(make_float2(PRT_Q2F_I2F(nearx, j), PRT_Q2F_PRMT(farx, j)), make_float2(sx, sx), make_float2(ax, fx));
            const float2 ty = __ffma2_rn(make_float2(PRT_Q2F_I2F(neary, j), PRT_Q2F_PRMT(fary, j)), make_float2(sy, sy), make_float2(ay, fy));
            const float2 tz = __ffma2_rn(make_float2(PRT_Q2F_I2F(nearz, j), PRT_Q2F_PRMT(farz, j)), make_float2(sz, sz), make_float2(az, fz));
            const float t0x = tx.x, t1x = tx.y, t0y = ty.x, t1y = ty.y, t0z = tz.x, t1z = tz.y;
#elif defined(__CUDA_ARCH__) && PRT_NODE_FFMA2 && PRT_NODE_PRMT == 0
            const float2 tx = __ffma2_rn(make_float2(PRT_Q2F_I2F(nearx, j), PRT_Q2F_I2F(farx, j)), make_float2(sx, sx), make_float2(ax, ax));
            const float2 ty = __ffma2_rn(make_float2(PRT_Q2F_I2F(neary, j), PRT_Q2F_I2F(fary, j)), make_float2(sy, sy), make_float2(ay, ay));
            const float2 tz = __ffma2_rn(make_float2(PRT_Q2F_I2F(nearz, j), PRT_Q2F_I2F(farz, j)), make_float2(sz, sz), make_float2(az, az));
            const float t0x = tx.x, t1x = tx.y, t0y = ty.x, t1y = ty.y, t0z = tz.x, t1z = tz.y;
#else
            const float t1x = PRT_PLANE(0, farx, j, sx, ax, fx), t1y = PRT_PLANE(1, fary, j, sy, ay, fy), t1z = PRT_PLANE(2, farz, j, sz, az, fz);
            const float t0x = PRT_PLANE(3, nearx, j, sx, ax, mx), t0y = PRT_PLANE(4, neary, j, sy, ay, my), t0z = PRT_PLANE(5, nearz, j, sz, az, mz);
#endif
            float tmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, RANGE ? tnear : 0.0f));
            float tmax = RANGE ? fminf(fminf(t1x, t1y), fminf(t1z, tfar)) : fminf(fminf(t1x, t1y), t1z);
            if (DOP) { tmin = fmaxf(tmin, t0s); tmax = fminf(tmax, t1s); }
            if (tmin <= tmax) hits |= 1u << (4 * h + j);
        }
    }
    return hits;
}

// run() results: RUNNING = interrupted for a refill (state kept); HIT = any-hit found (ANY only);
// EMPTY = stack and current groups exhausted (closest-hit result, if any, is in best_*)
enum { TRAV_RUNNING = 0, TRAV_HIT = 1, TRAV_EMPTY = 2 };

// Pinned ray/triangle decision on triangle `ti`: hit iff tnear < t <= tfar.
// tri_hit_regs: the same decision on a triangle whose three 16-byte words are already in registers (lets a caller issue the
// fetch before the ray direction is known)
PRT_HD bool tri_hit_regs(const u4 a, const u4 b, const u4 c, f3 o, f3 d, float tnear, float tfar, bool want_t, float &t, uint32_t &prim) {
    const f3 v0 = mk3(PRT_U2F(a.x), PRT_U2F(a.y), PRT_U2F(a.z));
    const f3 e1 = mk3(PRT_U2F(b.x), PRT_U2F(b.y), PRT_U2F(b.z));
    const f3 e2 = mk3(PRT_U2F(c.x), PRT_U2F(c.y), PRT_U2F(c.z));
    const f3 tv = sub3(o, v0);
    const f3 pv = cross3(d, e2);
    const fpair du = dot3_pair(e1, tv, pv);          // the two dot products with pv, and below the two with qv, share their packed ops
    const float det = du.x, U = du.y;
    const f3 qv = cross3(tv, e1);
    const fpair vt = dot3_pair(d, e2, qv);
    const float V = vt.x, T = vt.y;
    const uint32_t sgn = PRT_F2U(det) & 0x80000000u;
    const float ad = fabsf(det);
    const float Us = PRT_U2F(PRT_F2U(U) ^ sgn), Vs = PRT_U2F(PRT_F2U(V) ^ sgn), Ts = PRT_U2F(PRT_F2U(T) ^ sgn);
    const bool hit = (ad > 0.0f) && (Us >= 0.0f) && (Vs >= 0.0f) && (PRT_ADD(Us, Vs) <= ad) &&
                     (Ts > PRT_MUL(tnear, ad)) && (Ts <= PRT_MUL(tfar, ad));
    if (hit && want_t) { t = PRT_DIV(Ts, ad); prim = a.w; }
    return hit;
}
PRT_HD bool tri_hit(const Tri48 *tris, uint32_t ti, f3 o, f3 d, float tnear, float tfar, bool want_t, float &t, uint32_t &prim) {
    const char *tp = reinterpret_cast<const char *>(tris + ti);
    const u4 a = ld16(tp), b = ld16(tp + 16), c = ld16(tp + 32);
    return tri_hit_regs(a, b, c, o, d, tnear, tfar, want_t, t, prim);
}

struct Trav {
    // ray for the pinned triangle test
    f3 o, d;
    float tnear, tfar0;
    // box-test constants
    float idx, idy, idz;
    float tbox;          // current far limit for box culling (shrinks in closest-hit mode)
    uint32_t octinv4;
    u2 ng, tg;           // current node group (child base, hits|imask) and triangle group (tri base, bits)
    int sp;
    // closest-hit result
    float best_t;
    uint32_t best_prim, best_tri;
    uint32_t n_node_visits, n_tri_tests;   // algorithmic work counters (bench roofline)
    u2 stack[kStackEntries];

    // sets up the ray; traversal starts empty (use start_root() / start_group())
    PRT_HD void init(f3 org, f3 dir, float tn, float tf) {
        o = org; d = dir; tnear = tn; tfar0 = tf; tbox = tf;
        idx = safe_rcp(dir.x); idy = safe_rcp(dir.y); idz = safe_rcp(dir.z);
        // signs are taken from the clamped reciprocals so that +-0 components stay consistent with the slab order
        uint32_t oct = (idx < 0.f ? 4u : 0u) | (idy < 0.f ? 2u : 0u) | (idz < 0.f ? 1u : 0u);
        octinv4 = (7u - oct) * 0x01010101u;
        ng.x = 0u; ng.y = 0u;
        tg.x = 0u; tg.y = 0u;
        sp = 0;
        best_t = INFINITY; best_prim = 0xFFFFFFFFu; best_tri = 0xFFFFFFFFu;
    }
    PRT_HD void reset_counters() { n_node_visits = 0u; n_tri_tests = 0u; }
    // (x, y) is a stack-entry style group: y > 0x00FFFFFF -> node group (child base x, hit bits | imask);
    // otherwise a triangle group (first triangle x, bit per triangle).  (node, 0x80000000) starts at `node`.
    PRT_HD void start_group(uint32_t x, uint32_t y) { ng.x = x; ng.y = y; }
    PRT_HD void start_root() { start_group(0u, 0x80000000u); }

    // Intersects the 8 quantised child boxes of `node`; fills ng/tg.
    PRT_HD void visit_node(const Node8 *nodes, uint32_t node) {
        const char *np = reinterpret_cast<const char *>(nodes + node);
        const u4 n0 = ld16(np), n1 = ld16(np + 16), n2 = ld16(np + 32), n3 = ld16(np + 48), n4 = ld16(np + 64);
        const float sx = PRT_U2F((n0.w & 0xFFu) << 23) * idx;
        const float sy = PRT_U2F(((n0.w >> 8) & 0xFFu) << 23) * idy;
        const float sz = PRT_U2F(((n0.w >> 16) & 0xFFu) << 23) * idz;
        const float ax = (PRT_U2F(n0.x) - o.x) * idx;
        const float ay = (PRT_U2F(n0.y) - o.y) * idy;
        const float az = (PRT_U2F(n0.z) - o.z) * idz;
        const bool nx = idx < 0.f, ny = idy < 0.f, nz = idz < 0.f;
        uint32_t hitmask = 0u;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t meta4 = h ? n1.w : n1.z;
            const uint32_t inner = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t inner_mask4 = (inner >> 4) * 0xFFu;
            const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
            const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
            const uint32_t qlx = h ? n2.y : n2.x, qly = h ? n2.w : n2.z, qlz = h ? n3.y : n3.x;
            const uint32_t qhx = h ? n3.w : n3.z, qhy = h ? n4.y : n4.x, qhz = h ? n4.w : n4.z;
            const uint32_t nearx = nx ? qhx : qlx, farx = nx ? qlx : qhx;
            const uint32_t neary = ny ? qhy : qly, fary = ny ? qly : qhy;
            const uint32_t nearz = nz ? qhz : qlz, farz = nz ? qlz : qhz;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int sh = 8 * j;
                const float t0x = (float)((nearx >> sh) & 0xFFu) * sx + ax;
                const float t0y = (float)((neary >> sh) & 0xFFu) * sy + ay;
                const float t0z = (float)((nearz >> sh) & 0xFFu) * sz + az;
                const float t1x = (float)((farx >> sh) & 0xFFu) * sx + ax;
                const float t1y = (float)((fary >> sh) & 0xFFu) * sy + ay;
                const float t1z = (float)((farz >> sh) & 0xFFu) * sz + az;
                const float tmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tnear));
                const float tmax = fminf(fminf(t1x, t1y), fminf(t1z, tbox));
                if (tmin <= tmax) hitmask |= ((child_bits4 >> sh) & 0xFFu) << ((bit_index4 >> sh) & 0xFFu);
            }
        }
        ng.x = n1.x; ng.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
        tg.x = n1.y; tg.y = hitmask & 0x00FFFFFFu;
    }

    PRT_HD bool tri_test(const Tri48 *tris, uint32_t ti, bool want_t, float &t, uint32_t &prim) const {
        return tri_hit(tris, ti, o, d, tnear, tfar0, want_t, t, prim);
    }

    // Runs until the ray is finished, or -- when refill_thresh > 0 and more rays are waiting -- until fewer than
    // refill_thresh lanes of the warp are still traversing.  ANY: stop at the first hit.
    template <bool ANY>
    PRT_HD int run(const Node8 *nodes, const Tri48 *tris, int refill_thresh, bool more) {
        for (;;) {
            if (ng.y > 0x00FFFFFFu) {
                const uint32_t hits_imask = ng.y;
                const uint32_t bit = 31u - (uint32_t)PRT_CLZ(hits_imask);
                ng.y &= ~(1u << bit);
                if (ng.y > 0x00FFFFFFu) stack[sp++] = ng;
                const uint32_t slot = (bit - 24u) ^ (octinv4 & 7u);
                const uint32_t rel = (uint32_t)PRT_POPC(hits_imask & ~(0xFFFFFFFFu << slot));
                visit_node(nodes, ng.x + rel);
                n_node_visits++;
            } else {
                tg = ng; ng.y = 0u;
            }
            while (tg.y) {
                const uint32_t bit = (uint32_t)PRT_FFS(tg.y) - 1u;
                tg.y &= tg.y - 1u;
                float t; uint32_t prim;
                n_tri_tests++;
                if (tri_test(tris, tg.x + bit, !ANY, t, prim)) {
                    if (ANY) return TRAV_HIT;
                    if (t < best_t || (t == best_t && prim < best_prim)) { best_t = t; best_prim = prim; best_tri = tg.x + bit; tbox = t; }
                }
            }
            if (ng.y <= 0x00FFFFFFu) {
                if (sp == 0) return TRAV_EMPTY;
                ng = stack[--sp];
            }
            if (refill_thresh > 0 && more && PRT_ACTIVE_LANES() < refill_thresh) return TRAV_RUNNING;
        }
    }

    // unnormalised geometric normal (v1-v0)x(v2-v0) of the closest hit (Embree Ng convention)
    PRT_HD f3 hit_ng(const Tri48 *tris) const {
        const char *tp = reinterpret_cast<const char *>(tris + best_tri);
        const u4 b = ld16(tp + 16), c = ld16(tp + 32);
        return cross3(mk3(PRT_U2F(b.x), PRT_U2F(b.y), PRT_U2F(b.z)), mk3(PRT_U2F(c.x), PRT_U2F(c.y), PRT_U2F(c.z)));
    }
};

}  // namespace prt
