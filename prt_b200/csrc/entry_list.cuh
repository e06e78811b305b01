// entry_list.cuh -- per-origin entry lists (device only), shared by the bake kernels.
#pragma once
#include "traverse.cuh"
#include "horizon_math.cuh"
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
    float4 ca[kMaxCand];       // centre - origin (xyz), half extent z
    float4 cb[kMaxCand];       // half extent x, y (an aligned pair: the scan tests x and y with packed FFMA2), group x, group y (see Trav::start_group)
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
            W.ca[slot] = make_float4(cx, cy, cz, ez);
            W.cb[slot] = make_float4(ex, ey, __uint_as_float(gx), __uint_as_float(gy));
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
#ifndef PRT_RCP_FMNMX
#define PRT_RCP_FMNMX 1
#endif
__device__ __forceinline__ float rcp_box(float d) {
#if PRT_RCP_FMNMX
    d = copysignf(fmaxf(fabsf(d), 1e-18f), d);              // one FMNMX + one LOP3 (a NaN direction becomes 1e-18)
#else
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
#endif
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
#else
    return 1.0f / d;            // host build of the CPU test harness (tests/hostcheck)
#endif
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
// kHzBins azimuth bins in the local frame of the vertex; hz[b] = a conservative upper bound of sin(elevation above the tangent
// plane) of ALL geometry seen from the origin in that azimuth range.  A ray whose local z exceeds hz[bin] cannot hit anything and
// is visible without any traversal.  Every piece of geometry is covered by exactly one entry-list candidate (culled ones lie
// wholly below the tangent plane), so bounding every candidate bounds the scene.  The builder works *lazily*, nearest geometry
// first: every box (subtree or leaf) first gets a cheap bound --
//   cone around its bounding sphere, clamped by the slab bound z_max / d_min (largest height above the tangent plane over the
//   smallest distance from the origin: much tighter for the flat, surface-hugging boxes of a mesh, whose bounding sphere lifts
//   the horizon by their whole angular radius)
// -- and a box whose cheap bound cannot raise the map in any bin it spans is dropped at once, whatever its size (no refinement,
// no triangle bounds).  Of the boxes that could raise it:
//   leaf               exact maximum elevation of each of its triangles: maximum over the three edge arcs (end points and the
//                      interior critical point, which is the root of a LINEAR equation) or 1 if the vertical axis pierces it
//   far subtree        (small angular size) the cheap bound is merged
//   near subtree       the builder descends into it -- WITHOUT adding candidates -- until the pieces are small, are dropped or are
//                      triangles (budgeted; afterwards the exact box bound)
// "Cannot raise the map" is one range-minimum query: the map is published to shared memory together with its minima over 2, 4, 8
// and 16 consecutive bins (a sparse table, kHzLevels x 32 words), so the test is two loads whatever the azimuth span of the item.
// Azimuth is measured by the monotone "diamond" pseudo-angle p(x,y) in [0,4) (one division, no atan2); bins are uniform in p and
// the host bins the sample directions with the same formula.  Margins (2e-4 in sin-elevation, 0.02 in p >= 0.9 degree) absorb
// float rounding; rays inside the margin simply take the full path.  Lane b of the warp owns bin b while the map is built.
// Dropping only ever uses a published (possibly stale, i.e. lower) copy of the map, so it is conservative by construction.
// warp-collective: publishes the map (lane = bin) and its range minima
__device__ __forceinline__ void hz_publish(const float my, const int lane, uint32_t *hz) {
    float t = my;
    hz[lane] = __float_as_uint(t);
#pragma unroll
    for (int k = 1; k < kHzLevels; k++) {
        t = fminf(t, __shfl_sync(kFull, t, (lane + (1 << (k - 1))) & 31));
        hz[kHzBins * k + lane] = __float_as_uint(t);
    }
    __syncwarp();
}
// could the item raise the published map in any bin it spans?  (two loads: the range b0..b1 is covered by two table entries)
__device__ __forceinline__ bool hz_useful(const HzItem it, const uint32_t *hz) {
    if (!(it.v > 0.f)) return false;
    const int k = min(31 - __clz(it.b1 - it.b0 + 1), kHzLevels - 1);
    const float m = fminf(__uint_as_float(hz[kHzBins * k + (it.b0 & (kHzBins - 1))]),
                          __uint_as_float(hz[kHzBins * k + ((it.b1 - (1 << k) + 1) & (kHzBins - 1))]));
    return it.v > m;
}
// estimate of what merging the item would cost: the share of a bin's cosine-weighted samples between the published map (its minimum m
// over the item's bins) and the item's value, (v^2 - m^2), times the bins it spans -- x S / kHzBins = samples that would have to be traced
// (> 0 <=> hz_useful).  The range minimum overestimates the cost of a wide item; it is only used to decide which boxes are worth opening.
__device__ __forceinline__ float hz_gain(const HzItem it, const uint32_t *hz) {
    if (!(it.v > 0.f)) return 0.f;
    const int k = min(31 - __clz(it.b1 - it.b0 + 1), kHzLevels - 1);
    const float m = fminf(__uint_as_float(hz[kHzBins * k + (it.b0 & (kHzBins - 1))]),
                          __uint_as_float(hz[kHzBins * k + ((it.b1 - (1 << k) + 1) & (kHzBins - 1))]));
    return (it.v - m) * (it.v + m) * (float)(it.b1 - it.b0 + 1);
}
// warp-collective: every lane contributes one item (if `useful`); lane = bin keeps the running maximum.  Only the useful items
// go through the serial broadcast loop; the map is re-published when it may have changed.
__device__ __forceinline__ float hz_merge(float my, const HzItem it, const bool useful, const int lane, uint32_t *hz) {
    unsigned m = __ballot_sync(kFull, useful);              // also: every lane has finished its reads of hz[]
    if (!m) return my;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1u;
        const int b0 = __shfl_sync(kFull, it.b0, src), b1 = __shfl_sync(kFull, it.b1, src);
        const float v = __shfl_sync(kFull, it.v, src);
        if (((lane - b0) & (kHzBins - 1)) <= b1 - b0) my = fmaxf(my, v);
    }
    hz_publish(my, lane, hz);
    return my;
}

__device__ __forceinline__ HzItem hz_tri_item(const Tri48 *tris, const uint32_t ti, const f3 O, const Frame &fr) {
    const char *tp = reinterpret_cast<const char *>(tris + ti);
    const u4 a4 = ld16(tp), b4 = ld16(tp + 16), c4 = ld16(tp + 32);
    const f3 p0 = mk3(PRT_U2F(a4.x) - O.x, PRT_U2F(a4.y) - O.y, PRT_U2F(a4.z) - O.z);
    const f3 p1 = mk3(p0.x + PRT_U2F(b4.x), p0.y + PRT_U2F(b4.y), p0.z + PRT_U2F(b4.z));
    const f3 p2 = mk3(p0.x + PRT_U2F(c4.x), p0.y + PRT_U2F(c4.y), p0.z + PRT_U2F(c4.z));
    const f3 q0 = mk3(p0.x * fr.right.x + p0.y * fr.right.y + p0.z * fr.right.z, p0.x * fr.up.x + p0.y * fr.up.y + p0.z * fr.up.z, p0.x * fr.n.x + p0.y * fr.n.y + p0.z * fr.n.z);
    const f3 q1 = mk3(p1.x * fr.right.x + p1.y * fr.right.y + p1.z * fr.right.z, p1.x * fr.up.x + p1.y * fr.up.y + p1.z * fr.up.z, p1.x * fr.n.x + p1.y * fr.n.y + p1.z * fr.n.z);
    const f3 q2 = mk3(p2.x * fr.right.x + p2.y * fr.right.y + p2.z * fr.right.z, p2.x * fr.up.x + p2.y * fr.up.y + p2.z * fr.up.z, p2.x * fr.n.x + p2.y * fr.n.y + p2.z * fr.n.z);
    return hz_triangle(q0, q1, q2);
}

// host-side work counters of the CPU study harness (tests/hostcheck, tools/horizon_study.py); nothing in a product build
#ifndef PRT_HZ_STAT
#define PRT_HZ_STAT(counter, n)
#endif
#ifndef PRT_HZ_TRACE_FAR
#define PRT_HZ_TRACE_FAR(c, e, item)
#endif
constexpr int kHzQueue = 64;
constexpr int kHzTriQueue = 128;       // triangles waiting for a full-warp round: < 32 left over + at most 3 x 32 new per iteration
// `near2`: a box is "near" (gets refined) when d^2 < near2 * r^2, i.e. its angular radius exceeds asin(1/sqrt(near2)); `mid2`, `gain_min`: a
// smaller box (d^2 < mid2 * r^2) is refined too when merging its bound would leave more than gain_min x S / kHzBins samples to trace (hz_gain).
// hz: kHzWords words of shared memory; on return hz[0..kHzBins) is the map.
// slabs (optional): the oriented slab of every node (bvh8.h); tightens the bound of subtree boxes before they are merged or opened.
__device__ __forceinline__ void build_horizon(const EntryList &W, const int n_cand, const Node8 *nodes, const Tri48 *tris, const f3 O, const f3 N,
                                              const Frame &fr, uint32_t *hz, uint32_t *rq, uint32_t *tq, const int budget_iters, const float near2, const int lane,
                                              const Slab32 *slabs = nullptr, const float mid2 = 0.f, const float gain_min = 0.f) {
    const unsigned lt_mask = (1u << lane) - 1u;
    float my = 0.f;                                          // lane b owns bin b
    int rn = 0, tn = 0;
    hz_publish(my, lane, hz);
    // work item of a lane in one round: one box (subtree or leaf)
    // ---- pass 1: the entry-list candidates, deepest (nearest) first: they establish the horizon that lets the far boxes be
    //      dropped; pass 2: children of queued subtrees (4 nodes x 8 children per iteration) ---------------------------------
    int k0 = 0, budget = budget_iters;
    for (;;) {
        const bool pass1 = k0 < n_cand;
        if (!pass1 && rn == 0) break;
        bool valid = false, inner = false;
        f3 c = mk3(0.f, 0.f, 0.f), e = mk3(0.f, 0.f, 0.f);
        uint32_t gx = 0u, unary = 0u;
        if (pass1) {
            const int k = n_cand - 1 - (k0 + lane);
            if (k >= 0) {
                const float4 ca = W.ca[k], cb = W.cb[k];
                gx = __float_as_uint(cb.z);
                const uint32_t gy = __float_as_uint(cb.w);
                valid = true; inner = gy > 0x00FFFFFFu; unary = gy;
                c = mk3(ca.x, ca.y, ca.z); e = mk3(cb.x, cb.y, ca.w);
            }
            k0 += 32;
        } else {
            const int take = min(rn, 4), g = lane >> 3, slot = lane & 7;
            uint32_t node = 0xFFFFFFFFu;
            if (g < take) node = rq[rn - 1 - g];
            rn -= take;
            budget--;
            PRT_HZ_STAT(iterations, 1);
            __syncwarp();
            if (node != 0xFFFFFFFFu) {
                if (slot == 0) PRT_HZ_STAT(nodes_expanded, 1);
                const char *np = reinterpret_cast<const char *>(nodes + node);
                const u4 n0 = ld16(np), n1 = ld16(np + 16), n2 = ld16(np + 32), n3 = ld16(np + 48), n4 = ld16(np + 64);
                const int h = slot >> 2, sh = 8 * (slot & 3);
                const uint32_t meta = ((h ? n1.w : n1.z) >> sh) & 0xFFu;
                if (meta) {
                    const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23), sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
                    const float lox = __uint_as_float(n0.x) + (float)(((h ? n2.y : n2.x) >> sh) & 0xFFu) * sx;
                    const float loy = __uint_as_float(n0.y) + (float)(((h ? n2.w : n2.z) >> sh) & 0xFFu) * sy;
                    const float loz = __uint_as_float(n0.z) + (float)(((h ? n3.y : n3.x) >> sh) & 0xFFu) * sz;
                    const float hix = __uint_as_float(n0.x) + (float)(((h ? n3.w : n3.z) >> sh) & 0xFFu) * sx;
                    const float hiy = __uint_as_float(n0.y) + (float)(((h ? n4.y : n4.x) >> sh) & 0xFFu) * sy;
                    const float hiz = __uint_as_float(n0.z) + (float)(((h ? n4.w : n4.z) >> sh) & 0xFFu) * sz;
                    c = mk3(0.5f * (lox + hix) - O.x, 0.5f * (loy + hiy) - O.y, 0.5f * (loz + hiz) - O.z);
                    e = mk3(0.5f * (hix - lox), 0.5f * (hiy - loy), 0.5f * (hiz - loz));
                    e.x += 4e-7f * (fabsf(c.x) + e.x); e.y += 4e-7f * (fabsf(c.y) + e.y); e.z += 4e-7f * (fabsf(c.z) + e.z);
                    const float top = N.x * c.x + N.y * c.y + N.z * c.z + fabsf(N.x) * e.x + fabsf(N.y) * e.y + fabsf(N.z) * e.z;
                    const float far = fmaxf(fabsf(c.x) + e.x, fmaxf(fabsf(c.y) + e.y, fabsf(c.z) + e.z));
                    if (!(top < -1e-5f * far)) {
                        valid = true;
                        inner = (n0.w >> (24 + slot)) & 1u;
                        if (inner) gx = n1.x + __popc((n0.w >> 24) & ((1u << slot) - 1u));
                        else { gx = n1.y + (meta & 31u); unary = meta >> 5; }
                    }
                }
            }
        }
        // ---- classify the lane's box ------------------------------------------------------------------------------------
        HzItem it = hz_item(0.f, 0.f, false, 0.f);
        bool merge = false, push = false, leaf = false, nearb = false;
        float slab_v = 2.0f;
        if (valid) PRT_HZ_STAT(boxes_bounded, 1);
        if (valid && !(e.x < 1e30f)) { it = hz_item(0.f, 0.f, true, 1.0f); merge = true; }      // overflow candidate: unbounded
        else if (valid) {
            const float r2 = e.x * e.x + e.y * e.y + e.z * e.z, d2 = c.x * c.x + c.y * c.y + c.z * c.z;
            HzItem cb = hz_cheap_box(c, e, r2, d2, fr);
            if (hz_useful(cb, hz)) {
                if (!inner) leaf = true;
                else {
                    bool useful = true;
                    if (slabs) {
                        // the subtree is a sheet inside its box: bound the sheet (horizon_math.cuh)
                        const char *sp = reinterpret_cast<const char *>(slabs + gx);
                        const u4 s0 = ld16(sp), s1 = ld16(sp + 16);
                        const f3 m = mk3(PRT_U2F(s0.x), PRT_U2F(s0.y), PRT_U2F(s0.z));
                        const float mo = m.x * O.x + m.y * O.y + m.z * O.z;
                        slab_v = hz_slab_value(c, e, fr.n, m, PRT_U2F(s0.w) - mo, PRT_U2F(s1.x) - mo);
                        if (slab_v < cb.v) { cb.v = slab_v; useful = hz_useful(cb, hz); }
                    }
                    if (useful) {
                        // a mid-sized box whose bound would cost many samples is opened as well (mid2 = 0: rule off)
                        const bool mid = d2 < mid2 * r2 && hz_gain(cb, hz) > gain_min;
                        if (!(d2 < near2 * r2) && !mid) { it = cb; merge = true; PRT_HZ_TRACE_FAR(c, e, cb); }
                        else { nearb = true; push = budget > 0; }
                    }
                }
            }
        }
        const unsigned pb = __ballot_sync(kFull, push);
        const int pos = rn + __popc(pb & lt_mask);
        if (push && pos < kHzQueue) rq[pos] = gx;
        // near, but no budget / queue space left: exact bound of the box itself
        const bool boxed = nearb && !(push && pos < kHzQueue);
        if (__any_sync(kFull, boxed)) { if (boxed) { it = hz_box(c, e, fr); it.v = fminf(it.v, slab_v); merge = hz_useful(it, hz); } }
        rn = min(rn + __popc(pb), kHzQueue);
        my = hz_merge(my, it, merge, lane, hz);
        // leaves: their triangles are queued and bounded 32 at a time, so the (long) triangle bound always runs on a full warp;
        // every triangle keeps its own azimuth range
        for (uint32_t j = 0; j < 3u; j++) {
            const bool has = leaf && ((unary >> j) & 1u);
            const unsigned tb = __ballot_sync(kFull, has);
            if (!tb) break;
            if (has) tq[tn + __popc(tb & lt_mask)] = gx + j;
            tn += __popc(tb);
        }
        __syncwarp();
        while (tn >= 32) {
            tn -= 32;
            PRT_HZ_STAT(triangle_rounds, 1);
            const HzItem ti = hz_tri_item(tris, tq[tn + lane], O, fr);
            my = hz_merge(my, ti, hz_useful(ti, hz), lane, hz);
        }
        __syncwarp();
    }
    if (tn > 0) {
        PRT_HZ_STAT(triangle_rounds, 1);
        HzItem ti = hz_item(0.f, 0.f, false, 0.f);
        if (lane < tn) ti = hz_tri_item(tris, tq[lane], O, fr);
        my = hz_merge(my, ti, hz_useful(ti, hz), lane, hz);
    }
    hz[lane] = __float_as_uint(my);
    __syncwarp();
}

#ifndef PRT_WAVE_CULL
#define PRT_WAVE_CULL 0             // 1 (experimental, CPU-validated only): the candidate scan of bake_wave.cuh skips, warp-uniformly, the
#endif                              // candidates whose elevation bound lies below the lowest ray of the round (tools/entry_list_culling_study.py)
#if PRT_WAVE_CULL
// Optional culling data for the candidate scan: the cheap elevation bound (hz_cheap_box: no ray from the origin whose local z exceeds
// it can hit the box) of every candidate, stored as float bits in the build queue, which is free once the list is built.
static __device__ __noinline__ void entry_list_elevation_bounds(EntryList &W, const int n, const Frame &fr, const int lane) {
    for (int k = lane; k < n; k += 32) {
        const float4 a = W.ca[k], b = W.cb[k];
        float v = 2.0f;                                            // overflow candidate (unbounded): never skipped
        if (a.w < 1e30f) {
            const f3 c = mk3(a.x, a.y, a.z), e = mk3(b.x, b.y, a.w);
            v = hz_cheap_box(c, e, e.x * e.x + e.y * e.y + e.z * e.z, c.x * c.x + c.y * c.y + c.z * c.z, fr).v;
        }
        W.queue[k] = __float_as_uint(v);
    }
    __syncwarp();
}
// scan_entry_list with the warp-uniform skip: zmin = the lowest local z of the rays of this round
__device__ __forceinline__ void scan_entry_list_culled(const EntryList &W, const int n, const float idx, const float idy, const float idz, const float zmin, uint32_t m[3]) {
    const float aix = fabsf(idx), aiy = fabsf(idy), aiz = fabsf(idz);
#pragma unroll
    for (int w = 0; w < 3; w++) {
        uint32_t bits = 0u, bit = 1u;
        const int k1 = min(n, 32 * (w + 1));
        for (int k = 32 * w; k < k1; k++, bit += bit) {
            if (__uint_as_float(W.queue[k]) < zmin) continue;
            const float4 a = W.ca[k], b = W.cb[k];
            const float tx = a.x * idx, ty = a.y * idy, tz = a.z * idz;
            const float tmin = fmaxf(fmaxf(fmaf(-b.x, aix, tx), fmaf(-b.y, aiy, ty)), fmaf(-a.w, aiz, tz));
            const float tmax = fminf(fminf(fmaf(b.x, aix, tx), fmaf(b.y, aiy, ty)), fmaf(a.w, aiz, tz));
            if (tmin <= tmax && tmax >= 0.0f) bits |= bit;
        }
        m[w] = bits;
    }
}

#endif  // PRT_WAVE_CULL

// Tests the lane's ray (origin = list origin, interval [0, inf)) against all candidate boxes; one bit per candidate.
#ifndef PRT_SCAN_PACKED
#define PRT_SCAN_PACKED 1
#endif
#ifndef PRT_SCAN_UNROLL
#define PRT_SCAN_UNROLL 2
#endif
constexpr int kScanUnroll = PRT_SCAN_UNROLL;
__device__ __forceinline__ void scan_entry_list(const EntryList &W, const int n, const float idx, const float idy, const float idz, uint32_t m[3]) {
    const float aix = fabsf(idx), aiy = fabsf(idy), aiz = fabsf(idz);
#if defined(__CUDA_ARCH__) && PRT_SCAN_PACKED
    const float2 id_xy = make_float2(idx, idy), ai_xy = make_float2(aix, aiy), nai_xy = make_float2(-aix, -aiy);
#endif
#pragma unroll
    for (int w = 0; w < 3; w++) {
        uint32_t bits = 0u, bit = 1u;
        const int k1 = min(n, 32 * (w + 1));
#pragma unroll kScanUnroll
        for (int k = 32 * w; k < k1; k++) {
            const float4 a = W.ca[k], b = W.cb[k];
#if defined(__CUDA_ARCH__) && PRT_SCAN_PACKED
            // x and y of the slab test in packed mul / fma (sm_100a; the same IEEE operations as below): 6 instead of 9 per candidate
            const float2 txy = __fmul2_rn(make_float2(a.x, a.y), id_xy);
            const float tz = a.z * idz;
            const float2 lo = __ffma2_rn(make_float2(b.x, b.y), nai_xy, txy), hi = __ffma2_rn(make_float2(b.x, b.y), ai_xy, txy);
            const float tmin = fmaxf(fmaxf(lo.x, lo.y), fmaf(-a.w, aiz, tz));
            const float tmax = fminf(fminf(hi.x, hi.y), fmaf(a.w, aiz, tz));
#else
            const float tx = a.x * idx, ty = a.y * idy, tz = a.z * idz;
            const float tmin = fmaxf(fmaxf(fmaf(-b.x, aix, tx), fmaf(-b.y, aiy, ty)), fmaf(-a.w, aiz, tz));
            const float tmax = fminf(fminf(fmaf(b.x, aix, tx), fmaf(b.y, aiy, ty)), fmaf(a.w, aiz, tz));
#endif
            if (tmin <= tmax && tmax >= 0.0f) bits |= bit;
            bit += bit;
        }
        m[w] = bits;
    }
}

}  // namespace prt
