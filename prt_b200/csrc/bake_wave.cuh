// bake_wave.cuh -- the per-vertex work of the warp-local wavefront kernel (bake_wave.cu): everything one persistent warp does for
// one vertex, as an inline device function, so that the same code also runs on the CPU test harness (tests/hostcheck: plain C++
// against the warp emulator) where the CPU test-suite checks it against the oracle.  See bake_wave.cu for the design.
#pragma once
#include "kernels.h"
#include "entry_list.cuh"

namespace prt {

namespace {

constexpr int kMaxS = 8192;             // occlusion bitset capacity (samples per vertex)
#ifndef PRT_WAVE_CAP
#define PRT_WAVE_CAP 256            // node stack 320 + leaf stack 192 (same shared memory; the leaf stack never fills): 2.6 % SLOWER (round 2, session V)
#endif
#ifndef PRT_WAVE_SCAN_PUSH
#define PRT_WAVE_SCAN_PUSH 1
#endif
#ifndef PRT_WAVE_MINB
#define PRT_WAVE_MINB 7
#endif
#ifndef PRT_WAVE_BLOCK
#define PRT_WAVE_BLOCK 128
#endif
#ifndef PRT_WAVE_NCAP
#define PRT_WAVE_NCAP PRT_WAVE_CAP
#endif
#ifndef PRT_WAVE_LCAP
#define PRT_WAVE_LCAP PRT_WAVE_CAP
#endif
#ifndef PRT_WAVE_ROOM8
#define PRT_WAVE_ROOM8 2            // new rays are scanned while both stacks are at most ROOM8/8 full
#endif
constexpr int kNodeCap = PRT_WAVE_NCAP, kLeafCap = PRT_WAVE_LCAP;
static_assert(kNodeCap * 64 >= kMaxS, "the node stack doubles as the visibility permutation buffer");

// per-warp shared memory: this struct followed by the occlusion bitset (vis_words words rounded up to 16 bytes; bit i: the
// primary ray with PROCESSING index i is occluded -- an item carries that index, so a step can drop the items of occluded
// rays before it fetches anything, and the sample, node and triangle fetches of a live item are issued together) -- sized
// per launch so that 1024-sample bakes fit 8 CTAs per SM.  The visibility words of the C ABI are in reference order: they
// are permuted through the (then empty) node stack when a caller asks for them.
struct WaveShared {
    EntryList el;
    uint2 nq[kNodeCap];                 // (processing index of the ray, node index)
    uint2 lq[kLeafCap];                 // (processing index | triangle bits << 16, first triangle)
    uint32_t pend[64];                  // processing indices of rays that are not above the horizon, waiting for a scan round
};

// work counters of the CPU harness (tests/hostcheck, tools/wave_study.py): steps by kind, lanes per step, stack overflows; nothing in a
// product build
#ifndef PRT_WAVE_STAT
#define PRT_WAVE_STAT(counter, n)
#endif
// study hook of the CPU harness: may clear bits of the hit masks of a node step (models a tighter child test); nothing in a product build
#ifndef PRT_WAVE_NODE_STUDY
#define PRT_WAVE_NODE_STUDY(A, node, org, d, inner8, leaf8)
#endif

// 8 quantised child boxes of one node against a ray from the vertex (interval [0, inf)): traverse.cuh
__device__ __forceinline__ uint32_t node_slots_hit(const u4 n0, const u4 n2, const u4 n3, const u4 n4, const f3 o,
                                                   const float idx, const float idy, const float idz) {
    return node_slots_hit_t<false>(n0, n2, n3, n4, o, idx, idy, idz, 0.0f, 0.0f);
}

// rare overflow paths, kept out of line so that they do not occupy the instruction cache of the hot loop
__device__ __noinline__ bool fallback_subtree(const Node8 *nodes, const Tri48 *tris, const f3 org, const f3 d, const uint32_t child, uint32_t &nv, uint32_t &nt) {
    Trav tr; tr.reset_counters();
    tr.init(org, d, 0.0f, INFINITY); tr.start_group(child, 0x80000000u);
    const bool hit = tr.run<true>(nodes, tris, 0, false) == TRAV_HIT;
    nv += tr.n_node_visits; nt += tr.n_tri_tests;
    return hit;
}
__device__ __noinline__ bool fallback_leaf(const Tri48 *tris, const f3 org, const f3 d, const uint32_t tri0, uint32_t bits, uint32_t &nt) {
    while (bits) {
        const uint32_t b = (uint32_t)__ffs(bits) - 1u;
        bits &= bits - 1u;
        float t; uint32_t prim;
        nt++;
        if (tri_hit(tris, tri0 + b, org, d, 0.0f, INFINITY, false, t, prim)) return true;
    }
    return false;
}

// The visibility words of the C ABI are in reference order, the occlusion bitset in processing order: permute through `perm`
// (the node stack, empty by then).  Out of line: only callers that ask for visibility words pay for it, and its registers
// stay out of the traversal loop's allocation.
__device__ __noinline__ void write_vis_permuted(const float4 *samples, const uint32_t *occl, uint32_t *perm, uint32_t *row, const int S, const int words, const int lane) {
    for (int w = lane; w < words; w += 32) perm[w] = 0u;
    __syncwarp();
    for (int i = lane; i < S; i += 32) {
        if ((occl[i >> 5] >> lane) & 1u) continue;
        const uint32_t sref = __float_as_uint(__ldg(&samples[i].w)) & 0xFFFFFFu;
        atomicOr(&perm[sref >> 5], 1u << (sref & 31u));
    }
    __syncwarp();
    for (int w = lane; w < words; w += 32) row[w] = perm[w];
}

// One vertex: entry list, lockstep scan of the flagged samples, node / leaf steps until both stacks are empty, projection, row
// and (optional) visibility words.  W / occl: the warp's shared memory; lt_mask = (1 << lane) - 1; sgn = Condon-Shortley sign; the last four
// arguments are the work counters of an instrumented launch (COUNT).
template <int ORDER, bool TRACE, bool COUNT, bool DOP>
__device__ __forceinline__ void bake_wave_vertex(const BakeArgs &A, WaveShared &W, uint32_t *const occl, const uint32_t v, const int lane, const int S, const int words,
                                                 const unsigned lt_mask, const float sgn, unsigned long long &cand_tests,
                                                 unsigned long long &rays_scanned, uint32_t &node_visits, uint32_t &tri_tests) {
    constexpr int N2 = ORDER * ORDER;
    const uint32_t *need_row = (TRACE && A.need_bits) ? A.need_bits + (size_t)v * A.vis_words : nullptr;

    const float *pp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.pos) + (size_t)v * A.stride);
    const float *np = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.nrm) + (size_t)v * A.stride);
    const f3 N = mk3(__ldg(np), __ldg(np + 1), __ldg(np + 2));
    const f3 P = mk3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
    const Frame fr = make_frame(N);
    const f3 org = madd3(P, A.origin_eps, N);                       // raytracing.cpp:343

    for (int w = lane; w < words; w += 32) occl[w] = 0u;
    int n_cand = 0;
    if (TRACE) {
        n_cand = build_entry_list(A.nodes, org, N, W.el, lane);
#if PRT_WAVE_CULL
        entry_list_elevation_bounds(W.el, n_cand, fr, lane);
#endif
    }
    __syncwarp();

    if (TRACE) {
        int base = 0, nn = 0, ln = 0;             // warp-uniform: next sample, node-stack fill, leaf-stack fill
        int npend = 0;                            // warp-uniform: rays waiting in W.pend
        uint32_t m0 = 0u, m1 = 0u, m2 = 0u;       // candidate hits of the lane's scanned ray not yet queued
        uint32_t sproc = 0u;
        for (;;) {
            // ---- emit pending (ray, candidate) items while one more warp-wide append fits ----------------------
            bool pending = __any_sync(kFull, (m0 | m1 | m2) != 0u);
            while (pending && nn <= kNodeCap - 32 && ln <= kLeafCap - 32) {
                int k = -1;
                if (m0) { k = __ffs(m0) - 1; m0 &= m0 - 1u; }
                else if (m1) { k = 32 + __ffs(m1) - 1; m1 &= m1 - 1u; }
                else if (m2) { k = 64 + __ffs(m2) - 1; m2 &= m2 - 1u; }
                const bool has = k >= 0;
                const float4 g = W.el.cb[has ? k : 0];
                const uint32_t gx = __float_as_uint(g.z), gy = __float_as_uint(g.w);
                const bool leaf = has && gy <= 0x00FFFFFFu;
                const unsigned hb = __ballot_sync(kFull, has), lb = __ballot_sync(kFull, leaf), ib = hb & ~lb;
                if (leaf) W.lq[ln + __popc(lb & lt_mask)] = make_uint2(sproc | (gy << 16), gx);
                else if (has) W.nq[nn + __popc(ib & lt_mask)] = make_uint2(sproc, gx);
                ln += __popc(lb); nn += __popc(ib);
                pending = __any_sync(kFull, (m0 | m1 | m2) != 0u);
            }
            // ---- classify samples: a ray above the horizon of its azimuth bin is visible without any test ---------------
#ifdef PRT_WAVE_ROOM
            const bool room = nn <= PRT_WAVE_ROOM && ln <= PRT_WAVE_ROOM;              // absolute admission threshold (A/B builds)
#else
            const bool room = nn <= kNodeCap * PRT_WAVE_ROOM8 / 8 && ln <= kLeafCap * PRT_WAVE_ROOM8 / 8;
#endif
            if (!pending && room) {
                while (base < S && npend < 32) {
                    const int i = base + lane;
                    base += 32;
                    // samples the horizon pass proved visible are skipped (need bit = 0)
                    bool need = i < S;
                    if (need_row) need = need && ((__ldg(&need_row[i >> 5]) >> lane) & 1u);
                    const unsigned nb = __ballot_sync(kFull, need);
                    if (need) W.pend[npend + __popc(nb & lt_mask)] = (uint32_t)i;
                    npend += __popc(nb);
                }
                // ---- scan up to 32 waiting rays against the candidate boxes (lockstep) --------------------------------------
                if (npend >= 32 || (base >= S && npend > 0)) {
                    __syncwarp();
                    const int cnt = min(npend, 32);
                    npend -= cnt;
                    if (lane == 0) { PRT_WAVE_STAT(scan_steps, 1); PRT_WAVE_STAT(scan_lanes, cnt); }
#if PRT_WAVE_CULL
                    // candidates whose elevation bound lies below the lowest ray of the round are skipped by the whole warp
                    uint32_t ci = 0u;
                    float4 csmp = make_float4(0.f, 0.f, 1.f, 0.f);
                    if (lane < cnt) { ci = W.pend[npend + lane]; csmp = __ldg(&A.samples[ci]); }
                    float zmin = lane < cnt ? csmp.z : 2.0f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) zmin = fminf(zmin, __shfl_xor_sync(kFull, zmin, o));
                    if (lane < cnt) {
                        const f3 d = to_world(fr, mk3(csmp.x, csmp.y, csmp.z));   // raytracing.cpp:340
                        uint32_t cm[3];
                        scan_entry_list_culled(W.el, n_cand, rcp_box(d.x), rcp_box(d.y), rcp_box(d.z), zmin, cm);
                        m0 = cm[0]; m1 = cm[1]; m2 = cm[2];
                        sproc = ci;
                    }
#else
                    if (lane < cnt) {
                        const uint32_t i = W.pend[npend + lane];
                        const float4 smp = __ldg(&A.samples[i]);
                        const f3 d = to_world(fr, mk3(smp.x, smp.y, smp.z));   // raytracing.cpp:340
                        uint32_t cm[3];
                        scan_entry_list(W.el, n_cand, rcp_box(d.x), rcp_box(d.y), rcp_box(d.z), cm);
                        m0 = cm[0]; m1 = cm[1]; m2 = cm[2];
                        sproc = i;
                    }
#endif
                    if (COUNT) { cand_tests += (unsigned long long)n_cand * (unsigned long long)cnt; rays_scanned += (unsigned long long)cnt; }
                    __syncwarp();
                    continue;
                }
            }
            if (nn == 0 && ln == 0) { if (!pending && base >= S && npend == 0) break; else continue; }
            __syncwarp();
            if (ln >= 32 || nn == 0) {
                // ---- leaf step ------------------------------------------------------------------------------------
                const int cnt = min(ln, 32);
                ln -= cnt;
                if (lane == 0) { PRT_WAVE_STAT(leaf_steps, 1); PRT_WAVE_STAT(leaf_lanes, cnt); }
                if (lane < cnt) {
                    const uint2 it = W.lq[ln + lane];
                    const uint32_t oi = it.x & 0xFFFFu;
                    uint32_t bits = it.x >> 16;
                    if (bits && !((occl[oi >> 5] >> (oi & 31u)) & 1u)) {
                        // the first triangle and the sample are fetched together
                        uint32_t b = (uint32_t)__ffs(bits) - 1u;
                        bits &= bits - 1u;
                        const char *tp = reinterpret_cast<const char *>(A.tris + it.y + b);
                        u4 ta = ld16(tp), tb = ld16(tp + 16), tc = ld16(tp + 32);
                        const float4 smp = __ldg(&A.samples[oi]);
                        const f3 d = to_world(fr, mk3(smp.x, smp.y, smp.z));
                        for (;;) {
                            float t; uint32_t prim;
                            if (COUNT) tri_tests++;
                            if (tri_hit_regs(ta, tb, tc, org, d, 0.0f, INFINITY, false, t, prim)) {
                                atomicOr(&occl[oi >> 5], 1u << (oi & 31u));
                                break;
                            }
                            if (!bits) break;
                            b = (uint32_t)__ffs(bits) - 1u;
                            bits &= bits - 1u;
                            tp = reinterpret_cast<const char *>(A.tris + it.y + b);
                            ta = ld16(tp); tb = ld16(tp + 16); tc = ld16(tp + 32);
                        }
                    }
                }
            } else {
                // ---- node step ------------------------------------------------------------------------------------
                const int cnt = min(nn, 32);
                nn -= cnt;
                if (lane == 0) { PRT_WAVE_STAT(node_steps, 1); PRT_WAVE_STAT(node_lanes, cnt); }
                uint2 it = make_uint2(0u, 0u);
                bool has = lane < cnt;
                if (has) it = W.nq[nn + lane];
                __syncwarp();                   // all pops are done before anybody pushes
                uint32_t inner8 = 0u, leaf8 = 0u, child_base = 0u, tri_base = 0u, imask = 0u, meta_lo = 0u, meta_hi = 0u;
                f3 d = mk3(0.f, 0.f, 1.f);
                if (has) {
                    has = !((occl[it.x >> 5] >> (it.x & 31u)) & 1u);
                    if (has) {
                        const char *npn = reinterpret_cast<const char *>(A.nodes + it.y);
                        const u4 n0 = ld16(npn), n1 = ld16(npn + 16), n2 = ld16(npn + 32), n3 = ld16(npn + 48), n4 = ld16(npn + 64);
                        const float4 smp = __ldg(&A.samples[it.x]);       // issued together with the node fetch
                        d = to_world(fr, mk3(smp.x, smp.y, smp.z));
                        uint32_t hits;
                        if (DOP) {
                            // the node's fourth slab axis (32 bytes of a side array, fetched with the node)
                            const char *dpn = reinterpret_cast<const char *>(A.dops + it.y);
                            const u4 p0 = ld16(dpn), p1 = ld16(dpn + 16);
                            hits = node_slots_hit_t<false, true>(n0, n2, n3, n4, org, rcp_box(d.x), rcp_box(d.y), rcp_box(d.z), 0.0f, 0.0f, p0, p1, d);
                        } else hits = node_slots_hit(n0, n2, n3, n4, org, rcp_box(d.x), rcp_box(d.y), rcp_box(d.z));
                        imask = n0.w >> 24; child_base = n1.x; tri_base = n1.y; meta_lo = n1.z; meta_hi = n1.w;
                        inner8 = hits & imask; leaf8 = hits & ~imask;
                        PRT_WAVE_NODE_STUDY(A, it.y, org, d, inner8, leaf8);
                        if (COUNT) node_visits++;
                    }
                }
                // push hit children (any order: any-hit is order independent)
#if PRT_WAVE_SCAN_PUSH
                uint32_t tot;
                const uint32_t ex = warp_excl_scan_packed((uint32_t)__popc(inner8) | ((uint32_t)__popc(leaf8) << 16), lane, tot);
                if (nn + (int)(tot & 0xFFFFu) <= kNodeCap && ln + (int)(tot >> 16) <= kLeafCap) {
                    // everything fits (the common case): one packed warp scan gave every lane its write positions on both
                    // stacks, the lanes store their own children without further votes
                    int pi = nn + (int)(ex & 0xFFFFu), pl = ln + (int)(ex >> 16);
                    while (inner8) {
                        const uint32_t s = (uint32_t)__ffs(inner8) - 1u; inner8 &= inner8 - 1u;
                        W.nq[pi++] = make_uint2(it.x, child_base + __popc(imask & ((1u << s) - 1u)));
                    }
                    while (leaf8) {
                        const uint32_t s = (uint32_t)__ffs(leaf8) - 1u; leaf8 &= leaf8 - 1u;
                        const uint32_t meta = ((s < 4u ? meta_lo : meta_hi) >> (8u * (s & 3u))) & 0xFFu;
                        W.lq[pl++] = make_uint2(it.x | ((meta >> 5) << 16), tri_base + (meta & 31u));
                    }
                    nn += (int)(tot & 0xFFFFu); ln += (int)(tot >> 16);
                } else
#endif
                {
                while (__any_sync(kFull, inner8 != 0u)) {
                    const bool p = inner8 != 0u;
                    uint32_t child = 0u;
                    if (p) { const uint32_t s = (uint32_t)__ffs(inner8) - 1u; inner8 &= inner8 - 1u; child = child_base + __popc(imask & ((1u << s) - 1u)); }
                    const unsigned pb = __ballot_sync(kFull, p);
                    const int pos = nn + __popc(pb & lt_mask);
                    if (p) {
                        if (pos < kNodeCap) W.nq[pos] = make_uint2(it.x, child);
                        else {
                            // stack full: ordinary traversal of this subtree (rare)
                            PRT_WAVE_STAT(overflow_subtrees, 1);
                            if (fallback_subtree(A.nodes, A.tris, org, d, child, node_visits, tri_tests)) {
                                atomicOr(&occl[it.x >> 5], 1u << (it.x & 31u));
                                inner8 = 0u; leaf8 = 0u;
                            }
                        }
                    }
                    nn = min(nn + __popc(pb), kNodeCap);
                }
                while (__any_sync(kFull, leaf8 != 0u)) {
                    const bool p = leaf8 != 0u;
                    uint32_t tri0 = 0u, bits = 0u;
                    if (p) {
                        const uint32_t s = (uint32_t)__ffs(leaf8) - 1u; leaf8 &= leaf8 - 1u;
                        const uint32_t meta = ((s < 4u ? meta_lo : meta_hi) >> (8u * (s & 3u))) & 0xFFu;
                        tri0 = tri_base + (meta & 31u); bits = meta >> 5;
                    }
                    const unsigned pb = __ballot_sync(kFull, p);
                    const int pos = ln + __popc(pb & lt_mask);
                    if (p) {
                        if (pos < kLeafCap) W.lq[pos] = make_uint2(it.x | (bits << 16), tri0);
                        else {
                            PRT_WAVE_STAT(overflow_leaves, 1);
                            if (fallback_leaf(A.tris, org, d, tri0, bits, tri_tests)) {
                                atomicOr(&occl[it.x >> 5], 1u << (it.x & 31u));
                                leaf8 = 0u;
                            }
                        }
                    }
                    ln = min(ln + __popc(pb), kLeafCap);
                }
                }
            }
            __syncwarp();
        }
    }

    // ---- projection: L = Y_lm(dir) for every unoccluded sample (raytracing.cpp:226,257-261,348) ---------------
    float acc[N2];
#pragma unroll
    for (int k = 0; k < N2; k++) acc[k] = 0.f;
    for (int i = lane; i < S; i += 32) {
        if (TRACE && ((occl[i >> 5] >> lane) & 1u)) continue;           // one broadcast word per 32 samples
        const float4 smp = __ldg(&A.samples[i]);
        const f3 d = to_world(fr, mk3(smp.x, smp.y, smp.z));
        float y[N2];
        sh_eval<ORDER>(d.z, d.x, d.y, sgn, y);
#pragma unroll
        for (int k = 0; k < N2; k++) acc[k] += y[k];
    }
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < N2; k++) {
        const float s = warp_sum(acc[k]);
        if (lane == k) mine = s;
    }
    if (lane < N2) store_row(A, v, N2, lane, mine * A.inv_S);            // raytracing.cpp:350
    if (A.vis) {
        if (TRACE) write_vis_permuted(A.samples, occl, reinterpret_cast<uint32_t *>(W.nq), A.vis + (size_t)v * words, S, words, lane);
        else
        for (int w = lane; w < words; w += 32) {
            const int rem = S - 32 * w;
            const uint32_t valid = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
            A.vis[(size_t)v * words + w] = ~occl[w] & valid;
        }
    }
    __syncwarp();
}

}  // namespace

}  // namespace prt
