// bake_inter.cuh -- the per-vertex work of the interreflection kernel (bake_inter.cu): everything one persistent warp does for one
// vertex, as an inline device function, so that the same code also runs on the CPU test harness (tests/hostcheck: plain C++ against
// the warp emulator), where the CPU test-suite checks it against the oracle.  See bake_inter.cu for the design.
#pragma once
#include "kernels.h"
#include "traverse.cuh"
#include "entry_list.cuh"

#include <cuda_runtime.h>

namespace prt {

namespace {

#ifndef PRT_INTER_MINB
#define PRT_INTER_MINB 5
#endif
#ifndef PRT_INTER_CAP
#define PRT_INTER_CAP 192
#endif
#ifndef PRT_INTER_SHADE_MIN
#define PRT_INTER_SHADE_MIN 32
#endif
#ifndef PRT_INTER_ROOM8
#define PRT_INTER_ROOM8 4          // new primary rays are scanned while both stacks are at most ROOM8/8 full
#endif
constexpr int kSlots = 64;
constexpr int kCap = PRT_INTER_CAP;
constexpr uint32_t kFree = 0xFFFFFFFFu;
constexpr unsigned long long kNoHit = 0x7F800000FFFFFFFFull;        // (+inf, invalid prim)

struct InterShared {
    EntryList el;
    float4 od0[kSlots];                 // origin, tnear
    float4 od1[kSlots];                 // direction, unused
    unsigned long long best[kSlots];    // (closest t bits << 32) | prim
    uint32_t btri[kSlots];              // triangle slot of `best`
    int refc[kSlots];                   // outstanding work items of the slot's current segment
    uint32_t info[kSlots];              // processing index of the sample | segment << 24; kFree = empty
    uint2 nq[kCap];                     // (slot, node index)
    uint2 lq[kCap];                     // (slot | triangle bits << 16, first triangle)
};

__device__ __forceinline__ uint32_t node_slots_hit_range(const u4 n0, const u4 n2, const u4 n3, const u4 n4, const f3 o, const float idx,
                                                         const float idy, const float idz, const float tnear, const float tfar) {
    return node_slots_hit_t<true>(n0, n2, n3, n4, o, idx, idy, idz, tnear, tfar);          // traverse.cuh
}

// rare overflow path: ordinary closest-hit stack traversal of one subtree
__device__ __noinline__ void fallback_subtree_closest(const Node8 *nodes, const Tri48 *tris, InterShared &W, const uint32_t slot, const uint32_t child,
                                                      uint32_t &nv, uint32_t &nt) {
    const float4 a = W.od0[slot], b = W.od1[slot];
    Trav tr; tr.reset_counters();
    tr.init(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, INFINITY); tr.start_group(child, 0x80000000u);
    tr.run<false>(nodes, tris, 0, false);
    nv += tr.n_node_visits; nt += tr.n_tri_tests;
    if (tr.best_prim != 0xFFFFFFFFu) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(tr.best_t) << 32) | (unsigned long long)tr.best_prim;
        const unsigned long long old = atomicMin(&W.best[slot], key);
        if (key <= old) W.btri[slot] = tr.best_tri;       // alone in this path for the slot's key: see the note in the leaf step
    }
}
__device__ __noinline__ void fallback_leaf_closest(const Tri48 *tris, InterShared &W, const uint32_t slot, const uint32_t tri0, uint32_t bits, uint32_t &nt) {
    const float4 a = W.od0[slot], b = W.od1[slot];
    while (bits) {
        const uint32_t bb = (uint32_t)__ffs(bits) - 1u;
        bits &= bits - 1u;
        float t; uint32_t prim;
        nt++;
        if (tri_hit(tris, tri0 + bb, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, INFINITY, true, t, prim)) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)prim;
            const unsigned long long old = atomicMin(&W.best[slot], key);
            if (key <= old) W.btri[slot] = tri0 + bb;
        }
    }
}

// The triangle slot recorded next to the closest-hit key is written by whichever lane holds the final key; should a scheduling
// interleave ever leave a stale slot behind (its primitive id then differs from the key's), the hit is re-found by a plain traversal.
__device__ __noinline__ uint32_t repair_hit_triangle(const Node8 *nodes, const Tri48 *tris, const float4 a, const float4 b) {
    Trav tr; tr.reset_counters();
    tr.init(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, INFINITY); tr.start_root();
    tr.run<false>(nodes, tris, 0, false);
    return tr.best_tri;
}

// One vertex with n_need > 0 flagged samples (the caller skips the vertices the horizon pass finished).  lt_mask = (1 << lane) - 1,
// sgn = Condon-Shortley sign; the last four arguments are the work counters of an instrumented launch (COUNT).
template <int ORDER, bool COUNT>
__device__ __forceinline__ void bake_inter_vertex(const BakeArgs &A, InterShared &W, const uint32_t v, const int n_need, const int lane, const int S,
                                                  const int depth, const unsigned lt_mask, const float sgn, unsigned long long &cand_tests,
                                                  unsigned long long &rays_scanned, uint32_t &node_visits, uint32_t &tri_tests) {
    constexpr int N2 = ORDER * ORDER;
    const uint32_t *need_row = A.need_bits + (size_t)v * A.vis_words;

    const float *pp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.pos) + (size_t)v * A.stride);
    const float *np = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.nrm) + (size_t)v * A.stride);
    const f3 N = mk3(__ldg(np), __ldg(np + 1), __ldg(np + 2));
    const f3 P = mk3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
    const Frame fr = make_frame(N);
    const f3 org = madd3(P, A.origin_eps, N);                            // raytracing.cpp:343

    const int n_cand = build_entry_list(A.nodes, org, N, W.el, lane);
    W.info[lane] = kFree; W.info[lane + 32] = kFree;
    W.refc[lane] = 0; W.refc[lane + 32] = 0;
    __syncwarp();

    float acc[N2];
#pragma unroll
    for (int k = 0; k < N2; k++) acc[k] = 0.f;
    // primary rays the horizon pass proved free escape with weight 1 (raytracing.cpp:257-261)
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        if (i < S && !((__ldg(&need_row[base >> 5]) >> lane) & 1u)) {
            const float4 smp = __ldg(&A.samples[i]);
            const f3 dir = to_world(fr, mk3(smp.x, smp.y, smp.z));
            float y[N2];
            sh_eval<ORDER>(dir.z, dir.x, dir.y, sgn, y);
#pragma unroll
            for (int k = 0; k < N2; k++) acc[k] += y[k];
            if (A.vis) { const uint32_t sr = __float_as_uint(smp.w) & 0xFFFFFFu; atomicOr(&A.vis[(size_t)v * A.vis_words + (sr >> 5)], 1u << (sr & 31u)); }
        }
    }

    int nn = 0, ln = 0, nfree = kSlots, fetched = 0;      // warp-uniform
    int need_word = -1;
    uint32_t need_cur = 0u;
    uint32_t m0 = 0u, m1 = 0u, m2 = 0u, sslot = 0u;       // candidate hits of the lane's scanned primary ray not yet queued
    uint32_t guard = 0u;
    for (;;) {
        if (++guard > (1u << 24)) { if (lane == 0 && A.work) atomicAdd(&A.work[3], 1ull << 60); break; }     // never expected: bail out instead of hanging
        // ---- emit pending (slot, candidate) items while one more warp-wide append fits ------------------------------------
        bool pending = __any_sync(kFull, (m0 | m1 | m2) != 0u);
        while (pending && nn <= kCap - 32 && ln <= kCap - 32) {
            int k = -1;
            if (m0) { k = __ffs(m0) - 1; m0 &= m0 - 1u; }
            else if (m1) { k = 32 + __ffs(m1) - 1; m1 &= m1 - 1u; }
            else if (m2) { k = 64 + __ffs(m2) - 1; m2 &= m2 - 1u; }
            const bool has = k >= 0;
            const float4 g = W.el.cb[has ? k : 0];
            const uint32_t gx = __float_as_uint(g.z), gy = __float_as_uint(g.w);
            const bool leaf = has && gy <= 0x00FFFFFFu;
            const unsigned hb = __ballot_sync(kFull, has), lb = __ballot_sync(kFull, leaf), ib = hb & ~lb;
            if (leaf) W.lq[ln + __popc(lb & lt_mask)] = make_uint2(sslot | (gy << 16), gx);
            else if (has) W.nq[nn + __popc(ib & lt_mask)] = make_uint2(sslot, gx);
            ln += __popc(lb); nn += __popc(ib);
            pending = __any_sync(kFull, (m0 | m1 | m2) != 0u);
        }
        __syncwarp();

        // ---- shade: slots whose segment is finished (no outstanding item) ----------------------------------------------------
        // (batched: shading a couple of slots per iteration would run the long bounce code on a few lanes each time, so it waits
        //  until PRT_INTER_SHADE_MIN slots are finished or the stacks run low)
        bool shade_now = false;
        if (!pending && nn <= kCap - 64) {
            const unsigned f0 = __ballot_sync(kFull, W.info[lane] != kFree && W.refc[lane] == 0);
            const unsigned f1 = __ballot_sync(kFull, W.info[lane + 32] != kFree && W.refc[lane + 32] == 0);
            const int nfin = __popc(f0) + __popc(f1);
            shade_now = nfin > 0 && (nfin >= PRT_INTER_SHADE_MIN || nn + ln < 32);
        }
        if (shade_now) {
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t slot = (uint32_t)(lane + 32 * h);
                const uint32_t inf = W.info[slot];
                const bool fin = inf != kFree && W.refc[slot] == 0;
                if (!__any_sync(kFull, fin)) continue;
                bool push = false, freed = false;
                if (fin) {
                    const float4 a = W.od0[slot], b = W.od1[slot];
                    const f3 d = mk3(b.x, b.y, b.z);
                    const int seg = (int)(inf >> 24);
                    const float4 smp = __ldg(&A.samples[inf & 0xFFFFFFu]);
                    const uint32_t sidx = __float_as_uint(smp.w) & 0xFFFFFFu;
                    float Lw0 = 1.f, Lw1 = 1.f, Lw2 = 1.f;
                    for (int s = 0; s < seg; s++) { Lw0 = PRT_MUL(Lw0, A.albedo[0]); Lw1 = PRT_MUL(Lw1, A.albedo[1]); Lw2 = PRT_MUL(Lw2, A.albedo[2]); }
                    const unsigned long long key = W.best[slot];
                    freed = true;
                    if (key == kNoHit) {
                        // environment reached: L = Lw * Y_lm(dir), sh-space (z,x,y)  (raytracing.cpp:226,257-261)
                        float y[N2];
                        sh_eval<ORDER>(d.z, d.x, d.y, sgn, y);
#pragma unroll
                        for (int k = 0; k < N2; k++) acc[k] = fmaf(Lw0, y[k], acc[k]);
                        if (A.vis && seg == 0) atomicOr(&A.vis[(size_t)v * A.vis_words + (sidx >> 5)], 1u << (sidx & 31u));
                    } else if (seg < depth - 1) {
                        uint32_t bt = W.btri[slot];
                        if (ld16(reinterpret_cast<const char *>(A.tris + bt)).w != (uint32_t)key) bt = repair_hit_triangle(A.nodes, A.tris, a, b);
                        const char *tp = reinterpret_cast<const char *>(A.tris + bt);
                        const u4 b4 = ld16(tp + 16), c4 = ld16(tp + 32);
                        const f3 n = normalize3(cross3(mk3(PRT_U2F(b4.x), PRT_U2F(b4.y), PRT_U2F(b4.z)), mk3(PRT_U2F(c4.x), PRT_U2F(c4.y), PRT_U2F(c4.z))));   // :263-264
                        if (!(dot3(d, n) >= -1e-4f)) {                                      // :265
                            f3 pos = madd3(mk3(a.x, a.y, a.z), __uint_as_float((uint32_t)(key >> 32)), d);   // :266
                            float u, w;
                            rand2(A.seed, A.vid_base + global_row(A, v), sidx, (uint32_t)seg, 1u, u, w);    // :267
                            const f3 l = cosine_local(u, w);
                            const float pdf = PRT_DIV(l.z, kPiF);
                            const Frame fb = make_frame(n);
                            const f3 nd = to_world(fb, l);
                            if (!(pdf <= 1e-4f)) {                                          // :269
                                Lw0 = PRT_MUL(Lw0, A.albedo[0]); Lw1 = PRT_MUL(Lw1, A.albedo[1]); Lw2 = PRT_MUL(Lw2, A.albedo[2]);   // :271
                                const float sg = dot3(nd, n) < 0.0f ? -1.0f : 1.0f;         // :273
                                pos = madd3(pos, PRT_MUL(sg, A.bounce_eps), nd);            // :274
                                if (!(fmaxf(Lw0, fmaxf(Lw1, Lw2)) < 0.01f)) {               // :249 (checked at the top of the next iteration)
                                    W.od0[slot] = make_float4(pos.x, pos.y, pos.z, A.bounce_eps);   // :275
                                    W.od1[slot] = make_float4(nd.x, nd.y, nd.z, 0.f);
                                    W.best[slot] = kNoHit;
                                    W.info[slot] = (inf & 0xFFFFFFu) | ((uint32_t)(seg + 1) << 24);
                                    W.refc[slot] = 1;
                                    push = true; freed = false;
                                }
                            }
                        }
                    }
                    if (freed) W.info[slot] = kFree;
                }
                const unsigned pb = __ballot_sync(kFull, push), fb_ = __ballot_sync(kFull, freed);
                if (push) W.nq[nn + __popc(pb & lt_mask)] = make_uint2(slot, 0u);            // bounce rays start at the root
                nn += __popc(pb);
                nfree += __popc(fb_);
            }
            __syncwarp();
        }

        // ---- refill: the next flagged samples take free slots, 32 at a time (lockstep entry-list scan) ------------------------
        if (!pending && fetched < n_need && nn <= kCap * PRT_INTER_ROOM8 / 8 && ln <= kCap * PRT_INTER_ROOM8 / 8 && (nfree >= 32 || (nn == 0 && ln == 0 && nfree > 0))) {
            const int cnt = min(min(32, nfree), n_need - fetched);
            // the cnt next flagged samples, in processing order
            int taken = 0, my_k = -1;
            while (taken < cnt) {
                if (!need_cur) { need_cur = __ldg(&need_row[++need_word]); continue; }
                const int c = __popc(need_cur), take = min(c, cnt - taken);
                if (lane >= taken && lane < taken + take) my_k = need_word * 32 + (int)__fns(need_cur, 0u, lane - taken + 1);
                if (take == c) need_cur = 0u;
                else need_cur &= ~((2u << __fns(need_cur, 0u, take)) - 1u);
                taken += take;
            }
            fetched += cnt;
            // the cnt first free slots
            const unsigned f0 = __ballot_sync(kFull, W.info[lane] == kFree), f1 = __ballot_sync(kFull, W.info[lane + 32] == kFree);
            if (lane < cnt) {
                const int c0 = __popc(f0);
                const uint32_t slot = lane < c0 ? __fns(f0, 0u, lane + 1) : 32u + __fns(f1, 0u, lane - c0 + 1);
                const float4 smp = __ldg(&A.samples[my_k]);
                const f3 d = to_world(fr, mk3(smp.x, smp.y, smp.z));                       // raytracing.cpp:340
                uint32_t cm[3];
                scan_entry_list(W.el, n_cand, rcp_box(d.x), rcp_box(d.y), rcp_box(d.z), cm);
                m0 = cm[0]; m1 = cm[1]; m2 = cm[2];
                sslot = slot;
                W.od0[slot] = make_float4(org.x, org.y, org.z, 0.0f);
                W.od1[slot] = make_float4(d.x, d.y, d.z, 0.f);
                W.best[slot] = kNoHit;
                W.info[slot] = (uint32_t)my_k;                                            // segment 0
                W.refc[slot] = __popc(m0) + __popc(m1) + __popc(m2);
            }
            nfree -= cnt;
            if (COUNT) { cand_tests += (unsigned long long)n_cand * (unsigned long long)cnt; rays_scanned += (unsigned long long)cnt; }
            __syncwarp();
            continue;
        }
        if (nn == 0 && ln == 0) {
            if (!pending && fetched >= n_need && nfree == kSlots) break;
            continue;                                   // slots finished without items (or pending emission) are handled above
        }
        if (ln >= 32 || nn == 0) {
            // ---- leaf step ------------------------------------------------------------------------------------------------
            const int cnt = min(ln, 32);
            ln -= cnt;
            uint32_t slot = 0u, mytri = 0u;
            unsigned long long mykey = kNoHit;
            if (lane < cnt) {
                const uint2 it = W.lq[ln + lane];
                slot = it.x & 0xFFFFu;
                const float4 a = W.od0[slot], b = W.od1[slot];
                const f3 o = mk3(a.x, a.y, a.z), d = mk3(b.x, b.y, b.z);
                uint32_t bits = it.x >> 16;
                while (bits) {
                    const uint32_t bb = (uint32_t)__ffs(bits) - 1u;
                    bits &= bits - 1u;
                    float t; uint32_t prim;
                    if (COUNT) tri_tests++;
                    if (tri_hit(A.tris, it.y + bb, o, d, a.w, INFINITY, true, t, prim)) {
                        const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)prim;
                        if (key < mykey) { mykey = key; mytri = it.y + bb; }
                    }
                }
                if (mykey != kNoHit) atomicMin(&W.best[slot], mykey);
            }
            __syncwarp();
            // whoever holds the final key records its triangle (equal keys are the same triangle)
            if (lane < cnt) {
                if (mykey != kNoHit && W.best[slot] == mykey) W.btri[slot] = mytri;
                atomicSub(&W.refc[slot], 1);
            }
        } else {
            // ---- node step ------------------------------------------------------------------------------------------------
            const int cnt = min(nn, 32);
            nn -= cnt;
            uint2 it = make_uint2(0u, 0u);
            const bool has = lane < cnt;
            if (has) it = W.nq[nn + lane];
            __syncwarp();                   // all pops are done before anybody pushes
            uint32_t inner8 = 0u, leaf8 = 0u, child_base = 0u, tri_base = 0u, imask = 0u, meta_lo = 0u, meta_hi = 0u;
            int delta = 0;
            if (has) {
                const float4 a = W.od0[it.x], b = W.od1[it.x];
                const float tfar = __uint_as_float((uint32_t)(W.best[it.x] >> 32));
                const char *npn = reinterpret_cast<const char *>(A.nodes + it.y);
                const u4 n0 = ld16(npn), n1 = ld16(npn + 16), n2 = ld16(npn + 32), n3 = ld16(npn + 48), n4 = ld16(npn + 64);
                const uint32_t hits = node_slots_hit_range(n0, n2, n3, n4, mk3(a.x, a.y, a.z), rcp_box(b.x), rcp_box(b.y), rcp_box(b.z), a.w, tfar);
                imask = n0.w >> 24; child_base = n1.x; tri_base = n1.y; meta_lo = n1.z; meta_hi = n1.w;
                inner8 = hits & imask; leaf8 = hits & ~imask;
                if (COUNT) node_visits++;
                delta = __popc(hits) - 1;
                if (delta) atomicAdd(&W.refc[it.x], delta);       // before the pushes: the count never reaches zero early
            }
            __syncwarp();
            uint32_t tot;
            const uint32_t ex = warp_excl_scan_packed((uint32_t)__popc(inner8) | ((uint32_t)__popc(leaf8) << 16), lane, tot);
            if (nn + (int)(tot & 0xFFFFu) <= kCap && ln + (int)(tot >> 16) <= kCap) {
                // everything fits (the common case): one packed warp scan gave every lane its write positions on both stacks
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
            }
            while (__any_sync(kFull, inner8 != 0u)) {
                const bool p = inner8 != 0u;
                uint32_t child = 0u;
                if (p) { const uint32_t s = (uint32_t)__ffs(inner8) - 1u; inner8 &= inner8 - 1u; child = child_base + __popc(imask & ((1u << s) - 1u)); }
                const unsigned pb = __ballot_sync(kFull, p);
                const int pos = nn + __popc(pb & lt_mask);
                if (p) {
                    if (pos < kCap) W.nq[pos] = make_uint2(it.x, child);
                    else {
                        fallback_subtree_closest(A.nodes, A.tris, W, it.x, child, node_visits, tri_tests);
                        atomicSub(&W.refc[it.x], 1);
                    }
                }
                nn = min(nn + __popc(pb), kCap);
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
                    if (pos < kCap) W.lq[pos] = make_uint2(it.x | (bits << 16), tri0);
                    else {
                        fallback_leaf_closest(A.tris, W, it.x, tri0, bits, tri_tests);
                        atomicSub(&W.refc[it.x], 1);
                    }
                }
                ln = min(ln + __popc(pb), kCap);
            }
        }
        __syncwarp();
    }

    // reduce the lane partial sums and store row v (raytracing.cpp:350: /= S)
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < N2; k++) {
        const float s = warp_sum(acc[k]);
        if (lane == k) mine = s;
    }
    if (lane < N2) store_row(A, v, N2, lane, mine * A.inv_S);
    __syncwarp();
}

}  // namespace

}  // namespace prt
