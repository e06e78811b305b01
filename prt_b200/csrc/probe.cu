// probe.cu -- sm_100a kernels + C ABI for per-probe radiance-transfer capture and projection (BASELINE config 3).
//
//   prt_probe_capture   reference SH_volume::precompute, src/sh/volume.cpp:149-316: for every probe and every direction of a
//                       fixed set: closest hit (replaces the 64x64x6 G-buffer raster + glGetTexImage readback, :166-178,241-244),
//                       sky / back-face skips (:246-249), surfel cluster key (:205-223), transfer[probe][cluster] += SH9(d)*dOmega
//                       (:250-260); emits the CSR the viewer uploads (:265-295) and the surfel table (:301-312).
//   prt_probe_project   reference SH_volume::project_sh + precomp_projectSH.comp:32-143: CSR SpMV with surfel radiance,
//                       sinc window, Ramamoorthi-Hanrahan pack into 7 vec4 per probe.
//
// One CTA per probe (probes handed out in order by an atomic ticket): 512 threads trace the probe's <= 4096 rays (warp-local
// asynchronous wavefront, rays taken from a CTA-wide counter), the (cluster key, ray) pairs are bitonic-sorted in shared memory,
// every thread reduces a chunk of consecutive sorted slots (clusters spanning chunks are handed over through shared memory;
// fixed summation order) into the probe's slice of a staging area; an exclusive scan of the per-probe entry counts and a compaction pass (one
// warp per probe, coalesced copies) then give the probe-major, exactly sized CSR.  (A chained prefix inside the capture kernel
// serialised the probes.)
// Surfel ids are the rank of the cluster key among all keys (== std::map<std::array<int,4>> order of volume.cpp:204).
#include "../../include/prt_b200.h"
#include "abi_internal.h"
#include "bvh8.h"
#include "traverse.cuh"
#include "kernels.h"
#include "entry_list.cuh"

#include <algorithm>
#include <cuda_runtime.h>
#include <vector>

using namespace prt;


namespace {

constexpr int kMaxRays = 4096;
constexpr int kThreads = 512;
constexpr unsigned long long kInvalid = ~0ull;

struct CaptureArgs {
    const Node8 *nodes; const Tri48 *tris;
    const float *probe_pos; uint32_t n_probes;
    const float4 *dirs;          // xyz = direction, w = solid angle
    uint32_t n_dirs;
    int refill_thresh;
    int entry_list;              // per-probe entry list (tuning knob entry_list)
    const uint32_t *order;       // [n_dirs] trace slot -> ray index (directions sorted along a space-filling curve: coherent warps)
    uint32_t *ticket;            // [1] zeroed
    uint32_t *counts;            // [n_probes] entries (clusters) of each probe
    unsigned long long *ekeys;   // staging [n_probes][n_dirs]
    float *etransfer;            // staging [n_probes][n_dirs][9]
    float *eacc;                 // staging [n_probes][n_dirs][7] sum pos, sum normal, count
};

__device__ __forceinline__ bool cluster_key(f3 pos, f3 n, unsigned long long &key) {
    int dir = 0;
    const float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    if (ax > ay && ax > az) dir = n.x > 0 ? 0 : 1;
    if (ay > ax && ay > az) dir = n.y > 0 ? 2 : 3;
    if (az > ax && az > ay) dir = n.z > 0 ? 4 : 5;
    const float fx = floorf(pos.x), fy = floorf(pos.y), fz = floorf(pos.z);
    if (!(fabsf(fx) < 32768.f && fabsf(fy) < 32768.f && fabsf(fz) < 32768.f)) return false;
    const unsigned long long x = (unsigned long long)((int)fx + 32768), y = (unsigned long long)((int)fy + 32768), z = (unsigned long long)((int)fz + 32768);
    key = (x << 35) | (y << 19) | (z << 3) | (unsigned long long)dir;
    return true;
}

// ---- warp-local asynchronous closest-hit wavefront (same scheme as bake_inter.cu, one ray slot per lane) ---------------------
#ifndef PRT_PROBE_FIN_MIN
#define PRT_PROBE_FIN_MIN 16
#endif
constexpr int kWaveCap = 96;
constexpr unsigned long long kNoHit = 0x7F800000FFFFFFFFull;        // (+inf, invalid prim)
struct TraceShared {
    float4 dir[32];                     // direction of the slot's ray (origin = the probe, tnear = 0)
    unsigned long long best[32];        // (closest t bits << 32) | prim
    uint32_t btri[32];                  // triangle slot of `best`
    int refc[32];                       // outstanding work items
    uint32_t ray[32];                   // ray index, kFreeSlot = empty
    uint2 nq[kWaveCap];                 // (slot, node)
    uint2 lq[kWaveCap];                 // (slot | triangle bits << 16, first triangle)
};
constexpr uint32_t kFreeSlot = 0xFFFFFFFFu;

__device__ __forceinline__ float rcp_dir(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}
// 8 quantised child boxes of one node against the ray interval [0, tfar]; bit s = slot s hit (traverse.cuh)
__device__ __forceinline__ uint32_t node_hits(const u4 n0, const u4 n2, const u4 n3, const u4 n4, const f3 o, const float idx, const float idy,
                                              const float idz, const float tfar) {
    return node_slots_hit_t<true>(n0, n2, n3, n4, o, idx, idy, idz, 0.0f, tfar);
}
// rare paths, out of line: queue overflow (plain stack traversal of a subtree / leaf) and repair of a stale triangle slot
__device__ __noinline__ void wave_fallback_subtree(const Node8 *nodes, const Tri48 *tris, TraceShared &W, const f3 P, const uint32_t slot, const uint32_t child) {
    const float4 b = W.dir[slot];
    Trav tr; tr.reset_counters();
    tr.init(P, mk3(b.x, b.y, b.z), 0.0f, INFINITY); tr.start_group(child, 0x80000000u);
    tr.run<false>(nodes, tris, 0, false);
    if (tr.best_prim != 0xFFFFFFFFu) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(tr.best_t) << 32) | (unsigned long long)tr.best_prim;
        if (key <= atomicMin(&W.best[slot], key)) W.btri[slot] = tr.best_tri;
    }
}
__device__ __noinline__ void wave_fallback_leaf(const Tri48 *tris, TraceShared &W, const f3 P, const uint32_t slot, const uint32_t tri0, uint32_t bits) {
    const float4 b = W.dir[slot];
    while (bits) {
        const uint32_t bb = (uint32_t)__ffs(bits) - 1u;
        bits &= bits - 1u;
        float t; uint32_t prim;
        if (tri_hit(tris, tri0 + bb, P, mk3(b.x, b.y, b.z), 0.0f, INFINITY, true, t, prim)) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)prim;
            if (key <= atomicMin(&W.best[slot], key)) W.btri[slot] = tri0 + bb;
        }
    }
}
__device__ __noinline__ uint32_t wave_repair_triangle(const Node8 *nodes, const Tri48 *tris, const f3 P, const float4 b) {
    Trav tr; tr.reset_counters();
    tr.init(P, mk3(b.x, b.y, b.z), 0.0f, INFINITY); tr.start_root();
    tr.run<false>(nodes, tris, 0, false);
    return tr.best_tri;
}

__global__ void __launch_bounds__(kThreads, 2) probe_capture_kernel(const CaptureArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(smem);          // [4096] (cluster key << 12) | ray
    float *tt = reinterpret_cast<float *>(sk + kMaxRays);                          // [4096] hit distance by ray
    uint32_t *tri = reinterpret_cast<uint32_t *>(tt + kMaxRays);                   // [4096] hit triangle slot by ray
    float *s_lead = reinterpret_cast<float *>(tri + kMaxRays);                      // [kThreads][16] partial sums handed to an earlier chunk
    // s_counts / s_through (reduction phase) share their 4 KB with the probe's entry list (trace phase)
    __shared__ __align__(16) unsigned char s_aux[2 * kThreads * sizeof(uint32_t)];
    static_assert(sizeof(EntryList) <= sizeof(s_aux), "the entry list aliases the reduction counters");
    uint32_t *const s_counts = reinterpret_cast<uint32_t *>(s_aux), *const s_through = s_counts + kThreads;
    EntryList &EL = *reinterpret_cast<EntryList *>(s_aux);
    __shared__ uint32_t s_probe, s_next;
    __shared__ int s_ncand;
    const int tid = threadIdx.x;

    for (;;) {
        if (tid == 0) s_probe = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t p = s_probe;
        if (p >= A.n_probes) return;
        const f3 P = mk3(A.probe_pos[3 * p], A.probe_pos[3 * p + 1], A.probe_pos[3 * p + 2]);

        // ---- trace: asynchronous wavefront, one ray slot per lane -------------------------------------------------------------
        // Rays are handed out from a CTA-wide counter (Morton order of their directions).  Work is (slot, node) / (slot, leaf) items
        // on two warp-local stacks: a node step decodes one node per lane and pushes the child boxes the ray hits inside
        // [0, closest t], a leaf step runs the pinned triangle tests and lowers the slot's (t, prim) key with a 64-bit atomicMin;
        // a slot whose outstanding-item count reaches zero is finished, stored and refilled.  Per-ray stacks ran this phase at
        // 12.5 of 32 active lanes (profiles/r1_probe_capture_ncu_summary.txt).
        // Per-origin entry list (entry_list.cuh, as in the vertex bakes): all rays of a probe share its position, so the chain of nodes
        // containing it is expanded ONCE, by warp 0, into candidate boxes; a ray then starts at the candidates it hits instead of repeating
        // the descent from the root.  No tangent plane here (normal 0: nothing is culled).
        for (int i = tid; i < kMaxRays; i += kThreads) sk[i] = kInvalid;
        if (tid == 0) s_next = 0u;
        if (tid < 32) {
            const int nc = A.entry_list ? build_entry_list(A.nodes, P, mk3(0.f, 0.f, 0.f), EL, tid) : 0;
            if (tid == 0) s_ncand = nc;
        }
        __syncthreads();
        const int n_cand = s_ncand;
        {
            const int lane = tid & 31;
            const unsigned lt_mask = (1u << lane) - 1u;
            TraceShared &W = reinterpret_cast<TraceShared *>(s_lead)[tid >> 5];      // aliases the hand-over area of the reduction
            W.ray[lane] = kFreeSlot; W.refc[lane] = 0;
            __syncwarp();
            int nn = 0, ln = 0;
            bool exhausted = false;
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u;                                  // candidate hits of the lane's new ray not yet queued
            for (uint32_t guard = 0; guard < (1u << 22); guard++) {               // (bounded: a logic error must not hang the device)
                // emit pending (slot, candidate) items while one more warp-wide append fits
                bool pending = __any_sync(0xFFFFFFFFu, (m0 | m1 | m2) != 0u);
                while (pending && nn <= kWaveCap - 32 && ln <= kWaveCap - 32) {
                    int k = -1;
                    if (m0) { k = __ffs(m0) - 1; m0 &= m0 - 1u; }
                    else if (m1) { k = 32 + __ffs(m1) - 1; m1 &= m1 - 1u; }
                    else if (m2) { k = 64 + __ffs(m2) - 1; m2 &= m2 - 1u; }
                    const bool hasc = k >= 0;
                    const float4 g = EL.cb[hasc ? k : 0];
                    const uint32_t gx = __float_as_uint(g.z), gy = __float_as_uint(g.w);
                    const bool leafc = hasc && gy <= 0x00FFFFFFu;
                    const unsigned hb = __ballot_sync(0xFFFFFFFFu, hasc), lb = __ballot_sync(0xFFFFFFFFu, leafc), ib = hb & ~lb;
                    if (leafc) W.lq[ln + __popc(lb & lt_mask)] = make_uint2((uint32_t)lane | (gy << 16), gx);
                    else if (hasc) W.nq[nn + __popc(ib & lt_mask)] = make_uint2((uint32_t)lane, gx);
                    ln += __popc(lb); nn += __popc(ib);
                    pending = __any_sync(0xFFFFFFFFu, (m0 | m1 | m2) != 0u);
                }
                __syncwarp();
                // finished slots: store the hit, free the slot
                const uint32_t myray = W.ray[lane];
                const bool fin = myray != kFreeSlot && W.refc[lane] == 0;
                const int nfin = __popc(__ballot_sync(0xFFFFFFFFu, fin));
                if (fin && (nfin >= PRT_PROBE_FIN_MIN || nn + ln < 32)) {          // batched: the store code runs on many lanes at once
                    const unsigned long long key = W.best[lane];
                    if (key != kNoHit) {                                              // sky otherwise: volume.cpp:246
                        const float4 b = W.dir[lane];
                        uint32_t bt = W.btri[lane];
                        if (ld16(reinterpret_cast<const char *>(A.tris + bt)).w != (uint32_t)key) bt = wave_repair_triangle(A.nodes, A.tris, P, b);
                        const char *tp = reinterpret_cast<const char *>(A.tris + bt);
                        const u4 b4 = ld16(tp + 16), c4 = ld16(tp + 32);
                        const f3 n = normalize3(cross3(mk3(PRT_U2F(b4.x), PRT_U2F(b4.y), PRT_U2F(b4.z)), mk3(PRT_U2F(c4.x), PRT_U2F(c4.y), PRT_U2F(c4.z))));
                        const float t = __uint_as_float((uint32_t)(key >> 32));
                        const f3 pos = madd3(P, t, mk3(b.x, b.y, b.z));
                        unsigned long long ck;
                        if (!(dot3(sub3(pos, P), n) > 0.0f) && cluster_key(pos, n, ck)) {   // back face: volume.cpp:249
                            sk[myray] = (ck << 12) | (unsigned long long)myray;
                            tt[myray] = t; tri[myray] = bt;
                        }
                    }
                    W.ray[lane] = kFreeSlot;
                }
                __syncwarp();
                // refill free slots while rays remain and one warp-wide push fits
                const unsigned freeb = __ballot_sync(0xFFFFFFFFu, W.ray[lane] == kFreeSlot);
                if (freeb && !exhausted && !pending && nn <= kWaveCap - 32) {
                    uint32_t base = 0u;
                    if (lane == 0) base = atomicAdd(&s_next, (uint32_t)__popc(freeb));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    const uint32_t slot_rank = (uint32_t)__popc(freeb & lt_mask);
                    const bool take = ((freeb >> lane) & 1u) && base + slot_rank < A.n_dirs;
                    if (take) {
                        const uint32_t r = __ldg(&A.order[base + slot_rank]);          // results are stored by ray index, so the
                        const float4 dw = __ldg(&A.dirs[r]);                           // trace order does not change them
                        W.dir[lane] = make_float4(dw.x, dw.y, dw.z, 0.f);
                        W.best[lane] = kNoHit; W.ray[lane] = r;
                        if (n_cand > 0) {
                            uint32_t cm[3];
                            scan_entry_list(EL, n_cand, rcp_dir(dw.x), rcp_dir(dw.y), rcp_dir(dw.z), cm);
                            m0 = cm[0]; m1 = cm[1]; m2 = cm[2];
                            W.refc[lane] = __popc(m0) + __popc(m1) + __popc(m2);     // 0: the ray misses everything (sky), finished at once
                        } else W.refc[lane] = 1;
                    }
                    if (n_cand == 0) {
                        const unsigned tb = __ballot_sync(0xFFFFFFFFu, take);
                        if (take) W.nq[nn + __popc(tb & lt_mask)] = make_uint2((uint32_t)lane, 0u);      // no entry list: start at the root
                        nn += __popc(tb);
                    }
                    exhausted = base + (uint32_t)__popc(freeb) >= A.n_dirs;
                }
                __syncwarp();
                if (nn == 0 && ln == 0) {
                    if (__any_sync(0xFFFFFFFFu, (m0 | m1 | m2) != 0u)) continue;            // items of the rays just scanned are emitted first
                    if (exhausted && __all_sync(0xFFFFFFFFu, W.ray[lane] == kFreeSlot)) break;
                    if (exhausted && !__any_sync(0xFFFFFFFFu, W.ray[lane] != kFreeSlot && W.refc[lane] != 0)) continue;   // only finished slots left
                    continue;
                }
                if (ln >= 32 || nn == 0) {
                    // ---- leaf step ----
                    const int cnt = min(ln, 32);
                    ln -= cnt;
                    uint32_t slot = 0u, mytri = 0u;
                    unsigned long long mykey = kNoHit;
                    if (lane < cnt) {
                        const uint2 it = W.lq[ln + lane];
                        slot = it.x & 0xFFFFu;
                        const float4 b = W.dir[slot];
                        const f3 d = mk3(b.x, b.y, b.z);
                        uint32_t bits = it.x >> 16;
                        while (bits) {
                            const uint32_t bb = (uint32_t)__ffs(bits) - 1u;
                            bits &= bits - 1u;
                            float t; uint32_t prim;
                            if (tri_hit(A.tris, it.y + bb, P, d, 0.0f, INFINITY, true, t, prim)) {
                                const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)prim;
                                if (key < mykey) { mykey = key; mytri = it.y + bb; }
                            }
                        }
                        if (mykey != kNoHit) atomicMin(&W.best[slot], mykey);
                    }
                    __syncwarp();
                    if (lane < cnt) {
                        if (mykey != kNoHit && W.best[slot] == mykey) W.btri[slot] = mytri;      // the holder of the final key records its triangle
                        atomicSub(&W.refc[slot], 1);
                    }
                } else {
                    // ---- node step ----
                    const int cnt = min(nn, 32);
                    nn -= cnt;
                    uint2 it = make_uint2(0u, 0u);
                    const bool has = lane < cnt;
                    if (has) it = W.nq[nn + lane];
                    __syncwarp();                   // all pops are done before anybody pushes
                    uint32_t inner8 = 0u, leaf8 = 0u, child_base = 0u, tri_base = 0u, imask = 0u, meta_lo = 0u, meta_hi = 0u;
                    if (has) {
                        const float4 b = W.dir[it.x];
                        const float tfar = __uint_as_float((uint32_t)(W.best[it.x] >> 32));
                        const char *npn = reinterpret_cast<const char *>(A.nodes + it.y);
                        const u4 n0 = ld16(npn), n1 = ld16(npn + 16), n2 = ld16(npn + 32), n3 = ld16(npn + 48), n4 = ld16(npn + 64);
                        const uint32_t hits = node_hits(n0, n2, n3, n4, P, rcp_dir(b.x), rcp_dir(b.y), rcp_dir(b.z), tfar);
                        imask = n0.w >> 24; child_base = n1.x; tri_base = n1.y; meta_lo = n1.z; meta_hi = n1.w;
                        inner8 = hits & imask; leaf8 = hits & ~imask;
                        const int delta = __popc(hits) - 1;
                        if (delta) atomicAdd(&W.refc[it.x], delta);       // before the pushes: the count never reaches zero early
                    }
                    __syncwarp();
                    {
                        // positions on both stacks from one packed warp scan; when everything fits the lanes store their own children
                        uint32_t incl = (uint32_t)__popc(inner8) | ((uint32_t)__popc(leaf8) << 16);
                        const uint32_t mine = incl;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
                        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31), ex = incl - mine;
                        if (nn + (int)(tot & 0xFFFFu) <= kWaveCap && ln + (int)(tot >> 16) <= kWaveCap) {
                            int pi = nn + (int)(ex & 0xFFFFu), pl = ln + (int)(ex >> 16);
                            while (inner8) {
                                const uint32_t sl = (uint32_t)__ffs(inner8) - 1u; inner8 &= inner8 - 1u;
                                W.nq[pi++] = make_uint2(it.x, child_base + __popc(imask & ((1u << sl) - 1u)));
                            }
                            while (leaf8) {
                                const uint32_t sl = (uint32_t)__ffs(leaf8) - 1u; leaf8 &= leaf8 - 1u;
                                const uint32_t meta = ((sl < 4u ? meta_lo : meta_hi) >> (8u * (sl & 3u))) & 0xFFu;
                                W.lq[pl++] = make_uint2(it.x | ((meta >> 5) << 16), tri_base + (meta & 31u));
                            }
                            nn += (int)(tot & 0xFFFFu); ln += (int)(tot >> 16);
                        }
                    }
                    while (__any_sync(0xFFFFFFFFu, inner8 != 0u)) {
                        const bool pp = inner8 != 0u;
                        uint32_t child = 0u;
                        if (pp) { const uint32_t sl = (uint32_t)__ffs(inner8) - 1u; inner8 &= inner8 - 1u; child = child_base + __popc(imask & ((1u << sl) - 1u)); }
                        const unsigned pb = __ballot_sync(0xFFFFFFFFu, pp);
                        const int pos = nn + __popc(pb & lt_mask);
                        if (pp) {
                            if (pos < kWaveCap) W.nq[pos] = make_uint2(it.x, child);
                            else { wave_fallback_subtree(A.nodes, A.tris, W, P, it.x, child); atomicSub(&W.refc[it.x], 1); }
                        }
                        nn = min(nn + __popc(pb), kWaveCap);
                    }
                    while (__any_sync(0xFFFFFFFFu, leaf8 != 0u)) {
                        const bool pp = leaf8 != 0u;
                        uint32_t tri0 = 0u, bits = 0u;
                        if (pp) {
                            const uint32_t sl = (uint32_t)__ffs(leaf8) - 1u; leaf8 &= leaf8 - 1u;
                            const uint32_t meta = ((sl < 4u ? meta_lo : meta_hi) >> (8u * (sl & 3u))) & 0xFFu;
                            tri0 = tri_base + (meta & 31u); bits = meta >> 5;
                        }
                        const unsigned pb = __ballot_sync(0xFFFFFFFFu, pp);
                        const int pos = ln + __popc(pb & lt_mask);
                        if (pp) {
                            if (pos < kWaveCap) W.lq[pos] = make_uint2(it.x | (bits << 16), tri0);
                            else { wave_fallback_leaf(A.tris, W, P, it.x, tri0, bits); atomicSub(&W.refc[it.x], 1); }
                        }
                        ln = min(ln + __popc(pb), kWaveCap);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();

        // ---- bitonic sort of 4096 keys in shared memory ---------------------------------------------------------------------
        // One compare-exchange per thread and step.  Warp w owns pairs [128 w, 128 w + 128): for strides j <= 128 they all lie in
        // its own 256-key window, so only the 10 steps with j >= 256 need a block-wide barrier, the other 68 a __syncwarp.
        {
            constexpr int kPairsPerWarp = (kMaxRays / 2) / (kThreads / 32);
            const int wbase = (tid >> 5) * kPairsPerWarp + (tid & 31);
            bool prev_wide = false;                                              // (the barrier after the trace phase covers step 1)
            for (int k = 2; k <= kMaxRays; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const bool wide = j >= 2 * kPairsPerWarp;
                    if (wide || prev_wide) __syncthreads(); else __syncwarp();
                    prev_wide = wide;
#pragma unroll
                    for (int q = 0; q < kPairsPerWarp / 32; q++) {
                        const int t = wbase + 32 * q;
                        const int i = 2 * t - (t & (j - 1)), ixj = i + j;
                        const unsigned long long a = sk[i], b = sk[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
                    }
                }
            __syncthreads();
        }

        // ---- segment heads, ranks --------------------------------------------------------------------------------------------
        constexpr int kChunk = kMaxRays / kThreads;
        uint32_t heads = 0;
        for (int q = 0; q < kChunk; q++) {
            const int i = tid * kChunk + q;
            const unsigned long long a = sk[i];
            if (a != kInvalid && (i == 0 || (sk[i - 1] >> 12) != (a >> 12))) heads++;
        }
        s_counts[tid] = heads;
        __syncthreads();
        if (tid == 0) {
            uint32_t run = 0;
            for (int i = 0; i < kThreads; i++) { const uint32_t c = s_counts[i]; s_counts[i] = run; run += c; }
            A.counts[p] = run;
        }
        __syncthreads();
        const unsigned long long start = (unsigned long long)p * A.n_dirs;          // the probe's staging slice

        // ---- per-cluster reduction (volume.cpp:250-260) ------------------------------------------------------------------------
        // Every thread walks its kChunk consecutive sorted slots: clusters that end inside the chunk are written directly, the
        // slots before the first head ("lead") belong to a cluster begun in an earlier chunk and are handed over through shared
        // memory; the thread whose last cluster runs past its chunk collects the leads of the following chunks.  All lanes stay
        // busy whatever the cluster sizes, and the summation order is fixed by the layout (deterministic).
        auto flush = [&](const float *acc, const unsigned long long key, const uint32_t rank) {
            const unsigned long long e = start + rank;
            A.ekeys[e] = key;
#pragma unroll
            for (int c = 0; c < 9; c++) A.etransfer[9 * e + c] = acc[c];
#pragma unroll
            for (int c = 0; c < 7; c++) A.eacc[7 * e + c] = acc[9 + c];
        };
        float lead[16], acc[16];
#pragma unroll
        for (int c = 0; c < 16; c++) { lead[c] = 0.f; acc[c] = 0.f; }
        bool open = false, ended = false;                 // open: a cluster of this chunk is being summed; ended: ran into the invalid tail
        unsigned long long open_key = 0ull;
        uint32_t rank = s_counts[tid];
        for (int q = 0; q < kChunk; q++) {
            const int i = tid * kChunk + q;
            const unsigned long long a = sk[i];
            if (a == kInvalid) { ended = true; break; }
            const bool head = (i == 0) || ((sk[i - 1] >> 12) != (a >> 12));
            const uint32_t r = (uint32_t)(a & 0xFFFull);
            const float4 dw = __ldg(&A.dirs[r]);
            const f3 pos = madd3(P, tt[r], mk3(dw.x, dw.y, dw.z));
            const f3 tp = sub3(pos, P);
            const float m = fmaxf(fabsf(tp.x), fmaxf(fabsf(tp.y), fabsf(tp.z)));
            const f3 cc = mk3(PRT_DIV(tp.x, m), PRT_DIV(tp.y, m), PRT_DIV(tp.z, m));             // volume.cpp:250
            const f3 d = normalize3(mk3(cc.z, cc.x, cc.y));                                    // :255-256
            float c16[16];
            sh_eval<3>(d.x, d.y, d.z, 1.0f, c16);
#pragma unroll
            for (int c = 0; c < 9; c++) c16[c] *= dw.w;                                        // :260
            const char *tp48 = reinterpret_cast<const char *>(A.tris + tri[r]);
            const u4 b4 = ld16(tp48 + 16), c4 = ld16(tp48 + 32);
            const f3 n = normalize3(cross3(mk3(PRT_U2F(b4.x), PRT_U2F(b4.y), PRT_U2F(b4.z)), mk3(PRT_U2F(c4.x), PRT_U2F(c4.y), PRT_U2F(c4.z))));
            c16[9] = pos.x; c16[10] = pos.y; c16[11] = pos.z; c16[12] = n.x; c16[13] = n.y; c16[14] = n.z; c16[15] = 1.f;
            if (head) {
                if (open) flush(acc, open_key, rank++);
                open = true; open_key = a >> 12;
#pragma unroll
                for (int c = 0; c < 16; c++) acc[c] = c16[c];
            } else if (open) {
#pragma unroll
                for (int c = 0; c < 16; c++) acc[c] += c16[c];
            } else {
#pragma unroll
                for (int c = 0; c < 16; c++) lead[c] += c16[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 16; c++) s_lead[tid * 16 + c] = lead[c];
        s_through[tid] = (!open && !ended) ? 1u : 0u;      // the whole chunk continues an earlier cluster and runs on
        __syncthreads();
        if (open) {
            if (!ended) {
                for (int t = tid + 1; t < kThreads; t++) {
#pragma unroll
                    for (int c = 0; c < 16; c++) acc[c] += s_lead[t * 16 + c];
                    if (!s_through[t]) break;
                }
            }
            flush(acc, open_key, rank);
        }
        __syncthreads();
    }
}

// ---- exact CSR: exclusive scan of the per-probe counts (one CTA; probes are few) + compaction of the staging slices ---------
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t *counts, uint32_t n, unsigned long long *offsets) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (n + 1023u) / 1024u, lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += counts[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) { const unsigned long long c = part[i]; part[i] = run; run += c; }
        offsets[n] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) { offsets[i] = run; run += counts[i]; }
}
__global__ void __launch_bounds__(256) compact_csr_kernel(const uint32_t *counts, const unsigned long long *offsets, uint32_t n_probes, uint32_t n_dirs,
                                                          const unsigned long long *skeys, const float *stransfer, const float *sacc,
                                                          unsigned long long *ekeys, float *etransfer, float *eacc, uint32_t *range) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n_probes) return;
    const uint32_t cnt = counts[p];
    const unsigned long long src = (unsigned long long)p * n_dirs, dst = offsets[p];
    if (lane == 0) { range[2 * p] = (uint32_t)dst; range[2 * p + 1] = (uint32_t)(dst + cnt); }
    for (uint32_t i = lane; i < cnt; i += 32) ekeys[dst + i] = skeys[src + i];
    for (uint32_t i = lane; i < 9u * cnt; i += 32) etransfer[9 * dst + i] = stransfer[9 * src + i];
    for (uint32_t i = lane; i < 7u * cnt; i += 32) eacc[7 * dst + i] = sacc[7 * src + i];
}

// ---- global surfel ids: hash-set dedupe of cluster keys, host sort of the (few) distinct keys, rank lookup ----------------
__device__ __forceinline__ uint32_t hash64(unsigned long long k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; return (uint32_t)k; }
__global__ void hash_insert_kernel(const unsigned long long *keys, unsigned long long n, unsigned long long *table, uint32_t mask, int *overflow) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    uint32_t h = hash64(k) & mask;
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const unsigned long long old = atomicCAS(&table[h], kInvalid, k);
        if (old == kInvalid || old == k) return;
        h = (h + 1) & mask;
    }
    *overflow = 1;
}
__global__ void hash_compact_kernel(const unsigned long long *table, uint32_t size, unsigned long long *out, uint32_t *count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const unsigned long long k = table[i];
    if (k != kInvalid) out[atomicAdd(count, 1u)] = k;
}
__global__ void assign_ids_kernel(const unsigned long long *ekeys, unsigned long long n, const unsigned long long *sorted, uint32_t n_prim,
                                  const float *eacc, uint32_t *ids, double *sacc) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = ekeys[i];
    uint32_t lo = 0, hi = n_prim;
    while (lo + 1 < hi) { const uint32_t mid = (lo + hi) >> 1; if (sorted[mid] <= k) lo = mid; else hi = mid; }
    ids[i] = lo;
#pragma unroll
    for (int c = 0; c < 7; c++) atomicAdd(&sacc[7 * (size_t)lo + c], (double)eacc[7 * i + c]);
}
__global__ void finalize_surfels_kernel(const double *sacc, uint32_t n_prim, float *surfels) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_prim) return;
    const double cnt = sacc[7 * (size_t)s + 6];
    const double nx = sacc[7 * (size_t)s + 3] / cnt, ny = sacc[7 * (size_t)s + 4] / cnt, nz = sacc[7 * (size_t)s + 5] / cnt;
    const double il = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);                                      // volume.cpp:307-308
    float *o = surfels + 6 * (size_t)s;
    o[0] = (float)(sacc[7 * (size_t)s] / cnt); o[1] = (float)(sacc[7 * (size_t)s + 1] / cnt); o[2] = (float)(sacc[7 * (size_t)s + 2] / cnt);
    o[3] = (float)(nx * il); o[4] = (float)(ny * il); o[5] = (float)(nz * il);
}

// ---- projection: one warp per probe (precomp_projectSH.comp:51-139) -----------------------------------------------------------
__global__ void __launch_bounds__(256) probe_project_kernel(const uint32_t *range, const uint32_t *ids, const float *transfer, const float4 *radiance,
                                                            uint32_t n_probes, float4 *out) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n_probes) return;
    float L[27];
#pragma unroll
    for (int k = 0; k < 27; k++) L[k] = 0.f;
    for (uint32_t i = range[2 * p] + lane; i < range[2 * p + 1]; i += 32) {
        const float4 rad = __ldg(&radiance[ids[i]]);
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const float t = __ldg(&transfer[9 * (size_t)i + k]);
            L[3 * k] += t * rad.x; L[3 * k + 1] += t * rad.y; L[3 * k + 2] += t * rad.z;
        }
    }
#pragma unroll
    for (int k = 0; k < 27; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) L[k] += __shfl_xor_sync(0xFFFFFFFFu, L[k], o);
    if (lane == 0) {
        const float PI = 3.14159265359f;
        const float w1 = 3.f / PI * sinf(PI / 3), w2 = 3.f / 2 / PI * sinf(2 * PI / 3);                // :104-113
#pragma unroll
        for (int k = 3; k < 12; k++) L[k] *= w1;
#pragma unroll
        for (int k = 12; k < 27; k++) L[k] *= w2;
        const float c1 = 0.429043f, c2 = 0.511664f, c3 = 0.743125f, c4 = 0.886227f, c5 = 0.247708f;   // :23
        float4 *o = out + 7 * (size_t)p;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            o[ch] = make_float4(2 * c2 * L[9 + ch], 2 * c2 * L[3 + ch], 2 * c2 * L[6 + ch], c4 * L[ch] - c5 * L[18 + ch]);      // :118-128
            o[3 + ch] = make_float4(2 * c1 * L[12 + ch], 2 * c1 * L[21 + ch], 2 * c1 * L[15 + ch], c3 * L[18 + ch]);           // :130-136
        }
        o[6] = make_float4(c1 * L[24], c1 * L[25], c1 * L[26], 1.0f);                                                          // :138
    }
}

}  // namespace


#define PB_TRY(expr)                                                                                                    \
    do {                                                                                                                \
        cudaError_t e_ = (expr);                                                                                        \
        if (e_ != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

void prt_csr_destroy(prt_csr *c) {
    if (!c) return;
    cudaSetDevice(prt_ctx_device(c->ctx));
    cudaStream_t st = prt_ctx_stream(c->ctx);
    void *arrays[6] = { c->range, c->ids, c->transfer, c->surfels, c->keys, c->sums };
    for (void *a : arrays) if (a) cudaFreeAsync(a, st);          // stream-ordered: no device-wide synchronisation per array
    delete c;
}

// temporaries and results come from the device's stream-ordered memory pool (cudaMallocAsync on the context's stream): no
// device-wide synchronisation per free and no allocator lock shared by the host threads of a multi-GPU capture (group.cu)
#define PB_MALLOC(pp, n) cudaMallocAsync((void **)(pp), (n), st)
#define PB_FREE(p) do { if (p) cudaFreeAsync((p), st); } while (0)
int prt_probe_capture(prt_scene *scene, const float *probe_pos, uint32_t n_probes, const float *dirs, const float *weights,
                      uint32_t n_dirs, prt_csr **out) {
    if (!scene || !probe_pos || !dirs || !weights || !out || n_probes == 0 || n_dirs == 0) return prt_set_error(PRT_ERR_INVALID, "prt_probe_capture: bad argument");
    *out = nullptr;
    if (n_dirs > (uint32_t)kMaxRays) return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_probe_capture: at most 4096 directions per probe");
    const prt_scene_view sv = prt_scene_get_view(scene);
    PB_TRY(cudaSetDevice(prt_ctx_device(sv.ctx)));
    cudaStream_t st = prt_ctx_stream(sv.ctx);
    const unsigned long long capacity = (unsigned long long)n_probes * n_dirs;
    if (capacity >= 0xFFFFFFFFull) return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_probe_capture: more than 2^32 CSR entries");

    std::vector<float> dw(4 * (size_t)n_dirs);
    for (uint32_t i = 0; i < n_dirs; i++) { dw[4 * i] = dirs[3 * i]; dw[4 * i + 1] = dirs[3 * i + 1]; dw[4 * i + 2] = dirs[3 * i + 2]; dw[4 * i + 3] = weights[i]; }
    // trace order: directions sorted by the Morton code of their position on the unit sphere (10 bits per axis)
    std::vector<uint32_t> order(n_dirs);
    {
        std::vector<std::pair<uint32_t, uint32_t>> mk(n_dirs);
        auto spread = [](uint32_t v) { v &= 1023u; v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u; return v; };
        for (uint32_t i = 0; i < n_dirs; i++) {
            const float l = sqrtf(dirs[3 * i] * dirs[3 * i] + dirs[3 * i + 1] * dirs[3 * i + 1] + dirs[3 * i + 2] * dirs[3 * i + 2]);
            uint32_t q[3];
            for (int a = 0; a < 3; a++) {
                const float v = l > 0.f ? dirs[3 * i + a] / l : 0.f;
                q[a] = (uint32_t)std::min(1023.f, std::max(0.f, (v * 0.5f + 0.5f) * 1023.f));
            }
            mk[i] = {spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2), i};
        }
        std::sort(mk.begin(), mk.end());
        for (uint32_t i = 0; i < n_dirs; i++) order[i] = mk[i].second;
    }
    uint32_t *d_order = nullptr;
    float *d_pos = nullptr, *d_dirs = nullptr, *stransfer = nullptr, *sacc_stage = nullptr, *etransfer = nullptr, *eacc = nullptr;
    uint32_t *ticket = nullptr, *range = nullptr, *counts = nullptr;
    unsigned long long *offsets = nullptr, *skeys = nullptr, *ekeys = nullptr;
    int *overflow = nullptr;
    auto free_stage = [&]() { PB_FREE(skeys); PB_FREE(stransfer); PB_FREE(sacc_stage); skeys = nullptr; stransfer = nullptr; sacc_stage = nullptr; };
    auto cleanup = [&]() { free_stage(); PB_FREE(d_order); PB_FREE(d_pos); PB_FREE(d_dirs); PB_FREE(ticket); PB_FREE(counts); PB_FREE(offsets); PB_FREE(ekeys); PB_FREE(eacc); PB_FREE(overflow); };
    cudaError_t e = PB_MALLOC(&d_pos, sizeof(float) * 3 * (size_t)n_probes);
    if (e == cudaSuccess) e = PB_MALLOC(&d_dirs, sizeof(float) * 4 * (size_t)n_dirs);
    if (e == cudaSuccess) e = PB_MALLOC(&d_order, 4 * (size_t)n_dirs);
    if (e == cudaSuccess) e = PB_MALLOC(&ticket, 8);
    if (e == cudaSuccess) e = PB_MALLOC(&overflow, 4);
    if (e == cudaSuccess) e = PB_MALLOC(&counts, 4 * (size_t)n_probes);
    if (e == cudaSuccess) e = PB_MALLOC(&offsets, 8 * ((size_t)n_probes + 1));
    if (e == cudaSuccess) e = PB_MALLOC(&range, 8 * (size_t)n_probes);
    if (e == cudaSuccess) e = PB_MALLOC(&skeys, 8 * capacity);
    if (e == cudaSuccess) e = PB_MALLOC(&stransfer, 36 * capacity);
    if (e == cudaSuccess) e = PB_MALLOC(&sacc_stage, 28 * capacity);
    if (e != cudaSuccess) { cleanup(); PB_FREE(range); return prt_set_error(PRT_ERR_NOMEM, std::string("prt_probe_capture: ") + cudaGetErrorString(e)); }
    cudaMemcpyAsync(d_pos, probe_pos, sizeof(float) * 3 * (size_t)n_probes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_dirs, dw.data(), sizeof(float) * 4 * (size_t)n_dirs, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_order, order.data(), 4 * (size_t)n_dirs, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(ticket, 0, 8, st);
    cudaMemsetAsync(overflow, 0, 4, st);

    CaptureArgs A{};
    A.nodes = sv.nodes; A.tris = sv.tris; A.probe_pos = d_pos; A.n_probes = n_probes; A.dirs = (const float4 *)d_dirs; A.n_dirs = n_dirs; A.order = d_order; A.refill_thresh = prt_ctx_refill_thresh(sv.ctx); A.entry_list = prt_ctx_entry_list(sv.ctx);
    A.ticket = ticket; A.counts = counts; A.ekeys = skeys; A.etransfer = stransfer; A.eacc = sacc_stage;
    const size_t smem = (size_t)kMaxRays * (8 + 4 + 4) + std::max((size_t)kThreads * 16 * 4, sizeof(TraceShared) * (size_t)(kThreads / 32));
    static std::atomic<unsigned long long> configured{0};      // one bit per device
    PB_TRY(prt::ensure_dynamic_smem(probe_capture_kernel, (int)smem, configured));
    int per_sm = 1;
    PB_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_capture_kernel, kThreads, smem));
    const int grid = (int)std::min<unsigned long long>((unsigned long long)prt_ctx_sms(sv.ctx) * (unsigned long long)std::max(per_sm, 1), n_probes);
    // capture_ms = the kernels only (capture + scan, then the compaction); the exact-size allocations between them are host work
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    cudaEventRecord(e0, st);
    probe_capture_kernel<<<grid, kThreads, smem, st>>>(A);
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n_probes, offsets);
    cudaEventRecord(e1, st);
    unsigned long long nnz = 0; int ovf = 0;
    cudaMemcpyAsync(&nnz, offsets + n_probes, 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
        const unsigned long long nz = std::max<unsigned long long>(1, nnz);
        e = PB_MALLOC(&ekeys, 8 * nz);
        if (e == cudaSuccess) e = PB_MALLOC(&etransfer, 36 * nz);
        if (e == cudaSuccess) e = PB_MALLOC(&eacc, 28 * nz);
        if (e == cudaSuccess) {
            cudaEventRecord(e2, st);
            compact_csr_kernel<<<(unsigned)(((size_t)n_probes * 32 + 255) / 256), 256, 0, st>>>(counts, offsets, n_probes, n_dirs, skeys, stransfer, sacc_stage,
                                                                                          ekeys, etransfer, eacc, range);
            cudaEventRecord(e3, st);
            e = cudaStreamSynchronize(st);
        }
    }
    float ms = 0.f, ms2 = 0.f;
    if (e == cudaSuccess) { cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, e2, e3); ms += ms2; }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    if (e != cudaSuccess) { cleanup(); PB_FREE(range); PB_FREE(etransfer); return prt_set_error(PRT_ERR_CUDA, std::string("prt_probe_capture: ") + cudaGetErrorString(e)); }
    free_stage();

    // ---- global ids ----------------------------------------------------------------------------------------------------------
    prt_csr *c = new prt_csr();
    c->ctx = sv.ctx; c->n_probes = n_probes; c->nnz = nnz; c->range = range; c->transfer = etransfer; c->capture_ms = ms;
    uint32_t tbl_bits = 16;
    std::vector<unsigned long long> distinct;
    for (;;) {
        const uint32_t size = 1u << tbl_bits;
        unsigned long long *table = nullptr, *dlist = nullptr; uint32_t *dcount = nullptr;
        PB_MALLOC(&table, 8 * (size_t)size); PB_MALLOC(&dlist, 8 * (size_t)size); PB_MALLOC(&dcount, 4);
        cudaMemsetAsync(table, 0xFF, 8 * (size_t)size, st); cudaMemsetAsync(dcount, 0, 4, st); cudaMemsetAsync(overflow, 0, 4, st);
        if (nnz) hash_insert_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(ekeys, nnz, table, size - 1, overflow);
        hash_compact_kernel<<<(size + 255) / 256, 256, 0, st>>>(table, size, dlist, dcount);
        uint32_t cnt = 0;
        cudaMemcpyAsync(&cnt, dcount, 4, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(&ovf, overflow, 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        const bool too_full = ovf || cnt > size / 2;
        if (!too_full) { distinct.resize(cnt); if (cnt) cudaMemcpy(distinct.data(), dlist, 8 * (size_t)cnt, cudaMemcpyDeviceToHost); }
        PB_FREE(table); PB_FREE(dlist); PB_FREE(dcount);
        if (!too_full) break;
        if (++tbl_bits > 28) { cleanup(); prt_csr_destroy(c); return prt_set_error(PRT_ERR_NOMEM, "prt_probe_capture: too many distinct surfel clusters"); }
    }
    std::sort(distinct.begin(), distinct.end());
    c->n_prim = (uint32_t)distinct.size();
    double *sacc = nullptr;
    const size_t np = std::max<size_t>(1, distinct.size());
    e = PB_MALLOC(&c->keys, 8 * np);
    if (e == cudaSuccess) e = PB_MALLOC(&c->ids, 4 * std::max<unsigned long long>(1, nnz));
    if (e == cudaSuccess) e = PB_MALLOC(&c->surfels, 24 * np);
    if (e == cudaSuccess) e = PB_MALLOC(&sacc, 56 * np);
    if (e != cudaSuccess) { cleanup(); PB_FREE(sacc); prt_csr_destroy(c); return prt_set_error(PRT_ERR_NOMEM, "prt_probe_capture: cudaMalloc failed"); }
    if (!distinct.empty()) cudaMemcpyAsync(c->keys, distinct.data(), 8 * distinct.size(), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(sacc, 0, 56 * np, st);
    if (nnz) assign_ids_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(ekeys, nnz, c->keys, c->n_prim, eacc, c->ids, sacc);
    if (c->n_prim) finalize_surfels_kernel<<<(c->n_prim + 255) / 256, 256, 0, st>>>(sacc, c->n_prim, c->surfels);
    e = cudaStreamSynchronize(st);
    c->sums = sacc;
    cleanup();
    if (e != cudaSuccess) { prt_csr_destroy(c); return prt_set_error(PRT_ERR_CUDA, std::string("prt_probe_capture: ") + cudaGetErrorString(e)); }
    *out = c;
    return PRT_OK;
}

int prt_csr_sizes(const prt_csr *c, uint32_t *n_probes, uint64_t *nnz, uint32_t *n_surfels, double *capture_ms) {
    if (!c) return prt_set_error(PRT_ERR_INVALID, "prt_csr_sizes: null argument");
    if (n_probes) *n_probes = c->n_probes;
    if (nnz) *nnz = c->nnz;
    if (n_surfels) *n_surfels = c->n_prim;
    if (capture_ms) *capture_ms = c->capture_ms;
    return PRT_OK;
}

int prt_csr_download(const prt_csr *c, uint32_t *range, uint32_t *ids, float *transfer, float *surfels, uint64_t *keys) {
    if (!c) return prt_set_error(PRT_ERR_INVALID, "prt_csr_download: null argument");
    PB_TRY(cudaSetDevice(prt_ctx_device(c->ctx)));
    if (range) PB_TRY(cudaMemcpy(range, c->range, 8 * (size_t)c->n_probes, cudaMemcpyDeviceToHost));
    if (ids && c->nnz) PB_TRY(cudaMemcpy(ids, c->ids, 4 * c->nnz, cudaMemcpyDeviceToHost));
    if (transfer && c->nnz) PB_TRY(cudaMemcpy(transfer, c->transfer, 36 * c->nnz, cudaMemcpyDeviceToHost));
    if (surfels && c->n_prim) PB_TRY(cudaMemcpy(surfels, c->surfels, 24 * (size_t)c->n_prim, cudaMemcpyDeviceToHost));
    if (keys && c->n_prim) PB_TRY(cudaMemcpy(keys, c->keys, 8 * (size_t)c->n_prim, cudaMemcpyDeviceToHost));
    return PRT_OK;
}

}  // extern "C"

// device-resident projection for the per-frame pipeline (gi.cu): radiance [n_surfels] float4 -> out [n_probes][7] float4
cudaError_t prt_csr_project_device(const prt_csr *c, const float4 *d_radiance, float4 *d_out, cudaStream_t st) {
    probe_project_kernel<<<(unsigned)(((size_t)c->n_probes * 32 + 255) / 256), 256, 0, st>>>(c->range, c->ids, c->transfer, d_radiance, c->n_probes, d_out);
    return cudaGetLastError();
}

extern "C" {

int prt_csr_surfel_sums(const prt_csr *c, double *out_sums) {
    if (!c || !out_sums) return prt_set_error(PRT_ERR_INVALID, "prt_csr_surfel_sums: null argument");
    PB_TRY(cudaSetDevice(prt_ctx_device(c->ctx)));
    if (c->n_prim) {
        if (!c->sums) return prt_set_error(PRT_ERR_INVALID, "prt_csr_surfel_sums: this CSR was uploaded, not captured");
        PB_TRY(cudaMemcpy(out_sums, c->sums, 56 * (size_t)c->n_prim, cudaMemcpyDeviceToHost));
    }
    return PRT_OK;
}

int prt_csr_upload(prt_ctx *ctx, uint32_t n_probes, uint64_t nnz, uint32_t n_surfels, const uint32_t *range, const uint32_t *ids,
                   const float *transfer, const float *surfels, const uint64_t *keys, prt_csr **out) {
    if (!ctx || !out || n_probes == 0 || !range || (nnz && (!ids || !transfer)) || (n_surfels && !surfels))
        return prt_set_error(PRT_ERR_INVALID, "prt_csr_upload: bad argument");
    *out = nullptr;
    if (nnz >= 0xFFFFFFFFull) return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_csr_upload: more than 2^32 CSR entries");
    for (uint32_t p = 0; p < n_probes; p++)
        if (range[2 * p] > range[2 * p + 1] || range[2 * p + 1] > nnz) return prt_set_error(PRT_ERR_INVALID, "prt_csr_upload: probe range outside [0, nnz]");
    for (uint64_t i = 0; i < nnz; i++)
        if (ids[i] >= n_surfels) return prt_set_error(PRT_ERR_INVALID, "prt_csr_upload: surfel id out of range");
    PB_TRY(cudaSetDevice(prt_ctx_device(ctx)));
    cudaStream_t st = prt_ctx_stream(ctx);
    prt_csr *c = new prt_csr();
    c->ctx = ctx; c->n_probes = n_probes; c->nnz = nnz; c->n_prim = n_surfels;
    const size_t np = std::max<size_t>(1, n_surfels), nz = std::max<size_t>(1, (size_t)nnz);
    cudaError_t e = PB_MALLOC(&c->range, 8 * (size_t)n_probes);
    if (e == cudaSuccess) e = PB_MALLOC(&c->ids, 4 * nz);
    if (e == cudaSuccess) e = PB_MALLOC(&c->transfer, 36 * nz);
    if (e == cudaSuccess) e = PB_MALLOC(&c->surfels, 24 * np);
    if (e == cudaSuccess) e = PB_MALLOC(&c->keys, 8 * np);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaMemcpy(c->range, range, 8 * (size_t)n_probes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(c->ids, ids, 4 * (size_t)nnz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(c->transfer, transfer, 36 * (size_t)nnz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_surfels) e = cudaMemcpy(c->surfels, surfels, 24 * (size_t)n_surfels, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_surfels && keys) e = cudaMemcpy(c->keys, keys, 8 * (size_t)n_surfels, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_surfels && !keys) e = cudaMemset(c->keys, 0, 8 * (size_t)n_surfels);
    if (e != cudaSuccess) { prt_csr_destroy(c); return prt_set_error(PRT_ERR_CUDA, std::string("prt_csr_upload: ") + cudaGetErrorString(e)); }
    *out = c;
    return PRT_OK;
}

int prt_probe_project(const prt_csr *c, const float *radiance_rgba, float *out_sh_volumes) {
    if (!c || !radiance_rgba || !out_sh_volumes) return prt_set_error(PRT_ERR_INVALID, "prt_probe_project: null argument");
    PB_TRY(cudaSetDevice(prt_ctx_device(c->ctx)));
    cudaStream_t st = prt_ctx_stream(c->ctx);
    float4 *d_rad = nullptr, *d_out = nullptr;
    PB_TRY(PB_MALLOC(&d_rad, 16 * std::max<size_t>(1, c->n_prim)));
    PB_TRY(PB_MALLOC(&d_out, 112 * (size_t)c->n_probes));
    if (c->n_prim) cudaMemcpyAsync(d_rad, radiance_rgba, 16 * (size_t)c->n_prim, cudaMemcpyHostToDevice, st);
    probe_project_kernel<<<(unsigned)(((size_t)c->n_probes * 32 + 255) / 256), 256, 0, st>>>(c->range, c->ids, c->transfer, d_rad, c->n_probes, d_out);
    cudaError_t e = cudaMemcpyAsync(out_sh_volumes, d_out, 112 * (size_t)c->n_probes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    PB_FREE(d_rad); PB_FREE(d_out);
    if (e != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string("prt_probe_project: ") + cudaGetErrorString(e));
    return PRT_OK;
}

// probe grid (volume.cpp:83-90) and direction sets (light_probe.cpp:137-152; SH_function.h:96-112 + volume.cpp:251-254): host helpers
int prt_probe_positions(const int32_t res[3], const float size[3], float *out_pos) {
    if (!res || !size || !out_pos) return prt_set_error(PRT_ERR_INVALID, "prt_probe_positions: null argument");
    size_t k = 0;
    for (int z = 0; z < res[2]; z++)
        for (int y = 0; y < res[1]; y++)
            for (int x = 0; x < res[0]; x++, k++) {
                const int id[3] = {x, y, z};
                for (int a = 0; a < 3; a++) { const float ds = 2.0f / (float)res[a] * size[a]; out_pos[3 * k + a] = -size[a] + ds * (0.5f + (float)id[a]); }
            }
    return PRT_OK;
}
int prt_fibonacci_dirs(int32_t n, float *out_dirs) {
    if (n < 2 || !out_dirs) return prt_set_error(PRT_ERR_INVALID, "prt_fibonacci_dirs: bad argument");
    const double pi = 2 * acos(0.0), gold = 3 - sqrt(5.0);
    for (int i = 0; i < n; i++) {
        const double z = 1 - ((double)i / (double)(n - 1)) * 2, theta = pi * i * gold, r = sqrt(1 - z * z);
        out_dirs[3 * i] = (float)(cos(theta) * r); out_dirs[3 * i + 1] = (float)(sin(theta) * r); out_dirs[3 * i + 2] = (float)z;
    }
    return PRT_OK;
}
int prt_cube_dirs(int32_t res, float *out_dirs, float *out_weights) {
    if (res < 1 || !out_dirs || !out_weights) return prt_set_error(PRT_ERR_INVALID, "prt_cube_dirs: bad argument");
    size_t k = 0;
    for (int f = 0; f < 6; f++)
        for (int y = 0; y < res; y++)
            for (int x = 0; x < res; x++, k++) {
                const float u = (float)(((double)x + 0.5) / (double)res) * 2.0f - 1.0f, v = (float)(((double)y + 0.5) / (double)res) * 2.0f - 1.0f;
                float d[3];
                switch (f) {
                case 0: d[0] = 1.f; d[1] = -v; d[2] = -u; break;
                case 1: d[0] = -1.f; d[1] = -v; d[2] = u; break;
                case 2: d[0] = u; d[1] = 1.f; d[2] = v; break;
                case 3: d[0] = u; d[1] = -1.f; d[2] = -v; break;
                case 4: d[0] = u; d[1] = -v; d[2] = 1.f; break;
                default: d[0] = -u; d[1] = -v; d[2] = -1.f; break;
                }
                float w = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                w *= sqrtf(w);
                out_dirs[3 * k] = d[0]; out_dirs[3 * k + 1] = d[1]; out_dirs[3 * k + 2] = d[2];
                out_weights[k] = 4.0f / (float)res / (float)res / w;
            }
    return PRT_OK;
}

}  // extern "C"
