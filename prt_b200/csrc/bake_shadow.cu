// bake_shadow.cu -- the headline kernel: shadowed (and unshadowed Monte-Carlo) per-vertex SH transfer, order 1..5.
//
// Same contract as bake_kernel in bake.cu (reference bake_SH, src/raytracing/raytracing.cpp:320-360, with
// renderSH at depth 1, :228-278), restructured so that every phase runs with full warps:
//
//   per vertex (one persistent warp):
//     1. build_entry_list         flatten the chain of nodes containing the shared ray origin into <= 96 candidate
//                                 boxes in shared memory (entry_list.cuh)
//     2. scan   (lockstep)        32 rays at a time test every candidate box; each (ray, candidate) hit becomes an
//                                 independent any-hit sub-query, appended to a shared-memory queue -- leaf candidates
//                                 and subtree candidates in separate queues (warp-level compaction by ballot/popc)
//     3. drain leaf pairs         one pair per lane: <= 3 pinned triangle tests, sets the ray's occlusion bit
//     4. drain subtree pairs      one pair per lane, CWBVH traversal of the subtree; idle lanes pull the next pair
//     5. project (lockstep)       every unoccluded sample direction is projected on the SH basis in registers,
//                                 warp-shuffle reduction, one row store
//
// A ray is occluded iff any of its pairs reports a hit, and hits are decided by the pinned triangle test only, so
// the result (visibility bits, coefficients) is identical to the plain traversal.
#include "kernels.h"
#include "entry_list.cuh"

namespace prt {

namespace {

constexpr int kMaxS = 8192;             // occlusion bitset capacity (samples per vertex) of this kernel
constexpr int kLeafCap = 256, kInnerCap = 256;

struct ShadowShared {
    EntryList el;                       // candidate boxes (+ build queue)
    uint32_t occl[kMaxS / 32];          // bit s (reference sample index): primary ray s is occluded
    uint32_t lq[kLeafCap];              // (processing index << 7) | candidate
    uint32_t iq[kInnerCap];
};

__device__ __forceinline__ float fast_rcp(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    return __frcp_rn(d);
}

template <int ORDER, bool TRACE>
__global__ void __launch_bounds__(256) bake_shadow_kernel(const BakeArgs A) {
    constexpr int N2 = ORDER * ORDER;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ShadowShared &W = reinterpret_cast<ShadowShared *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float sgn = A.cs_phase ? -1.0f : 1.0f;
    const int S = A.S, words = A.vis_words;
    unsigned long long cand_tests = 0ull;
    uint32_t node_visits = 0u, tri_tests = 0u;

    for (;;) {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(A.counter, 1u);
        v = __shfl_sync(kFull, v, 0);
        if (v >= A.n_verts) break;

        const float *pp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.pos) + (size_t)v * A.stride);
        const float *np = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.nrm) + (size_t)v * A.stride);
        const f3 N = mk3(__ldg(np), __ldg(np + 1), __ldg(np + 2));
        const f3 P = mk3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
        const Frame fr = make_frame(N);
        const f3 org = madd3(P, A.origin_eps, N);                       // raytracing.cpp:343

        for (int w = lane; w < words; w += 32) W.occl[w] = 0u;
        int n_cand = 0;
        if (TRACE) {
            n_cand = build_entry_list(A.nodes, org, N, W.el, lane);     // ends with __syncwarp()
            cand_tests += (unsigned long long)n_cand * (unsigned long long)S;
        }
        __syncwarp();

        if (TRACE) {
            int base = 0, ltail = 0, itail = 0;       // warp-uniform
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u;       // candidate hits of the lane's current ray not yet queued
            uint32_t sproc = 0u;                      // processing-order index of the lane's current ray
            for (;;) {
                // ---- 2a. queue pending (ray, candidate) pairs while there is room for one more warp-wide append
                bool pending = __any_sync(kFull, (m0 | m1 | m2) != 0u);
                while (pending && ltail <= kLeafCap - 32 && itail <= kInnerCap - 32) {
                    int k = -1;
                    if (m0) { k = __ffs(m0) - 1; m0 &= m0 - 1u; }
                    else if (m1) { k = 32 + __ffs(m1) - 1; m1 &= m1 - 1u; }
                    else if (m2) { k = 64 + __ffs(m2) - 1; m2 &= m2 - 1u; }
                    const bool has = k >= 0;
                    const bool leaf = has && __float_as_uint(W.el.cb[has ? k : 0].w) <= 0x00FFFFFFu;
                    const unsigned hb = __ballot_sync(kFull, has), lb = __ballot_sync(kFull, leaf), ib = hb & ~lb;
                    const uint32_t pair = (sproc << 7) | (uint32_t)(has ? k : 0);
                    if (leaf) W.lq[ltail + __popc(lb & lt_mask)] = pair;
                    else if (has) W.iq[itail + __popc(ib & lt_mask)] = pair;
                    ltail += __popc(lb); itail += __popc(ib);
                    pending = __any_sync(kFull, (m0 | m1 | m2) != 0u);
                }
                // ---- 2b. scan the next 32 rays against the candidate boxes (lockstep)
                if (!pending && base < S && ltail <= kLeafCap - 32 && itail <= kInnerCap - 32) {
                    const int i = base + lane;
                    base += 32;
                    if (i < S) {
                        const float4 smp = __ldg(&A.samples[i]);
                        const f3 d = to_world(fr, mk3(smp.x, smp.y, smp.z));   // raytracing.cpp:340
                        uint32_t cm[3];
                        scan_entry_list(W.el, n_cand, fast_rcp(d.x), fast_rcp(d.y), fast_rcp(d.z), cm);
                        m0 = cm[0]; m1 = cm[1]; m2 = cm[2];
                        sproc = (uint32_t)i;
                    }
                    continue;
                }
                __syncwarp();
                // ---- 3. leaf pairs: <= 3 triangles each ---------------------------------------------------------
                for (int q = lane; q < ltail; q += 32) {
                    const uint32_t pair = W.lq[q];
                    const float4 smp = __ldg(&A.samples[pair >> 7]);
                    const uint32_t sref = __float_as_uint(smp.w) & 0xFFFFFFu;
                    if ((W.occl[sref >> 5] >> (sref & 31u)) & 1u) continue;
                    const f3 d = to_world(fr, mk3(smp.x, smp.y, smp.z));
                    const float4 g = W.el.cb[pair & 127u];
                    const uint32_t tri0 = __float_as_uint(g.z);
                    uint32_t bits = __float_as_uint(g.w);
                    while (bits) {
                        const uint32_t b = (uint32_t)__ffs(bits) - 1u;
                        bits &= bits - 1u;
                        float t; uint32_t prim;
                        tri_tests++;
                        if (tri_hit(A.tris, tri0 + b, org, d, 0.0f, INFINITY, false, t, prim)) {
                            atomicOr(&W.occl[sref >> 5], 1u << (sref & 31u));
                            break;
                        }
                    }
                }
                ltail = 0;
                __syncwarp();
                // ---- 4. subtree pairs: idle lanes pull the next pair ------------------------------------------------
                {
                    int head = 0;
                    bool active = false;
                    uint32_t sref = 0u;
                    Trav tr;
                    tr.reset_counters();
                    for (;;) {
                        const unsigned idle = __ballot_sync(kFull, !active);
                        if (idle && head < itail) {
                            const int q = head + __popc(idle & lt_mask);
                            if (!active && q < itail) {
                                const uint32_t pair = W.iq[q];
                                const float4 smp = __ldg(&A.samples[pair >> 7]);
                                sref = __float_as_uint(smp.w) & 0xFFFFFFu;
                                if (!((W.occl[sref >> 5] >> (sref & 31u)) & 1u)) {
                                    tr.init(org, to_world(fr, mk3(smp.x, smp.y, smp.z)), 0.0f, INFINITY);
                                    const float4 g = W.el.cb[pair & 127u];
                                    tr.start_group(__float_as_uint(g.z), __float_as_uint(g.w));
                                    active = true;
                                }
                            }
                            head += __popc(idle);
                        }
                        if (!__any_sync(kFull, active)) { if (head >= itail) break; else continue; }
                        if (!active) continue;
                        const int rc = tr.template run<true>(A.nodes, A.tris, A.refill_thresh, head < itail);
                        if (rc == TRAV_RUNNING) continue;
                        if (rc == TRAV_HIT) atomicOr(&W.occl[sref >> 5], 1u << (sref & 31u));
                        active = false;
                    }
                    node_visits += tr.n_node_visits; tri_tests += tr.n_tri_tests;
                }
                itail = 0;
                __syncwarp();
                if (base >= S && !__any_sync(kFull, (m0 | m1 | m2) != 0u)) break;
            }
        }

        // ---- 5. projection: L = Y_lm(dir) for every unoccluded sample (raytracing.cpp:226,257-261,348) ----------
        float acc[N2];
#pragma unroll
        for (int k = 0; k < N2; k++) acc[k] = 0.f;
        for (int i = lane; i < S; i += 32) {
            const float4 smp = __ldg(&A.samples[i]);
            const uint32_t sref = __float_as_uint(smp.w) & 0xFFFFFFu;
            if ((W.occl[sref >> 5] >> (sref & 31u)) & 1u) continue;
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
        if (lane < N2) A.out[(size_t)v * N2 + lane] = mine * A.inv_S;       // raytracing.cpp:350
        if (A.vis) {
            for (int w = lane; w < words; w += 32) {
                const int rem = S - 32 * w;
                const uint32_t valid = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
                A.vis[(size_t)v * words + w] = ~W.occl[w] & valid;
            }
        }
        __syncwarp();
    }
    if (A.work) {
        const unsigned long long nv = warp_sum_u64(node_visits), nt = warp_sum_u64(tri_tests);
        if (lane == 0) { atomicAdd(&A.work[0], nv); atomicAdd(&A.work[1], nt); atomicAdd(&A.work[2], cand_tests); }
    }
}

template <int ORDER, bool TRACE>
cudaError_t launch_shadow_t(const BakeArgs &A, int *grid, int block, int n_sms, cudaStream_t st) {
    const size_t smem = sizeof(ShadowShared) * (size_t)(block / 32);
    static bool configured = false;   // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(bake_shadow_kernel<ORDER, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ShadowShared) * 8));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (*grid <= 0) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bake_shadow_kernel<ORDER, TRACE>, block, smem);
        if (e != cudaSuccess) return e;
        *grid = n_sms * (per_sm > 0 ? per_sm : 1);
    }
    const int warps_per_block = block / 32;
    const long long need = ((long long)A.n_verts + warps_per_block - 1) / warps_per_block;
    if (need < *grid) *grid = (int)(need > 0 ? need : 1);
    bake_shadow_kernel<ORDER, TRACE><<<*grid, block, smem, st>>>(A);
    return cudaGetLastError();
}

template <int ORDER>
cudaError_t launch_shadow_o(const BakeArgs &A, bool trace, int *grid, int block, int n_sms, cudaStream_t st) {
    return trace ? launch_shadow_t<ORDER, true>(A, grid, block, n_sms, st) : launch_shadow_t<ORDER, false>(A, grid, block, n_sms, st);
}

}  // namespace

int bake_shadow_max_samples() { return kMaxS; }

cudaError_t launch_bake_shadow(const BakeArgs &A, int order, bool trace, int *grid, int block, int n_sms, cudaStream_t st) {
    switch (order) {
    case 1: return launch_shadow_o<1>(A, trace, grid, block, n_sms, st);
    case 2: return launch_shadow_o<2>(A, trace, grid, block, n_sms, st);
    case 3: return launch_shadow_o<3>(A, trace, grid, block, n_sms, st);
    case 4: return launch_shadow_o<4>(A, trace, grid, block, n_sms, st);
    case 5: return launch_shadow_o<5>(A, trace, grid, block, n_sms, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace prt
