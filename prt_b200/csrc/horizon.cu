// horizon.cu -- horizon pass of the shadowed per-vertex transfer bake (first of two kernels, see bake_wave.cu).
//
// One persistent warp per vertex:
//   1. build_entry_list   candidate boxes for the shared ray origin (entry_list.cuh)
//   2. build_horizon      conservative bound of sin(elevation) of all geometry per azimuth bin (kHzBins = 32 or 64), refined through near
//                         subtrees down to exact triangle bounds
//   3. classify           sample i needs tracing iff its local z does not exceed the horizon of its bin; the need bits of the
//                         vertex (processing order) and their count go to global memory for the traversal pass
//   4. vertices with no sample to trace are finished here: every sample is visible, so the row is the plain cosine-weighted
//      projection (reference raytracing.cpp:257-261,348-350 with every ray escaping) and the visibility words are all ones
//
// The pass is a pure culling device: a ray it marks "free" provably cannot hit anything, every other ray goes through the pinned
// triangle test in the traversal pass, so results are identical with or without it (tests/test_gpu_parity.py).
#include "kernels.h"
#include "entry_list.cuh"

namespace prt {

namespace {

#ifndef PRT_HZ_MINB
#define PRT_HZ_MINB 9               // CTAs per SM the register allocation aims at (56 registers; 10 CTAs at 48 registers measured 3 % slower)
#endif

struct HorizonShared {
    EntryList el;
    uint32_t hz[kHzWords];             // the map (first kHzBins words) and its range-minimum table
    uint32_t rq[kHzQueue];
    uint32_t tq[kHzTriQueue];
};

template <int ORDER>
__global__ void __launch_bounds__(128, PRT_HZ_MINB) horizon_kernel(const BakeArgs A) {
    constexpr int N2 = ORDER * ORDER;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HorizonShared &W = reinterpret_cast<HorizonShared *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const float sgn = A.cs_phase ? -1.0f : 1.0f;
    const int S = A.S, words = A.vis_words;

    for (;;) {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(A.counter + 1, 1u);          // counter[1]: this pass, counter[0]: traversal pass
        v = __shfl_sync(kFull, v, 0);
        if (v >= A.n_verts) break;

        const float *pp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.pos) + (size_t)v * A.stride);
        const float *np = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.nrm) + (size_t)v * A.stride);
        const f3 N = mk3(__ldg(np), __ldg(np + 1), __ldg(np + 2));
        const f3 P = mk3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
        const Frame fr = make_frame(N);
        const f3 org = madd3(P, A.origin_eps, N);                       // raytracing.cpp:343

        const int n_cand = build_entry_list(A.nodes, org, N, W.el, lane);
        build_horizon(W.el, n_cand, A.nodes, A.tris, org, N, fr, W.hz, W.rq, W.tq, A.horizon_budget, A.horizon_near2, lane);

        uint32_t *row = A.need_bits + (size_t)v * words;
        uint32_t total = 0u;
        for (int base = 0; base < S; base += 32) {
            const int i = base + lane;
            bool need = false;
            if (i < S) {
                const float4 smp = __ldg(&A.samples[i]);
                need = !(smp.z > __uint_as_float(W.hz[__float_as_uint(smp.w) >> 24]));
            }
            const unsigned nb = __ballot_sync(kFull, need);
            if (lane == 0) row[base >> 5] = nb;
            total += __popc(nb);
        }
        if (lane == 0) A.need_count[v] = total;
        if (total == 0u) {
            float acc[N2];
#pragma unroll
            for (int k = 0; k < N2; k++) acc[k] = 0.f;
            for (int i = lane; i < S; i += 32) {
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
            if (lane < N2) A.out[(size_t)v * N2 + lane] = mine * A.inv_S;
            if (A.vis) {
                for (int w = lane; w < words; w += 32) {
                    const int rem = S - 32 * w;
                    A.vis[(size_t)v * words + w] = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
                }
            }
        }
        __syncwarp();
    }
}

// Files every unfinished vertex under one of four cost classes (the quartile of S its need count falls in).  The traversal pass
// walks the classes heaviest first, so that the warps still busy when the work runs out hold cheap vertices (shorter tail).
// Thread per vertex, one atomic per warp and class; the order inside a class stays close to the (Morton) vertex order.
__global__ void __launch_bounds__(256) work_list_kernel(const uint32_t *need_count, const uint32_t n, const uint32_t S, uint32_t *list, uint32_t *class_count) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t cls = 4u;
    if (v < n) {
        const uint32_t q = 4u * need_count[v];
        if (q) cls = q > 3u * S ? 0u : q > 2u * S ? 1u : q > S ? 2u : 3u;
    }
#pragma unroll
    for (uint32_t c = 0; c < 4u; c++) {
        const unsigned m = __ballot_sync(kFull, cls == c);
        if (!m) continue;
        const int leader = __ffs(m) - 1;
        uint32_t base = 0u;
        if (lane == leader) base = atomicAdd(class_count + c, (uint32_t)__popc(m));
        base = __shfl_sync(kFull, base, leader);
        if (cls == c) list[(size_t)c * n + base + __popc(m & ((1u << lane) - 1u))] = v;
    }
}

template <int ORDER>
cudaError_t launch_horizon_t(const BakeArgs &A, int *grid, int n_sms, cudaStream_t st) {
    const int block = 128;
    const size_t smem = sizeof(HorizonShared) * (size_t)(block / 32);
    if (*grid <= 0) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, horizon_kernel<ORDER>, block, smem);
        if (e != cudaSuccess) return e;
        *grid = n_sms * (per_sm > 0 ? per_sm : 1);
    }
    const long long need = ((long long)A.n_verts + 3) / 4;
    if (need < *grid) *grid = (int)(need > 0 ? need : 1);
    horizon_kernel<ORDER><<<*grid, block, smem, st>>>(A);
    if (A.work_list) work_list_kernel<<<(A.n_verts + 255u) / 256u, 256, 0, st>>>(A.need_count, A.n_verts, (uint32_t)A.S, A.work_list, A.counter + 4);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_horizon(const BakeArgs &A, int order, int *grid, int n_sms, cudaStream_t st) {
    switch (order) {
    case 1: return launch_horizon_t<1>(A, grid, n_sms, st);
    case 2: return launch_horizon_t<2>(A, grid, n_sms, st);
    case 3: return launch_horizon_t<3>(A, grid, n_sms, st);
    case 4: return launch_horizon_t<4>(A, grid, n_sms, st);
    case 5: return launch_horizon_t<5>(A, grid, n_sms, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace prt
