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
#include "horizon.cuh"

namespace prt {

namespace {

#ifndef PRT_HZ_MINB
#define PRT_HZ_MINB 9               // CTAs per SM the register allocation aims at (56 registers; 10 CTAs at 48 registers measured 3 % slower)
#endif

template <int ORDER>
__global__ void __launch_bounds__(128, PRT_HZ_MINB) horizon_kernel(const BakeArgs A) {
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
        horizon_vertex<ORDER>(A, W, v, lane, S, words, sgn);
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
