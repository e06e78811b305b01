// horizon.cu -- horizon pass of the shadowed per-vertex transfer bake (first of two kernels, see bake_wave.cu).
//
// One persistent warp per vertex:
//   1. build_entry_list   candidate boxes for the shared ray origin (entry_list.cuh)
//   2. build_horizon      conservative bound of sin(elevation) of all geometry per azimuth bin (kHzBins = 32 or 64), refined through near
//                         subtrees down to exact triangle bounds; a subtree is bounded by its box CUT BY ITS ORIENTED SLAB (Slab32, bvh8.h:
//                         hz_slab_value) and opened when merging that bound would leave many samples to trace (hz_gain)
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
#define PRT_HZ_MINB 8               // CTAs per SM the register allocation aims at (64 registers; with the slab bound of round 2: 9 CTAs at 56 registers
                                    // 5.70 ms, 8 CTAs 5.58 ms on the headline bake; round 1, before it: 9 CTAs best, 10 CTAs at 48 registers 3 % slower)
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

// Work list of the traversal pass: the unfinished vertices in DESCENDING order of their need count (a counting sort over kWorkBuckets
// buckets of need / S), so that the persistent warps take the heaviest vertices first and the warps still busy when the work runs out
// hold cheap ones (longest-processing-time-first: the makespan of a small shard is bounded by the mean load plus one LIGHT vertex).
// list[0 .. total): vertex indices; aux = [hist kWorkBuckets][offsets kWorkBuckets] (hist pre-zeroed); total -> class_count[0].
#ifndef PRT_WORK_BUCKETS
#define PRT_WORK_BUCKETS 256         // <= 256; measured on the 68 k-vertex shard of an 8-GPU bake (list off: 6.54 ms): 4 buckets 6.45, 32: 6.34, 256: 6.31
#endif
constexpr uint32_t kWorkBuckets = 256;                 // size of the histogram arrays
constexpr uint32_t kWorkUsed = PRT_WORK_BUCKETS;
static_assert(kWorkUsed >= 1 && kWorkUsed <= kWorkBuckets, "PRT_WORK_BUCKETS out of range");
__device__ __forceinline__ uint32_t work_bucket(const uint32_t need, const uint32_t S) {
    return (kWorkUsed - 1u) - min(kWorkUsed - 1u, (uint32_t)(((unsigned long long)need * kWorkUsed) / (S + 1u)));   // heavy -> bucket 0
}
__global__ void __launch_bounds__(256) work_hist_kernel(const uint32_t *need_count, const uint32_t n, const uint32_t S, uint32_t *hist) {
    __shared__ uint32_t h[kWorkBuckets];
    h[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const uint32_t need = need_count[v];
        if (need) atomicAdd(&h[work_bucket(need, S)], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(256) work_scan_kernel(const uint32_t *hist, uint32_t *offsets, uint32_t *total) {
    __shared__ uint32_t s[kWorkBuckets];
    const uint32_t mine = hist[threadIdx.x];
    s[threadIdx.x] = mine;
    __syncthreads();
    for (uint32_t o = 1; o < kWorkBuckets; o <<= 1) {
        const uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    offsets[threadIdx.x] = s[threadIdx.x] - mine;
    if (threadIdx.x == kWorkBuckets - 1u) *total = s[threadIdx.x];
}
__global__ void __launch_bounds__(256) work_scatter_kernel(const uint32_t *need_count, const uint32_t n, const uint32_t S, uint32_t *offsets, uint32_t *list) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t need = need_count[v];
    if (need) list[atomicAdd(&offsets[work_bucket(need, S)], 1u)] = v;
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
    if (A.work_list) {
        uint32_t *hist = A.work_list + A.n_verts, *offsets = hist + kWorkBuckets;          // zeroed by the caller (abi.cu)
        const unsigned g = (unsigned)min((A.n_verts + 255u) / 256u, 1024u);
        work_hist_kernel<<<g, 256, 0, st>>>(A.need_count, A.n_verts, (uint32_t)A.S, hist);
        work_scan_kernel<<<1, 256, 0, st>>>(hist, offsets, A.counter + 4);
        work_scatter_kernel<<<(A.n_verts + 255u) / 256u, 256, 0, st>>>(A.need_count, A.n_verts, (uint32_t)A.S, offsets, A.work_list);
    }
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
