// kernels.h -- launch interface between the C-ABI layer (abi.cu) and the kernel translation units.
#pragma once
#include "bvh8.h"
#include "horizon_math.cuh"     // kHzBins
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace prt {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: `done` (one static per kernel instantiation) remembers the
// devices it has been set on, so a second GPU in the same process (multi-GPU driver, a second context) gets it too
template <class K>
inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, std::atomic<unsigned long long> &done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}


struct BakeArgs {
    const Node8 *nodes;
    const Tri48 *tris;
    const Slab32 *slabs;        // optional [n_nodes]: oriented slab of every node (horizon pass)
    const Dop32 *dops;          // optional [n_nodes]: fourth slab axis of the node test (traversal pass; nullable by the tuning knob wave_dop)
    const float *pos, *nrm;     // device; consecutive vertices `stride` bytes apart
    size_t stride;
    uint32_t n_verts, vid_base;
    const float4 *samples;      // [S] (local dir xyz, w = reference sample index s | azimuth bin << 24), processing order
    int S;
    float inv_S;
    float *out;                 // [n_verts][order^2]
    uint32_t *vis;              // optional [n_verts][vis_words], pre-zeroed
    int vis_words;
    uint32_t *counter;          // persistent-warp work counter, pre-zeroed
    unsigned long long *work;   // optional [3]: node visits, triangle tests, candidate-box tests (pre-zeroed)
    int entry_list;             // 1: per-origin entry lists (bake.cu), 0: every ray starts at the root
    int horizon;                // 1: per-origin horizon map (bake_wave.cu): rays above it skip traversal
    int horizon_budget;         // refinement iterations (4 nodes each) the horizon builder may spend per vertex
    float horizon_near2;        // refine boxes with d^2 < near2 * r^2 (angular radius above asin(1/sqrt(near2)))
    float horizon_mid2;         // ... and boxes with d^2 < mid2 * r^2 whose bound would leave more than horizon_gain (in units of
    float horizon_gain;         //     S / kHzBins samples) to trace; mid2 = 0: rule off
    uint32_t *need_bits;        // horizon pass output / traversal pass input: [n_verts][vis_words], bit i of a row = sample with
                                // processing index i is NOT above the horizon and must be traced
    uint32_t *need_count;       // [n_verts] number of such samples; 0 = the horizon pass already wrote the vertex's row
    uint32_t *work_list;        // optional [n_verts + 512]: the unfinished vertices in descending order of their need count (counting
                                // sort after the horizon pass, horizon.cu; length in counter[4]; the tail holds the sort's histogram) --
                                // the traversal pass takes the heaviest vertices first, so the warps that finish last hold cheap ones
    uint32_t seed;
    int depth;                  // path segments = bounces + 1
    float albedo[3];
    float origin_eps, bounce_eps;
    int cs_phase;
    int refill_thresh;          // refill idle lanes when fewer than this many lanes are traversing (0: static rounds)
    // ---- output placement / sharding (all zero: rows packed [n_verts][order^2] in launch order) ------------------------------
    uint32_t out_stride;        // floats between the rows of consecutive vertices; 0 = order^2 (60 B / 4 = 15 writes straight into
                                // a Mesh::Vert array at sh_coeff, gl.h:76-80)
    uint32_t shard_world, shard_rank;   // this launch bakes shard `rank` of `world` interleaved kShardChunk-vertex chunks of a longer
                                // vertex list: local vertex v is global vertex global_row(v); the bounce RNG is keyed by the GLOBAL id
    int out_global;             // rows are stored at the global index (a full-size buffer per GPU) instead of the local one
    int n_peer;                 // fused gather: every finished row is also stored into these peer-GPU buffers (P2P over NVLink),
    float *out_peer[7];         // so that no collective has to follow the kernel
};

constexpr uint32_t kShardChunk = 64;   // vertices per interleaved chunk of a sharded bake (prt_group_bake_transfer, prt_b200/dist.py)

#if defined(__CUDACC__) || defined(PRT_HOSTCHECK)
// global index of local vertex v of a sharded launch
__host__ __device__ __forceinline__ uint32_t global_row(const BakeArgs &A, const uint32_t v) {
    return A.shard_world > 1u ? ((v / kShardChunk) * A.shard_world + A.shard_rank) * kShardChunk + (v % kShardChunk) : v;
}
// coefficient k of vertex v (the lanes k < order^2 of the warp that baked v call this)
__device__ __forceinline__ void store_row(const BakeArgs &A, const uint32_t v, const int n2, const int k, const float val) {
    const size_t at = (size_t)(A.out_global ? global_row(A, v) : v) * (A.out_stride ? A.out_stride : (uint32_t)n2) + (size_t)k;
    A.out[at] = val;
    for (int p = 0; p < A.n_peer; p++) A.out_peer[p][at] = val;
}
#endif

// mode: 0 shadowed, 1 interreflect, 2 unshadowed Monte-Carlo, 3 unshadowed analytic
// *grid <= 0: one persistent wave, grid = n_sms x occupancy(kernel, block); the grid used is written back
cudaError_t launch_bake(const BakeArgs &, int order, int mode, int *grid, int block, int n_sms, cudaStream_t);
// horizon pass (horizon.cu): per-vertex horizon map, classification of every sample, rows of fully visible vertices
cudaError_t launch_bake_inter(const BakeArgs &, int order, int *grid, int n_sms, cudaStream_t);
int bake_inter_max_samples();
cudaError_t launch_horizon(const BakeArgs &, int order, int *grid, int n_sms, cudaStream_t);
// warp-local wavefront kernel (bake_wave.cu), same modes / limits
cudaError_t launch_bake_wave(const BakeArgs &, int order, bool trace, int *grid, int block, int n_sms, cudaStream_t);
int bake_wave_max_samples();
int bake_wave_block();   // threads per CTA the wavefront kernel is compiled for
// streaming L2 prefetch of two arrays (the BVH) ahead of a bake on a cold L2
cudaError_t launch_l2_prefetch(const void *a, size_t bytes_a, const void *b, size_t bytes_b, int n_sms, cudaStream_t);
cudaError_t launch_trace_any(const Node8 *, const Tri48 *, const float *rays, uint32_t n, uint8_t *out, cudaStream_t);
cudaError_t launch_trace_closest(const Node8 *, const Tri48 *, const float *rays, uint32_t n, float *out_t,
                                 uint32_t *out_prim, float *out_ng, cudaStream_t);

}  // namespace prt
