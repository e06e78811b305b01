// bake_inter.cu -- interreflected per-vertex transfer (BASELINE config 4) as a warp-local asynchronous wavefront.
//
// Replaces bake_SH + renderSH with depth > 1 (reference src/raytracing/raytracing.cpp:228-278,320-360).  Runs after the horizon
// pass (horizon.cu), which has settled every primary ray that provably escapes; the flagged samples of a vertex become paths.
//
// One persistent warp per vertex keeps up to 64 paths alive in shared-memory slots (origin, direction, tnear, the closest hit so
// far, a count of outstanding work items).  Work is (slot, node) and (slot, leaf) items on two shared-memory stacks:
//   node step   each lane decodes ONE 80-byte node for its item, tests the 8 child boxes against [tnear, closest t], pushes the hits
//   leaf step   each lane runs the pinned triangle tests of ONE leaf; a hit lowers the slot's (t, prim) key with a 64-bit atomicMin
//   shade       a slot whose outstanding-item count reaches zero has finished its segment: a miss adds Lw * Y(dir) to the lane's
//               accumulators and frees the slot, a hit bounces (raytracing.cpp:263-275) and restarts the slot at the BVH root
//   refill      free slots take the next flagged samples 32 at a time: lockstep scan of the per-origin entry list (entry_list.cuh)
// so the lanes of a step always hold 32 items whatever the individual paths are doing.  The per-ray stack kernel it replaces
// (bake.cu) ran at 6.7 of 32 active lanes on this workload (profiles/r1_interreflect_ncu_summary.txt).
//
// Closest-hit order: (t, prim) keys are totally ordered, so the result does not depend on the order in which items are processed
//  and is the same hit the oracle's stack traversal finds.  A full stack falls back to an ordinary stack
// traversal of the subtree (noinline, cold), so capacity never affects results.
#include "bake_inter.cuh"

namespace prt {

namespace {

// COUNT: carry the work counters (instrumented launches only; the timed variant keeps those registers free)
template <int ORDER, bool COUNT>
__global__ void __launch_bounds__(128, PRT_INTER_MINB) bake_inter_kernel(const BakeArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    InterShared &W = reinterpret_cast<InterShared *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float sgn = A.cs_phase ? -1.0f : 1.0f;
    const int S = A.S, depth = A.depth;
    unsigned long long cand_tests = 0ull, rays_scanned = 0ull;
    uint32_t node_visits = 0u, tri_tests = 0u;

    for (;;) {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(A.counter, 1u);
        v = __shfl_sync(kFull, v, 0);
        if (v >= A.n_verts) break;
        const int n_need = (int)__ldg(&A.need_count[v]);
        if (n_need == 0) continue;                                           // finished by the horizon pass
        bake_inter_vertex<ORDER, COUNT>(A, W, v, n_need, lane, S, depth, lt_mask, sgn, cand_tests, rays_scanned, node_visits, tri_tests);
    }
    if (COUNT && A.work) {
        const unsigned long long nv = warp_sum_u64(node_visits), nt = warp_sum_u64(tri_tests);
        if (lane == 0) { atomicAdd(&A.work[0], nv); atomicAdd(&A.work[1], nt); atomicAdd(&A.work[2], cand_tests); atomicAdd(&A.work[3], rays_scanned); }
    }
}

template <int ORDER, bool COUNT>
cudaError_t launch_inter_tc(const BakeArgs &A, int *grid, int n_sms, cudaStream_t st) {
    const int block = 128;
    const size_t smem = sizeof(InterShared) * (size_t)(block / 32);
    static std::atomic<unsigned long long> configured{0};   // per instantiation, one bit per device
    {
        cudaError_t e = ensure_dynamic_smem(bake_inter_kernel<ORDER, COUNT>, (int)smem, configured);
        if (e != cudaSuccess) return e;
    }
    if (*grid <= 0) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bake_inter_kernel<ORDER, COUNT>, block, smem);
        if (e != cudaSuccess) return e;
        *grid = n_sms * (per_sm > 0 ? per_sm : 1);
    }
    const long long need = ((long long)A.n_verts + 3) / 4;
    if (need < *grid) *grid = (int)(need > 0 ? need : 1);
    bake_inter_kernel<ORDER, COUNT><<<*grid, block, smem, st>>>(A);
    return cudaGetLastError();
}

template <int ORDER>
cudaError_t launch_inter_t(const BakeArgs &A, int *grid, int n_sms, cudaStream_t st) {
    return A.work ? launch_inter_tc<ORDER, true>(A, grid, n_sms, st) : launch_inter_tc<ORDER, false>(A, grid, n_sms, st);
}

}  // namespace

int bake_inter_max_samples() { return 1 << 16; }      // slot words carry 16-bit queue fields; sample indices are 24-bit

cudaError_t launch_bake_inter(const BakeArgs &A, int order, int *grid, int n_sms, cudaStream_t st) {
    switch (order) {
    case 1: return launch_inter_t<1>(A, grid, n_sms, st);
    case 2: return launch_inter_t<2>(A, grid, n_sms, st);
    case 3: return launch_inter_t<3>(A, grid, n_sms, st);
    case 4: return launch_inter_t<4>(A, grid, n_sms, st);
    case 5: return launch_inter_t<5>(A, grid, n_sms, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace prt
