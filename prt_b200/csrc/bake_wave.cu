// bake_wave.cu -- warp-local wavefront kernel for shadowed per-vertex SH transfer (the headline kernel).
//
// Contract: reference bake_SH (src/raytracing/raytracing.cpp:320-360) with renderSH at depth 1 (:228-278); same
// results as bake_kernel (bake.cu).  Any-hit visibility is order independent, so instead of one stack per ray the
// warp keeps two work stacks in shared memory and runs every step in lockstep:
//
//   per vertex (one persistent warp; the per-vertex code is bake_wave_vertex in bake_wave.cuh, an inline device function that the
//   CPU test harness also runs, unmodified, on a warp emulator -- tests/test_wave_emulated.py):
//     build_entry_list      chain of nodes containing the shared origin -> candidate boxes (entry_list.cuh)
//     scan                  32 rays x all candidate boxes; every hit becomes an item
//                             subtree candidate  -> node stack  (ray, node)
//                             leaf candidate     -> leaf stack  (ray, first triangle, count bits)
//     node step             pops 32 (ray, node) items: each lane decodes ONE 80-byte node, tests its 8 child boxes -- three box axes and a
//                           fourth slab axis along the node's mean normal (Dop32, bvh8.h: a ray grazing the surface passes through many
//                           boxes beside the sheet of triangles in them; -21 % node visits, -24 % triangle tests) -- and pushes
//                           (ray, child) / (ray, leaf) items (one packed warp scan)
//     leaf step             pops 32 (ray, leaf) items: each lane runs <= 3 pinned triangle tests, sets the occlusion bit
//     project               every unoccluded sample direction -> SH basis in registers, shuffle reduction, row store
//
// Items of rays that are already known to be occluded are dropped when popped.  If a stack is full the lane falls
// back to an ordinary stack traversal of that subtree (Trav::run), so capacity never affects results.
//
// When the horizon pass (horizon.cu) ran first, only the samples it flagged are scanned; vertices it finished are skipped.
// Measured tuning (profiles/r1_final_ncu_summary.txt): 7 CTAs of 4 warps per SM at 72 registers without spills (the work
// counters live only in the COUNT variant used by instrumented launches), stacks of 256 items, new rays admitted only while
// both stacks are at most a quarter full.
#include "bake_wave.cuh"

namespace prt {

namespace {

// COUNT: carry the work counters (an instrumented, untimed launch of bench.py); the timed variant keeps those registers free
// DOP: the node test includes the fourth slab axis of the node (A.dops, bvh8.h Dop32)
template <int ORDER, bool TRACE, bool COUNT, bool DOP>
__global__ void __launch_bounds__(PRT_WAVE_BLOCK, PRT_WAVE_MINB) bake_wave_kernel(const BakeArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = A.S, words = A.vis_words;
    const size_t wstride = sizeof(WaveShared) + 4 * (size_t)((words + 3) & ~3);
    WaveShared &W = *reinterpret_cast<WaveShared *>(smem_raw + wstride * (threadIdx.x >> 5));
    uint32_t *const occl = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(&W) + sizeof(WaveShared));
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float sgn = A.cs_phase ? -1.0f : 1.0f;
    unsigned long long cand_tests = 0ull, rays_scanned = 0ull;
    uint32_t node_visits = 0u, tri_tests = 0u;

    // work list of the horizon pass (horizon.cu): the unfinished vertices, heaviest first
    const bool listed = TRACE && A.work_list != nullptr;

    for (;;) {
        uint32_t v = 0;
        if (lane == 0) {
            v = atomicAdd(A.counter, 1u);
            if (listed) {
                // the list's length is re-read per vertex (an L2 hit on lane 0 only) rather than kept in a register
                const uint32_t total = __ldcg(A.counter + 4);
                v = v < total ? __ldcg(A.work_list + v) : 0xFFFFFFFFu;
            }
        }
        v = __shfl_sync(kFull, v, 0);
        if (v >= A.n_verts) break;
        if (TRACE && !listed && A.need_count && __ldg(&A.need_count[v]) == 0u) continue;     // finished by the horizon pass
        bake_wave_vertex<ORDER, TRACE, COUNT, DOP>(A, W, occl, v, lane, S, words, lt_mask, sgn, cand_tests, rays_scanned, node_visits, tri_tests);
    }
    if (COUNT && A.work) {
        const unsigned long long nv = warp_sum_u64(node_visits), nt = warp_sum_u64(tri_tests);
        if (lane == 0) { atomicAdd(&A.work[0], nv); atomicAdd(&A.work[1], nt); atomicAdd(&A.work[2], cand_tests); atomicAdd(&A.work[3], rays_scanned); }
    }
}

template <int ORDER, bool TRACE, bool COUNT, bool DOP>
cudaError_t launch_wave_t(const BakeArgs &A, int *grid, int block, int n_sms, cudaStream_t st) {
    const size_t smem = (sizeof(WaveShared) + 4 * (size_t)((A.vis_words + 3) & ~3)) * (size_t)(block / 32);
    static std::atomic<unsigned long long> configured{0};   // per instantiation, one bit per device
    {
        cudaError_t e = ensure_dynamic_smem(bake_wave_kernel<ORDER, TRACE, COUNT, DOP>, (int)((sizeof(WaveShared) + kMaxS / 8) * (PRT_WAVE_BLOCK / 32)), configured);
        if (e != cudaSuccess) return e;
    }
    if (*grid <= 0) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bake_wave_kernel<ORDER, TRACE, COUNT, DOP>, block, smem);
        if (e != cudaSuccess) return e;
        *grid = n_sms * (per_sm > 0 ? per_sm : 1);
    }
    const int warps_per_block = block / 32;
    const long long need = ((long long)A.n_verts + warps_per_block - 1) / warps_per_block;
    if (need < *grid) *grid = (int)(need > 0 ? need : 1);
    bake_wave_kernel<ORDER, TRACE, COUNT, DOP><<<*grid, block, smem, st>>>(A);
    return cudaGetLastError();
}

template <int ORDER>
cudaError_t launch_wave_o(const BakeArgs &A, bool trace, int *grid, int block, int n_sms, cudaStream_t st) {
    if (!trace) return A.work ? launch_wave_t<ORDER, false, true, false>(A, grid, block, n_sms, st) : launch_wave_t<ORDER, false, false, false>(A, grid, block, n_sms, st);
    if (A.dops) return A.work ? launch_wave_t<ORDER, true, true, true>(A, grid, block, n_sms, st) : launch_wave_t<ORDER, true, false, true>(A, grid, block, n_sms, st);
    return A.work ? launch_wave_t<ORDER, true, true, false>(A, grid, block, n_sms, st) : launch_wave_t<ORDER, true, false, false>(A, grid, block, n_sms, st);
}

}  // namespace

int bake_wave_max_samples() { return kMaxS; }
int bake_wave_block() { return PRT_WAVE_BLOCK; }

cudaError_t launch_bake_wave(const BakeArgs &A, int order, bool trace, int *grid, int block, int n_sms, cudaStream_t st) {
    switch (order) {
    case 1: return launch_wave_o<1>(A, trace, grid, block, n_sms, st);
    case 2: return launch_wave_o<2>(A, trace, grid, block, n_sms, st);
    case 3: return launch_wave_o<3>(A, trace, grid, block, n_sms, st);
    case 4: return launch_wave_o<4>(A, trace, grid, block, n_sms, st);
    case 5: return launch_wave_o<5>(A, trace, grid, block, n_sms, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace prt
