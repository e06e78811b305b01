// horizon.cuh -- the per-vertex work of the horizon pass (horizon.cu): everything one persistent warp does for one vertex, as an
// inline device function, so that the same code also runs on the CPU test harness (tests/hostcheck: plain C++ against the warp
// emulator).  See horizon.cu for the design.
#pragma once
#include "kernels.h"
#include "entry_list.cuh"

namespace prt {

namespace {

struct HorizonShared {
    EntryList el;
    uint32_t hz[kHzWords];             // the map (first kHzBins words) and its range-minimum table
    uint32_t rq[kHzQueue];
    uint32_t tq[kHzTriQueue];
};

// One vertex: entry list, horizon map, need bits + count of the samples that are not above the map; a vertex without such a
// sample is finished here (row = the plain cosine-weighted projection, visibility words all ones).
template <int ORDER>
__device__ __forceinline__ void horizon_vertex(const BakeArgs &A, HorizonShared &W, const uint32_t v, const int lane, const int S, const int words,
                                               const float sgn) {
    constexpr int N2 = ORDER * ORDER;
    const float *pp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.pos) + (size_t)v * A.stride);
    const float *np = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.nrm) + (size_t)v * A.stride);
    const f3 N = mk3(__ldg(np), __ldg(np + 1), __ldg(np + 2));
    const f3 P = mk3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2));
    const Frame fr = make_frame(N);
    const f3 org = madd3(P, A.origin_eps, N);                       // raytracing.cpp:343

    // The map is built in the tangent frame and compared with the LOCAL z and azimuth bin of the samples, which is only valid when
    // the frame is orthonormal, i.e. |N| = 1: frame(N) (raytracing.cpp:101-107) has |up| = |N|, so a non-unit normal (the reference
    // passes assimp's normals through un-normalised, model.cpp:27) shears the ray directions against the sample table.  |N|^2 within
    // 2e-5 of 1 moves a direction by < 1e-5 in sin(elevation), a twentieth of the map's margin; any other normal (also NaN) gets no
    // map: every sample is traced by the traversal pass, whose arithmetic does not depend on |N|.
    const bool unit = fabsf(dot3(N, N) - 1.0f) <= 2e-5f;
    if (unit) {
        const int n_cand = build_entry_list(A.nodes, org, N, W.el, lane);
        build_horizon(W.el, n_cand, A.nodes, A.tris, org, N, fr, W.hz, W.rq, W.tq, A.horizon_budget, A.horizon_near2, lane, A.slabs, A.horizon_mid2, A.horizon_gain);
    }

    uint32_t *row = A.need_bits + (size_t)v * words;
    uint32_t total = 0u;
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        bool need = false;
        if (i < S) {
            const float4 smp = __ldg(&A.samples[i]);
            need = !unit || !(smp.z > __uint_as_float(W.hz[__float_as_uint(smp.w) >> 24]));
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
        if (lane < N2) store_row(A, v, N2, lane, mine * A.inv_S);
        if (A.vis) {
            for (int w = lane; w < words; w += 32) {
                const int rem = S - 32 * w;
                A.vis[(size_t)v * words + w] = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
            }
        }
    }
    __syncwarp();
}

}  // namespace

}  // namespace prt
