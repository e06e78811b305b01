// bake.cu -- sm_100a kernels for per-vertex diffuse SH transfer and batched ray queries.
//
// bake_kernel replaces the body of bake_SH (reference src/raytracing/raytracing.cpp:320-360) together with
// renderSH (:228-278): one persistent warp per vertex pulls vertices from an atomic counter, its 32 lanes
// trace the vertex's S stratified cosine-weighted rays (any-hit for the last path segment, closest-hit for
// interreflection segments), evaluate the SH basis of every escaping direction in registers and reduce
// across the warp with shuffles.  Lanes whose ray has terminated are refilled with the next samples of the
// same vertex (warp-level ray compaction), so accumulators never mix vertices.
//
// One trace per sample is shared by all order^2 coefficients (the reference re-traces per coefficient,
// raytracing.cpp:332-349; see DESIGN.md section 2 for why results are unchanged).
#include "kernels.h"
#include "traverse.cuh"
#include "entry_list.cuh"

#include <cuda_runtime.h>

namespace prt {

namespace {

// MODE 0: shadowed (depth 1, any-hit only)   MODE 1: interreflection   MODE 2: unshadowed Monte-Carlo (V == 1)
template <int ORDER, int MODE>
__global__ void __launch_bounds__(256) bake_kernel(const BakeArgs A) {
    constexpr int N2 = ORDER * ORDER;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EntryList &W = reinterpret_cast<EntryList *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float sgn = A.cs_phase ? -1.0f : 1.0f;
    const int depth = MODE == 1 ? A.depth : 1;
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

        // horizon pass ran first (interreflection): vertices with nothing to trace are finished, and of the others only the
        // flagged samples are traced -- every other primary ray provably escapes with weight 1 (raytracing.cpp:257-261)
        const uint32_t *need_row = (MODE == 1 && A.need_bits) ? A.need_bits + (size_t)v * A.vis_words : nullptr;
        int n_need = A.S;
        if (need_row) {
            n_need = (int)__ldg(&A.need_count[v]);
            if (n_need == 0) continue;
        }

        int n_cand = 0;
        if (MODE != 2 && A.entry_list) n_cand = build_entry_list(A.nodes, org, N, W, lane);
        cand_tests += (unsigned long long)n_cand * (unsigned long long)n_need;

        float acc[N2];
#pragma unroll
        for (int k = 0; k < N2; k++) acc[k] = 0.f;

        if (need_row) {
            for (int base = 0; base < A.S; base += 32) {
                const int i = base + lane;
                if (i < A.S && !((__ldg(&need_row[base >> 5]) >> lane) & 1u)) {
                    const float4 smp = __ldg(&A.samples[i]);
                    const f3 dir = to_world(fr, mk3(smp.x, smp.y, smp.z));
                    float y[N2];
                    sh_eval<ORDER>(dir.z, dir.x, dir.y, sgn, y);
#pragma unroll
                    for (int k = 0; k < N2; k++) acc[k] += y[k];
                    if (A.vis) { const uint32_t sr = __float_as_uint(smp.w) & 0xFFFFFFu; atomicOr(&A.vis[(size_t)v * A.vis_words + (sr >> 5)], 1u << (sr & 31u)); }
                }
            }
        }
        int need_word = -1, fetched = 0;      // cursor over the need bits (warp-uniform)
        uint32_t need_cur = 0u;

        int next = 0;
        bool active = false;
        Trav tr;
        tr.reset_counters();
        uint32_t cm[3] = {0u, 0u, 0u};   // candidate boxes hit by the lane's current primary ray, not yet traversed
        uint32_t sidx = 0;      // reference sample index s = i*samples_v + j of the lane's current path
        int seg = 0;            // path segment
        f3 pos = org;
        float Lw0 = 1.f, Lw1 = 1.f, Lw2 = 1.f;

        for (;;) {
            const unsigned idle = __ballot_sync(kFull, !active);
            if (need_row) {
                if (idle && fetched < n_need) {
                    // hand the next popc(idle) flagged samples to the idle lanes, in processing order
                    const int n_idle = __popc(idle), my_rank = __popc(idle & lt_mask);
                    int taken = 0, my_k = -1;
                    while (taken < n_idle && fetched + taken < n_need) {
                        if (!need_cur) { need_cur = __ldg(&need_row[++need_word]); continue; }
                        const int c = __popc(need_cur), take = min(c, n_idle - taken);
                        if (!active && my_rank >= taken && my_rank < taken + take) my_k = need_word * 32 + (int)__fns(need_cur, 0u, my_rank - taken + 1);
                        if (take == c) need_cur = 0u;
                        else need_cur &= ~((2u << __fns(need_cur, 0u, take)) - 1u);
                        taken += take;
                    }
                    fetched += taken;
                    if (my_k >= 0) {
                        const float4 smp = __ldg(&A.samples[my_k]);
                        sidx = __float_as_uint(smp.w) & 0xFFFFFFu;
                        const f3 dir = to_world(fr, mk3(smp.x, smp.y, smp.z));
                        pos = org; seg = 0; Lw0 = Lw1 = Lw2 = 1.f;
                        tr.init(pos, dir, 0.0f, INFINITY);
                        if (n_cand) scan_entry_list(W, n_cand, tr.idx, tr.idy, tr.idz, cm);
                        else tr.start_root();
                        active = true;
                    }
                }
                next = fetched < n_need ? 0 : A.S;               // "more samples waiting" for the refill test below
            } else
            if (idle && next < A.S) {
                const int k = next + __popc(idle & lt_mask);
                if (!active && k < A.S) {
                    const float4 smp = __ldg(&A.samples[k]);
                    sidx = __float_as_uint(smp.w) & 0xFFFFFFu;
                    const f3 dir = to_world(fr, mk3(smp.x, smp.y, smp.z));   // raytracing.cpp:340
                    pos = org; seg = 0; Lw0 = Lw1 = Lw2 = 1.f;
                    tr.init(pos, dir, 0.0f, INFINITY);
                    if (MODE != 2) {
                        if (n_cand) scan_entry_list(W, n_cand, tr.idx, tr.idy, tr.idz, cm);
                        else tr.start_root();
                    }
                    active = true;
                }
                next += __popc(idle);
            }
            if (!__any_sync(kFull, active)) break;
            if (!active) continue;

            const bool more = next < A.S;
            const bool closest = MODE == 1 && seg < depth - 1;
            int rc = TRAV_EMPTY;
            if (MODE != 2) {
                for (;;) {
                    rc = closest ? tr.template run<false>(A.nodes, A.tris, A.refill_thresh, more)
                                 : tr.template run<true>(A.nodes, A.tris, A.refill_thresh, more);
                    if (rc != TRAV_EMPTY) break;
                    // stack exhausted: continue with the next candidate the ray hit (deepest = nearest first)
                    int k;
                    if (cm[2]) { const int b = 31 - __clz(cm[2]); cm[2] &= ~(1u << b); k = 64 + b; }
                    else if (cm[1]) { const int b = 31 - __clz(cm[1]); cm[1] &= ~(1u << b); k = 32 + b; }
                    else if (cm[0]) { const int b = 31 - __clz(cm[0]); cm[0] &= ~(1u << b); k = b; }
                    else break;
                    const float4 g = W.cb[k];
                    tr.start_group(__float_as_uint(g.z), __float_as_uint(g.w));
                }
            }
            if (rc == TRAV_RUNNING) continue;
            const bool hit = closest ? (tr.best_prim != 0xFFFFFFFFu) : (rc == TRAV_HIT);

            if (!hit) {
                // environment reached: L = Lw * Y_lm(dir), sh-space (z,x,y)  (raytracing.cpp:226,257-261)
                float y[N2];
                sh_eval<ORDER>(tr.d.z, tr.d.x, tr.d.y, sgn, y);
#pragma unroll
                for (int k = 0; k < N2; k++) acc[k] = fmaf(Lw0, y[k], acc[k]);
                if (A.vis && seg == 0) atomicOr(&A.vis[(size_t)v * A.vis_words + (sidx >> 5)], 1u << (sidx & 31u));
                active = false;
                continue;
            }
            cm[0] = cm[1] = cm[2] = 0u;
            if (!closest) { active = false; continue; }                            // absorbed on the last segment
            {
                const f3 n = normalize3(tr.hit_ng(A.tris));                         // :263-264
                if (dot3(tr.d, n) >= -1e-4f) { active = false; continue; }          // :265
                pos = madd3(pos, tr.best_t, tr.d);                                  // :266
                float u, w;
                rand2(A.seed, A.vid_base + global_row(A, v), sidx, (uint32_t)seg, 1u, u, w);        // :267
                const f3 l = cosine_local(u, w);
                const float pdf = PRT_DIV(l.z, kPiF);
                const Frame fb = make_frame(n);
                const f3 nd = to_world(fb, l);
                if (pdf <= 1e-4f) { active = false; continue; }                     // :269
                Lw0 = PRT_MUL(Lw0, A.albedo[0]); Lw1 = PRT_MUL(Lw1, A.albedo[1]); Lw2 = PRT_MUL(Lw2, A.albedo[2]);  // :271
                const float sg = dot3(nd, n) < 0.0f ? -1.0f : 1.0f;                 // :273
                pos = madd3(pos, PRT_MUL(sg, A.bounce_eps), nd);                    // :274
                seg++;
                if (fmaxf(Lw0, fmaxf(Lw1, Lw2)) < 0.01f) { active = false; continue; }   // :249 (checked at loop top)
                tr.init(pos, nd, A.bounce_eps, INFINITY);                           // :275
                tr.start_root();
            }
        }

        // reduce the lane partial sums and store row v (raytracing.cpp:350: /= S)
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < N2; k++) {
            const float s = warp_sum(acc[k]);
            if (lane == k) mine = s;
        }
        if (lane < N2) store_row(A, v, N2, lane, mine * A.inv_S);
        node_visits += tr.n_node_visits; tri_tests += tr.n_tri_tests;
    }
    if (A.work) {
        // per-lane 32-bit counters are summed per warp at the end of its persistent loop
        const unsigned long long nv = warp_sum_u64(node_visits), nt = warp_sum_u64(tri_tests);
        if (lane == 0) { atomicAdd(&A.work[0], nv); atomicAdd(&A.work[1], nt); atomicAdd(&A.work[2], cand_tests); }
    }
}

// rotate_cos_lobe(norm) * INV_PI (reference src/scene/model.cpp:29-31): analytic unshadowed transfer
template <int ORDER>
__global__ void unshadowed_analytic_kernel(const BakeArgs A) {
    constexpr int N2 = ORDER * ORDER;
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.n_verts) return;
    const float *np = reinterpret_cast<const float *>(reinterpret_cast<const char *>(A.nrm) + (size_t)v * A.stride);
    const float nx = np[0], ny = np[1], nz = np[2];
    float y[N2];
    sh_eval<ORDER>(nz, nx, ny, A.cs_phase ? -1.0f : 1.0f, y);
    const float lobe[5] = { 1.0f, 2.0f / 3.0f, 0.25f, 0.0f, -1.0f / 24.0f };
    int k = 0;
#pragma unroll
    for (int l = 0; l < ORDER; l++)
        for (int m = -l; m <= l; m++, k++) store_row(A, v, N2, k, lobe[l] * y[k]);
}

__global__ void __launch_bounds__(128) trace_any_kernel(const Node8 *nodes, const Tri48 *tris, const float4 *rays, uint32_t n, uint8_t *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = rays[2 * (size_t)i], b = rays[2 * (size_t)i + 1];
    Trav t;
    t.reset_counters();
    t.init(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w);
    t.start_root();
    out[i] = t.run<true>(nodes, tris, 0, false) == TRAV_HIT ? 1 : 0;
}

__global__ void __launch_bounds__(128) trace_closest_kernel(const Node8 *nodes, const Tri48 *tris, const float4 *rays, uint32_t n,
                                                            float *out_t, uint32_t *out_prim, float *out_ng) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = rays[2 * (size_t)i], b = rays[2 * (size_t)i + 1];
    Trav t;
    t.reset_counters();
    t.init(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w);
    t.start_root();
    t.run<false>(nodes, tris, 0, false);
    const bool hit = t.best_prim != 0xFFFFFFFFu;
    f3 g = mk3(0.f, 0.f, 0.f);
    if (hit) g = t.hit_ng(tris);
    out_t[i] = hit ? t.best_t : INFINITY;
    out_prim[i] = hit ? t.best_prim : 0xFFFFFFFFu;
    if (out_ng) { out_ng[3 * (size_t)i] = g.x; out_ng[3 * (size_t)i + 1] = g.y; out_ng[3 * (size_t)i + 2] = g.z; }
}

// Pulls the BVH into L2 ahead of the first pass with streaming prefetches (one 128-byte line per thread and iteration): a bake that
// starts on a cold L2 otherwise pays one demand miss per line, serialised behind the traversal's dependent loads -- negligible for a
// 50 ms bake, a few per cent of the 6 ms shard of an 8-GPU bake.  Only for scenes that fit the L2 (abi.cu decides).
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const char *a, const size_t na, const char *b, const size_t nb) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 128, first = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    for (size_t o = first; o < na; o += stride) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
    for (size_t o = first; o < nb; o += stride) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
}

template <int ORDER, int MODE>
cudaError_t launch_persistent(const BakeArgs &A, int *grid, int block, int n_sms, cudaStream_t st) {
    const size_t smem = sizeof(EntryList) * (size_t)(block / 32);
    if (*grid <= 0) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bake_kernel<ORDER, MODE>, block, smem);
        if (e != cudaSuccess) return e;
        *grid = n_sms * (per_sm > 0 ? per_sm : 1);
    }
    // never launch more warps than vertices
    const int warps_per_block = block / 32;
    const long long need = ((long long)A.n_verts + warps_per_block - 1) / warps_per_block;
    if (need < *grid) *grid = (int)(need > 0 ? need : 1);
    bake_kernel<ORDER, MODE><<<*grid, block, smem, st>>>(A);
    return cudaGetLastError();
}

template <int ORDER>
cudaError_t launch_order(const BakeArgs &A, int mode, int *grid, int block, int n_sms, cudaStream_t st) {
    switch (mode) {
    case 0: return launch_persistent<ORDER, 0>(A, grid, block, n_sms, st);
    case 1: return launch_persistent<ORDER, 1>(A, grid, block, n_sms, st);
    case 2: return launch_persistent<ORDER, 2>(A, grid, block, n_sms, st);
    default:
        *grid = (int)((A.n_verts + 127) / 128);
        unshadowed_analytic_kernel<ORDER><<<*grid, 128, 0, st>>>(A);
        return cudaGetLastError();
    }
}

}  // namespace

cudaError_t launch_bake(const BakeArgs &A, int order, int mode, int *grid, int block, int n_sms, cudaStream_t st) {
    switch (order) {
    case 1: return launch_order<1>(A, mode, grid, block, n_sms, st);
    case 2: return launch_order<2>(A, mode, grid, block, n_sms, st);
    case 3: return launch_order<3>(A, mode, grid, block, n_sms, st);
    case 4: return launch_order<4>(A, mode, grid, block, n_sms, st);
    case 5: return launch_order<5>(A, mode, grid, block, n_sms, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_l2_prefetch(const void *a, size_t na, const void *b, size_t nb, int n_sms, cudaStream_t st) {
    l2_prefetch_kernel<<<n_sms * 8, 256, 0, st>>>((const char *)a, na, (const char *)b, nb);
    return cudaGetLastError();
}

cudaError_t launch_trace_any(const Node8 *nodes, const Tri48 *tris, const float *rays, uint32_t n, uint8_t *out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    trace_any_kernel<<<(n + 127) / 128, 128, 0, st>>>(nodes, tris, reinterpret_cast<const float4 *>(rays), n, out);
    return cudaGetLastError();
}

cudaError_t launch_trace_closest(const Node8 *nodes, const Tri48 *tris, const float *rays, uint32_t n, float *out_t,
                                 uint32_t *out_prim, float *out_ng, cudaStream_t st) {
    if (!n) return cudaSuccess;
    trace_closest_kernel<<<(n + 127) / 128, 128, 0, st>>>(nodes, tris, reinterpret_cast<const float4 *>(rays), n, out_t, out_prim, out_ng);
    return cudaGetLastError();
}

}  // namespace prt
