// tests/hostcheck/hostcheck_warp.cpp -- TEST TOOLING ONLY (never linked into libprt_b200.so).
// Runs the product's warp-cooperative horizon code -- build_entry_list + build_horizon of prt_b200/csrc/entry_list.cuh, the body
// of horizon_kernel (horizon.cu) up to the classification -- UNMODIFIED on the CPU: the header is compiled as plain C++ against the
// warp emulator (warp_emu.h: one thread per lane, collectives are rendezvous).  The CPU test-suite checks the maps against the
// oracle's per-ray visibility (tests/test_horizon_math.py): a sample the map frees must be visible.
#define PRT_HOSTCHECK 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#include "warp_emu.h"
#include <atomic>
// work counters of the builder (entry_list.cuh calls PRT_HZ_STAT from the lanes that do the work)
namespace { struct HzStats { std::atomic<uint64_t> iterations{0}, nodes_expanded{0}, boxes_bounded{0}, triangle_rounds{0}; } g_hz_stats; }
#define PRT_HZ_STAT(counter, n) (g_hz_stats.counter.fetch_add((n), std::memory_order_relaxed))
// far boxes whose cheap bound was merged into the map: (centre, half extents, value, first bin, last bin), for the bound studies
#include <mutex>
#include <vector>
namespace { std::mutex g_trace_mu; std::vector<float> g_trace; bool g_trace_on = false; }
#define PRT_HZ_TRACE_FAR(c, e, item) do { if (g_trace_on) { std::lock_guard<std::mutex> lk(g_trace_mu); \
    const float rec[9] = {(c).x, (c).y, (c).z, (e).x, (e).y, (e).z, (item).v, (float)(item).b0, (float)(item).b1}; g_trace.insert(g_trace.end(), rec, rec + 9); } } while (0)
#include "../../prt_b200/csrc/bvh8.h"
// step / overflow counters of the traversal pass (bake_wave.cuh calls PRT_WAVE_STAT)
namespace { struct WaveStats { std::atomic<uint64_t> leaf_steps{0}, leaf_lanes{0}, node_steps{0}, node_lanes{0}, scan_steps{0}, scan_lanes{0}, overflow_subtrees{0}, overflow_leaves{0}; } g_wave_stats; }
extern "C" void hc_wave_step_stats(uint64_t *out, int reset) {
    std::atomic<uint64_t> *a[8] = {&g_wave_stats.leaf_steps, &g_wave_stats.leaf_lanes, &g_wave_stats.node_steps,
                                   &g_wave_stats.node_lanes, &g_wave_stats.scan_steps, &g_wave_stats.scan_lanes, &g_wave_stats.overflow_subtrees, &g_wave_stats.overflow_leaves};
    for (int i = 0; i < 8; i++) { out[i] = *a[i]; if (reset) *a[i] = 0; }
}
#define PRT_WAVE_STAT(counter, n) (g_wave_stats.counter.fetch_add((n), std::memory_order_relaxed))
// ---- study: a fourth slab axis per node (the node's mean normal m, shared by its eight children; every child carries its extent along m,
// quantised to 8 bits over the node's own extent) tested together with the three box axes -- how many hit children would it remove?
// mode 0: off, 1: quantised (255 steps, one step of padding either side), 2: exact child extents (upper bound of what the idea can do)
#include "../../prt_b200/csrc/kernels.h"
#include "../../prt_b200/csrc/traverse.cuh"
namespace {
int g_dop_mode = 0;
std::vector<float> g_dop;          // [n_nodes][8][2] child extents along the parent's m
std::atomic<uint64_t> g_dop_tests{0}, g_dop_culled_inner{0}, g_dop_culled_leaf{0};
}
namespace prt { namespace {
inline void dop_study(const BakeArgs &A, uint32_t node, const f3 o, const f3 d, uint32_t &inner8, uint32_t &leaf8) {
    if (!g_dop_mode || !A.slabs || g_dop.empty()) return;
    const Slab32 &sl = A.slabs[node];
    if (sl.mx == 0.f && sl.my == 0.f && sl.mz == 0.f) return;
    const Node8 &nd = A.nodes[node];
    const double so = (double)sl.mx * o.x + (double)sl.my * o.y + (double)sl.mz * o.z, sd = (double)sl.mx * d.x + (double)sl.my * d.y + (double)sl.mz * d.z;
    const float lo3[3] = {nd.px, nd.py, nd.pz};
    const float sc[3] = {PRT_U2F((uint32_t)nd.ex << 23), PRT_U2F((uint32_t)nd.ey << 23), PRT_U2F((uint32_t)nd.ez << 23)};
    const uint8_t *ql[3] = {nd.qlox, nd.qloy, nd.qloz}, *qh[3] = {nd.qhix, nd.qhiy, nd.qhiz};
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    const double T = std::max((double)sl.d1 - (double)sl.d0, 1e-30), step = T / 255.0;
    uint32_t bits = inner8 | leaf8;
    while (bits) {
        const int s = __builtin_ctz(bits); bits &= bits - 1u;
        g_dop_tests++;
        double t0 = 0.0, t1 = INFINITY;
        for (int a = 0; a < 3; a++) {
            const double id = 1.0 / (std::fabs(dd[a]) < 1e-18 ? 1e-18 : dd[a]);
            const double ta = (lo3[a] + ql[a][s] * sc[a] - oo[a]) * id, tb = (lo3[a] + qh[a][s] * sc[a] - oo[a]) * id;
            t0 = std::max(t0, std::min(ta, tb)); t1 = std::min(t1, std::max(ta, tb));
        }
        double lo = g_dop[((size_t)node * 8 + s) * 2], hi = g_dop[((size_t)node * 8 + s) * 2 + 1];
        if (g_dop_mode == 1) {
            lo = sl.d0 + (std::floor((lo - sl.d0) / step) - 1.0) * step; hi = sl.d0 + (std::ceil((hi - sl.d0) / step) + 1.0) * step;
        }
        // s(t) = so + t sd must meet [lo, hi] for some t in [t0, t1]
        const double a = so + t0 * sd, b = so + (std::isfinite(t1) ? t1 * sd : (sd > 0 ? INFINITY : sd < 0 ? -INFINITY : 0.0));
        if (std::max(a, b) < lo || std::min(a, b) > hi) {
            if ((inner8 >> s) & 1u) { inner8 &= ~(1u << s); g_dop_culled_inner++; } else { leaf8 &= ~(1u << s); g_dop_culled_leaf++; }
        }
    }
}
} }
#define PRT_WAVE_NODE_STUDY(A, node, org, d, inner8, leaf8) dop_study(A, node, org, d, inner8, leaf8)
// fourth slab axis of the node test in the traversal pass (Dop32): on by default, as in the product
namespace { bool g_wave_dop = true; }
extern "C" void hc_wave_dop(int on) { g_wave_dop = on != 0; }
extern "C" void hc_dop_study(void *h, int mode, uint64_t *out) {
    using namespace prt;
    HostBVH8 *b = (HostBVH8 *)h;
    if (out) { out[0] = g_dop_tests; out[1] = g_dop_culled_inner; out[2] = g_dop_culled_leaf; }
    g_dop_tests = 0; g_dop_culled_inner = 0; g_dop_culled_leaf = 0;
    g_dop_mode = mode;
    if (!mode || !b || !b->slabs || !g_dop.empty()) return;
    // triangle range below every node (children are emitted after their parent)
    std::vector<uint32_t> lo(b->n_nodes, 0xFFFFFFFFu), hi(b->n_nodes, 0u);
    for (uint32_t x = b->n_nodes; x-- > 0;) {
        const Node8 &nd = b->nodes[x];
        uint32_t rank = 0;
        for (int s = 0; s < 8; s++) {
            if (!nd.meta[s]) continue;
            if ((nd.imask >> s) & 1) { const uint32_t ch = nd.child_base + rank++; lo[x] = std::min(lo[x], lo[ch]); hi[x] = std::max(hi[x], hi[ch]); }
            else { const uint32_t t0 = nd.tri_base + (nd.meta[s] & 31u), cnt = (uint32_t)__builtin_popcount(nd.meta[s] >> 5); lo[x] = std::min(lo[x], t0); hi[x] = std::max(hi[x], t0 + cnt); }
        }
    }
    g_dop.assign((size_t)b->n_nodes * 16, 0.f);
    for (uint32_t x = 0; x < b->n_nodes; x++) {
        const Node8 &nd = b->nodes[x];
        const Slab32 &sl = b->slabs[x];
        uint32_t rank = 0;
        for (int s = 0; s < 8; s++) {
            if (!nd.meta[s]) continue;
            uint32_t t0, t1;
            if ((nd.imask >> s) & 1) { const uint32_t ch = nd.child_base + rank++; t0 = lo[ch]; t1 = hi[ch]; }
            else { t0 = nd.tri_base + (nd.meta[s] & 31u); t1 = t0 + (uint32_t)__builtin_popcount(nd.meta[s] >> 5); }
            double mn = 3e38, mx = -3e38;
            for (uint32_t t = t0; t < t1; t++) {
                const Tri48 &T = b->tris[t];
                const double v[3][3] = {{T.v0x, T.v0y, T.v0z}, {T.v0x + T.e1x, T.v0y + T.e1y, T.v0z + T.e1z}, {T.v0x + T.e2x, T.v0y + T.e2y, T.v0z + T.e2z}};
                for (int k = 0; k < 3; k++) { const double q = sl.mx * v[k][0] + sl.my * v[k][1] + sl.mz * v[k][2]; mn = std::min(mn, q); mx = std::max(mx, q); }
            }
            g_dop[((size_t)x * 8 + s) * 2] = (float)(mn - b->pad); g_dop[((size_t)x * 8 + s) * 2 + 1] = (float)(mx + b->pad);
        }
    }
}
#include "../../prt_b200/csrc/entry_list.cuh"
#include "../../prt_b200/csrc/horizon.cuh"
#include "../../prt_b200/csrc/bake_wave.cuh"
#include "../../prt_b200/csrc/bake_inter.cuh"
#include <thread>
#include <vector>

using namespace prt;

// oriented slabs of the nodes (bvh8.h) in the horizon builder: on by default, as in the product; the studies switch them off to compare
namespace { bool g_use_slabs = true; int g_mid100 = 24; float g_gain = 0.2f; }
// fourth slab axis of the node test in the traversal pass: on by default, as in the product
extern "C" void hc_wave_dop(int on);
extern "C" void hc_use_slabs(int on) { g_use_slabs = on != 0; }
// refinement rule for mid-sized boxes (build_horizon: mid2, gain_min): angular radius x 100 (0 = rule off), gain in units of S / 32 samples
extern "C" void hc_horizon_mid(int mid100, float gain) { g_mid100 = mid100; g_gain = gain; }
namespace { float mid2_of(int mid100) { if (mid100 <= 0) return 0.f; const float s = sinf(0.01f * (float)mid100); return 1.0f / (s * s); } }

// h: HostBVH8* from hc_build.  pos / nrm: n x 3 floats.  near100: horizon_near (angular radius x 100, rad); budget: horizon_budget.
// out_hz: n x kHzBins floats (the map), out_ncand: n (entry-list candidates).
// stats (optional, 4 x uint64): refinement iterations, nodes expanded, boxes bounded, triangle rounds -- summed over the n vertices.
extern "C" void hc_horizon_maps(void *h, const float *pos, const float *nrm, uint32_t n, float origin_eps, int budget, int near100,
                                float *out_hz, int *out_ncand, uint64_t *stats) {
    HostBVH8 *b = (HostBVH8 *)h;
    warp_emu::State state;
    warp_emu::g_state = &state;
    static HorizonShared W;                 // the per-warp shared memory of horizon_kernel (horizon.cuh)
    g_hz_stats.iterations = 0; g_hz_stats.nodes_expanded = 0; g_hz_stats.boxes_bounded = 0; g_hz_stats.triangle_rounds = 0;
    const float sn = sinf(0.01f * (float)near100), near2 = 1.0f / (sn * sn);
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++) {
        lanes.emplace_back([&, lane]() {
            warp_emu::t_lane = lane;
            for (uint32_t v = 0; v < n; v++) {
                const f3 N = mk3(nrm[3 * v], nrm[3 * v + 1], nrm[3 * v + 2]);
                const f3 P = mk3(pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]);
                const Frame fr = make_frame(N);
                const f3 org = madd3(P, origin_eps, N);
                const int n_cand = build_entry_list(b->nodes, org, N, W.el, lane);
                build_horizon(W.el, n_cand, b->nodes, b->tris, org, N, fr, W.hz, W.rq, W.tq, budget, near2, lane, g_use_slabs ? b->slabs : nullptr, mid2_of(g_mid100), g_gain);
                out_hz[(size_t)v * kHzBins + lane] = __uint_as_float(W.hz[lane]);
                if (lane == 0) out_ncand[v] = n_cand;
                __syncwarp();
            }
        });
    }
    for (auto &t : lanes) t.join();
    warp_emu::g_state = nullptr;
    if (stats) {
        // iterations and triangle rounds are counted by all 32 lanes of the warp
        stats[0] = g_hz_stats.iterations / 32; stats[1] = g_hz_stats.nodes_expanded; stats[2] = g_hz_stats.boxes_bounded;
        stats[3] = g_hz_stats.triangle_rounds / 32;
    }
}

// bound study: the far boxes merged while building the map of ONE vertex; returns the number of records (9 floats each) written
extern "C" uint32_t hc_horizon_trace_far(void *h, const float *pos, const float *nrm, float origin_eps, int budget, int near100, float *out, uint32_t cap) {
    g_trace.clear(); g_trace_on = true;
    float hz[kHzBins]; int nc;
    hc_horizon_maps(h, pos, nrm, 1, origin_eps, budget, near100, hz, &nc, nullptr);
    g_trace_on = false;
    const uint32_t n = (uint32_t)std::min<size_t>(g_trace.size() / 9, cap);
    std::memcpy(out, g_trace.data(), (size_t)n * 9 * sizeof(float));
    return n;
}

// ---- the traversal pass: bake_wave_vertex of prt_b200/csrc/bake_wave.cuh (what one persistent warp of bake_wave_kernel does for
// one vertex), unmodified, on the warp emulator.  samples: S x 4 floats in PROCESSING order (local direction, w = reference sample
// index | azimuth bin << 24 -- the table abi.cu uploads); need_bits: optional [n][words] flags of the horizon pass (processing
// order); out: [n][order^2] rows; vis: optional [n][words] visibility words (reference order, pre-zeroed).
namespace {
struct WaveSharedHost { WaveShared W; uint32_t occl[kMaxS / 32]; };

// work (optional, 4 x uint64): node visits, triangle tests, entry-list box tests, rays scanned -- the counters of an instrumented
// launch (COUNT variant of the kernel), summed over lanes as the kernel does
template <int ORDER, bool COUNT>
void run_wave(const BakeArgs &A, uint64_t *work) {
    warp_emu::State state;
    warp_emu::g_state = &state;
    static WaveSharedHost sh;
    std::atomic<uint64_t> w_nv{0}, w_nt{0}, w_cand{0}, w_scanned{0};
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++) {
        lanes.emplace_back([&, lane]() {
            warp_emu::t_lane = lane;
            unsigned long long cand = 0ull, scanned = 0ull;
            uint32_t nv = 0u, nt = 0u;
            const float sgn = A.cs_phase ? -1.0f : 1.0f;
            for (uint32_t v = 0; v < A.n_verts; v++)
                if (A.dops) bake_wave_vertex<ORDER, true, COUNT, true>(A, sh.W, sh.occl, v, lane, A.S, A.vis_words, (1u << lane) - 1u, sgn, cand, scanned, nv, nt);
                else bake_wave_vertex<ORDER, true, COUNT, false>(A, sh.W, sh.occl, v, lane, A.S, A.vis_words, (1u << lane) - 1u, sgn, cand, scanned, nv, nt);
            w_nv += nv; w_nt += nt;
            if (lane == 0) { w_cand += cand; w_scanned += scanned; }      // warp-uniform counters
        });
    }
    for (auto &t : lanes) t.join();
    warp_emu::g_state = nullptr;
    if (work) { work[0] = w_nv; work[1] = w_nt; work[2] = w_cand; work[3] = w_scanned; }
}
}

extern "C" int hc_bake_wave(void *h, const float *pos, const float *nrm, uint32_t n, const float *samples, int S, int order,
                            const uint32_t *need_bits, float origin_eps, int cs_phase, float *out, uint32_t *vis, uint64_t *work) {
    if (S < 1 || S > kMaxS || order < 1 || order > 5) return -1;
    HostBVH8 *b = (HostBVH8 *)h;
    BakeArgs A{};
    A.nodes = b->nodes; A.tris = b->tris; A.pos = pos; A.nrm = nrm; A.stride = 12; A.n_verts = n;
    A.samples = reinterpret_cast<const float4 *>(samples); A.S = S; A.inv_S = 1.0f / (float)S;
    A.out = out; A.vis = vis; A.vis_words = (S + 31) / 32;
    A.need_bits = const_cast<uint32_t *>(need_bits);
    A.origin_eps = origin_eps; A.cs_phase = cs_phase;
    A.slabs = b->slabs;                 // the traversal pass does not read them; the DOP study hook does
    A.dops = g_wave_dop ? b->dops : nullptr;
    switch (order) {
    case 1: run_wave<1, false>(A, nullptr); break;
    case 2: run_wave<2, false>(A, nullptr); break;
    case 3: if (work) run_wave<3, true>(A, work); else run_wave<3, false>(A, nullptr); break;
    case 4: run_wave<4, false>(A, nullptr); break;
    default: run_wave<5, false>(A, nullptr); break;
    }
    return 0;
}

// ---- interreflection: bake_inter_vertex of prt_b200/csrc/bake_inter.cuh (one persistent warp of bake_inter_kernel, one vertex),
// unmodified, on the warp emulator.  need_bits [n][words] / need_count [n]: output of the horizon pass (or all samples flagged).
namespace {
template <int ORDER>
void run_inter(const BakeArgs &A) {
    warp_emu::State state;
    warp_emu::g_state = &state;
    static InterShared W;
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++) {
        lanes.emplace_back([&, lane]() {
            warp_emu::t_lane = lane;
            unsigned long long cand = 0ull, scanned = 0ull;
            uint32_t nv = 0u, nt = 0u;
            const float sgn = A.cs_phase ? -1.0f : 1.0f;
            for (uint32_t v = 0; v < A.n_verts; v++) {
                const int n_need = (int)A.need_count[v];
                if (n_need == 0) continue;                      // the kernel skips the vertices the horizon pass finished
                bake_inter_vertex<ORDER, false>(A, W, v, n_need, lane, A.S, A.depth, (1u << lane) - 1u, sgn, cand, scanned, nv, nt);
            }
        });
    }
    for (auto &t : lanes) t.join();
    warp_emu::g_state = nullptr;
}
}

extern "C" int hc_bake_inter(void *h, const float *pos, const float *nrm, uint32_t n, uint32_t vid_base, const float *samples, int S, int order,
                             const uint32_t *need_bits, const uint32_t *need_count, uint32_t seed, int bounces, const float *albedo,
                             float origin_eps, float bounce_eps, float *out, uint32_t *vis) {
    if (S < 1 || S > 4096 || order < 1 || order > 5 || !need_bits || !need_count) return -1;
    HostBVH8 *b = (HostBVH8 *)h;
    BakeArgs A{};
    A.nodes = b->nodes; A.tris = b->tris; A.pos = pos; A.nrm = nrm; A.stride = 12; A.n_verts = n; A.vid_base = vid_base;
    A.samples = reinterpret_cast<const float4 *>(samples); A.S = S; A.inv_S = 1.0f / (float)S;
    A.out = out; A.vis = vis; A.vis_words = (S + 31) / 32;
    A.need_bits = const_cast<uint32_t *>(need_bits); A.need_count = const_cast<uint32_t *>(need_count);
    A.seed = seed; A.depth = bounces + 1;
    A.albedo[0] = albedo[0]; A.albedo[1] = albedo[1]; A.albedo[2] = albedo[2];
    A.origin_eps = origin_eps; A.bounce_eps = bounce_eps;
    switch (order) {
    case 1: run_inter<1>(A); break;
    case 2: run_inter<2>(A); break;
    case 3: run_inter<3>(A); break;
    case 4: run_inter<4>(A); break;
    default: run_inter<5>(A); break;
    }
    return 0;
}

// study helper: the entry list of ONE vertex (candidate boxes relative to the ray origin): out = n_cand x 8 floats
// (centre xyz, half extents xyz, group x, group y as raw bits); returns n_cand
extern "C" int hc_entry_list(void *h, const float *pos, const float *nrm, float origin_eps, float *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    warp_emu::State state;
    warp_emu::g_state = &state;
    static EntryList el;
    int n_cand = 0;
    const f3 N = mk3(nrm[0], nrm[1], nrm[2]);
    const f3 org = madd3(mk3(pos[0], pos[1], pos[2]), origin_eps, N);
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++)
        lanes.emplace_back([&, lane]() { warp_emu::t_lane = lane; const int n = build_entry_list(b->nodes, org, N, el, lane); if (lane == 0) n_cand = n; });
    for (auto &t : lanes) t.join();
    warp_emu::g_state = nullptr;
    for (int k = 0; k < n_cand; k++) {
        const float rec[8] = {el.ca[k].x, el.ca[k].y, el.ca[k].z, el.cb[k].x, el.cb[k].y, el.ca[k].w, el.cb[k].z, el.cb[k].w};     // centre, half extents x y z, group
        std::memcpy(out + 8 * k, rec, sizeof rec);
    }
    return n_cand;
}

// ---- the horizon pass as the kernel runs it: horizon_vertex of prt_b200/csrc/horizon.cuh (entry list, map, need bits and counts,
// rows + visibility words of the vertices it finishes), unmodified, on the warp emulator
namespace {
template <int ORDER>
void run_horizon(const BakeArgs &A) {
    warp_emu::State state;
    warp_emu::g_state = &state;
    static HorizonShared W;
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++) {
        lanes.emplace_back([&, lane]() {
            warp_emu::t_lane = lane;
            const float sgn = A.cs_phase ? -1.0f : 1.0f;
            for (uint32_t v = 0; v < A.n_verts; v++) horizon_vertex<ORDER>(A, W, v, lane, A.S, A.vis_words, sgn);
        });
    }
    for (auto &t : lanes) t.join();
    warp_emu::g_state = nullptr;
}
}

extern "C" int hc_horizon_pass(void *h, const float *pos, const float *nrm, uint32_t n, const float *samples, int S, int order, int budget,
                               int near100, float origin_eps, int cs_phase, float *out, uint32_t *vis, uint32_t *need_bits, uint32_t *need_count) {
    if (S < 1 || order < 1 || order > 5 || !need_bits || !need_count) return -1;
    HostBVH8 *b = (HostBVH8 *)h;
    BakeArgs A{};
    A.nodes = b->nodes; A.tris = b->tris; A.pos = pos; A.nrm = nrm; A.stride = 12; A.n_verts = n;
    A.samples = reinterpret_cast<const float4 *>(samples); A.S = S; A.inv_S = 1.0f / (float)S;
    A.out = out; A.vis = vis; A.vis_words = (S + 31) / 32;
    A.need_bits = need_bits; A.need_count = need_count;
    A.origin_eps = origin_eps; A.cs_phase = cs_phase;
    A.horizon_budget = budget;
    { const float sn = sinf(0.01f * (float)near100); A.horizon_near2 = 1.0f / (sn * sn); }
    A.slabs = g_use_slabs ? b->slabs : nullptr; A.horizon_mid2 = mid2_of(g_mid100); A.horizon_gain = g_gain;
    switch (order) {
    case 1: run_horizon<1>(A); break;
    case 2: run_horizon<2>(A); break;
    case 3: run_horizon<3>(A); break;
    case 4: run_horizon<4>(A); break;
    default: run_horizon<5>(A); break;
    }
    return 0;
}
