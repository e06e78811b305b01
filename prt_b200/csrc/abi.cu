// abi.cu -- the C ABI of libprt_b200.so (include/prt_b200.h): contexts, scenes, ray queries and the
// per-vertex transfer bake.  Host C++ only calls CUDA from here; there is no CPU compute path -- every entry
// point fails with PRT_ERR_CUDA when no usable GPU is present.
#include "../../include/prt_b200.h"
#include "bvh8.h"
#include "kernels.h"
#include "prt_math.cuh"
#include "abi_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace prt;

namespace {
thread_local std::string g_err;
int set_err(int code, const std::string &m) { g_err = m; return code; }
#define CU_TRY(expr)                                                                                       \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess) return set_err(PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
}  // namespace

struct prt_ctx {
    int device = 0;
    int n_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, evh = nullptr;
    // tuning
    int block = 256, ctas_per_sm = 0 /* 0 = occupancy */, refill_thresh = 8, count_work = 0, entry_list = 1, pair_queue = 2, horizon = 1, horizon_budget = 64, horizon_near = 157, horizon_mid = 24, horizon_gain = 64 /* tenths of a sample */, horizon_slabs = 1, wave_dop = 1, work_list_on = -1 /* -1 = auto */, l2_prefetch = 0;
    // cached sample table (the device copy is only replaced after the bake that last read it has finished: ev_tab)
    DevBuf samples; int s_ru = -1, s_rv = -1, s_jit = -1; uint32_t s_seed = 0;
    std::vector<float> h_samples;
    cudaEvent_t ev_tab = nullptr; bool tab_in_use = false;
    DevBuf counter, d_pos, d_nrm, d_out, d_vis, d_rays, d_res, need_bits, need_count, work_list;
    prt_bake_stats stats{};
    bool stats_pending = false, work_pending = false;
    // grow-only scratch buffers and a phase timer for the other entry points (env.cu, volume.cu): no cudaMalloc / cudaFree per call
    DevBuf scratch[8];
    cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr; bool phase_timed = false;
};

struct prt_scene {
    prt_ctx *ctx = nullptr;
    Node8 *d_nodes = nullptr;
    Tri48 *d_tris = nullptr;
    Slab32 *d_slabs = nullptr;          // oriented slab of every node (horizon pass)
    Dop32 *d_dops = nullptr;            // fourth slab axis of every node (traversal pass)
    prt_scene_info info{};
};

int prt_set_error(int code, const std::string &msg) { return set_err(code, msg); }
void *prt_ctx_scratch(prt_ctx *c, int slot, size_t bytes) {
    if (!c || slot < 0 || slot >= 8) return nullptr;
    if (c->scratch[slot].reserve(bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return c->scratch[slot].p;
}
void prt_ctx_timer_begin(prt_ctx *c, cudaStream_t st) { cudaEventRecord(c->ev_p0, st); c->phase_timed = false; }
void prt_ctx_timer_end(prt_ctx *c, cudaStream_t st) { cudaEventRecord(c->ev_p1, st); c->phase_timed = true; }
cudaStream_t prt_ctx_stream(prt_ctx *c) { return c->stream; }
int prt_ctx_sms(const prt_ctx *c) { return c->n_sms; }
int prt_ctx_refill_thresh(const prt_ctx *c) { return c->refill_thresh; }
int prt_ctx_entry_list(const prt_ctx *c) { return c->entry_list; }
prt_scene_view prt_scene_get_view(prt_scene *s) { return prt_scene_view{s->d_nodes, s->d_tris, s->ctx}; }

// host BVH build shared by prt_scene_create and prt_group_scene_create (one build, one upload per GPU)
int prt_build_host_bvh(const float *pos, size_t stride, uint32_t nv, const uint32_t *idx, uint32_t nt, HostBVH8 *h) {
    if (!pos || !idx || nv == 0 || nt == 0) return set_err(PRT_ERR_INVALID, "prt_scene_create: empty mesh");
    if (stride == 0) stride = 12;
    if (stride % 4) return set_err(PRT_ERR_INVALID, "prt_scene_create: stride must be a multiple of 4");
    char err[256] = {0};
    if (build_bvh8(pos, stride, nv, idx, nt, h, err, sizeof err) != 0) return set_err(PRT_ERR_BUILD, err);
    return PRT_OK;
}

int prt_scene_from_host_bvh(prt_ctx *c, const HostBVH8 *hp, prt_scene **out) {
    const HostBVH8 &h = *hp;
    CU_TRY(cudaSetDevice(c->device));
    prt_scene *s = new prt_scene();
    s->ctx = c;
    auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMalloc(&s->d_nodes, sizeof(Node8) * (size_t)h.n_nodes);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_tris, sizeof(Tri48) * (size_t)h.n_tris);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_nodes, h.nodes, sizeof(Node8) * (size_t)h.n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->d_tris, h.tris, sizeof(Tri48) * (size_t)h.n_tris, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && h.slabs) {
        e = cudaMalloc(&s->d_slabs, sizeof(Slab32) * (size_t)h.n_nodes);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_slabs, h.slabs, sizeof(Slab32) * (size_t)h.n_nodes, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && h.dops) {
        e = cudaMalloc(&s->d_dops, sizeof(Dop32) * (size_t)h.n_nodes);
        if (e == cudaSuccess) e = cudaMemcpy(s->d_dops, h.dops, sizeof(Dop32) * (size_t)h.n_nodes, cudaMemcpyHostToDevice);
    }
    s->info.n_tris = h.n_tris; s->info.n_nodes = h.n_nodes; s->info.max_depth = h.max_depth;
    s->info.node_bytes = sizeof(Node8) * (uint64_t)h.n_nodes; s->info.tri_bytes = sizeof(Tri48) * (uint64_t)h.n_tris;
    s->info.build_seconds = h.build_seconds; s->info.sah_cost = h.sah_cost;
    s->info.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (e != cudaSuccess) {
        if (s->d_nodes) cudaFree(s->d_nodes);
        if (s->d_tris) cudaFree(s->d_tris);
        if (s->d_slabs) cudaFree(s->d_slabs);
        if (s->d_dops) cudaFree(s->d_dops);
        delete s;
        return set_err(e == cudaErrorMemoryAllocation ? PRT_ERR_NOMEM : PRT_ERR_CUDA, std::string("prt_scene_create: ") + cudaGetErrorString(e));
    }
    *out = s;
    return PRT_OK;
}

extern "C" {

const char *prt_last_error(void) { return g_err.c_str(); }
int prt_abi_version(void) { return PRT_B200_ABI_VERSION; }

int prt_ctx_create(int device_id, prt_ctx **out) {
    if (!out) return set_err(PRT_ERR_INVALID, "prt_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(PRT_ERR_CUDA, std::string("prt_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device_id < 0) CU_TRY(cudaGetDevice(&device_id));
    if (device_id >= n) return set_err(PRT_ERR_INVALID, "prt_ctx_create: device id out of range");
    CU_TRY(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10) return set_err(PRT_ERR_CUDA, "prt_ctx_create: device is not sm_100-class; kernels are built for sm_100a only");
    prt_ctx *c = new prt_ctx();
    c->device = device_id; c->n_sms = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // stream-ordered allocations (probe capture) keep up to 16 GB cached in the device's default pool instead of returning it
        // to the driver at every synchronisation
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess && pool) {
            uint64_t keep = 16ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else cudaGetLastError();
    }
    CU_TRY(cudaEventCreate(&c->ev0)); CU_TRY(cudaEventCreate(&c->ev1));
    CU_TRY(cudaEventCreate(&c->ev2)); CU_TRY(cudaEventCreate(&c->ev3)); CU_TRY(cudaEventCreate(&c->evh));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_tab, cudaEventDisableTiming));
    CU_TRY(cudaEventCreate(&c->ev_p0)); CU_TRY(cudaEventCreate(&c->ev_p1));
    *out = c;
    return PRT_OK;
}

void prt_ctx_destroy(prt_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    c->samples.release(); c->counter.release(); c->d_pos.release(); c->d_nrm.release(); c->d_out.release();
    c->d_vis.release(); c->d_rays.release(); c->d_res.release(); c->need_bits.release(); c->need_count.release(); c->work_list.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->evh) cudaEventDestroy(c->evh);
    if (c->ev_tab) cudaEventDestroy(c->ev_tab);
    if (c->ev_p0) cudaEventDestroy(c->ev_p0);
    if (c->ev_p1) cudaEventDestroy(c->ev_p1);
    for (DevBuf &b : c->scratch) b.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int prt_ctx_device(const prt_ctx *c) { return c ? c->device : -1; }

int prt_ctx_last_kernel_ms(const prt_ctx *c, double *ms) {
    if (!c || !ms) return set_err(PRT_ERR_INVALID, "prt_ctx_last_kernel_ms: null argument");
    *ms = 0.0;
    if (!c->phase_timed) return PRT_OK;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaEventSynchronize(c->ev_p1));
    float f = 0.f;
    CU_TRY(cudaEventElapsedTime(&f, c->ev_p0, c->ev_p1));
    *ms = f;
    return PRT_OK;
}

int prt_ctx_set_tuning(prt_ctx *c, const char *name, int value) {
    if (!c || !name) return set_err(PRT_ERR_INVALID, "prt_ctx_set_tuning: null argument");
    std::string n(name);
    if (n == "block") { if (value < 32 || value > 256 || value % 32) return set_err(PRT_ERR_INVALID, "block must be a multiple of 32 in [32,256]"); c->block = value; }
    else if (n == "ctas_per_sm") { if (value < 0 || value > 32) return set_err(PRT_ERR_INVALID, "ctas_per_sm must be in [0,32]"); c->ctas_per_sm = value; }
    else if (n == "refill_thresh") { if (value < 0 || value > 32) return set_err(PRT_ERR_INVALID, "refill_thresh must be in [0,32]"); c->refill_thresh = value; }
    else if (n == "count_work") c->count_work = value ? 1 : 0;
    else if (n == "entry_list") c->entry_list = value ? 1 : 0;
    else if (n == "horizon") c->horizon = value ? 1 : 0;
    else if (n == "work_list") c->work_list_on = value < 0 ? -1 : value ? 1 : 0;
    else if (n == "l2_prefetch") c->l2_prefetch = value > 0 ? 1 : 0;
    else if (n == "horizon_near") { if (value < 5 || value > 157) return set_err(PRT_ERR_INVALID, "horizon_near (angular radius x100, rad) must be in [5,157]"); c->horizon_near = value; }
    else if (n == "horizon_mid") { if (value < 0 || value > 157) return set_err(PRT_ERR_INVALID, "horizon_mid (angular radius x100, rad; 0 = off) must be in [0,157]"); c->horizon_mid = value; }
    else if (n == "horizon_gain") { if (value < 0 || value > 1000000) return set_err(PRT_ERR_INVALID, "horizon_gain (tenths of a sample) must be in [0,1000000]"); c->horizon_gain = value; }
    else if (n == "horizon_slabs") c->horizon_slabs = value ? 1 : 0;
    else if (n == "wave_dop") c->wave_dop = value ? 1 : 0;
    else if (n == "horizon_budget") { if (value < 0 || value > 4096) return set_err(PRT_ERR_INVALID, "horizon_budget must be in [0,4096]"); c->horizon_budget = value; }
    else if (n == "pair_queue") { if (value != 0 && value != 2) return set_err(PRT_ERR_INVALID, "pair_queue must be 0 (per-ray stacks) or 2 (wavefront)"); c->pair_queue = value; }
    else return set_err(PRT_ERR_INVALID, "prt_ctx_set_tuning: unknown knob " + n);
    return PRT_OK;
}

int prt_scene_create(prt_ctx *c, const float *pos, size_t stride, uint32_t nv, const uint32_t *idx, uint32_t nt, prt_scene **out) {
    if (!c || !out) return set_err(PRT_ERR_INVALID, "prt_scene_create: null argument");
    *out = nullptr;
    HostBVH8 h;
    int rc = prt_build_host_bvh(pos, stride, nv, idx, nt, &h);
    if (rc) return rc;
    rc = prt_scene_from_host_bvh(c, &h, out);
    free_bvh8(&h);
    return rc;
}

void prt_scene_destroy(prt_scene *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->d_nodes) cudaFree(s->d_nodes);
    if (s->d_tris) cudaFree(s->d_tris);
    if (s->d_slabs) cudaFree(s->d_slabs);
    if (s->d_dops) cudaFree(s->d_dops);
    delete s;
}

int prt_scene_get_info(const prt_scene *s, prt_scene_info *out) {
    if (!s || !out) return set_err(PRT_ERR_INVALID, "prt_scene_get_info: null argument");
    *out = s->info;
    return PRT_OK;
}

int prt_trace_any_hit(prt_scene *s, const float *rays, uint32_t n, uint8_t *out_hit) {
    if (!s || (n && (!rays || !out_hit))) return set_err(PRT_ERR_INVALID, "prt_trace_any_hit: null argument");
    if (!n) return PRT_OK;
    prt_ctx *c = s->ctx;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(c->d_rays.reserve((size_t)n * 32));
    CU_TRY(c->d_res.reserve((size_t)n));
    CU_TRY(cudaMemcpyAsync(c->d_rays.p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(launch_trace_any(s->d_nodes, s->d_tris, (const float *)c->d_rays.p, n, (uint8_t *)c->d_res.p, c->stream));
    CU_TRY(cudaMemcpyAsync(out_hit, c->d_res.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return PRT_OK;
}

int prt_trace_closest_hit(prt_scene *s, const float *rays, uint32_t n, float *out_t, uint32_t *out_prim, float *out_ng) {
    if (!s || (n && (!rays || !out_t || !out_prim))) return set_err(PRT_ERR_INVALID, "prt_trace_closest_hit: null argument");
    if (!n) return PRT_OK;
    prt_ctx *c = s->ctx;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(c->d_rays.reserve((size_t)n * 32));
    CU_TRY(c->d_res.reserve((size_t)n * 20));
    float *dt = (float *)c->d_res.p; uint32_t *dp = (uint32_t *)(dt + n); float *dn = (float *)(dp + n);
    CU_TRY(cudaMemcpyAsync(c->d_rays.p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(launch_trace_closest(s->d_nodes, s->d_tris, (const float *)c->d_rays.p, n, dt, dp, dn, c->stream));
    CU_TRY(cudaMemcpyAsync(out_t, dt, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(out_prim, dp, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_ng) CU_TRY(cudaMemcpyAsync(out_ng, dn, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return PRT_OK;
}

void prt_bake_params_default(prt_bake_params *p) {
    if (!p) return;
    p->order = 3; p->samples_u = 32; p->samples_v = 32; p->seed = 0x50525400u; p->bounces = 0;
    p->albedo[0] = p->albedo[1] = p->albedo[2] = 1.0f;
    p->origin_eps = 1e-4f; p->bounce_eps = 1e-5f; p->mode = PRT_SHADOWED; p->cs_phase = 1; p->jitter = 1;
}

}  // extern "C"

namespace {

int check_params(const prt_bake_params *p) {
    if (!p) return set_err(PRT_ERR_INVALID, "bake: params is null");
    if (p->order < 1 || p->order > 5) return set_err(PRT_ERR_INVALID, "bake: order must be 1..5 bands");
    if (p->samples_u < 1 || p->samples_v < 1 || (int64_t)p->samples_u * p->samples_v > (1 << 24)) return set_err(PRT_ERR_INVALID, "bake: bad sample counts");
    if (p->mode < 0 || p->mode > 3) return set_err(PRT_ERR_INVALID, "bake: bad mode");
    if (p->bounces < 0 || p->bounces > 64) return set_err(PRT_ERR_INVALID, "bake: bounces must be 0..64");
    return PRT_OK;
}

// host sample table in reference order (s = i*Rv + j): raytracing.cpp:338-340 with the shared Philox jitter
void host_sample_table(const prt_bake_params *p, float *uv, float *dirs) {
    const int Ru = p->samples_u, Rv = p->samples_v;
    for (int i = 0; i < Ru; i++)
        for (int j = 0; j < Rv; j++) {
            const int s = i * Rv + j;
            float x1 = 0.5f, x2 = 0.5f;
            if (p->jitter) rand2(p->seed, (uint32_t)s, 0u, 0u, 0u, x1, x2);
            const float u = PRT_DIV(PRT_ADD((float)i, x1), (float)Ru);
            const float v = PRT_DIV(PRT_ADD((float)j, x2), (float)Rv);
            const f3 l = cosine_local(u, v);
            if (uv) { uv[2 * s] = u; uv[2 * s + 1] = v; }
            if (dirs) { dirs[3 * s] = l.x; dirs[3 * s + 1] = l.y; dirs[3 * s + 2] = l.z; }
        }
}

uint32_t morton2(uint32_t i, uint32_t j) {
    auto spread = [](uint32_t x) { x &= 0xFFFF; x = (x | (x << 8)) & 0x00FF00FF; x = (x | (x << 4)) & 0x0F0F0F0F; x = (x | (x << 2)) & 0x33333333; x = (x | (x << 1)) & 0x55555555; return x; };
    return spread(j) | (spread(i) << 1);
}

// Device sample table in *processing* order: strata sorted along a Morton curve over (i,j) so that any run of 32
// consecutive samples is a compact bundle of directions (coherent warps); w carries the reference index s.
int ensure_samples(prt_ctx *c, const prt_bake_params *p, cudaStream_t st) {
    if (c->samples.p && c->s_ru == p->samples_u && c->s_rv == p->samples_v && c->s_jit == p->jitter && c->s_seed == p->seed) return PRT_OK;
    const int S = p->samples_u * p->samples_v;
    std::vector<float> dirs(3 * (size_t)S);
    host_sample_table(p, nullptr, dirs.data());
    std::vector<std::pair<uint32_t, uint32_t>> key(S);
    for (int i = 0; i < p->samples_u; i++)
        for (int j = 0; j < p->samples_v; j++) key[i * p->samples_v + j] = { morton2((uint32_t)i, (uint32_t)j), (uint32_t)(i * p->samples_v + j) };
    std::sort(key.begin(), key.end());
    std::vector<float> tab(4 * (size_t)S);
    for (int k = 0; k < S; k++) {
        const uint32_t s = key[k].second;
        tab[4 * k] = dirs[3 * s]; tab[4 * k + 1] = dirs[3 * s + 1]; tab[4 * k + 2] = dirs[3 * s + 2];
        // w = reference sample index (24 bits) | azimuth bin of the local direction (5-6 bits, horizon map of horizon.cu)
        // diamond pseudo-angle in [0,4), same formula as hz_pang (entry_list.cuh); kHzBins / 4 bins per unit
        const double lx = dirs[3 * s], ly = dirs[3 * s + 1], den = std::fabs(lx) + std::fabs(ly);
        double pa = den > 0.0 ? ly / den : 0.0;
        pa = lx < 0.0 ? 2.0 - pa : (ly < 0.0 ? 4.0 + pa : pa);
        int bin = (int)std::floor(pa * (double)(prt::kHzBins / 4));
        bin = std::min(prt::kHzBins - 1, std::max(0, bin));
        const uint32_t w = s | ((uint32_t)bin << 24);
        std::memcpy(&tab[4 * k + 3], &w, 4);
    }
    // the previous bake (possibly on another stream) may still be reading the old table: wait for it, then upload in stream order
    // from a host copy that stays alive in the context
    if (c->tab_in_use) CU_TRY(cudaEventSynchronize(c->ev_tab));
    CU_TRY(c->samples.reserve(sizeof(float) * 4 * (size_t)S));
    c->h_samples.swap(tab);
    CU_TRY(cudaMemcpyAsync(c->samples.p, c->h_samples.data(), sizeof(float) * 4 * (size_t)S, cudaMemcpyHostToDevice, st));
    c->s_ru = p->samples_u; c->s_rv = p->samples_v; c->s_jit = p->jitter; c->s_seed = p->seed;
    return PRT_OK;
}

}  // namespace

// The bake on device-resident vertices.  e0 / e1 (optional) bracket the kernels for prt_ctx_last_bake_stats; `place` (optional) says
// where rows go (row stride, shard of a multi-GPU bake, peer buffers of the fused gather): abi_internal.h.
int prt_bake_device(prt_ctx *c, prt_scene *sc, const float *d_pos, const float *d_nrm, size_t stride, uint32_t n, uint32_t vid_base,
                    const prt_bake_params *p, float *d_out, uint32_t *d_vis, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1,
                    const prt_row_placement *place) {
    int rc = check_params(p);
    if (rc) return rc;
    const bool needs_scene = p->mode == PRT_SHADOWED || p->mode == PRT_INTERREFLECT;
    if (needs_scene && !sc) return set_err(PRT_ERR_INVALID, "bake: this mode needs a scene");
    if (sc && sc->ctx != c) return set_err(PRT_ERR_INVALID, "bake: scene belongs to another context");
    if (stride == 0) stride = 12;
    if (stride % 4) return set_err(PRT_ERR_INVALID, "bake: stride must be a multiple of 4");
    c->stats = prt_bake_stats{};
    if (n == 0) return PRT_OK;
    if (!d_pos || !d_nrm || !d_out) return set_err(PRT_ERR_INVALID, "bake: null buffer");
    CU_TRY(cudaSetDevice(c->device));
    rc = ensure_samples(c, p, st);
    if (rc) return rc;
    const int S = p->samples_u * p->samples_v;
    BakeArgs A{};
    A.nodes = sc ? sc->d_nodes : nullptr; A.tris = sc ? sc->d_tris : nullptr;
    A.pos = d_pos; A.nrm = d_nrm; A.stride = stride; A.n_verts = n; A.vid_base = vid_base;
    A.samples = (const float4 *)c->samples.p; A.S = S; A.inv_S = 1.0f / (float)S;
    A.out = d_out; A.vis = d_vis; A.vis_words = (S + 31) / 32;
    CU_TRY(c->counter.reserve(256));
    A.counter = (uint32_t *)c->counter.p;
    A.work = c->count_work ? (unsigned long long *)((char *)c->counter.p + 64) : nullptr;
    A.seed = p->seed; A.depth = p->bounces + 1;
    A.albedo[0] = p->albedo[0]; A.albedo[1] = p->albedo[1]; A.albedo[2] = p->albedo[2];
    A.origin_eps = p->origin_eps; A.bounce_eps = p->bounce_eps; A.cs_phase = p->cs_phase;
    A.refill_thresh = c->refill_thresh;
    if (place) {
        const uint32_t n2 = (uint32_t)(p->order * p->order);
        if (place->out_stride_floats && place->out_stride_floats < n2) return set_err(PRT_ERR_INVALID, "bake: output row stride is shorter than a row");
        if (place->n_peer < 0 || place->n_peer > 7) return set_err(PRT_ERR_INVALID, "bake: at most 7 peer buffers");
        A.out_stride = place->out_stride_floats; A.shard_world = place->shard_world; A.shard_rank = place->shard_rank;
        A.out_global = place->out_global; A.n_peer = place->n_peer;
        for (int i = 0; i < place->n_peer; i++) A.out_peer[i] = place->out_peer[i];
    }
    A.entry_list = c->entry_list;
    A.horizon = c->horizon;
    A.horizon_budget = c->horizon_budget;
    { const float sn = sinf(0.01f * (float)c->horizon_near); A.horizon_near2 = 1.0f / (sn * sn); }
    if (c->horizon_mid > 0) { const float sn = sinf(0.01f * (float)c->horizon_mid); A.horizon_mid2 = 1.0f / (sn * sn); }
    A.horizon_gain = 0.1f * (float)c->horizon_gain * (float)kHzBins / (float)S;         // samples -> units of S / kHzBins (hz_gain)
    A.slabs = (sc && c->horizon_slabs) ? sc->d_slabs : nullptr;
    A.dops = (sc && c->wave_dop) ? sc->d_dops : nullptr;
    CU_TRY(cudaMemsetAsync(A.counter, 0, 128, st));
    if (d_vis) CU_TRY(cudaMemsetAsync(d_vis, 0, (size_t)n * A.vis_words * 4, st));
    uint32_t prefetched = 0;
    int mode = p->mode == PRT_SHADOWED ? 0 : p->mode == PRT_INTERREFLECT ? 1 : p->mode == PRT_UNSHADOWED ? 2 : 3;
    if (mode == 1 && p->bounces == 0) mode = 0;
    const int grid = c->ctas_per_sm > 0 ? c->n_sms * c->ctas_per_sm : 0;
    if (e0) CU_TRY(cudaEventRecord(e0, st));
    // optional (knob l2_prefetch, default off): stream the BVH into L2 before the first pass.  Measured on the 68 k-vertex shard of an
    // 8-GPU bake with a flushed L2: 6.647 ms with and without -- the demand misses of a cold start are already hidden by the other warps
    if (sc && needs_scene) {
        const bool on = c->l2_prefetch > 0;
        if (on) { CU_TRY(launch_l2_prefetch(sc->d_nodes, sc->info.node_bytes, sc->d_tris, sc->info.tri_bytes, c->n_sms, st)); prefetched = 1; }
    }
    int used_grid = grid;
    const bool fast_ok = (mode == 0 || mode == 2) && c->entry_list && S <= bake_wave_max_samples();
    uint32_t launches = 1;
    if (fast_ok && c->pair_queue == 2) {
        A.need_bits = nullptr; A.need_count = nullptr;
        if (mode == 0 && c->horizon) {
            // pass 1: horizon map + classification; finishes the vertices whose samples are all provably visible
            CU_TRY(c->need_bits.reserve((size_t)n * A.vis_words * 4));
            CU_TRY(c->need_count.reserve((size_t)n * 4));
            A.need_bits = (uint32_t *)c->need_bits.p; A.need_count = (uint32_t *)c->need_count.p;
            // heaviest-first work list (a counting sort of the need counts, horizon.cu): shortens the tail of the persistent grid.  Measured
            // on the bench mesh, list on vs off: 543 k vertices 0.0 %, 272 k -0.1 %, 136 k -0.7 %, 68 k (the shard of an 8-GPU bake) -3.2 %;
            // "auto" sorts below 256 vertices per resident warp (7 CTAs x 4 warps per SM), where the tail is a visible share of the launch
            if (c->work_list_on > 0 || (c->work_list_on < 0 && (uint64_t)n < 256ull * 28ull * (uint64_t)c->n_sms)) {
                CU_TRY(c->work_list.reserve(((size_t)n + 512) * 4));
                A.work_list = (uint32_t *)c->work_list.p;
                CU_TRY(cudaMemsetAsync(A.work_list + n, 0, 512 * 4, st));        // histogram + offsets of the counting sort
            }
            int hgrid = 0;
            CU_TRY(launch_horizon(A, p->order, &hgrid, c->n_sms, st));
            if (e0) CU_TRY(cudaEventRecord(c->evh, st));
            launches = A.work_list ? 5 : 2;        // horizon_kernel (+ the three kernels of the work-list sort) + bake_wave_kernel
        }
        CU_TRY(launch_bake_wave(A, p->order, mode == 0, &used_grid, bake_wave_block(), c->n_sms, st));
        c->stats.block = (uint32_t)bake_wave_block();
    }
    else {
        A.need_bits = nullptr; A.need_count = nullptr;
        if (mode == 1 && c->horizon && c->entry_list) {
            // interreflection: the horizon pass settles the primary rays that provably escape (and whole vertices)
            CU_TRY(c->need_bits.reserve((size_t)n * A.vis_words * 4));
            CU_TRY(c->need_count.reserve((size_t)n * 4));
            A.need_bits = (uint32_t *)c->need_bits.p; A.need_count = (uint32_t *)c->need_count.p;
            int hgrid = 0;
            CU_TRY(launch_horizon(A, p->order, &hgrid, c->n_sms, st));
            if (e0) CU_TRY(cudaEventRecord(c->evh, st));
            launches = 2;
        }
        if (A.need_bits && c->pair_queue == 2 && S <= bake_inter_max_samples()) {
            CU_TRY(launch_bake_inter(A, p->order, &used_grid, c->n_sms, st));
            c->stats.block = 128;
        } else
            CU_TRY(launch_bake(A, p->order, mode, &used_grid, c->block, c->n_sms, st));
    }
    if (e1) CU_TRY(cudaEventRecord(e1, st));
    CU_TRY(cudaEventRecord(c->ev_tab, st)); c->tab_in_use = true;
    c->stats.rays = (mode == 0 || mode == 1) ? (uint64_t)n * (uint64_t)S : 0;
    c->stats.launches = launches + prefetched; c->stats.grid = (uint32_t)used_grid; if (!c->stats.block) c->stats.block = (uint32_t)c->block;
    c->stats_pending = e0 && e1;
    c->work_pending = c->count_work != 0;
    return PRT_OK;
}

extern "C" {

int prt_bake_transfer_device(prt_ctx *c, prt_scene *sc, const float *d_pos, const float *d_nrm, size_t stride, uint32_t n,
                             uint32_t vid_base, const prt_bake_params *p, float *d_out, uint32_t *d_vis, void *stream) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device: ctx is null");
    return prt_bake_device(c, sc, d_pos, d_nrm, stride, n, vid_base, p, d_out, d_vis, (cudaStream_t)stream, c->ev1, c->ev2, nullptr);
}

int prt_bake_transfer_device_shard(prt_ctx *c, prt_scene *sc, const float *d_pos, const float *d_nrm, size_t stride, uint32_t n,
                                   uint32_t shard_world, uint32_t shard_rank, const prt_bake_params *p, float *d_out, uint32_t *d_vis, void *stream) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_shard: ctx is null");
    if (shard_world < 1 || shard_rank >= shard_world) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_shard: bad world / rank");
    prt_row_placement pl{};
    pl.shard_world = shard_world; pl.shard_rank = shard_rank;
    return prt_bake_device(c, sc, d_pos, d_nrm, stride, n, 0u, p, d_out, d_vis, (cudaStream_t)stream, c->ev1, c->ev2, &pl);
}

int prt_bake_transfer_device_shard_fused(prt_ctx *c, prt_scene *sc, const float *d_pos, const float *d_nrm, size_t stride, uint32_t n,
                                         uint32_t shard_world, uint32_t shard_rank, const prt_bake_params *p, float *d_rows_full,
                                         float *const *peer_rows_full, int32_t n_peers, void *stream) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_shard_fused: ctx is null");
    if (shard_world < 1 || shard_rank >= shard_world) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_shard_fused: bad world / rank");
    if (n_peers < 0 || n_peers > 7 || (n_peers && !peer_rows_full)) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_shard_fused: 0..7 peer buffers");
    prt_row_placement pl{};
    pl.shard_world = shard_world; pl.shard_rank = shard_rank; pl.out_global = 1; pl.n_peer = n_peers;
    for (int i = 0; i < n_peers; i++) {
        if (!peer_rows_full[i]) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_shard_fused: null peer buffer");
        pl.out_peer[i] = peer_rows_full[i];
    }
    return prt_bake_device(c, sc, d_pos, d_nrm, stride, n, 0u, p, d_rows_full, nullptr, (cudaStream_t)stream, c->ev1, c->ev2, &pl);
}

int prt_device_alloc(prt_ctx *c, size_t bytes, void **out) {
    if (!c || !out || !bytes) return set_err(PRT_ERR_INVALID, "prt_device_alloc: bad argument");
    *out = nullptr;
    CU_TRY(cudaSetDevice(c->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(PRT_ERR_NOMEM, std::string("prt_device_alloc: ") + cudaGetErrorString(e)); }
    CU_TRY(cudaMemset(p, 0, bytes));
    *out = p;
    return PRT_OK;
}
int prt_device_free(prt_ctx *c, void *p) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_device_free: ctx is null");
    if (!p) return PRT_OK;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaFree(p));
    return PRT_OK;
}
int prt_ipc_export(prt_ctx *c, const void *d_ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!c || !d_ptr || !handle) return set_err(PRT_ERR_INVALID, "prt_ipc_export: null argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    std::memcpy(handle, &h, 64);
    return PRT_OK;
}
int prt_ipc_open(prt_ctx *c, const uint8_t handle[64], void **out) {
    if (!c || !handle || !out) return set_err(PRT_ERR_INVALID, "prt_ipc_open: null argument");
    *out = nullptr;
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return PRT_OK;
}
int prt_ipc_close(prt_ctx *c, void *d_ptr) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_ipc_close: ctx is null");
    if (!d_ptr) return PRT_OK;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaIpcCloseMemHandle(d_ptr));
    return PRT_OK;
}

int prt_bake_transfer_device_strided(prt_ctx *c, prt_scene *sc, const float *d_pos, const float *d_nrm, size_t stride, uint32_t n,
                                     uint32_t vid_base, const prt_bake_params *p, float *d_out, size_t out_stride_bytes, uint32_t *d_vis, void *stream) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_strided: ctx is null");
    if (out_stride_bytes % 4) return set_err(PRT_ERR_INVALID, "prt_bake_transfer_device_strided: out_stride_bytes must be a multiple of 4");
    prt_row_placement pl{};
    pl.out_stride_floats = (uint32_t)(out_stride_bytes / 4);
    return prt_bake_device(c, sc, d_pos, d_nrm, stride, n, vid_base, p, d_out, d_vis, (cudaStream_t)stream, c->ev1, c->ev2, &pl);
}

int prt_bake_transfer(prt_ctx *c, prt_scene *sc, const float *pos, const float *nrm, size_t stride, uint32_t n, uint32_t vid_base,
                      const prt_bake_params *p, float *out, uint32_t *out_vis) {
    if (!c) return set_err(PRT_ERR_INVALID, "prt_bake_transfer: ctx is null");
    int rc = check_params(p);
    if (rc) return rc;
    if (n == 0) return PRT_OK;
    if (!pos || !nrm || !out) return set_err(PRT_ERR_INVALID, "prt_bake_transfer: null buffer");
    if (stride == 0) stride = 12;
    if (stride % 4) return set_err(PRT_ERR_INVALID, "prt_bake_transfer: stride must be a multiple of 4");
    CU_TRY(cudaSetDevice(c->device));
    const int n2 = p->order * p->order, S = p->samples_u * p->samples_v, words = (S + 31) / 32;
    // the vertices keep the caller's layout on the device.  Interleaved Mesh::Vert arrays (nrm inside the stride of pos, gl.h:76-80)
    // are uploaded ONCE; separate position / normal arrays are two uploads
    const size_t span = (size_t)(n - 1) * stride + 12;
    const ptrdiff_t delta = reinterpret_cast<const char *>(nrm) - reinterpret_cast<const char *>(pos);
    const bool interleaved = delta >= 12 && (size_t)delta + 12 <= stride;
    CU_TRY(c->d_pos.reserve(interleaved ? span + (size_t)delta : span));
    if (!interleaved) CU_TRY(c->d_nrm.reserve(span));
    CU_TRY(c->d_out.reserve((size_t)n * n2 * 4));
    if (out_vis) CU_TRY(c->d_vis.reserve((size_t)n * words * 4));
    cudaStream_t st = c->stream;
    CU_TRY(cudaEventRecord(c->ev0, st));
    CU_TRY(cudaMemcpyAsync(c->d_pos.p, pos, interleaved ? span + (size_t)delta : span, cudaMemcpyHostToDevice, st));
    if (!interleaved) CU_TRY(cudaMemcpyAsync(c->d_nrm.p, nrm, span, cudaMemcpyHostToDevice, st));
    const float *dp = (const float *)c->d_pos.p;
    const float *dn = interleaved ? reinterpret_cast<const float *>(reinterpret_cast<const char *>(c->d_pos.p) + delta) : (const float *)c->d_nrm.p;
    rc = prt_bake_device(c, sc, dp, dn, stride, n, vid_base, p, (float *)c->d_out.p, out_vis ? (uint32_t *)c->d_vis.p : nullptr, st, c->ev1, c->ev2, nullptr);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(out, c->d_out.p, (size_t)n * n2 * 4, cudaMemcpyDeviceToHost, st));
    if (out_vis) CU_TRY(cudaMemcpyAsync(out_vis, c->d_vis.p, (size_t)n * words * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(c->ev3, st));
    CU_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); c->stats.h2d_ms = ms;
    CU_TRY(cudaEventElapsedTime(&ms, c->ev1, c->ev2)); c->stats.kernel_ms = ms;
    if (c->stats.launches >= 2) { CU_TRY(cudaEventElapsedTime(&ms, c->ev1, c->evh)); c->stats.horizon_ms = ms; }
    CU_TRY(cudaEventElapsedTime(&ms, c->ev2, c->ev3)); c->stats.d2h_ms = ms;
    c->stats.h2d_bytes = interleaved ? (uint64_t)span + (uint64_t)delta : 2 * (uint64_t)span;
    c->stats.d2h_bytes = (uint64_t)n * n2 * 4 + (out_vis ? (uint64_t)n * words * 4 : 0);
    c->stats_pending = false;
    return PRT_OK;
}

int prt_ctx_last_bake_stats(const prt_ctx *cc, prt_bake_stats *out) {
    if (!cc || !out) return set_err(PRT_ERR_INVALID, "prt_ctx_last_bake_stats: null argument");
    prt_ctx *c = const_cast<prt_ctx *>(cc);
    if (c->stats_pending) {
        CU_TRY(cudaSetDevice(c->device));
        CU_TRY(cudaEventSynchronize(c->ev2));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, c->ev1, c->ev2));
        c->stats.kernel_ms = ms;
        if (c->stats.launches >= 2) { CU_TRY(cudaEventElapsedTime(&ms, c->ev1, c->evh)); c->stats.horizon_ms = ms; }
        c->stats_pending = false;
    }
    if (c->work_pending) {
        CU_TRY(cudaSetDevice(c->device));
        CU_TRY(cudaEventSynchronize(c->ev2));
        unsigned long long w[4] = {0, 0, 0, 0};
        CU_TRY(cudaMemcpy(w, (char *)c->counter.p + 64, 32, cudaMemcpyDeviceToHost));
        c->stats.node_visits = w[0]; c->stats.tri_tests = w[1]; c->stats.cand_tests = w[2]; c->stats.rays_traversed = w[3];
        c->work_pending = false;
    }
    *out = c->stats;
    return PRT_OK;
}

int prt_scatter_sh9(const float *coeffs, int32_t order, uint32_t n, void *verts, size_t vstride, size_t sh_off) {
    if (!coeffs || !verts || order < 3) return set_err(PRT_ERR_INVALID, "prt_scatter_sh9: needs order >= 3 rows and non-null buffers");
    const int n2 = order * order;
    for (uint32_t i = 0; i < n; i++)
        std::memcpy((char *)verts + (size_t)i * vstride + sh_off, coeffs + (size_t)i * n2, 9 * sizeof(float));
    return PRT_OK;
}

int prt_bake_sample_table(const prt_bake_params *p, float *uv, float *dirs) {
    int rc = check_params(p);
    if (rc) return rc;
    host_sample_table(p, uv, dirs);
    return PRT_OK;
}

}  // extern "C"
