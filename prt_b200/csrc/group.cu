// group.cu -- single-process multi-GPU driver behind the C ABI (include/prt_b200.h, prt_group_*).
//
// The reference has no multi-GPU path; its per-vertex loop `std::for_each(par, verts...)`
// (reference src/raytracing/raytracing.cpp:328) is embarrassingly parallel over vertices given a replicated scene, which is what is
// sharded here: one context per GPU, the BVH built ONCE on the host and uploaded to every GPU, the vertex list dealt to the GPUs in
// interleaved 64-vertex chunks (occlusion cost varies over a surface; round-robin chunks of the caller's -- ideally Morton --
// order give every GPU a statistically identical slice).  The coefficient rows come together in one of three ways:
//   PRT_GATHER_P2P   fused compute + collective: the projection epilogue of the bake kernels stores every finished row straight
//                    into the full-size row buffer of every peer GPU through P2P-mapped pointers (NVLink / NVSwitch), so the
//                    transfer overlaps the traversal row by row and NO collective follows the kernel
//   PRT_GATHER_NCCL  the checked baseline: rows packed per rank, one in-place ncclAllGather inside ncclGroupStart/End
//                    (ncclCommInitAll communicator), then a permutation kernel back to vertex order
//   PRT_GATHER_NONE  rows stay on the GPU that baked them
// In every mode each GPU copies ITS OWN rows to the caller's host buffer over its own PCIe link (strided 2-D copies straight into
// vertex order), so the host result never funnels through one GPU.
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a process that already loaded
// NCCL (torch) shares that copy, and a machine without it still gets the P2P and NONE modes.
#include "../../include/prt_b200.h"
#include "abi_internal.h"
#include "kernels.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <thread>
#include <vector>

using namespace prt;

namespace {

// minimal NCCL declarations (stable since NCCL 2.0; nccl.h is not needed to build)
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7;
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool load(std::string &why) {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) { why = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define PRT_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym)); if (!field) { why = std::string("NCCL symbol missing: ") + sym; return false; }
        PRT_SYM(CommInitAll, "ncclCommInitAll") PRT_SYM(CommDestroy, "ncclCommDestroy") PRT_SYM(AllGather, "ncclAllGather")
        PRT_SYM(GroupStart, "ncclGroupStart") PRT_SYM(GroupEnd, "ncclGroupEnd") PRT_SYM(CommGetAsyncError, "ncclCommGetAsyncError")
        PRT_SYM(GetErrorString, "ncclGetErrorString") PRT_SYM(GetVersion, "ncclGetVersion")
#undef PRT_SYM
        return true;
    }
};

struct Buf {
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

struct Member {
    prt_ctx *ctx = nullptr;
    int device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    Buf in_pos, in_nrm, rows, packed;       // rows: result rows (local or full-size), packed: [g][per_rank][n2] for the NCCL baseline
};

// dst[global_row] = src[rank][local] : packed all-gather layout -> vertex order
__global__ void unshard_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, const uint32_t per_rank, const uint32_t world, const uint32_t n2,
                                    const unsigned long long total) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t k = (uint32_t)(i % n2);
    const unsigned long long lv = i / n2;
    const uint32_t rank = (uint32_t)(lv / per_rank), v = (uint32_t)(lv % per_rank);
    const unsigned long long g = ((unsigned long long)(v / kShardChunk) * world + rank) * kShardChunk + (v % kShardChunk);
    dst[g * n2 + k] = src[i];
}

// merged CSR: surfel ids of one member's segment -> global ids, probe ranges shifted by the segment's offset
__global__ void remap_ids_kernel(uint32_t *ids, const uint32_t *__restrict__ remap, const unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = remap[ids[i]];
}
__global__ void shift_range_kernel(uint32_t *range2, const uint32_t n2, const uint32_t base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) range2[i] += base;
}

}  // namespace

struct prt_group {
    int n = 0;
    std::vector<Member> m;
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    bool nccl_ok = false, p2p_ok = false;
    std::string nccl_why, p2p_why;
    // layout of the last bake (for prt_group_download_rows)
    uint32_t last_n = 0, last_n2 = 0, last_v_pad = 0; int last_mode = PRT_GATHER_NONE;
};

struct prt_group_scene {
    prt_group *g = nullptr;
    std::vector<prt_scene *> sc;
    prt_scene_info info{};
    double upload_seconds_max = 0.0;
};

#define GR_TRY(expr)                                                                                                   \
    do {                                                                                                               \
        cudaError_t e_ = (expr);                                                                                       \
        if (e_ != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int prt_group_create(const int *device_ids, int n_devices, prt_group **out) {
    if (!out) return prt_set_error(PRT_ERR_INVALID, "prt_group_create: out is null");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 8) return prt_set_error(PRT_ERR_INVALID, "prt_group_create: 1..8 devices");
    prt_group *g = new prt_group();
    g->n = n_devices;
    g->m.resize(n_devices);
    bool distinct = true;
    for (int i = 0; i < n_devices; i++) {
        for (int j = 0; j < i; j++) distinct = distinct && device_ids[i] != device_ids[j];
        int rc = prt_ctx_create(device_ids[i], &g->m[i].ctx);
        if (rc) { prt_group_destroy(g); return rc; }
        Member &M = g->m[i];
        M.device = prt_ctx_device(M.ctx);
        cudaSetDevice(M.device);
        cudaError_t e = cudaStreamCreateWithFlags(&M.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&M.e0);
        if (e == cudaSuccess) e = cudaEventCreate(&M.e1);
        if (e == cudaSuccess) e = cudaEventCreate(&M.e2);
        if (e == cudaSuccess) e = cudaEventCreate(&M.e3);
        if (e != cudaSuccess) { prt_group_destroy(g); return prt_set_error(PRT_ERR_CUDA, std::string("prt_group_create: ") + cudaGetErrorString(e)); }
    }
    // peer access between every pair (NVSwitch: every GPU reaches every peer at full NVLink rate)
    g->p2p_ok = true;
    for (int i = 0; i < n_devices && g->p2p_ok; i++)
        for (int j = 0; j < n_devices; j++) {
            if (g->m[i].device == g->m[j].device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->m[i].device, g->m[j].device);
            if (!can) { g->p2p_ok = false; g->p2p_why = "no peer access between devices " + std::to_string(g->m[i].device) + " and " + std::to_string(g->m[j].device); break; }
            cudaSetDevice(g->m[i].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(g->m[j].device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { g->p2p_ok = false; g->p2p_why = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); cudaGetLastError(); break; }
        }
    // NCCL communicator over the devices (needs distinct devices)
    if (!distinct) g->nccl_why = "a device is listed twice (NCCL needs distinct devices)";
    else if (g->nccl.load(g->nccl_why)) {
        g->comms.assign(n_devices, nullptr);
        std::vector<int> devs(n_devices);
        for (int i = 0; i < n_devices; i++) devs[i] = g->m[i].device;
        const ncclResult_t r = g->nccl.CommInitAll(g->comms.data(), n_devices, devs.data());
        if (r == 0) g->nccl_ok = true;
        else { g->nccl_why = std::string("ncclCommInitAll: ") + g->nccl.GetErrorString(r); g->comms.clear(); }
    }
    *out = g;
    return PRT_OK;
}

void prt_group_destroy(prt_group *g) {
    if (!g) return;
    if (g->nccl_ok) for (ncclComm_t c : g->comms) if (c) g->nccl.CommDestroy(c);
    for (Member &M : g->m) {
        if (!M.ctx) continue;
        cudaSetDevice(M.device);
        M.in_pos.release(); M.in_nrm.release(); M.rows.release(); M.packed.release();
        if (M.e0) cudaEventDestroy(M.e0);
        if (M.e1) cudaEventDestroy(M.e1);
        if (M.e2) cudaEventDestroy(M.e2);
        if (M.e3) cudaEventDestroy(M.e3);
        if (M.st) cudaStreamDestroy(M.st);
        prt_ctx_destroy(M.ctx);
    }
    delete g;
}

int prt_group_size(const prt_group *g) { return g ? g->n : 0; }
prt_ctx *prt_group_ctx(prt_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->m[i].ctx : nullptr; }

int prt_group_capabilities(const prt_group *g, int *p2p, int *nccl, int *nccl_version, char *why, size_t why_len) {
    if (!g) return prt_set_error(PRT_ERR_INVALID, "prt_group_capabilities: group is null");
    if (p2p) *p2p = g->p2p_ok ? 1 : 0;
    if (nccl) *nccl = g->nccl_ok ? 1 : 0;
    if (nccl_version) { *nccl_version = 0; if (g->nccl_ok) g->nccl.GetVersion(nccl_version); }
    if (why && why_len) {
        const std::string w = (g->p2p_ok ? std::string() : "p2p: " + g->p2p_why + "; ") + (g->nccl_ok ? std::string() : "nccl: " + g->nccl_why);
        std::strncpy(why, w.c_str(), why_len - 1); why[why_len - 1] = 0;
    }
    return PRT_OK;
}

int prt_group_set_tuning(prt_group *g, const char *name, int value) {
    if (!g) return prt_set_error(PRT_ERR_INVALID, "prt_group_set_tuning: group is null");
    for (Member &M : g->m) { const int rc = prt_ctx_set_tuning(M.ctx, name, value); if (rc) return rc; }
    return PRT_OK;
}

int prt_group_scene_create(prt_group *g, const float *pos, size_t stride, uint32_t nv, const uint32_t *idx, uint32_t nt, prt_group_scene **out) {
    if (!g || !out) return prt_set_error(PRT_ERR_INVALID, "prt_group_scene_create: null argument");
    *out = nullptr;
    HostBVH8 h;
    int rc = prt_build_host_bvh(pos, stride, nv, idx, nt, &h);          // ONE build ...
    if (rc) return rc;
    prt_group_scene *gs = new prt_group_scene();
    gs->g = g; gs->sc.assign(g->n, nullptr);
    std::vector<int> rcs(g->n, 0);
    std::vector<std::string> errs(g->n);
    std::vector<std::thread> th;
    for (int i = 0; i < g->n; i++)                                       // ... one upload per GPU, in parallel over the PCIe links
        th.emplace_back([&, i]() { rcs[i] = prt_scene_from_host_bvh(g->m[i].ctx, &h, &gs->sc[i]); if (rcs[i]) errs[i] = prt_last_error(); });
    for (auto &t : th) t.join();
    free_bvh8(&h);
    for (int i = 0; i < g->n; i++)
        if (rcs[i]) { const int r = rcs[i]; const std::string e = errs[i]; prt_group_scene_destroy(gs); return prt_set_error(r, e); }
    prt_scene_get_info(gs->sc[0], &gs->info);
    for (int i = 0; i < g->n; i++) { prt_scene_info si; prt_scene_get_info(gs->sc[i], &si); gs->upload_seconds_max = std::max(gs->upload_seconds_max, si.upload_seconds); }
    gs->info.upload_seconds = gs->upload_seconds_max;
    *out = gs;
    return PRT_OK;
}

void prt_group_scene_destroy(prt_group_scene *gs) {
    if (!gs) return;
    for (prt_scene *s : gs->sc) prt_scene_destroy(s);
    delete gs;
}

int prt_group_scene_get_info(const prt_group_scene *gs, prt_scene_info *out) {
    if (!gs || !out) return prt_set_error(PRT_ERR_INVALID, "prt_group_scene_get_info: null argument");
    *out = gs->info;
    return PRT_OK;
}
prt_scene *prt_group_scene_member(prt_group_scene *gs, int i) { return (gs && i >= 0 && i < (int)gs->sc.size()) ? gs->sc[i] : nullptr; }

int prt_group_bake_transfer(prt_group *g, prt_group_scene *gs, const float *pos, const float *nrm, size_t stride, uint32_t n,
                            const prt_bake_params *p, float *out, int gather, prt_group_stats *stats) {
    if (!g || !p) return prt_set_error(PRT_ERR_INVALID, "prt_group_bake_transfer: null argument");
    if (gs && gs->g != g) return prt_set_error(PRT_ERR_INVALID, "prt_group_bake_transfer: scene belongs to another group");
    if (p->order < 1 || p->order > 5) return prt_set_error(PRT_ERR_INVALID, "bake: order must be 1..5 bands");
    if (stride == 0) stride = 12;
    if (stride % 4) return prt_set_error(PRT_ERR_INVALID, "prt_group_bake_transfer: stride must be a multiple of 4");
    if (gather == PRT_GATHER_AUTO) gather = g->n == 1 ? PRT_GATHER_NONE : g->p2p_ok ? PRT_GATHER_P2P : g->nccl_ok ? PRT_GATHER_NCCL : PRT_GATHER_NONE;
    if (gather == PRT_GATHER_P2P && !g->p2p_ok) return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_group_bake_transfer: P2P gather unavailable: " + g->p2p_why);
    if (gather == PRT_GATHER_NCCL && !g->nccl_ok) return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_group_bake_transfer: NCCL gather unavailable: " + g->nccl_why);
    if (gather < PRT_GATHER_NONE || gather > PRT_GATHER_P2P) return prt_set_error(PRT_ERR_INVALID, "prt_group_bake_transfer: bad gather mode");
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (n == 0) return PRT_OK;
    if (!pos || !nrm) return prt_set_error(PRT_ERR_INVALID, "prt_group_bake_transfer: null buffer");

    const auto t_start = std::chrono::steady_clock::now();
    const uint32_t W = (uint32_t)g->n, n2 = (uint32_t)(p->order * p->order);
    const uint32_t n_chunks = (n + kShardChunk - 1) / kShardChunk;
    const uint32_t chunks_pad = ((n_chunks + W - 1) / W) * W;
    const uint32_t chunks_per_rank = chunks_pad / W, per_rank = chunks_per_rank * kShardChunk;
    const uint64_t v_pad = (uint64_t)chunks_pad * kShardChunk;
    const ptrdiff_t delta = reinterpret_cast<const char *>(nrm) - reinterpret_cast<const char *>(pos);
    const bool interleaved = delta >= 12 && (size_t)delta + 12 <= stride;
    const size_t chunk_bytes = (size_t)kShardChunk * stride, row_bytes = (size_t)n2 * 4;

    // valid vertices of rank r: its chunks r, r+W, ... below n_chunks; the last global chunk may be partial
    auto rank_count = [&](uint32_t r) -> uint32_t {
        if (r >= n_chunks) return 0u;
        const uint32_t owned = (n_chunks - 1 - r) / W + 1;                       // chunks with index < n_chunks
        const uint32_t last = r + (owned - 1) * W;                               // its last chunk
        const uint32_t tail = last == n_chunks - 1 ? n - last * kShardChunk : kShardChunk;
        return (owned - 1) * kShardChunk + tail;
    };

    std::vector<int> rcs(W, 0);
    std::vector<std::string> errs(W);
    auto fail = [&](uint32_t r, int code, const std::string &msg) { rcs[r] = code; errs[r] = msg; };
#define TH_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { fail(r, PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); return; } } while (0)

    // ---- phase A (one host thread per GPU): buffers, upload of the shard, bake --------------------------------------------------
    auto phase_a = [&](uint32_t r) {
        Member &M = g->m[r];
        TH_TRY(cudaSetDevice(M.device));
        const uint32_t cnt = rank_count(r);
        const size_t in_bytes = (size_t)per_rank * stride + (interleaved ? (size_t)delta : 0) + 16;
        TH_TRY(M.in_pos.reserve(in_bytes));
        if (!interleaved) TH_TRY(M.in_nrm.reserve(in_bytes));
        TH_TRY(M.rows.reserve((gather == PRT_GATHER_NONE ? (size_t)per_rank : (size_t)v_pad) * row_bytes));
        if (gather == PRT_GATHER_NCCL) TH_TRY(M.packed.reserve((size_t)W * per_rank * row_bytes));
        TH_TRY(cudaEventRecord(M.e0, M.st));
        if (cnt) {
            // chunks r, r+W, ... of the caller's arrays -> contiguous shard on the device: one 2-D copy (a row per chunk) for all
            // chunks but the array's very last one, which is copied to the end of its last vertex only (the caller's array is not
            // known to extend a full stride beyond it).  An interleaved vertex array (Mesh::Vert) is copied once.
            const uint32_t owned = (cnt + kShardChunk - 1) / kShardChunk;
            const bool owns_last = r + (owned - 1) * W == n_chunks - 1;
            const uint32_t rows2d = owns_last ? owned - 1 : owned;
            const char *hp = reinterpret_cast<const char *>(pos) + (size_t)r * chunk_bytes;
            const char *hn = reinterpret_cast<const char *>(nrm) + (size_t)r * chunk_bytes;
            if (rows2d) {
                TH_TRY(cudaMemcpy2DAsync(M.in_pos.p, chunk_bytes, hp, (size_t)W * chunk_bytes, chunk_bytes, rows2d, cudaMemcpyHostToDevice, M.st));
                if (!interleaved) TH_TRY(cudaMemcpy2DAsync(M.in_nrm.p, chunk_bytes, hn, (size_t)W * chunk_bytes, chunk_bytes, rows2d, cudaMemcpyHostToDevice, M.st));
            }
            if (owns_last) {
                const uint32_t in_chunk = cnt - rows2d * kShardChunk;                       // 1..64 vertices
                const size_t off_d = (size_t)rows2d * chunk_bytes, off_h = (size_t)rows2d * W * chunk_bytes;
                const size_t nb = (size_t)(in_chunk - 1) * stride + 12;
                TH_TRY(cudaMemcpyAsync((char *)M.in_pos.p + off_d, hp + off_h, interleaved ? nb + (size_t)delta : nb, cudaMemcpyHostToDevice, M.st));
                if (!interleaved) TH_TRY(cudaMemcpyAsync((char *)M.in_nrm.p + off_d, hn + off_h, nb, cudaMemcpyHostToDevice, M.st));
            }
        }
        prt_row_placement pl{};
        pl.shard_world = W; pl.shard_rank = r;
        float *d_out = (float *)M.rows.p;
        if (gather == PRT_GATHER_P2P) {
            pl.out_global = 1;
            for (uint32_t q = 0; q < W; q++) if (q != r) pl.out_peer[pl.n_peer++] = (float *)g->m[q].rows.p;
        } else if (gather == PRT_GATHER_NCCL) d_out = (float *)M.packed.p + (size_t)r * per_rank * n2;
        TH_TRY(cudaEventRecord(M.e1, M.st));
        const float *dp = (const float *)M.in_pos.p;
        const float *dn = interleaved ? reinterpret_cast<const float *>((const char *)M.in_pos.p + delta) : (const float *)M.in_nrm.p;
        if (cnt) {
            const int rc = prt_bake_device(M.ctx, gs ? gs->sc[r] : nullptr, dp, dn, stride, cnt, 0u, p, d_out, nullptr, M.st, nullptr, nullptr, &pl);
            if (rc) { fail(r, rc, prt_last_error()); return; }
        }
        TH_TRY(cudaEventRecord(M.e2, M.st));
    };
    // P2P mode: peers' buffers must exist before anybody launches -> allocate first
    if (gather == PRT_GATHER_P2P)
        for (uint32_t r = 0; r < W; r++) {
            GR_TRY(cudaSetDevice(g->m[r].device));
            GR_TRY(g->m[r].rows.reserve((size_t)v_pad * row_bytes));
        }
    {
        std::vector<std::thread> th;
        for (uint32_t r = 0; r < W; r++) th.emplace_back(phase_a, r);
        for (auto &t : th) t.join();
    }
    for (uint32_t r = 0; r < W; r++) if (rcs[r]) return prt_set_error(rcs[r], errs[r]);

    // ---- phase B: the NCCL baseline gather (main thread, one group call over all communicators) ----------------------------------
    if (gather == PRT_GATHER_NCCL) {
        ncclResult_t nr = g->nccl.GroupStart();
        for (uint32_t r = 0; r < W && nr == 0; r++) {
            float *base = (float *)g->m[r].packed.p;
            nr = g->nccl.AllGather(base + (size_t)r * per_rank * n2, base, (size_t)per_rank * n2, kNcclFloat32, g->comms[r], g->m[r].st);   // in place
        }
        const ncclResult_t ne = g->nccl.GroupEnd();
        if (nr == 0) nr = ne;
        for (uint32_t r = 0; r < W && nr == 0; r++) { ncclResult_t ae = 0; nr = g->nccl.CommGetAsyncError(g->comms[r], &ae); if (nr == 0) nr = ae; }
        if (nr != 0) return prt_set_error(PRT_ERR_CUDA, std::string("ncclAllGather: ") + g->nccl.GetErrorString(nr));
        for (uint32_t r = 0; r < W; r++) {
            GR_TRY(cudaSetDevice(g->m[r].device));
            const unsigned long long total = (unsigned long long)W * per_rank * n2;
            unshard_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, g->m[r].st>>>((const float *)g->m[r].packed.p, (float *)g->m[r].rows.p, per_rank, W, n2, total);
            GR_TRY(cudaGetLastError());
        }
    }

    // ---- phase C (thread per GPU): every GPU copies its own rows to the host, straight into vertex order ---------------------------
    auto phase_c = [&](uint32_t r) {
        Member &M = g->m[r];
        TH_TRY(cudaSetDevice(M.device));
        TH_TRY(cudaEventRecord(M.e3, M.st));
        const uint32_t cnt = rank_count(r);
        if (out && cnt) {
            const uint32_t full = cnt / kShardChunk, tail = cnt % kShardChunk;
            const size_t crow = (size_t)kShardChunk * row_bytes;                      // bytes of one chunk of rows
            char *ho = reinterpret_cast<char *>(out) + (size_t)r * crow;
            const char *src; size_t spitch;
            if (gather == PRT_GATHER_NONE) { src = (const char *)M.rows.p; spitch = crow; }
            else { src = (const char *)M.rows.p + (size_t)r * crow; spitch = (size_t)W * crow; }   // full-size buffer in vertex order
            if (full) TH_TRY(cudaMemcpy2DAsync(ho, (size_t)W * crow, src, spitch, crow, full, cudaMemcpyDeviceToHost, M.st));
            if (tail) TH_TRY(cudaMemcpyAsync(ho + (size_t)full * W * crow, src + (size_t)full * spitch, (size_t)tail * row_bytes, cudaMemcpyDeviceToHost, M.st));
        }
        TH_TRY(cudaStreamSynchronize(M.st));
    };
    {
        std::vector<std::thread> th;
        for (uint32_t r = 0; r < W; r++) th.emplace_back(phase_c, r);
        for (auto &t : th) t.join();
    }
    for (uint32_t r = 0; r < W; r++) if (rcs[r]) return prt_set_error(rcs[r], errs[r]);
#undef TH_TRY
    g->last_n = n; g->last_n2 = n2; g->last_v_pad = (uint32_t)v_pad; g->last_mode = gather;

    if (stats) {
        stats->n_devices = (uint32_t)W; stats->gather_mode = gather;
        stats->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        for (uint32_t r = 0; r < W; r++) {
            cudaSetDevice(g->m[r].device);
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, g->m[r].e0, g->m[r].e1);
            cudaEventElapsedTime(&b, g->m[r].e1, g->m[r].e2);
            cudaEventElapsedTime(&c, g->m[r].e2, g->m[r].e3);
            stats->h2d_ms[r] = a; stats->kernel_ms[r] = b; stats->gather_ms[r] = c;
            stats->kernel_ms_max = std::max(stats->kernel_ms_max, (double)b);
            stats->gather_ms_max = std::max(stats->gather_ms_max, (double)c);
            stats->vertices[r] = rank_count(r);
        }
        stats->gather_bytes_per_gpu = gather == PRT_GATHER_NONE ? 0ull : (uint64_t)(W - 1) * per_rank * row_bytes;
        stats->h2d_bytes = (uint64_t)n * (interleaved ? stride : 24);
        stats->d2h_bytes = out ? (uint64_t)n * row_bytes : 0ull;
    }
    return PRT_OK;
}

// SH_volume::precompute (volume.cpp:149-316) over the group: contiguous probe ranges per GPU (probes keep the x-fastest order of
// volume.cpp:83-90, so the concatenation of the members' CSR slices IS the whole CSR), every member captures its slice, and the
// slices are merged ON THE DEVICE of `target`: the (few thousand) cluster keys and surfel accumulators go through the host to fix
// the global surfel ids (rank in the union of the key sets) and the surfel table; the heavy arrays -- ids, 36-byte transfer rows,
// ranges -- move GPU to GPU (cudaMemcpyPeerAsync over NVLink) straight into their place in the merged arrays, followed by one id
// remap kernel per segment.  The result is identical to a single-GPU capture of all probes (tests/test_gpu_group.py).
int prt_group_probe_capture(prt_group *g, prt_group_scene *gs, const float *probe_pos, uint32_t n_probes, const float *dirs, const float *weights,
                            uint32_t n_dirs, int target, prt_csr **out, double *capture_ms_max, double *merge_ms) {
    if (!g || !gs || gs->g != g || !probe_pos || !dirs || !weights || !out || n_probes == 0 || target < 0 || target >= g->n)
        return prt_set_error(PRT_ERR_INVALID, "prt_group_probe_capture: bad argument");
    *out = nullptr;
    const int W = g->n;
    const uint32_t per = (n_probes + (uint32_t)W - 1) / (uint32_t)W;
    std::vector<prt_csr *> part(W, nullptr);
    std::vector<int> rcs(W, 0);
    std::vector<std::string> errs(W);
    {
        std::vector<std::thread> th;
        for (int r = 0; r < W; r++)
            th.emplace_back([&, r]() {
                const uint32_t lo = std::min((uint32_t)r * per, n_probes), hi = std::min((uint32_t)(r + 1) * per, n_probes);
                if (hi <= lo) return;
                rcs[r] = prt_probe_capture(gs->sc[r], probe_pos + 3 * (size_t)lo, hi - lo, dirs, weights, n_dirs, &part[r]);
                if (rcs[r]) errs[r] = prt_last_error();
            });
        for (auto &t : th) t.join();
    }
    auto cleanup = [&]() { for (prt_csr *c : part) prt_csr_destroy(c); };
    for (int r = 0; r < W; r++) if (rcs[r]) { const int rc = rcs[r]; const std::string e = errs[r]; cleanup(); return prt_set_error(rc, e); }
    const auto t0 = std::chrono::steady_clock::now();
    // ---- global surfel ids: union of the members' (sorted) key sets; accumulators summed per key ------------------------------------
    std::vector<std::vector<unsigned long long>> keys(W);
    std::vector<std::vector<double>> sums(W);
    std::vector<unsigned long long> all;
    double cap_ms = 0.0;
    unsigned long long nnz = 0;
    for (int r = 0; r < W; r++) {
        if (!part[r]) continue;
        cap_ms = std::max(cap_ms, part[r]->capture_ms);
        nnz += part[r]->nnz;
        keys[r].resize(part[r]->n_prim); sums[r].resize((size_t)part[r]->n_prim * 7);
        cudaSetDevice(g->m[r].device);
        if (part[r]->n_prim) {
            cudaMemcpy(keys[r].data(), part[r]->keys, 8 * (size_t)part[r]->n_prim, cudaMemcpyDeviceToHost);
            cudaMemcpy(sums[r].data(), part[r]->sums, 56 * (size_t)part[r]->n_prim, cudaMemcpyDeviceToHost);
        }
        all.insert(all.end(), keys[r].begin(), keys[r].end());
    }
    if (nnz >= 0xFFFFFFFFull) { cleanup(); return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_group_probe_capture: merged CSR exceeds 2^32 entries"); }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    const uint32_t n_prim = (uint32_t)all.size();
    std::vector<double> gsum((size_t)n_prim * 7, 0.0);
    std::vector<std::vector<uint32_t>> remap(W);
    for (int r = 0; r < W; r++) {
        remap[r].resize(keys[r].size());
        for (size_t k = 0; k < keys[r].size(); k++) {
            const uint32_t gidx = (uint32_t)(std::lower_bound(all.begin(), all.end(), keys[r][k]) - all.begin());
            remap[r][k] = gidx;
            for (int q = 0; q < 7; q++) gsum[(size_t)gidx * 7 + q] += sums[r][k * 7 + q];
        }
    }
    std::vector<float> surfels((size_t)n_prim * 6);
    for (uint32_t k = 0; k < n_prim; k++) {                                       // volume.cpp:301-312: mean position, normalised mean normal
        const double cnt = gsum[(size_t)k * 7 + 6];
        double nm[3] = { gsum[(size_t)k * 7 + 3] / cnt, gsum[(size_t)k * 7 + 4] / cnt, gsum[(size_t)k * 7 + 5] / cnt };
        const double ln = std::sqrt(nm[0] * nm[0] + nm[1] * nm[1] + nm[2] * nm[2]);
        for (int q = 0; q < 3; q++) { surfels[(size_t)k * 6 + q] = (float)(gsum[(size_t)k * 7 + q] / cnt); surfels[(size_t)k * 6 + 3 + q] = (float)(nm[q] / ln); }
    }
    // ---- merged arrays on the target GPU --------------------------------------------------------------------------------------------
    Member &T = g->m[target];
    if (cudaSetDevice(T.device) != cudaSuccess) { cleanup(); return prt_set_error(PRT_ERR_CUDA, "prt_group_probe_capture: cudaSetDevice"); }
    prt_csr *c = new prt_csr();
    c->ctx = T.ctx; c->n_probes = n_probes; c->nnz = nnz; c->n_prim = n_prim; c->capture_ms = cap_ms;
    const size_t np = std::max<size_t>(1, n_prim), nz = std::max<size_t>(1, (size_t)nnz);
    cudaError_t e = cudaMallocAsync((void **)&c->range, 8 * (size_t)n_probes, T.st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&c->ids, 4 * nz, T.st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&c->transfer, 36 * nz, T.st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&c->surfels, 24 * np, T.st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&c->keys, 8 * np, T.st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&c->sums, 56 * np, T.st);
    uint32_t *d_remap = nullptr;
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_remap, 4 * np * (size_t)W, T.st);
    unsigned long long base = 0;
    for (int r = 0; r < W && e == cudaSuccess; r++) {
        if (!part[r]) continue;
        const uint32_t lo = (uint32_t)r * per;
        const unsigned long long nr = part[r]->nnz;
        if (nr) {
            e = cudaMemcpyPeerAsync(c->ids + base, T.device, part[r]->ids, g->m[r].device, 4 * (size_t)nr, T.st);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(c->transfer + 9 * base, T.device, part[r]->transfer, g->m[r].device, 36 * (size_t)nr, T.st);
        }
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(c->range + 2 * (size_t)lo, T.device, part[r]->range, g->m[r].device, 8 * (size_t)part[r]->n_probes, T.st);
        if (e == cudaSuccess && !remap[r].empty()) e = cudaMemcpyAsync(d_remap + (size_t)r * np, remap[r].data(), 4 * remap[r].size(), cudaMemcpyHostToDevice, T.st);
        if (e == cudaSuccess && nr) remap_ids_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, T.st>>>(c->ids + base, d_remap + (size_t)r * np, nr);
        if (e == cudaSuccess && base) shift_range_kernel<<<(2 * part[r]->n_probes + 255) / 256, 256, 0, T.st>>>(c->range + 2 * (size_t)lo, 2 * part[r]->n_probes, (uint32_t)base);
        base += nr;
    }
    if (e == cudaSuccess && n_prim) {
        e = cudaMemcpyAsync(c->surfels, surfels.data(), 24 * (size_t)n_prim, cudaMemcpyHostToDevice, T.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->keys, all.data(), 8 * (size_t)n_prim, cudaMemcpyHostToDevice, T.st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->sums, gsum.data(), 56 * (size_t)n_prim, cudaMemcpyHostToDevice, T.st);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(T.st);
    if (d_remap) cudaFreeAsync(d_remap, T.st);
    cleanup();
    if (e != cudaSuccess) { prt_csr_destroy(c); return prt_set_error(PRT_ERR_CUDA, std::string("prt_group_probe_capture: ") + cudaGetErrorString(e)); }
    if (capture_ms_max) *capture_ms_max = cap_ms;
    if (merge_ms) *merge_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = c;
    return PRT_OK;
}

int prt_group_download_rows(prt_group *g, int member, float *out) {
    if (!g || !out || member < 0 || member >= g->n) return prt_set_error(PRT_ERR_INVALID, "prt_group_download_rows: bad argument");
    if (g->last_mode == PRT_GATHER_NONE) return prt_set_error(PRT_ERR_INVALID, "prt_group_download_rows: the last bake did not gather");
    GR_TRY(cudaSetDevice(g->m[member].device));
    GR_TRY(cudaMemcpy(out, g->m[member].rows.p, (size_t)g->last_n * g->last_n2 * 4, cudaMemcpyDeviceToHost));
    return PRT_OK;
}

const float *prt_group_rows_device(prt_group *g, int member) {
    if (!g || member < 0 || member >= g->n || g->last_mode == PRT_GATHER_NONE) return nullptr;
    return (const float *)g->m[member].rows.p;
}

}  // extern "C"
