// cache.cu -- on-disk cache of baked results.  Host code only (compiled with the library).
//
// The reference re-bakes at every start (App::setup -> sh_volume.bake(), src/platform/app.cpp:52; the per-vertex bake is behind a
// UI button, app.cpp:112-113) and keeps results only in GL objects.  A baked result is a pure function of (mesh, bake parameters),
// so it is stored under that key: a fixed little-endian header {magic, version, kind, mesh hash, parameter block, payload size,
// payload hash} followed by the raw rows.  A load whose key differs is a MISS (re-bake), a damaged file is an error.
#include "../../include/prt_b200.h"
#include "abi_internal.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr char kMagic[8] = {'P', 'R', 'T', 'B', '2', '0', '0', '\0'};
constexpr uint32_t kVersion = 2u, kKindTransfer = 1u, kKindCsr = 2u;

struct Header {
    char magic[8];
    uint32_t version, kind;
    uint64_t mesh_hash, config_hash;
    uint64_t dims[4];            // transfer: n_verts, n_coeffs, 0, 0     csr: n_probes, nnz, n_surfels, 0
    uint64_t payload_bytes, payload_hash;
};
static_assert(sizeof(Header) == 80, "cache header layout");

uint64_t fnv1a(const void *data, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

uint64_t params_hash(const prt_bake_params *p) {
    // field by field: padding never enters the key
    uint64_t h = fnv1a(&p->order, 4);
    h = fnv1a(&p->samples_u, 4, h); h = fnv1a(&p->samples_v, 4, h); h = fnv1a(&p->seed, 4, h); h = fnv1a(&p->bounces, 4, h);
    h = fnv1a(p->albedo, 12, h); h = fnv1a(&p->origin_eps, 4, h); h = fnv1a(&p->bounce_eps, 4, h);
    h = fnv1a(&p->mode, 4, h); h = fnv1a(&p->cs_phase, 4, h); h = fnv1a(&p->jitter, 4, h);
    return h;
}

struct Part { const void *p; size_t n; };

int save(const char *path, Header h, const std::vector<Part> &parts) {
    h.payload_bytes = 0; h.payload_hash = 0xcbf29ce484222325ull;
    for (const Part &q : parts) { h.payload_bytes += q.n; h.payload_hash = fnv1a(q.p, q.n, h.payload_hash); }
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) return prt_set_error(PRT_ERR_IO, std::string("cache: cannot open ") + tmp + " for writing");
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1;
    for (const Part &q : parts) ok = ok && (q.n == 0 || std::fwrite(q.p, 1, q.n, f) == q.n);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok || std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return prt_set_error(PRT_ERR_IO, std::string("cache: write to ") + path + " failed"); }
    return PRT_OK;
}

// reads and checks the header; PRT_ERR_CACHE_MISS when the file is absent or was written for another key
int open_checked(const char *path, uint32_t kind, uint64_t mesh_hash, uint64_t config_hash, Header &h, FILE *&f) {
    f = std::fopen(path, "rb");
    if (!f) return prt_set_error(PRT_ERR_CACHE_MISS, std::string("cache: ") + path + " does not exist");
    if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, kMagic, 8) != 0) { std::fclose(f); f = nullptr; return prt_set_error(PRT_ERR_IO, std::string("cache: ") + path + " is not a prt_b200 cache file"); }
    if (h.version != kVersion || h.kind != kind || h.mesh_hash != mesh_hash || h.config_hash != config_hash) {
        std::fclose(f); f = nullptr;
        return prt_set_error(PRT_ERR_CACHE_MISS, std::string("cache: ") + path + " was written for a different mesh, parameter set or format version");
    }
    return PRT_OK;
}

int read_parts(FILE *f, const Header &h, const std::vector<std::pair<void *, size_t>> &parts, const char *path) {
    uint64_t total = 0, hash = 0xcbf29ce484222325ull;
    bool ok = true;
    for (const auto &q : parts) {
        total += q.second;
        ok = ok && (q.second == 0 || std::fread(q.first, 1, q.second, f) == q.second);
        if (ok) hash = fnv1a(q.first, q.second, hash);
    }
    char extra;
    ok = ok && total == h.payload_bytes && std::fread(&extra, 1, 1, f) == 0;
    std::fclose(f);
    if (!ok || hash != h.payload_hash) return prt_set_error(PRT_ERR_IO, std::string("cache: ") + path + " is truncated or corrupt");
    return PRT_OK;
}

}  // namespace

extern "C" {

uint64_t prt_hash_bytes(const void *data, size_t n_bytes, uint64_t seed) { return fnv1a(data, n_bytes, seed ? seed : 0xcbf29ce484222325ull); }

// The baked rows are a function of positions, NORMALS (ray origin P + eps N, cosine frame) and triangles: all three are hashed.
// nrm_xyz may be NULL for results that do not depend on normals (a probe capture keys on positions + triangles only).
uint64_t prt_mesh_hash(const float *pos_xyz, const float *nrm_xyz, size_t stride_bytes, uint32_t n_verts, const uint32_t *tri_idx, uint32_t n_tris) {
    if (!pos_xyz || (!tri_idx && n_tris)) return 0;
    if (stride_bytes == 0) stride_bytes = 12;
    uint64_t h = fnv1a(&n_verts, 4);
    h = fnv1a(&n_tris, 4, h);
    for (uint32_t i = 0; i < n_verts; i++) h = fnv1a(reinterpret_cast<const char *>(pos_xyz) + (size_t)i * stride_bytes, 12, h);
    if (nrm_xyz) {
        const uint32_t tag = 0x4E524D31u;   // "NRM1"
        h = fnv1a(&tag, 4, h);
        for (uint32_t i = 0; i < n_verts; i++) h = fnv1a(reinterpret_cast<const char *>(nrm_xyz) + (size_t)i * stride_bytes, 12, h);
    }
    return fnv1a(tri_idx, 12 * (size_t)n_tris, h);
}

int prt_cache_save_transfer(const char *path, uint64_t mesh_hash, uint32_t n_verts, const prt_bake_params *p, const float *coeffs) {
    if (!path || !p || (n_verts && !coeffs) || p->order < 1 || p->order > 5) return prt_set_error(PRT_ERR_INVALID, "prt_cache_save_transfer: bad argument");
    Header h{};
    std::memcpy(h.magic, kMagic, 8); h.version = kVersion; h.kind = kKindTransfer; h.mesh_hash = mesh_hash; h.config_hash = params_hash(p);
    h.dims[0] = n_verts; h.dims[1] = (uint64_t)(p->order * p->order);
    return save(path, h, {{coeffs, (size_t)n_verts * h.dims[1] * 4}});
}

int prt_cache_load_transfer(const char *path, uint64_t mesh_hash, uint32_t n_verts, const prt_bake_params *p, float *out_coeffs) {
    if (!path || !p || (n_verts && !out_coeffs) || p->order < 1 || p->order > 5) return prt_set_error(PRT_ERR_INVALID, "prt_cache_load_transfer: bad argument");
    Header h; FILE *f = nullptr;
    const int rc = open_checked(path, kKindTransfer, mesh_hash, params_hash(p), h, f);
    if (rc) return rc;
    if (h.dims[0] != n_verts || h.dims[1] != (uint64_t)(p->order * p->order)) { std::fclose(f); return prt_set_error(PRT_ERR_CACHE_MISS, "prt_cache_load_transfer: vertex count differs"); }
    return read_parts(f, h, {{out_coeffs, (size_t)n_verts * h.dims[1] * 4}}, path);
}

int prt_cache_save_csr(const char *path, uint64_t mesh_hash, uint64_t config_hash, uint32_t n_probes, uint64_t nnz, uint32_t n_surfels,
                       const uint32_t *range, const uint32_t *ids, const float *transfer, const float *surfels, const uint64_t *keys) {
    if (!path || !range || (nnz && (!ids || !transfer)) || (n_surfels && (!surfels || !keys))) return prt_set_error(PRT_ERR_INVALID, "prt_cache_save_csr: bad argument");
    Header h{};
    std::memcpy(h.magic, kMagic, 8); h.version = kVersion; h.kind = kKindCsr; h.mesh_hash = mesh_hash; h.config_hash = config_hash;
    h.dims[0] = n_probes; h.dims[1] = nnz; h.dims[2] = n_surfels;
    return save(path, h, {{range, 8 * (size_t)n_probes}, {ids, 4 * (size_t)nnz}, {transfer, 36 * (size_t)nnz}, {surfels, 24 * (size_t)n_surfels}, {keys, 8 * (size_t)n_surfels}});
}

int prt_cache_csr_sizes(const char *path, uint64_t mesh_hash, uint64_t config_hash, uint32_t *n_probes, uint64_t *nnz, uint32_t *n_surfels) {
    if (!path || !n_probes || !nnz || !n_surfels) return prt_set_error(PRT_ERR_INVALID, "prt_cache_csr_sizes: null argument");
    Header h; FILE *f = nullptr;
    const int rc = open_checked(path, kKindCsr, mesh_hash, config_hash, h, f);
    if (rc) return rc;
    std::fclose(f);
    *n_probes = (uint32_t)h.dims[0]; *nnz = h.dims[1]; *n_surfels = (uint32_t)h.dims[2];
    return PRT_OK;
}

int prt_cache_load_csr(const char *path, uint64_t mesh_hash, uint64_t config_hash, uint32_t n_probes, uint64_t nnz, uint32_t n_surfels,
                       uint32_t *range, uint32_t *ids, float *transfer, float *surfels, uint64_t *keys) {
    if (!path || !range || (nnz && (!ids || !transfer)) || (n_surfels && (!surfels || !keys))) return prt_set_error(PRT_ERR_INVALID, "prt_cache_load_csr: bad argument");
    Header h; FILE *f = nullptr;
    const int rc = open_checked(path, kKindCsr, mesh_hash, config_hash, h, f);
    if (rc) return rc;
    if (h.dims[0] != n_probes || h.dims[1] != nnz || h.dims[2] != n_surfels) { std::fclose(f); return prt_set_error(PRT_ERR_INVALID, "prt_cache_load_csr: buffer sizes differ from prt_cache_csr_sizes"); }
    return read_parts(f, h, {{range, 8 * (size_t)n_probes}, {ids, 4 * (size_t)nnz}, {transfer, 36 * (size_t)nnz}, {surfels, 24 * (size_t)n_surfels}, {keys, 8 * (size_t)n_surfels}}, path);
}

}  // extern "C"
