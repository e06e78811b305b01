// abi_internal.h -- helpers shared by the translation units that implement the C ABI.
#pragma once
#include "../../include/prt_b200.h"
#include "bvh8.h"
#include <cuda_runtime.h>
#include <string>

int prt_set_error(int code, const std::string &msg);   // records the thread-local message, returns code
cudaStream_t prt_ctx_stream(prt_ctx *);                // the context's own stream
int prt_ctx_sms(const prt_ctx *);
int prt_ctx_refill_thresh(const prt_ctx *);            // tuning knob: refill idle lanes when fewer than this many are traversing
int prt_ctx_entry_list(const prt_ctx *);               // tuning knob: per-origin entry lists (vertex bakes, probe capture)

// grow-only device scratch owned by the context (8 slots; nullptr when the allocation fails) and the phase timer behind
// prt_ctx_last_kernel_ms: entry points outside the bake bracket their kernels with begin / end on the stream they launch on
void *prt_ctx_scratch(prt_ctx *, int slot, size_t bytes);
void prt_ctx_timer_begin(prt_ctx *, cudaStream_t);
void prt_ctx_timer_end(prt_ctx *, cudaStream_t);

// host BVH build / per-GPU upload, shared by prt_scene_create and the multi-GPU driver (group.cu)
int prt_build_host_bvh(const float *pos, size_t stride, uint32_t n_verts, const uint32_t *idx, uint32_t n_tris, prt::HostBVH8 *out);
int prt_scene_from_host_bvh(prt_ctx *, const prt::HostBVH8 *, prt_scene **out);

// where the rows of a bake go (all zero: packed [n][order^2] in launch order) -- see BakeArgs in kernels.h
struct prt_row_placement {
    uint32_t out_stride_floats;         // 0 = order^2
    uint32_t shard_world, shard_rank;   // interleaved 64-vertex chunks of a longer list; keys the bounce RNG by the global vertex id
    int out_global;                     // rows stored at the global index
    int n_peer; float *out_peer[7];     // fused gather: rows also stored into these peer-GPU buffers
};
int prt_bake_device(prt_ctx *, prt_scene *, const float *d_pos, const float *d_nrm, size_t stride, uint32_t n, uint32_t vid_base,
                    const prt_bake_params *, float *d_out, uint32_t *d_vis, cudaStream_t, cudaEvent_t e0, cudaEvent_t e1,
                    const prt_row_placement *);
int prt_ctx_device_id(const prt_ctx *);

// device-side view of a scene for the other translation units
struct prt_scene_view { const prt::Node8 *nodes; const prt::Tri48 *tris; prt_ctx *ctx; };
prt_scene_view prt_scene_get_view(prt_scene *);

// CSR of a probe capture (probe.cu); device pointers
struct prt_csr {
    prt_ctx *ctx = nullptr;
    uint32_t n_probes = 0, n_prim = 0;
    unsigned long long nnz = 0;
    uint32_t *range = nullptr, *ids = nullptr;
    float *transfer = nullptr, *surfels = nullptr;
    unsigned long long *keys = nullptr;   // sorted distinct cluster keys
    double *sums = nullptr;               // [n_prim][7] sum of hit positions, sum of hit normals, hit count
    double capture_ms = 0.0;
};
cudaError_t prt_csr_project_device(const prt_csr *, const float4 *d_radiance, float4 *d_out, cudaStream_t);
