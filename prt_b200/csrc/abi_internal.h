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
