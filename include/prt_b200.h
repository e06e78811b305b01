/*
 * prt_b200.h -- C ABI of the B200-native PRT precomputation library (libprt_b200.so).
 *
 * Drop-in boundary for the precomputation hot path of lvjiahui/PRT.  The reference has no FFI of its
 * own: its boundary is a handful of C++ entry points called from App.  Every
 * function below names the reference interface it replaces (file:line under the reference tree); the
 * header-only C++ shim include/prt_b200_shim.hpp re-creates those C++ signatures on top of this ABI.
 *
 * Conventions: plain pointers and sizes only; opaque handles; int status (0 = ok, <0 = error, message
 * via prt_last_error(), thread-local); no exceptions cross the boundary; blocking unless a stream is
 * passed to a *_device entry point.  One handle must not be used from two threads concurrently;
 * different handles are independent.  There is NO CPU fallback: every entry point fails with
 * PRT_ERR_CUDA when no sm_100-class device is usable.
 *
 * Host-pointer entry points copy inputs host->device and results device->host themselves (this is the
 * "e2e" path of bench.py).  *_device entry points take device pointers and a cudaStream_t (as void*).
 */
#ifndef PRT_B200_H
#define PRT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRT_B200_ABI_VERSION 2

enum {
    PRT_OK = 0,
    PRT_ERR_INVALID = -1,   /* bad argument */
    PRT_ERR_CUDA = -2,      /* CUDA runtime / no device */
    PRT_ERR_NOMEM = -3,
    PRT_ERR_BUILD = -4,     /* BVH build failed (bad indices, depth limit) */
    PRT_ERR_UNSUPPORTED = -5,
    PRT_ERR_IO = -6,         /* cache file unreadable, truncated or corrupt */
    PRT_ERR_CACHE_MISS = -7  /* no cache file for this (mesh, parameters) key: bake again */
};

typedef struct prt_ctx prt_ctx;
typedef struct prt_scene prt_scene;

const char *prt_last_error(void);
int prt_abi_version(void);

/* Process-global "RTCDevice device = initializeDevice()" (raytracing.cpp:42-52, light_probe.cpp:28-38):
 * one context per GPU.  device_id < 0 -> current device. */
int prt_ctx_create(int device_id, prt_ctx **out);
void prt_ctx_destroy(prt_ctx *);
int prt_ctx_device(const prt_ctx *);
/* CUDA-event duration of the kernels of the most recent prt_env_* / prt_brdf_lut / prt_volume_weights call on this context (those
 * entry points also accept NULL host outputs: the result then stays in the context's device scratch -- timing runs) */
int prt_ctx_last_kernel_ms(const prt_ctx *, double *ms);
/* name/value tuning knobs for experiments; unknown names and out-of-range values fail.  None of them changes a result
 * (tests/test_gpu_parity.py), only how the work is organised:
 *   horizon 0/1 (1)            per-origin horizon pass before the shadowed / interreflected bake
 *   horizon_near 5..157 (157)  subtrees of angular radius above value/100 rad are always refined by the horizon builder
 *   horizon_mid 0..157 (24)    ... and those above this radius when merging their bound would leave more than
 *   horizon_gain (64)          this many tenths of a sample to trace (estimate; 0 = every such box); horizon_mid 0 = rule off
 *   horizon_slabs 0/1 (1)      the builder bounds a subtree by its oriented slab (mean normal of its triangles) inside its box
 *   wave_dop 0/1 (1)           node test of the traversal pass with a fourth slab axis (the node's mean normal; children's extents quantised)
 *   horizon_budget 0..4096 (64) refinement iterations (4 nodes each) per vertex
 *   work_list -1/0/1 (-1)      traversal pass walks the vertices heaviest first (counting sort of the need counts); -1 = below ~1 M vertices
 *   l2_prefetch 0/1 (0)        stream the BVH into L2 before the first pass (measured: no effect, the cold-start misses are hidden)
 *   entry_list, pair_queue (0 per-ray stacks / 2 wavefront), refill_thresh, block, ctas_per_sm   the per-ray fallback kernel (S > 8192, horizon off)
 *   count_work 0/1 (0)         instrumented launch filling the work counters of prt_bake_stats */
int prt_ctx_set_tuning(prt_ctx *, const char *name, int value);

/* RTScene::RTScene(Mesh&) / RTScene(Model&) (raytracing.cpp:58-94, light_probe.cpp:44-87): copies
 * positions (pos_stride_bytes apart; 60 for Mesh::Vert, gl.h:76-80) and uint32 index triples, builds the
 * 8-wide compressed BVH on the host and uploads it to the context's GPU.  ~RTScene -> prt_scene_destroy. */
int prt_scene_create(prt_ctx *, const float *pos_xyz, size_t pos_stride_bytes, uint32_t n_verts,
                     const uint32_t *tri_idx, uint32_t n_tris, prt_scene **out);
void prt_scene_destroy(prt_scene *);

typedef struct {
    uint32_t n_tris, n_nodes;      /* 80-byte wide nodes, 48-byte triangles */
    uint32_t max_depth;
    uint32_t reserved;
    uint64_t node_bytes, tri_bytes;
    double build_seconds, upload_seconds;
    double sah_cost;
} prt_scene_info;
int prt_scene_get_info(const prt_scene *, prt_scene_info *out);

/* ---- ray queries: struct Ray::{any_hit, first_hit, hit_normal} (light_probe.cpp:95-133) -------------
 * rays: n x 8 floats (org.xyz, tnear, dir.xyz, tfar); dir is NOT normalised by the library (segment
 * tests pass probe-voxel with tfar = 1, light_probe.cpp:250).  A hit needs tnear < t <= tfar.
 * any-hit : out_hit[i] = 1/0                         (rtcOccluded1, light_probe.cpp:128)
 * closest : out_t[i] (inf on miss), out_prim[i] (0xFFFFFFFF on miss), out_ng[i][3] = unnormalised
 *           geometric normal (v1-v0)x(v2-v0)           (rtcIntersect1, light_probe.cpp:119; Ng :124)      */
int prt_trace_any_hit(prt_scene *, const float *rays, uint32_t n, uint8_t *out_hit);
int prt_trace_closest_hit(prt_scene *, const float *rays, uint32_t n, float *out_t, uint32_t *out_prim, float *out_ng);

/* ---- per-vertex diffuse SH transfer: bake_SH(Mesh&) (raytracing.cpp:320-360) ----------------------- */
enum { PRT_UNSHADOWED = 0, PRT_SHADOWED = 1, PRT_INTERREFLECT = 2, PRT_UNSHADOWED_ANALYTIC = 3 };

typedef struct {
    int32_t order;        /* bands; coefficients = order^2; reference: file-scope "order"+1 = 3 (raytracing.cpp:320) */
    int32_t samples_u;    /* radial strata;  reference App::sh_resolution = 32 (app.h:71) */
    int32_t samples_v;    /* angular strata; reference App::sh_resolution = 32 */
    uint32_t seed;        /* strata-jitter / bounce RNG seed (Philox4x32-10) */
    int32_t bounces;      /* B; path depth = B+1 = App::max_path_length-1 (raytracing.cpp:345, app.h:70) */
    float albedo[3];      /* App::albedo (app.h:55) */
    float origin_eps;     /* 1e-4 (raytracing.cpp:343) */
    float bounce_eps;     /* 1e-5 (raytracing.cpp:235) */
    int32_t mode;         /* PRT_SHADOWED etc. */
    int32_t cs_phase;     /* 1 (default): Condon-Shortley sign (-1)^m on the basis, as google/spherical-harmonics' sh::EvalSH has it --
                             what bake_SH itself computes (raytracing.cpp:226); 0: the sign-free convention of the reference's own
                             SH_function.h / common/SH.glsl, which the probe path and the viewer's SH_Irad() use */
    int32_t jitter;       /* 1: jittered strata (raytracing.cpp:338-339); 0: stratum centres */
} prt_bake_params;

/* The reference's defaults: order 3, 32 x 32 jittered strata, no bounce (max_path_length 2), albedo 1, eps 1e-4 / 1e-5, shadowed,
 * cs_phase 1 (bake_SH evaluates the basis with sh::EvalSH). */
void prt_bake_params_default(prt_bake_params *);

/* Host buffers in, host buffers out.  pos/nrm point at the first vertex's position / normal, consecutive
 * vertices stride_bytes apart (60 for Mesh::Vert).  out_coeffs [n_verts][order^2] floats, row-major,
 * k = l(l+1)+m.  out_vis optional: [n_verts][ceil(S/32)] uint32, bit s set <=> primary ray s (s = i*samples_v+j)
 * is unoccluded.  vertex_id_base: id of vertex 0 for the bounce RNG key (sharded bakes).  scene may be NULL for the
 * unshadowed modes. */
int prt_bake_transfer(prt_ctx *, prt_scene *, const float *pos, const float *nrm, size_t stride_bytes,
                      uint32_t n_verts, uint32_t vertex_id_base, const prt_bake_params *,
                      float *out_coeffs, uint32_t *out_vis);

/* Same with device-resident inputs/outputs, asynchronous on `stream` (a cudaStream_t). */
int prt_bake_transfer_device(prt_ctx *, prt_scene *, const float *d_pos, const float *d_nrm, size_t stride_bytes,
                             uint32_t n_verts, uint32_t vertex_id_base, const prt_bake_params *,
                             float *d_out_coeffs, uint32_t *d_out_vis, void *stream);

/* One shard of a multi-process bake (one process per GPU, e.g. under torchrun; prt_group_* is the single-process driver): the n_verts
 * device-resident vertices are shard `shard_rank` of `shard_world` interleaved 64-vertex chunks of a longer list -- local vertex v is
 * list position ((v / 64) * shard_world + shard_rank) * 64 + v % 64, which keys the bounce RNG, so an interreflection bake does
 * not depend on the number of ranks.  Rows are written packed in local order (the layout an in-place all-gather wants). */
int prt_bake_transfer_device_shard(prt_ctx *, prt_scene *, const float *d_pos, const float *d_nrm, size_t stride_bytes, uint32_t n_verts,
                                   uint32_t shard_world, uint32_t shard_rank, const prt_bake_params *,
                                   float *d_out_coeffs, uint32_t *d_out_vis, void *stream);

/* The same shard with the gather FUSED into the kernels, for one process per GPU: rows are stored at their list position in this
 * rank's full-size buffer d_rows_full [padded list][order^2] AND in the full-size buffers of the peer ranks (device pointers valid
 * in this process: opened with prt_ipc_open from handles the peers exported with prt_ipc_export) -- P2P stores over NVLink from the
 * projection epilogue, so no collective follows the kernel; the ranks only need a barrier before anybody READS the gathered rows.
 * The padded list has ceil(ceil(n_list / 64) / world) * world * 64 rows. */
int prt_bake_transfer_device_shard_fused(prt_ctx *, prt_scene *, const float *d_pos, const float *d_nrm, size_t stride_bytes, uint32_t n_verts,
                                         uint32_t shard_world, uint32_t shard_rank, const prt_bake_params *, float *d_rows_full,
                                         float *const *peer_rows_full, int32_t n_peers, void *stream);
/* plain device allocations (cudaMalloc, zeroed) that can be exported to other processes, and CUDA IPC export / open / close */
int prt_device_alloc(prt_ctx *, size_t bytes, void **out_d_ptr);
int prt_device_free(prt_ctx *, void *d_ptr);
int prt_ipc_export(prt_ctx *, const void *d_ptr, uint8_t handle[64]);
int prt_ipc_open(prt_ctx *, const uint8_t handle[64], void **out_d_ptr);
int prt_ipc_close(prt_ctx *, void *d_ptr);

/* Same, with the rows written out_stride_bytes apart instead of packed: d_out_coeffs + 24 bytes into a device-resident Mesh::Vert
 * array with out_stride_bytes = 60 and order 3 fills sh_coeff[9] of every vertex in place (gl.h:76-80) -- the device-side half of
 * Mesh::update (gl.cpp:255-268) for a VBO mapped with cudaGraphicsGLRegisterBuffer: no host hop, no scatter pass. */
int prt_bake_transfer_device_strided(prt_ctx *, prt_scene *, const float *d_pos, const float *d_nrm, size_t stride_bytes,
                                     uint32_t n_verts, uint32_t vertex_id_base, const prt_bake_params *,
                                     float *d_out_coeffs, size_t out_stride_bytes, uint32_t *d_out_vis, void *stream);

/* Mesh::Vert scatter for the viewer (gl.h:76-80, gl.cpp:255-268): writes order-3 rows into sh_coeff[9] of an
 * interleaved 60-byte vertex array on the host. */
int prt_scatter_sh9(const float *coeffs, int32_t order, uint32_t n_verts, void *mesh_verts, size_t vert_stride_bytes,
                    size_t sh_offset_bytes);

/* Sample table used by the bake: uv[S][2], local_dirs[S][3] in reference order s = i*samples_v + j (host). */
int prt_bake_sample_table(const prt_bake_params *, float *uv, float *local_dirs);

/* Per-call statistics of the most recent bake on this context */
typedef struct {
    double kernel_ms;        /* all kernels of the bake (horizon pass + traversal/projection), CUDA events on the launching stream */
    double h2d_ms, d2h_ms;   /* host-pointer entry point only */
    uint64_t rays;           /* primary rays issued */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t launches;       /* kernels launched by the call */
    uint32_t grid, block;
    uint64_t node_visits;    /* only with tuning knob count_work=1: 80-byte node fetches ... */
    uint64_t tri_tests;      /* ... and 48-byte triangle fetches of the launch (algorithmic traversal work) */
    uint64_t cand_tests;     /* 32-byte entry-list candidate boxes tested (shared memory) */
    uint64_t rays_traversed; /* rays NOT resolved by the horizon map (0 when the kernel variant has no horizon map) */
    double horizon_ms;       /* part of kernel_ms spent in the horizon pass (0 when it did not run) */
} prt_bake_stats;
int prt_ctx_last_bake_stats(const prt_ctx *, prt_bake_stats *out);

/* ---- multi-GPU driver: one process, 1..8 GPUs of one box (BASELINE config 5) ---------------------------------
 * The reference's per-vertex loop (std::for_each(par, verts...), raytracing.cpp:328) sharded over GPUs: one context per device, the
 * BVH built once and replicated, the vertex list dealt out in interleaved 64-vertex chunks (callers should pass vertices in a
 * space-filling-curve order), coefficient rows gathered on every GPU and returned to the host in the caller's vertex order.
 * Results are bit-identical to a single-GPU prt_bake_transfer of the same list (each vertex is reduced by one warp in a fixed order;
 * the bounce RNG is keyed by the position in the list). */
typedef struct prt_group prt_group;
typedef struct prt_group_scene prt_group_scene;
enum {
    PRT_GATHER_NONE = 0,   /* rows stay on the GPU that baked them (the host still receives all rows) */
    PRT_GATHER_NCCL = 1,   /* in-place ncclAllGather (ncclCommInitAll communicator over the group) + permutation to vertex order */
    PRT_GATHER_P2P = 2,    /* fused: the bake kernels store every finished row into every peer GPU's buffer (P2P over NVLink) */
    PRT_GATHER_AUTO = -1   /* P2P if every pair has peer access, else NCCL, else NONE */
};
/* contexts on the listed devices, peer access between every pair, an NCCL communicator over them (NCCL is loaded at run time with
 * dlopen; without it, or with a device listed twice, only the P2P / NONE modes work -- prt_group_capabilities says which and why) */
int prt_group_create(const int *device_ids, int n_devices, prt_group **out);
void prt_group_destroy(prt_group *);
int prt_group_size(const prt_group *);
prt_ctx *prt_group_ctx(prt_group *, int member);
int prt_group_capabilities(const prt_group *, int *p2p, int *nccl, int *nccl_version, char *why, size_t why_len);
int prt_group_set_tuning(prt_group *, const char *name, int value);            /* prt_ctx_set_tuning on every member */
/* RTScene(Mesh&) once for the whole group: ONE host BVH build, one upload per GPU (parallel over the PCIe links) */
int prt_group_scene_create(prt_group *, const float *pos_xyz, size_t pos_stride_bytes, uint32_t n_verts,
                           const uint32_t *tri_idx, uint32_t n_tris, prt_group_scene **out);
void prt_group_scene_destroy(prt_group_scene *);
int prt_group_scene_get_info(const prt_group_scene *, prt_scene_info *out);      /* upload_seconds = slowest GPU */
prt_scene *prt_group_scene_member(prt_group_scene *, int member);
typedef struct {
    uint32_t n_devices; int32_t gather_mode;     /* the mode used (AUTO resolved) */
    double wall_ms;                              /* host clock around the whole call */
    double kernel_ms_max, gather_ms_max;         /* slowest GPU: bake kernels; collective + permutation (NCCL mode; ~0 otherwise) */
    double h2d_ms[8], kernel_ms[8], gather_ms[8];/* per GPU, CUDA events on its stream */
    uint32_t vertices[8];                        /* vertices baked per GPU */
    uint64_t gather_bytes_per_gpu;               /* row bytes every GPU receives from its peers */
    uint64_t h2d_bytes, d2h_bytes;               /* summed over the GPUs */
} prt_group_stats;
/* bake_SH over all GPUs of the group: host vertices in (layout as prt_bake_transfer; an interleaved Mesh::Vert array is uploaded
 * once), out_coeffs [n_verts][order^2] on the host (may be NULL: rows stay on the GPUs).  Every GPU uploads its own shard and
 * downloads its own rows over its own PCIe link; pinned host buffers make those copies asynchronous.  stats may be NULL. */
int prt_group_bake_transfer(prt_group *, prt_group_scene *, const float *pos, const float *nrm, size_t stride_bytes, uint32_t n_verts,
                            const prt_bake_params *, float *out_coeffs, int gather_mode, prt_group_stats *stats);
/* the full [n_verts][order^2] rows as gathered on one member GPU by the last bake (modes NCCL / P2P): device pointer / host copy */
const float *prt_group_rows_device(prt_group *, int member);
int prt_group_download_rows(prt_group *, int member, float *out_coeffs);

/* SH_volume::precompute (volume.cpp:149-316) over the group: contiguous probe ranges per GPU, the per-GPU CSR slices merged on the device
 * of member `target` (GPU-to-GPU copies + an id remap kernel; only the few thousand cluster keys and surfel accumulators pass through
 * the host).  The resulting prt_csr lives on prt_group_ctx(group, target) and equals a single-GPU prt_probe_capture of all probes. */
typedef struct prt_csr prt_csr;
int prt_group_probe_capture(prt_group *, prt_group_scene *, const float *probe_pos_xyz, uint32_t n_probes, const float *dirs_xyz,
                            const float *solid_angles, uint32_t n_dirs, int target, prt_csr **out, double *capture_kernel_ms_max, double *merge_ms);

/* ---- image-based lighting (BASELINE config 2) ---------------------------------------------------------------------
 * Texture semantics the GL driver leaves open are pinned: FP32 storage, GL cube (sc,tc) table ==
 * cubeCoordToWorld (SH_function.h:104-109), bilinear centres at (i+0.5)/N, seamless re-projection of off-face taps,
 * 2x2-box mips, trilinear = lerp of two levels.  Cube buffers are [6][n][n][3] floats, faces +X,-X,+Y,-Y,+Z,-Z; mip
 * chains are stored level after level. */
typedef struct prt_env prt_env;

/* load_hdr + LightProbe::equirectangular_to_cubemap + CubeMap::generateMipmap (util.cpp:6-24, gl.cpp:581-591,454-460;
 * app.cpp:41-46): equirect RGB float image (top row first, as stb_image returns it) -> GPU-resident cube + full mip chain.
 * cube_size = 512 in the reference (app.cpp:44). */
int prt_env_create(prt_ctx *, const float *equirect_rgb, int w, int h, int cube_size, prt_env **out);
void prt_env_destroy(prt_env *);
int prt_env_levels(const prt_env *);
int prt_env_get_cube(prt_env *, int level, float *out_rgb);
/* LightProbe::irradiance (gl.cpp:569-579, irradiance.frag:9-43); n_out = 32 in the reference (app.cpp:55) */
int prt_env_irradiance(prt_env *, int n_out, float *out_rgb);
/* LightProbe::prefilter (gl.cpp:546-567, prefilter.frag:64-107); reference: n_out 256, mips 5, 1024 samples (app.cpp:58).
 * out_rgb holds the mips one after another: sum over m of 6*(n_out>>m)^2*3 floats; roughness = m/(mips-1). */
int prt_env_prefilter(prt_env *, int n_out, int mips, int n_samples, float *out_rgb);
/* brdfLUT.render_to(screen_quad) with brdf.frag (app.cpp:61-63, brdf.frag:69-113): out_rg[h][w][2] = (A,B),
 * NdotV = (x+0.5)/w, roughness = (y+0.5)/h; reference 512 x 512, 1024 samples. */
int prt_brdf_lut(prt_ctx *, int w, int h, int n_samples, float *out_rg);
/* environment -> SH coefficients, out_rgb_coeffs[order^2][3] (k = l(l+1)+m, sh-space (z,x,y)).
 * method 0: lat-long quadrature of bak/projectSH.comp:63-151 (size = 256 phi columns, theta step 2pi/size);
 * method 1: cube-texel quadrature of bak/image_projectSH.comp:64-133 (size = 64, weight |p|^-3 * 4/size^2). */
int prt_env_project_sh(prt_env *, int order, int method, int size, float *out_rgb_coeffs);
/* Ramamoorthi-Hanrahan polynomial pack consumed by the viewer's SH_Irad (common/SH.glsl:17-36; precomp_projectSH.comp:
 * 23,118-139): L9_rgb[9][3] -> out28 = Ar,Ag,Ab,Br,Bg,Bb (vec4 each) then C = (rgb,1). Host helper. */
int prt_sh_pack_rh(const float *L9_rgb, float *out28);

/* ---- light-probe GI: per-probe transfer capture and projection (BASELINE config 3) ----------------------------------
 * SH_volume::precompute (volume.cpp:149-316) with the 64x64x6 G-buffer raster + readback replaced by closest-hit rays:
 * for every probe and every direction d_i (solid angle w_i): sky and back-face hits are skipped (:246-249), the hit is
 * binned into its surfel cluster (floor(pos), signed principal axis of the normal; :205-223) and
 * transfer[probe][cluster] += SH9(sh-space direction) * w_i (:250-260).  Result = the CSR the viewer uploads
 * (probe_range / ID_buffer / transfer_buffer, :265-295) plus the surfel table (primitive_buffer, :301-312), kept on the GPU.
 * Surfel ids are the rank of the cluster key in (x,y,z,direction) lexicographic order (the reference numbers clusters
 * first-seen, which depends on probe order; parity is defined up to that permutation).  n_dirs <= 4096. */
int prt_probe_capture(prt_scene *, const float *probe_pos_xyz, uint32_t n_probes, const float *dirs_xyz, const float *solid_angles,
                      uint32_t n_dirs, prt_csr **out);
void prt_csr_destroy(prt_csr *);
int prt_csr_sizes(const prt_csr *, uint32_t *n_probes, uint64_t *nnz, uint32_t *n_surfels, double *capture_kernel_ms);
/* range[n_probes][2] (start,end), ids[nnz], transfer[nnz][9], surfels[n_surfels][6] (mean position, normalised mean normal),
 * keys[n_surfels] (packed cluster keys, ascending); any pointer may be NULL */
int prt_csr_download(const prt_csr *, uint32_t *range, uint32_t *ids, float *transfer, float *surfels, uint64_t *keys);
/* Partial captures (multi-GPU: every rank captures a slice of the probes, volume.cpp:83-90 order) are merged on the host:
 * the union of the key sets gives the global surfel ids, and the surfel table needs the un-normalised accumulators --
 * sums[n_surfels][7] = sum of hit positions, sum of hit normals, hit count (volume.cpp:229-236,301-312). */
int prt_csr_surfel_sums(const prt_csr *, double *out_sums);
/* Builds a device CSR from host arrays (a merged capture, or one read back from a cache) so prt_probe_project can run on it.
 * keys may be NULL.  Ranges and ids are validated. */
int prt_csr_upload(prt_ctx *, uint32_t n_probes, uint64_t nnz, uint32_t n_surfels, const uint32_t *range, const uint32_t *ids,
                   const float *transfer, const float *surfels, const uint64_t *keys, prt_csr **out);
/* SH_volume::project_sh + precomp_projectSH.comp:32-143: radiance_rgba[n_surfels][4] -> out[n_probes][7][4]
 * (Ar,Ag,Ab,Br,Bg,Bb,C of common/SH.glsl:1-7), FP32 instead of the reference's RGBA16F storage */
int prt_probe_project(const prt_csr *, const float *radiance_rgba, float *out_sh_volumes);
/* host helpers: probe grid positions (volume.cpp:83-90, x fastest), get_dirs (light_probe.cpp:137-152) and the cube-texel
 * direction set of cubeCoordToWorld (SH_function.h:96-112) with the reference's solid angle 4/res^2/|c|^3 (volume.cpp:251-254) */
int prt_probe_positions(const int32_t res[3], const float scene_size[3], float *out_pos_xyz);
int prt_fibonacci_dirs(int32_t n, float *out_dirs_xyz);
int prt_cube_dirs(int32_t res, float *out_dirs_xyz /*[6*res*res][3]*/, float *out_solid_angles);

/* ---- per-frame probe pipeline: SH_volume::relight + project_sh (volume.cpp:357-452) ----------------------
 * relight.comp:68-82 lights every surfel (sky light with the shadow-map test of common/paral_shadow.glsl, spot light and point
 * light of common/light.glsl, albedo of colored_wall.glsl unless given, optional SH_Irad feedback of common/SH.glsl from the
 * volumes of the previous round, temporal blend temp_weight), precomp_projectSH.comp projects the radiance into per-probe SH,
 * transfer2volume.comp:36-147 blends the 8 surrounding probes into every voxel with the calculate_weight masks.  State stays in
 * HBM inside prt_gi; prt_gi_step runs n_rounds rounds (one round = one displayed frame of app.cpp:164-166) back to back.
 * Storage is FP32 (reference: RGBA16F volumes).  Layouts: radiance [n_surfels][4]; probe_sh [pz][py][px][7][4];
 * volumes [vz][vy][vx][7][4] (Ar,Ag,Ab,Br,Bg,Bb,C of common/SH.glsl:1-7). */
typedef struct prt_relight_params {
    float cast_intensity[3], cast_position[3], cast_direction[3], cast_cutoff;   /* CastLight  light.glsl:1-6;  volume.cpp:362-365 */
    float ambient_intensity[3], ambient_position[3];                             /* PointLight light.glsl:8-11; volume.cpp:366-367 */
    float sky_intensity[3], sky_direction[3];                                    /* ParalLight light.glsl:13-16; volume.cpp:374-375 */
    float light_space_matrix[16];                                                /* column-major like glm; volume.cpp:373 */
    int32_t multi_bounce;                                                        /* app.h:33 */
    float atten, sh_shift, temp_weight;                                          /* app.h:34,31; relight.comp:12 (0.1) */
} prt_relight_params;
/* Paral_Shadow::set_dir (gl.cpp:620-631): direction and ortho(-30,30,-30,30,0.1,60) * lookAt(30 d, 0, up) from (up, dir) in [0,1]^2 */
int prt_paral_shadow_matrix(float up, float dir, float out_direction[3], float out_matrix[16]);
/* Paral_Shadow::render (gl.cpp:633-648) without a rasteriser: depth[j*size+i] in [0,1] = closest hit of the ray through the
 * centre of texel (i,j) of an orthographic (affine) light matrix; 1 = nothing.  size <= 16384 (reference: 4096, gl.h:314). */
int prt_shadow_map(prt_scene *, const float matrix[16], int32_t size, float *out_depth);
typedef struct prt_gi prt_gi;
/* the CSR is borrowed and must outlive the object; w0123/w4567 from prt_volume_weights; prod(probe_res) == captured probes */
int prt_gi_create(const prt_csr *, const int32_t probe_res[3], const int32_t volume_res[3], const float scene_size[3],
                  const float *w0123, const float *w4567, prt_gi **out);
void prt_gi_destroy(prt_gi *);
int prt_gi_set_shadow_map(prt_gi *, const float *depth /*NULL: unshadowed*/, int32_t size);
int prt_gi_set_albedo(prt_gi *, const float *albedo_rgb /*[n_surfels][3]; NULL: colored_wall.glsl*/);
int prt_gi_set_radiance(prt_gi *, const float *radiance_rgba);
int prt_gi_step(prt_gi *, const prt_relight_params *, int32_t n_rounds);
int prt_gi_download(const prt_gi *, float *radiance_rgba, float *probe_sh, float *volumes);   /* any pointer may be NULL */

/* ---- progressive preview tracer: raytrace / renderAO / renderNormal (raytracing.cpp:162-222,280-317) --------
 * The film holds App::pixels_w (sum rgb, count) and App::pixels (RGBA8) in HBM; prt_raytrace adds n_frames samples per pixel
 * (one call of the reference's raytrace() per frame).  Camera vectors as in the reference's Camera (Front/Up/Right unit vectors,
 * Zoom in degrees); ray (i,j) = Front - Up/2 - Right/2 + j/h Up + i/w Right with |Up| = 2 tan(Zoom/2), |Right| = |Up| w/h, row j = 0
 * at the bottom.  Bounce randoms: Philox stream 2 keyed (seed; pixel, frame, bounce). */
typedef struct prt_camera { float position[3], front[3], up[3], right[3], zoom_deg; } prt_camera;
enum { PRT_RAYTRACE_AO = 0, PRT_RAYTRACE_NORMAL = 1 };
typedef struct prt_film prt_film;
int prt_film_create(prt_ctx *, int32_t width, int32_t height, prt_film **out);
void prt_film_destroy(prt_film *);
int prt_film_reset(prt_film *);                                    /* camera.dirty: clears the accumulators (raytracing.cpp:282-285) */
int prt_raytrace(prt_scene *, prt_film *, const prt_camera *, int32_t max_path_length /*app.h: 3*/, const float albedo[3],
                 int32_t gamma, int32_t mode, uint32_t seed, int32_t n_frames);
int prt_film_download(const prt_film *, float *accum /*[h*w][4]*/, uint8_t *pixels_rgba8 /*[h*w][4]*/);   /* either may be NULL */

/* ---- on-disk cache of baked results --------------------------------------------------------------------
 * The reference re-bakes at every start (app.cpp:52) and keeps results only in GL objects.  Files: 80-byte header {magic
 * "PRTB200", version, kind, mesh hash, key hash, dims, payload bytes, payload FNV-1a} + raw rows.  Loads return
 * PRT_ERR_CACHE_MISS when the file is absent or keyed differently (other mesh / parameters / version) and PRT_ERR_IO when it is
 * damaged.  Host-only calls (no GPU needed). */
uint64_t prt_hash_bytes(const void *data, size_t n_bytes, uint64_t seed /*0: FNV offset basis; chain calls by passing the last result*/);
uint64_t prt_mesh_hash(const float *pos_xyz, const float *nrm_xyz, size_t stride_bytes, uint32_t n_verts, const uint32_t *tri_idx, uint32_t n_tris);
int prt_cache_save_transfer(const char *path, uint64_t mesh_hash, uint32_t n_verts, const prt_bake_params *, const float *coeffs);
int prt_cache_load_transfer(const char *path, uint64_t mesh_hash, uint32_t n_verts, const prt_bake_params *, float *out_coeffs);
/* config_hash: the caller's hash of everything else the capture depends on (probe positions, directions, solid angles) */
int prt_cache_save_csr(const char *path, uint64_t mesh_hash, uint64_t config_hash, uint32_t n_probes, uint64_t nnz, uint32_t n_surfels,
                       const uint32_t *range, const uint32_t *ids, const float *transfer, const float *surfels, const uint64_t *keys);
int prt_cache_csr_sizes(const char *path, uint64_t mesh_hash, uint64_t config_hash, uint32_t *n_probes, uint64_t *nnz, uint32_t *n_surfels);
int prt_cache_load_csr(const char *path, uint64_t mesh_hash, uint64_t config_hash, uint32_t n_probes, uint64_t nnz, uint32_t n_surfels,
                       uint32_t *range, uint32_t *ids, float *transfer, float *surfels, uint64_t *keys);

/* Volume_weight calculate_weight(Model&, ivec3 probe_res, ivec3 volume_res, vec3 scene_size) (light_probe.h:14-18,
 * light_probe.cpp:156-367): per voxel, the trilinear weights of its 8 surrounding probes masked by segment visibility and
 * renormalised, after moving "inside" voxels (score > 0.2 from 100 closest-hit rays) to their least-inside neighbour.
 * w0123 / w4567: [rx*ry*rz][4] floats, index (z*ry + y)*rx + x, corner order of the diagram at light_probe.cpp:269-294
 * (what SH_volume::set_visibility uploads, volume.cpp:318-333).  inside_score optional [rx*ry*rz] (NaN where no ray hit). */
int prt_volume_weights(prt_scene *, const int32_t probe_res[3], const int32_t volume_res[3], const float scene_size[3],
                       float *w0123, float *w4567, float *inside_score);

#ifdef __cplusplus
}
#endif
#endif
