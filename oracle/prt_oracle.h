/*
 * oracle/prt_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm for the PRT precomputation hot path
 * (lvjiahui/PRT).  It exists to CHECK the sm_100a kernels in
 * prt_b200/csrc; nothing in the product library links, includes or calls it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY STATUS: the spherical-harmonic basis and cubeCoordToWorld are pinned against the
 * reference's own header (src/sh/SH_function.h compiled in place, oracle/_ref, see
 * oracle/ref_shfun.cpp).  Everything that the reference delegates to components that are NOT in
 * /root/reference -- Intel Embree 3 (CMakeLists.txt:6; BVH + ray/triangle arithmetic),
 * google/spherical-harmonics + Eigen (raytracing.cpp:10; sh::EvalSH) and the OpenGL 4.5 driver
 * (raster rules, texture filtering) -- is "PARITY UNPINNED": the reference ships no tests, golden
 * vectors or fixtures, so those parts are anchored on the analytic
 * known-answer tests of SURVEY section 8c (tests/test_oracle_kat.py).
 */
#ifndef PRT_ORACLE_H
#define PRT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct prt_o_scene prt_o_scene;

/* RTScene(Mesh&) (raytracing.cpp:58-94): copy positions + index triples, build a BVH. */
prt_o_scene *prt_o_scene_create(const float *pos, size_t pos_stride_bytes, uint32_t n_verts,
                                const uint32_t *tri_idx, uint32_t n_tris);
void prt_o_scene_destroy(prt_o_scene *);
uint32_t prt_o_scene_ntris(const prt_o_scene *);

/* struct Ray::any_hit / first_hit (light_probe.cpp:95-133). use_bvh = 0 -> brute force over all
 * triangles (validates that the BVH never culls a triangle the pinned test accepts). */
int prt_o_any_hit(const prt_o_scene *, const float org[3], const float dir[3], float tnear, float tfar, int use_bvh);
/* returns 1 on hit; t, primitive index, unnormalised Ng = (v1-v0)x(v2-v0) */
int prt_o_closest_hit(const prt_o_scene *, const float org[3], const float dir[3], float tnear, float tfar,
                      int use_bvh, float *t, uint32_t *prim, float ng[3]);

enum { PRT_O_UNSHADOWED = 0, PRT_O_SHADOWED = 1, PRT_O_INTERREFLECT = 2, PRT_O_UNSHADOWED_ANALYTIC = 3 };

typedef struct {
    int32_t order;       /* bands: coefficients = order^2, 1..5 */
    int32_t samples_u;   /* radial strata  (reference: sh_resolution, app.h:71) */
    int32_t samples_v;   /* angular strata */
    uint32_t seed;
    int32_t bounces;     /* B: path depth = B+1 = max_path_length-1 (raytracing.cpp:345) */
    float albedo[3];     /* app.h:55 */
    float origin_eps;    /* 1e-4, raytracing.cpp:343 */
    float bounce_eps;    /* 1e-5, raytracing.cpp:235 */
    int32_t mode;
    int32_t cs_phase;
    int32_t jitter;      /* 1: (i+xi)/R jittered strata (raytracing.cpp:338-339); 0: stratum centres */
} prt_o_bake_params;

/* Sample table shared by every vertex and coefficient: uv[2*s..], local dirs[3*s..], s = i*samples_v + j */
void prt_o_sample_table(const prt_o_bake_params *, float *uv, float *local_dirs);

/* bake_SH (raytracing.cpp:320-360) with one trace per sample shared by all coefficients.
 * faithful != 0 re-traces once per coefficient (the reference's cost model; same results).
 * out_coeffs[n][order^2]; out_vis (optional) [n][ceil(S/32)] uint32 words, bit s = 1 <=> primary ray s is
 * UNOCCLUDED; n_threads <= 0 -> all cores.  counters (optional) [2] = {rays traced, path segments}. */
int prt_o_bake_transfer(const prt_o_scene *, const float *pos, const float *nrm, size_t stride_bytes,
                        uint32_t n_verts, uint32_t vertex_id_base, const prt_o_bake_params *,
                        float *out_coeffs, uint32_t *out_vis, int n_threads, int faithful, uint64_t *counters);

/* bake_SH in the reference's own loop order, consuming a given random() sequence (see bake.c); returns the values consumed */
uint64_t prt_o_bake_transfer_ref_order(const prt_o_scene *, const float *pos, const float *nrm, size_t stride_bytes, uint32_t n_verts, int order, int res,
                                       int max_path_length, const float albedo[3], int cs_phase, const float *rnd, uint64_t n_rnd, int u_first,
                                       uint32_t seed, uint32_t vertex_id_base, float *out_coeffs);

/* ---- image-based lighting (oracle/env.c); cube buffers: levels concatenated, [6][n][n][3] floats per level ---- */
size_t prt_o_cube_floats(int n0, int levels);
int prt_o_cube_levels(int n0);
void prt_o_cube_sample(const float *cube, int n0, int levels, const float d[3], float lod, float out[3]);
void prt_o_env_equirect_to_cube(const float *eq, int w, int h, int n0, int levels, float *cube);
void prt_o_equirect_uv(const float v[3], float uv[2]);
void prt_o_env_irradiance_dir(const float *cube, int n0, int levels, const float P[3], float out[3]);
void prt_o_env_prefilter_dir(const float *cube, int n0, int levels, const float P[3], float roughness, int n_samples, float out[3]);
void prt_o_env_irradiance(const float *cube, int n0, int levels, int n_out, float *out);
void prt_o_env_prefilter(const float *cube, int n0, int levels, int n_out, int mips, int n_samples, float *out);
void prt_o_brdf_lut(int w, int h, int n_samples, float *out);
void prt_o_env_project_sh(const float *cube, int n0, int levels, int order, int method, int size, float *out);
void prt_o_sh_pack_rh(const float *L, float *out28);

/* ---- probe capture / projection (oracle/probe.c) ---- */
typedef struct prt_o_csr prt_o_csr;
void prt_o_fibonacci_dirs(int n, float *dirs);
void prt_o_cube_dirs(int res, float *dirs, float *weights);
void prt_o_probe_positions(const int res[3], const float size[3], float *pos);
prt_o_csr *prt_o_probe_capture(const prt_o_scene *, const float *probe_pos, uint32_t n_probes, const float *dirs,
                               const float *weights, uint32_t n_dirs);
void prt_o_csr_sizes(const prt_o_csr *, uint32_t *nnz, uint32_t *n_prim);
void prt_o_csr_get(const prt_o_csr *, uint32_t *range, uint32_t *ids, float *transfer, float *surfels, uint64_t *keys);
void prt_o_csr_get_sums(const prt_o_csr *, double *sums /*[n_prim][7]: sum pos, sum normal, count*/);
void prt_o_csr_destroy(prt_o_csr *);
void prt_o_probe_project(const prt_o_csr *, const float *radiance_rgba, float *out);
void prt_o_project_arrays(const uint32_t *range, uint32_t n_probes, const uint32_t *ids, const float *transfer, const float *radiance_rgba, float *out);
/* calculate_weight (light_probe.cpp:156-367); w0123/w4567 [n_voxels][4], index (z*ry+y)*rx+x; score_out optional [n_voxels] */
void prt_o_volume_weights(const prt_o_scene *, const int probe_res[3], const int volume_res[3], const float scene_size[3],
                          float *w0123, float *w4567, float *score_out);

/* ---- per-frame probe pipeline (oracle/gi.c): sky shadow map, relight.comp, transfer2volume.comp ---- */
typedef struct {
    float cast_intensity[3], cast_position[3], cast_direction[3], cast_cutoff;   /* CastLight   (light.glsl:1-6)  */
    float ambient_intensity[3], ambient_position[3];                             /* PointLight  (light.glsl:8-11) */
    float sky_intensity[3], sky_direction[3];                                    /* ParalLight  (light.glsl:13-16) */
    float light_space_matrix[16];                                                /* column-major (glm) */
    int32_t multi_bounce;
    float atten, sh_shift, temp_weight;
} prt_o_relight_params;
void prt_o_paral_shadow_matrix(float up, float dir, float out_dir[3], float out_matrix[16]);
int prt_o_shadow_map(const prt_o_scene *, const float matrix[16], int size, float *depth);
void prt_o_relight(const prt_o_relight_params *, uint32_t n_surfels, const float *surfels, const float *albedo, const float *depth,
                   int shadow_size, const float *volumes, const int volume_res[3], const float scene_size[3], float *radiance);
void prt_o_transfer_to_volume(const float *probe_sh, const int probe_res[3], const float *w0123, const float *w4567,
                              const int volume_res[3], float *out);

/* ---- progressive AO / normal preview (oracle/raytrace.c; raytracing.cpp:162-222,280-317) ---- */
typedef struct { float position[3], front[3], up[3], right[3], zoom_deg; } prt_o_camera;
/* mode 0 = renderAO, 1 = renderNormal; accum [h*w][4] (sum rgb, count) is updated, pixels [h*w][4] RGBA8 written */
void prt_o_raytrace(const prt_o_scene *, const prt_o_camera *, int w, int h, int max_path_length, const float albedo[3], int gamma,
                    int mode, uint32_t seed, uint32_t frame, float *accum, uint8_t *pixels);

void prt_o_sh_eval(int order, int cs_phase, const float dir_sh[3], float *out);
void prt_o_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void prt_o_sincos2pi(float v, float *s, float *c);
/* frame(N) and cosineSampleHemisphere(u, v, N) (raytracing.cpp:101-107,130-160): local direction, world direction,
 * frame columns (right, up, N) and the pdf z / PI */
void prt_o_cosine_world(float u, float v, const float N[3], float local[3], float world[3], float frame9[9], float *pdf);
int prt_o_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
