// gi.cu -- sm_100a kernels + C ABI for the reference's per-frame probe pipeline:
//
//   Paral_Shadow::set_dir / render  src/opengl/gl.cpp:620-648         sky-light matrix; depth map = closest hit of one ray per texel centre
//   SH_volume::relight              src/sh/volume.cpp:357-386 -> relight.comp:68-82 (lights common/light.glsl:18-46, shadow lookup
//                                   common/paral_shadow.glsl:4-36, albedo colored_wall.glsl:3-11, SH feedback common/SH.glsl:17-36)
//   SH_volume::project_sh           src/sh/volume.cpp:388-452 -> precomp_projectSH.comp (probe.cu) + transfer2volume.comp:36-147
//
// The reference runs one relight -> project -> transfer2volume round per displayed frame (app.cpp:164-166) with everything resident
// in GL objects.  Here the state lives in a prt_gi object in HBM and prt_gi_step runs any number of rounds back to back on the
// context's stream with no host synchronisation in between.  Storage is FP32 (the reference's volumes are RGBA16F).  All arithmetic
// that the CPU oracle (oracle/gi.c) restates is explicitly rounded, so relight and transfer2volume are bit-identical to it.
#include "../../include/prt_b200.h"
#include "abi_internal.h"
#include "kernels.h"
#include "prt_math.cuh"

#include <cuda_runtime.h>
#include <math.h>
#include <vector>

using namespace prt;

struct prt_gi {
    prt_ctx *ctx = nullptr;
    const prt_csr *csr = nullptr;          // borrowed: must outlive the object
    int pres[3] = {0, 0, 0}, vres[3] = {0, 0, 0};
    float scene_size[3] = {0, 0, 0};
    size_t n_vox = 0;
    float4 *w0123 = nullptr, *w4567 = nullptr, *radiance = nullptr, *probe_sh = nullptr, *volumes = nullptr;
    float *albedo = nullptr, *depth = nullptr;
    int shadow_size = 0;
    unsigned long long rounds = 0;
};

namespace {

struct RelightArgs {
    prt_relight_params P;
    const float *surfels, *albedo, *depth;
    const float4 *volumes;
    float4 *radiance;
    uint32_t n;
    int shadow_size, vres[3];
    float scene_size[3];
    int feedback;
};

__device__ __forceinline__ float dot3p(const float a0, const float a1, const float a2, const float b0, const float b1, const float b2) {
    return PRT_FMA(a2, b2, PRT_FMA(a1, b1, PRT_MUL(a0, b0)));
}

// paral_shadow.glsl:4-36 with a nearest, clamp-to-border(1.0) depth texture (gl.cpp:603-607)
__device__ __forceinline__ float shadow_calc(const RelightArgs &A, const float pos[3], const float n[3]) {
    if (!A.depth) return 0.0f;
    const float *m = A.P.light_space_matrix;
    float p[3];
#pragma unroll
    for (int r = 0; r < 3; r++) p[r] = PRT_FMA(m[r], pos[0], PRT_FMA(m[4 + r], pos[1], PRT_FMA(m[8 + r], pos[2], m[12 + r])));
    const float w = PRT_FMA(m[3], pos[0], PRT_FMA(m[7], pos[1], PRT_FMA(m[11], pos[2], m[15])));
#pragma unroll
    for (int r = 0; r < 3; r++) p[r] = PRT_FMA(PRT_DIV(p[r], w), 0.5f, 0.5f);
    float closest = 1.0f;
    if (p[0] >= 0.0f && p[0] < 1.0f && p[1] >= 0.0f && p[1] < 1.0f) {
        const int i = min((int)PRT_MUL(p[0], (float)A.shadow_size), A.shadow_size - 1), j = min((int)PRT_MUL(p[1], (float)A.shadow_size), A.shadow_size - 1);
        closest = __ldg(&A.depth[(size_t)j * A.shadow_size + i]);
    }
    const float ndl = dot3p(n[0], n[1], n[2], A.P.sky_direction[0], A.P.sky_direction[1], A.P.sky_direction[2]);
    const float bias = PRT_MUL(0.1f, fmaxf(PRT_MUL(0.05f, PRT_SUB(1.0f, ndl)), 0.005f));
    float shadow = (PRT_SUB(p[2], bias) > closest) ? 1.0f : 0.0f;
    if (p[2] > 1.0f) shadow = 0.0f;
    return shadow;
}

// SH.glsl:17-36: trilinear fetch (GL_LINEAR, CLAMP_TO_EDGE; volume.cpp:33-41) of the 7 packed volumes + the R-H polynomial
__device__ __forceinline__ void sh_irad(const RelightArgs &A, const float n[3], const float pos[3], float out[3]) {
    const float N[4] = {n[2], n[0], n[1], 1.0f};
    int i0[3], i1[3];
    float f[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float c = PRT_DIV(PRT_SUB(pos[a], PRT_MUL(-1.0f, A.scene_size[a])), PRT_MUL(2.0f, A.scene_size[a]));
        const float u = PRT_FMA(c, (float)A.vres[a], -0.5f), fl = floorf(u);
        f[a] = PRT_SUB(u, fl);
        const int i = (int)fl;
        i0[a] = min(max(i, 0), A.vres[a] - 1);
        i1[a] = min(max(i + 1, 0), A.vres[a] - 1);
    }
    float t[28];
#pragma unroll
    for (int k = 0; k < 28; k++) t[k] = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
        const float w = PRT_MUL(PRT_MUL(dx ? f[0] : PRT_SUB(1.0f, f[0]), dy ? f[1] : PRT_SUB(1.0f, f[1])), dz ? f[2] : PRT_SUB(1.0f, f[2]));
        const size_t vox = ((size_t)(dz ? i1[2] : i0[2]) * A.vres[1] + (dy ? i1[1] : i0[1])) * A.vres[0] + (dx ? i1[0] : i0[0]);
#pragma unroll
        for (int q = 0; q < 7; q++) {
            const float4 v = __ldg(&A.volumes[7 * vox + q]);
            t[4 * q] = PRT_FMA(w, v.x, t[4 * q]); t[4 * q + 1] = PRT_FMA(w, v.y, t[4 * q + 1]);
            t[4 * q + 2] = PRT_FMA(w, v.z, t[4 * q + 2]); t[4 * q + 3] = PRT_FMA(w, v.w, t[4 * q + 3]);
        }
    }
    const float BN[4] = {PRT_MUL(N[0], N[1]), PRT_MUL(N[0], N[2]), PRT_MUL(N[1], N[2]), PRT_MUL(N[2], N[2])};
    const float cc = PRT_FMA(N[0], N[0], -PRT_MUL(N[1], N[1]));
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float *Aq = t + 4 * ch, *Bq = t + 4 * (3 + ch);
        const float x = PRT_FMA(Aq[3], N[3], PRT_FMA(Aq[2], N[2], PRT_FMA(Aq[1], N[1], PRT_MUL(Aq[0], N[0]))));
        const float y = PRT_FMA(Bq[3], BN[3], PRT_FMA(Bq[2], BN[2], PRT_FMA(Bq[1], BN[1], PRT_MUL(Bq[0], BN[0]))));
        const float z = PRT_MUL(t[24 + ch], cc);
        out[ch] = fmaxf(PRT_ADD(PRT_ADD(x, y), z), 0.0f);
    }
}

__device__ __forceinline__ float len3p(const float v[3]) { return PRT_SQRT(PRT_FMA(v[2], v[2], PRT_FMA(v[1], v[1], PRT_MUL(v[0], v[0])))); }

// relight.comp:68-82, one thread per surfel
__global__ void __launch_bounds__(128) relight_kernel(const RelightArgs A) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    const prt_relight_params &P = A.P;
    float pos[3], N[3], alb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { pos[k] = __ldg(&A.surfels[6 * (size_t)s + k]); N[k] = __ldg(&A.surfels[6 * (size_t)s + 3 + k]); }
    if (A.albedo) {
#pragma unroll
        for (int k = 0; k < 3; k++) alb[k] = __ldg(&A.albedo[3 * (size_t)s + k]);
    } else {                                                              // colored_wall.glsl:3-11
        alb[0] = alb[1] = alb[2] = 0.4f;
        if (pos[0] > 5.9f) {
            const int sel = ((int)PRT_ADD(PRT_DIV(pos[1], 6.0f), 100.0f) + (int)PRT_ADD(PRT_DIV(pos[2], 6.0f), 100.0f)) % 3;
#pragma unroll
            for (int k = 0; k < 3; k++) alb[k] = (k == sel) ? PRT_ADD(0.1f, 0.7f) : 0.1f;
        }
    }
    const float shadow = shadow_calc(A, pos, N);
    const float sky_cos = fmaxf(dot3p(P.sky_direction[0], P.sky_direction[1], P.sky_direction[2], N[0], N[1], N[2]), 0.0f);   // light.glsl:43-46
    float cast[3] = {0.f, 0.f, 0.f};
    {                                                                     // light.glsl:18-31
        const float d[3] = {PRT_SUB(P.cast_position[0], pos[0]), PRT_SUB(P.cast_position[1], pos[1]), PRT_SUB(P.cast_position[2], pos[2])};
        const float dist = len3p(d), inv = PRT_DIV(1.0f, dist);
        const float ld[3] = {PRT_MUL(d[0], inv), PRT_MUL(d[1], inv), PRT_MUL(d[2], inv)};
        const float nd[3] = {-P.cast_direction[0], -P.cast_direction[1], -P.cast_direction[2]};
        const float ninv = PRT_DIV(1.0f, len3p(nd));
        const float theta = PRT_FMA(ld[2], PRT_MUL(nd[2], ninv), PRT_FMA(ld[1], PRT_MUL(nd[1], ninv), PRT_MUL(ld[0], PRT_MUL(nd[0], ninv))));
        if (theta > P.cast_cutoff) {
            const float icos = fmaxf(dot3p(ld[0], ld[1], ld[2], N[0], N[1], N[2]), 0.0f);
            const float soft = PRT_DIV(PRT_SUB(theta, P.cast_cutoff), PRT_SUB(1.0f, P.cast_cutoff));
#pragma unroll
            for (int k = 0; k < 3; k++) cast[k] = PRT_DIV(PRT_MUL(PRT_MUL(soft, P.cast_intensity[k]), icos), PRT_MUL(dist, dist));
        }
    }
    float amb[3] = {0.f, 0.f, 0.f};
    if (P.ambient_intensity[0] != 0.f || P.ambient_intensity[1] != 0.f || P.ambient_intensity[2] != 0.f) {                     // light.glsl:33-41
        const float d[3] = {PRT_SUB(P.ambient_position[0], pos[0]), PRT_SUB(P.ambient_position[1], pos[1]), PRT_SUB(P.ambient_position[2], pos[2])};
        const float dist = len3p(d), inv = PRT_DIV(1.0f, dist);
        const float icos = fmaxf(PRT_FMA(PRT_MUL(d[2], inv), N[2], PRT_FMA(PRT_MUL(d[1], inv), N[1], PRT_MUL(PRT_MUL(d[0], inv), N[0]))), 0.0f);
#pragma unroll
        for (int k = 0; k < 3; k++) amb[k] = PRT_DIV(PRT_MUL(P.ambient_intensity[k], icos), PRT_MUL(dist, dist));
    }
    float irr[3] = {0.f, 0.f, 0.f};
    if (A.feedback) {
        const float q[3] = {PRT_FMA(P.sh_shift, N[0], pos[0]), PRT_FMA(P.sh_shift, N[1], pos[1]), PRT_FMA(P.sh_shift, N[2], pos[2])};
        sh_irad(A, N, q, irr);
    }
    const float4 old = A.radiance[s];
    const float o[3] = {old.x, old.y, old.z};
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float rad = PRT_MUL(alb[k], PRT_ADD(PRT_ADD(PRT_MUL(PRT_SUB(1.0f, shadow), PRT_MUL(P.sky_intensity[k], sky_cos)), cast[k]), amb[k]));
        if (A.feedback) rad = PRT_ADD(rad, PRT_DIV(PRT_MUL(PRT_MUL(alb[k], P.atten), irr[k]), kPiF));
        r[k] = PRT_FMA(P.temp_weight, rad, PRT_MUL(PRT_SUB(1.0f, P.temp_weight), o[k]));
    }
    A.radiance[s] = make_float4(r[0], r[1], r[2], 1.0f);
}

struct VolArgs { int pres[3], vres[3]; const float4 *probe_sh, *w0123, *w4567; float4 *out; };

// transfer2volume.comp:36-147: one thread per (voxel, packed volume q)
__global__ void __launch_bounds__(256) transfer_to_volume_kernel(const VolArgs A, const size_t n_vox) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t v = gid / 7;
    const int q = (int)(gid % 7);
    if (v >= n_vox) return;
    const int id[3] = {(int)(v % A.vres[0]), (int)((v / A.vres[0]) % A.vres[1]), (int)(v / ((size_t)A.vres[0] * A.vres[1]))};
    int anchor[3];
#pragma unroll
    for (int a = 0; a < 3; a++)
        anchor[a] = (int)floorf(PRT_SUB(PRT_MUL(PRT_DIV(PRT_ADD((float)id[a], 0.5f), (float)A.vres[a]), (float)A.pres[a]), 0.5f));
    const float4 wa = __ldg(&A.w0123[v]), wb = __ldg(&A.w4567[v]);
    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    const int off[8][3] = {{0, 0, 1}, {1, 0, 1}, {1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int px = anchor[0] + off[c][0], py = anchor[1] + off[c][1], pz = anchor[2] + off[c][2];
        if (px < 0 || py < 0 || pz < 0 || px >= A.pres[0] || py >= A.pres[1] || pz >= A.pres[2]) continue;
        const float4 s = __ldg(&A.probe_sh[7 * (((size_t)pz * A.pres[1] + py) * A.pres[0] + px) + q]);
        acc.x = PRT_FMA(w[c], s.x, acc.x); acc.y = PRT_FMA(w[c], s.y, acc.y); acc.z = PRT_FMA(w[c], s.z, acc.z); acc.w = PRT_FMA(w[c], s.w, acc.w);
    }
    A.out[7 * v + q] = acc;
}

// Paral_Shadow::render (gl.cpp:633-648) as one closest-hit ray per texel centre
__global__ void shadow_rays_kernel(float *rays, int size, f3 O, f3 U, f3 V, f3 D) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)size * size) return;
    const int i = (int)(idx % size), j = (int)(idx / size);
    const float x = PRT_SUB(PRT_MUL(PRT_DIV(PRT_ADD((float)i, 0.5f), (float)size), 2.0f), 1.0f);
    const float y = PRT_SUB(PRT_MUL(PRT_DIV(PRT_ADD((float)j, 0.5f), (float)size), 2.0f), 1.0f);
    float *r = rays + 8 * idx;
    r[0] = PRT_FMA(y, V.x, PRT_FMA(x, U.x, O.x)); r[1] = PRT_FMA(y, V.y, PRT_FMA(x, U.y, O.y)); r[2] = PRT_FMA(y, V.z, PRT_FMA(x, U.z, O.z));
    r[3] = 0.0f; r[4] = D.x; r[5] = D.y; r[6] = D.z; r[7] = 1.0f;
}
__global__ void shadow_depth_kernel(const float *t, const uint32_t *prim, float *depth, size_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) depth[idx] = prim[idx] != 0xFFFFFFFFu ? t[idx] : 1.0f;
}

int light_rays(const float m[16], double O[3], double U[3], double V[3], double D[3]) {
    if (m[3] != 0.f || m[7] != 0.f || m[11] != 0.f || m[15] != 1.f) return -1;
    double a[3][3], inv[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = m[4 * c + r];
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (det == 0.0) return -1;
    inv[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det; inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det; inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
    inv[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det; inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det; inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
    inv[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det; inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det; inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
    const double rhs[3] = {0.0 - m[12], 0.0 - m[13], -1.0 - m[14]};
    for (int r = 0; r < 3; r++) {
        O[r] = inv[r][0] * rhs[0] + inv[r][1] * rhs[1] + inv[r][2] * rhs[2];
        U[r] = inv[r][0]; V[r] = inv[r][1]; D[r] = 2.0 * inv[r][2];
    }
    return 0;
}

}  // namespace

#define GI_TRY(expr)                                                                                                    \
    do {                                                                                                                \
        cudaError_t e_ = (expr);                                                                                        \
        if (e_ != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int prt_paral_shadow_matrix(float up, float dir, float out_direction[3], float out_matrix[16]) {
    if (!out_direction || !out_matrix) return prt_set_error(PRT_ERR_INVALID, "prt_paral_shadow_matrix: null argument");
    const double PI = 3.14159265359;                                   // util.h:6
    const double theta = PI * (double)up, phi = 2.0 * PI * (double)dir;
    const double d[3] = {sin(theta) * sin(phi), cos(theta), sin(theta) * cos(phi)};                 // gl.cpp:623-625
    double upv[3] = {0.0, 1.0, 0.0};
    if (up < 0.1f || up > 0.9f) { upv[1] = 0.0; upv[2] = 1.0; }                                      // gl.cpp:628
    const double eye[3] = {30.0 * d[0], 30.0 * d[1], 30.0 * d[2]};
    double f[3] = {-eye[0], -eye[1], -eye[2]};
    const double fl = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int k = 0; k < 3; k++) f[k] /= fl;
    double s[3] = {f[1] * upv[2] - f[2] * upv[1], f[2] * upv[0] - f[0] * upv[2], f[0] * upv[1] - f[1] * upv[0]};
    const double sl = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int k = 0; k < 3; k++) s[k] /= sl;
    const double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    const double view[4][4] = {{s[0], u[0], -f[0], 0}, {s[1], u[1], -f[1], 0}, {s[2], u[2], -f[2], 0},
                               {-(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2]), -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]),
                                f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2], 1}};                  // glm::lookAt, [column][row]
    const double l = -30, r = 30, b = -30, t = 30, n = 0.1, fa = 60;                                 // gl.cpp:593,626
    const double proj[4][4] = {{2 / (r - l), 0, 0, 0}, {0, 2 / (t - b), 0, 0}, {0, 0, -2 / (fa - n), 0},
                               {-(r + l) / (r - l), -(t + b) / (t - b), -(fa + n) / (fa - n), 1}};   // glm::ortho
    for (int c = 0; c < 4; c++)
        for (int rr = 0; rr < 4; rr++) {
            double a = 0;
            for (int k = 0; k < 4; k++) a += proj[k][rr] * view[c][k];
            out_matrix[4 * c + rr] = (float)a;
        }
    for (int k = 0; k < 3; k++) out_direction[k] = (float)d[k];
    return PRT_OK;
}

int prt_shadow_map(prt_scene *scene, const float matrix[16], int32_t size, float *out_depth) {
    if (!scene || !matrix || !out_depth || size <= 0 || size > 16384) return prt_set_error(PRT_ERR_INVALID, "prt_shadow_map: bad argument");
    double O[3], U[3], V[3], D[3];
    if (light_rays(matrix, O, U, V, D)) return prt_set_error(PRT_ERR_UNSUPPORTED, "prt_shadow_map: the light matrix must be affine and invertible (orthographic light)");
    const prt_scene_view sv = prt_scene_get_view(scene);
    GI_TRY(cudaSetDevice(prt_ctx_device(sv.ctx)));
    cudaStream_t st = prt_ctx_stream(sv.ctx);
    const size_t n = (size_t)size * size;
    float *rays = nullptr, *t = nullptr, *depth = nullptr; uint32_t *prim = nullptr;
    cudaError_t e = cudaMalloc(&rays, 32 * n);
    if (e == cudaSuccess) e = cudaMalloc(&t, 4 * n);
    if (e == cudaSuccess) e = cudaMalloc(&prim, 4 * n);
    if (e == cudaSuccess) e = cudaMalloc(&depth, 4 * n);
    if (e == cudaSuccess) {
        const unsigned g = (unsigned)((n + 255) / 256);
        shadow_rays_kernel<<<g, 256, 0, st>>>(rays, size, mk3((float)O[0], (float)O[1], (float)O[2]), mk3((float)U[0], (float)U[1], (float)U[2]),
                                               mk3((float)V[0], (float)V[1], (float)V[2]), mk3((float)D[0], (float)D[1], (float)D[2]));
        e = launch_trace_closest(sv.nodes, sv.tris, rays, (uint32_t)n, t, prim, nullptr, st);
        if (e == cudaSuccess) { shadow_depth_kernel<<<g, 256, 0, st>>>(t, prim, depth, n); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_depth, depth, 4 * n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(rays); cudaFree(t); cudaFree(prim); cudaFree(depth);
    if (e != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string("prt_shadow_map: ") + cudaGetErrorString(e));
    return PRT_OK;
}

void prt_gi_destroy(prt_gi *g) {
    if (!g) return;
    cudaSetDevice(prt_ctx_device(g->ctx));
    cudaFree(g->w0123); cudaFree(g->w4567); cudaFree(g->radiance); cudaFree(g->probe_sh); cudaFree(g->volumes); cudaFree(g->albedo); cudaFree(g->depth);
    delete g;
}

int prt_gi_create(const prt_csr *csr, const int32_t probe_res[3], const int32_t volume_res[3], const float scene_size[3],
                  const float *w0123, const float *w4567, prt_gi **out) {
    if (!csr || !probe_res || !volume_res || !scene_size || !w0123 || !w4567 || !out) return prt_set_error(PRT_ERR_INVALID, "prt_gi_create: null argument");
    *out = nullptr;
    for (int a = 0; a < 3; a++)
        if (probe_res[a] <= 0 || volume_res[a] <= 0 || !(scene_size[a] > 0.f)) return prt_set_error(PRT_ERR_INVALID, "prt_gi_create: resolutions and scene size must be positive");
    if ((unsigned long long)probe_res[0] * probe_res[1] * probe_res[2] != csr->n_probes)
        return prt_set_error(PRT_ERR_INVALID, "prt_gi_create: probe_res does not match the number of captured probes");
    GI_TRY(cudaSetDevice(prt_ctx_device(csr->ctx)));
    prt_gi *g = new prt_gi();
    g->ctx = csr->ctx; g->csr = csr;
    for (int a = 0; a < 3; a++) { g->pres[a] = probe_res[a]; g->vres[a] = volume_res[a]; g->scene_size[a] = scene_size[a]; }
    g->n_vox = (size_t)volume_res[0] * volume_res[1] * volume_res[2];
    const size_t np = std::max<size_t>(1, csr->n_prim);
    cudaError_t e = cudaMalloc(&g->w0123, 16 * g->n_vox);
    if (e == cudaSuccess) e = cudaMalloc(&g->w4567, 16 * g->n_vox);
    if (e == cudaSuccess) e = cudaMalloc(&g->radiance, 16 * np);
    if (e == cudaSuccess) e = cudaMalloc(&g->probe_sh, 112 * (size_t)csr->n_probes);
    if (e == cudaSuccess) e = cudaMalloc(&g->volumes, 112 * g->n_vox);
    if (e == cudaSuccess) e = cudaMemcpy(g->w0123, w0123, 16 * g->n_vox, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->w4567, w4567, 16 * g->n_vox, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(g->radiance, 0, 16 * np);
    if (e == cudaSuccess) e = cudaMemset(g->probe_sh, 0, 112 * (size_t)csr->n_probes);
    if (e == cudaSuccess) e = cudaMemset(g->volumes, 0, 112 * g->n_vox);
    if (e != cudaSuccess) { prt_gi_destroy(g); return prt_set_error(PRT_ERR_CUDA, std::string("prt_gi_create: ") + cudaGetErrorString(e)); }
    *out = g;
    return PRT_OK;
}

int prt_gi_set_shadow_map(prt_gi *g, const float *depth, int32_t size) {
    if (!g || (depth && (size <= 0 || size > 16384))) return prt_set_error(PRT_ERR_INVALID, "prt_gi_set_shadow_map: bad argument");
    GI_TRY(cudaSetDevice(prt_ctx_device(g->ctx)));
    GI_TRY(cudaStreamSynchronize(prt_ctx_stream(g->ctx)));
    cudaFree(g->depth); g->depth = nullptr; g->shadow_size = 0;
    if (!depth) return PRT_OK;
    GI_TRY(cudaMalloc(&g->depth, 4 * (size_t)size * size));
    GI_TRY(cudaMemcpy(g->depth, depth, 4 * (size_t)size * size, cudaMemcpyHostToDevice));
    g->shadow_size = size;
    return PRT_OK;
}

int prt_gi_set_albedo(prt_gi *g, const float *albedo_rgb) {
    if (!g) return prt_set_error(PRT_ERR_INVALID, "prt_gi_set_albedo: null argument");
    GI_TRY(cudaSetDevice(prt_ctx_device(g->ctx)));
    GI_TRY(cudaStreamSynchronize(prt_ctx_stream(g->ctx)));
    cudaFree(g->albedo); g->albedo = nullptr;
    if (!albedo_rgb || g->csr->n_prim == 0) return PRT_OK;
    GI_TRY(cudaMalloc(&g->albedo, 12 * (size_t)g->csr->n_prim));
    GI_TRY(cudaMemcpy(g->albedo, albedo_rgb, 12 * (size_t)g->csr->n_prim, cudaMemcpyHostToDevice));
    return PRT_OK;
}

int prt_gi_set_radiance(prt_gi *g, const float *radiance_rgba) {
    if (!g || !radiance_rgba) return prt_set_error(PRT_ERR_INVALID, "prt_gi_set_radiance: null argument");
    GI_TRY(cudaSetDevice(prt_ctx_device(g->ctx)));
    GI_TRY(cudaStreamSynchronize(prt_ctx_stream(g->ctx)));
    if (g->csr->n_prim) GI_TRY(cudaMemcpy(g->radiance, radiance_rgba, 16 * (size_t)g->csr->n_prim, cudaMemcpyHostToDevice));
    return PRT_OK;
}

int prt_gi_step(prt_gi *g, const prt_relight_params *P, int32_t n_rounds) {
    if (!g || !P || n_rounds < 0) return prt_set_error(PRT_ERR_INVALID, "prt_gi_step: bad argument");
    GI_TRY(cudaSetDevice(prt_ctx_device(g->ctx)));
    cudaStream_t st = prt_ctx_stream(g->ctx);
    RelightArgs A{};
    A.P = *P;
    A.surfels = g->csr->surfels; A.albedo = g->albedo; A.depth = g->depth; A.volumes = g->volumes; A.radiance = g->radiance;
    A.n = g->csr->n_prim; A.shadow_size = g->shadow_size; A.feedback = P->multi_bounce ? 1 : 0;
    VolArgs Vv{};
    for (int a = 0; a < 3; a++) { A.vres[a] = g->vres[a]; A.scene_size[a] = g->scene_size[a]; Vv.pres[a] = g->pres[a]; Vv.vres[a] = g->vres[a]; }
    Vv.probe_sh = g->probe_sh; Vv.w0123 = g->w0123; Vv.w4567 = g->w4567; Vv.out = g->volumes;
    for (int32_t it = 0; it < n_rounds; it++) {                                       // app.cpp:164-166: relight(); project_sh();
        if (A.n) relight_kernel<<<(A.n + 127) / 128, 128, 0, st>>>(A);
        GI_TRY(prt_csr_project_device(g->csr, g->radiance, g->probe_sh, st));
        transfer_to_volume_kernel<<<(unsigned)((g->n_vox * 7 + 255) / 256), 256, 0, st>>>(Vv, g->n_vox);
        GI_TRY(cudaGetLastError());
    }
    GI_TRY(cudaStreamSynchronize(st));
    g->rounds += (unsigned long long)n_rounds;
    return PRT_OK;
}

int prt_gi_download(const prt_gi *g, float *radiance_rgba, float *probe_sh, float *volumes) {
    if (!g) return prt_set_error(PRT_ERR_INVALID, "prt_gi_download: null argument");
    GI_TRY(cudaSetDevice(prt_ctx_device(g->ctx)));
    if (radiance_rgba && g->csr->n_prim) GI_TRY(cudaMemcpy(radiance_rgba, g->radiance, 16 * (size_t)g->csr->n_prim, cudaMemcpyDeviceToHost));
    if (probe_sh) GI_TRY(cudaMemcpy(probe_sh, g->probe_sh, 112 * (size_t)g->csr->n_probes, cudaMemcpyDeviceToHost));
    if (volumes) GI_TRY(cudaMemcpy(volumes, g->volumes, 112 * g->n_vox, cudaMemcpyDeviceToHost));
    return PRT_OK;
}

}  // extern "C"
