// env.cu -- sm_100a kernels for the image-based-lighting passes (BASELINE config 2) and their C ABI.
//
//   prt_env_create        equirect -> 6 x N^2 cube + 2x2-box mip chain   reference src/shaders/rectangle2cube.frag:7-22,
//                                                                         LightProbe::equirectangular_to_cubemap gl.cpp:581-591,
//                                                                         CubeMap::generateMipmap gl.cpp:454-460
//   prt_env_irradiance    cosine convolution, 252 x 63 samples / texel   src/shaders/irradiance.frag:9-43, gl.cpp:569-579
//   prt_env_prefilter     GGX prefilter, 1024 Hammersley samples / texel  src/shaders/prefilter.frag:10-107, gl.cpp:546-567
//   prt_brdf_lut          split-sum BRDF LUT                              src/shaders/brdf.frag:9-113, app.cpp:61-63
//   prt_env_project_sh    env cube -> SH (lat-long / cube-texel rules)    src/shaders/bak/projectSH.comp:63-151,
//                                                                         src/shaders/bak/image_projectSH.comp:64-133
//
// Texture semantics the GL driver leaves open are pinned exactly as in oracle/env.c (FP32 storage, major-axis face
// selection, bilinear with centres at (i+0.5)/N, off-face taps re-projected to the neighbouring face, trilinear = lerp of two
// levels).  The cube lives in HBM as float4 texels (one aligned 16-byte load per tap; 512^2 x 6 + mips = 33.5 MB, L2-resident).
#include "../../include/prt_b200.h"
#include "abi_internal.h"

#include <cuda_runtime.h>
#include <math.h>
#include <vector>

namespace {

constexpr float kPi = 3.14159265359f;   // the shaders' PI

struct CubeView {
    const float4 *base;
    int n0, levels;
};
__host__ __device__ inline size_t level_off(int n0, int l) {
    size_t o = 0;
    for (int i = 0; i < l; i++) { size_t n = (size_t)(n0 >> i); o += 6 * n * n; }
    return o;
}

__device__ __forceinline__ float3 face_dir(int f, float u, float v) {
    switch (f) {
    case 0: return make_float3(1.f, -v, -u);
    case 1: return make_float3(-1.f, -v, u);
    case 2: return make_float3(u, 1.f, v);
    case 3: return make_float3(u, -1.f, -v);
    case 4: return make_float3(u, -v, 1.f);
    default: return make_float3(-u, -v, -1.f);
    }
}
__device__ __forceinline__ void dir_face(float3 d, int &f, float &s, float &t) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    float ma, sc, tc;
    if (ax >= ay && ax >= az) { ma = ax; if (d.x >= 0) { f = 0; sc = -d.z; tc = -d.y; } else { f = 1; sc = d.z; tc = -d.y; } }
    else if (ay >= az) { ma = ay; if (d.y >= 0) { f = 2; sc = d.x; tc = d.z; } else { f = 3; sc = d.x; tc = -d.z; } }
    else { ma = az; if (d.z >= 0) { f = 4; sc = d.x; tc = -d.y; } else { f = 5; sc = -d.x; tc = -d.y; } }
    s = 0.5f * (sc / ma + 1.0f);
    t = 0.5f * (tc / ma + 1.0f);
}
__device__ __forceinline__ float4 texel(const float4 *lvl, int n, int f, int i, int j) {
    if (i < 0 || j < 0 || i >= n || j >= n) {
        const float3 d = face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n - 1.0f);
        float s, t;
        dir_face(d, f, s, t);
        i = min(n - 1, max(0, (int)floorf(s * (float)n)));
        j = min(n - 1, max(0, (int)floorf(t * (float)n)));
    }
    return __ldg(&lvl[(size_t)f * n * n + (size_t)j * n + i]);
}
__device__ __forceinline__ float3 cube_bilinear(const CubeView &c, int l, float3 d) {
    int f; float s, t;
    dir_face(d, f, s, t);
    const int n = c.n0 >> l;
    const float4 *lvl = c.base + level_off(c.n0, l);
    const float x = s * (float)n - 0.5f, y = t * (float)n - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
    const int i = (int)x0, j = (int)y0;
    const float4 a = texel(lvl, n, f, i, j), b = texel(lvl, n, f, i + 1, j), cc = texel(lvl, n, f, i, j + 1), e = texel(lvl, n, f, i + 1, j + 1);
    float3 r;
    { const float top = a.x + fx * (b.x - a.x), bot = cc.x + fx * (e.x - cc.x); r.x = top + fy * (bot - top); }
    { const float top = a.y + fx * (b.y - a.y), bot = cc.y + fx * (e.y - cc.y); r.y = top + fy * (bot - top); }
    { const float top = a.z + fx * (b.z - a.z), bot = cc.z + fx * (e.z - cc.z); r.z = top + fy * (bot - top); }
    return r;
}
__device__ __forceinline__ float3 cube_sample(const CubeView &c, float3 d, float lod) {
    if (!(lod > 0.0f)) lod = 0.0f;
    if (lod > (float)(c.levels - 1)) lod = (float)(c.levels - 1);
    const int l0 = (int)floorf(lod);
    const float f = lod - (float)l0;
    float3 r = cube_bilinear(c, l0, d);
    if (f > 0.0f && l0 + 1 < c.levels) {
        const float3 r1 = cube_bilinear(c, l0 + 1, d);
        r.x += f * (r1.x - r.x); r.y += f * (r1.y - r.y); r.z += f * (r1.z - r.z);
    }
    return r;
}
__device__ __forceinline__ float3 normalize3f(float3 v) {
    const float inv = 1.0f / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return make_float3(v.x * inv, v.y * inv, v.z * inv);
}
__device__ __forceinline__ void tangent_frame(float3 N, float thr, float3 &right, float3 &up) {
    const float3 u0 = fabsf(N.z) < thr ? make_float3(0.f, 0.f, 1.f) : make_float3(1.f, 0.f, 0.f);
    right = normalize3f(make_float3(u0.y * N.z - u0.z * N.y, u0.z * N.x - u0.x * N.z, u0.x * N.y - u0.y * N.x));
    up = make_float3(N.y * right.z - N.z * right.y, N.z * right.x - N.x * right.z, N.x * right.y - N.y * right.x);
}

// ---- equirect -> cube, mips ---------------------------------------------------------------------------------------------
__global__ void equirect_to_cube_kernel(const float *eq, int w, int h, int n0, float4 *cube) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)6 * n0 * n0) return;
    const int i = (int)(idx % n0), j = (int)((idx / n0) % n0), f = (int)(idx / ((size_t)n0 * n0));
    const float3 v = normalize3f(face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n0 - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n0 - 1.0f));
    const float u = atan2f(-v.x, -v.z) * 0.1591f + 0.5f;                      // rectangle2cube.frag:9-14
    const float vv = acosf(fminf(1.0f, fmaxf(-1.0f, v.y))) * 0.3183f;
    const float x = u * (float)w - 0.5f, y = vv * (float)h - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
    const int i0 = min(w - 1, max(0, (int)x0)), i1 = min(w - 1, max(0, (int)x0 + 1));
    const int j0 = min(h - 1, max(0, (int)y0)), j1 = min(h - 1, max(0, (int)y0 + 1));
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = eq[((size_t)j0 * w + i0) * 3 + k], b = eq[((size_t)j0 * w + i1) * 3 + k];
        const float c = eq[((size_t)j1 * w + i0) * 3 + k], e = eq[((size_t)j1 * w + i1) * 3 + k];
        const float top = a + fx * (b - a), bot = c + fx * (e - c);
        o[k] = top + fy * (bot - top);
    }
    cube[idx] = make_float4(o[0], o[1], o[2], 1.0f);
}
__global__ void cube_downsample_kernel(const float4 *src, int np, float4 *dst) {
    const int n = np >> 1;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)6 * n * n) return;
    const int i = (int)(idx % n), j = (int)((idx / n) % n), f = (int)(idx / ((size_t)n * n));
    const float4 *p = src + (size_t)f * np * np + (size_t)(2 * j) * np + 2 * i;
    const float4 a = p[0], b = p[1], c = p[np], e = p[np + 1];
    dst[idx] = make_float4(0.25f * ((a.x + b.x) + (c.x + e.x)), 0.25f * ((a.y + b.y) + (c.y + e.y)), 0.25f * ((a.z + b.z) + (c.z + e.z)), 1.0f);
}
__global__ void cube_to_rgb_kernel(const float4 *src, size_t n, float *dst) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float4 v = src[idx];
    dst[3 * idx] = v.x; dst[3 * idx + 1] = v.y; dst[3 * idx + 2] = v.z;
}

__device__ __forceinline__ float3 block_sum3(float3 v, float *sm) {
    // 128-thread block reduction (warp shuffles + one shared-memory stage)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xFFFFFFFFu, v.x, o); v.y += __shfl_xor_sync(0xFFFFFFFFu, v.y, o); v.z += __shfl_xor_sync(0xFFFFFFFFu, v.z, o);
    }
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { sm[3 * w] = v.x; sm[3 * w + 1] = v.y; sm[3 * w + 2] = v.z; }
    __syncthreads();
    float3 r = make_float3(0.f, 0.f, 0.f);
    for (int k = 0; k < nw; k++) { r.x += sm[3 * k]; r.y += sm[3 * k + 1]; r.z += sm[3 * k + 2]; }
    __syncthreads();
    return r;
}

// ---- irradiance: one CTA per output texel, 252 x 63 tangent-space samples (irradiance.frag:25-40) ------------------------
__global__ void __launch_bounds__(128) irradiance_kernel(CubeView c, int n_out, const float *phis, int n_phi, const float *thetas, int n_theta, float *out) {
    __shared__ float sm[12];
    const int idx = blockIdx.x;
    const int i = idx % n_out, j = (idx / n_out) % n_out, f = idx / (n_out * n_out);
    const float3 N = normalize3f(face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n_out - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n_out - 1.0f));
    float3 right, up;
    tangent_frame(N, 0.99f, right, up);
    float3 acc = make_float3(0.f, 0.f, 0.f);
    const int total = n_phi * n_theta;
    for (int s = threadIdx.x; s < total; s += blockDim.x) {
        const float phi = phis[s / n_theta], theta = thetas[s % n_theta];
        float st, ct, sp, cp;
        sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
        const float tx = st * cp, ty = st * sp, tz = ct;
        const float3 d = make_float3(tx * right.x + ty * up.x + tz * N.x, tx * right.y + ty * up.y + tz * N.y, tx * right.z + ty * up.z + tz * N.z);
        const float3 col = cube_sample(c, d, 0.0f);
        const float w = ct * st;
        acc.x += col.x * w; acc.y += col.y * w; acc.z += col.z * w;
    }
    const float3 r = block_sum3(acc, sm);
    if (threadIdx.x == 0) {
        const float k = kPi * kPi * (1.0f / (float)total);
        out[3 * (size_t)idx] = r.x * k; out[3 * (size_t)idx + 1] = r.y * k; out[3 * (size_t)idx + 2] = r.z * k;
    }
}

__device__ __forceinline__ float radical_inverse(uint32_t bits) { return (float)__brev(bits) * 2.3283064365386963e-10f; }
__device__ __forceinline__ float3 sample_ggx(float xi_x, float xi_y, float3 N, float3 t, float3 b, float roughness) {
    const float a = roughness * roughness;
    const float phi = 2.0f * kPi * xi_x;
    const float ct = sqrtf((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    const float st = sqrtf(1.0f - ct * ct);
    float sp, cp;
    sincosf(phi, &sp, &cp);
    const float hx = cp * st, hy = sp * st, hz = ct;
    return normalize3f(make_float3(t.x * hx + b.x * hy + N.x * hz, t.y * hx + b.y * hy + N.y * hz, t.z * hx + b.z * hy + N.z * hz));
}

// ---- GGX prefilter: one warp per output texel, lanes stride the Hammersley set (prefilter.frag:64-107) --------------------
__global__ void __launch_bounds__(256) prefilter_kernel(CubeView c, int n, float roughness, int n_samples, float *out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= 6 * n * n) return;
    const int i = warp % n, j = (warp / n) % n, f = warp / (n * n);
    const float3 N = normalize3f(face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n - 1.0f));
    float3 t, b;
    tangent_frame(N, 0.999f, t, b);
    const float a = roughness * roughness, a2 = a * a;
    const float sa_texel = 4.0f * kPi / (6.0f * 512.0f * 512.0f);            // prefilter.frag:93-94 (resolution literal 512)
    float3 acc = make_float3(0.f, 0.f, 0.f);
    float wsum = 0.f;
    for (int s = lane; s < n_samples; s += 32) {
        const float3 H = sample_ggx((float)s / (float)n_samples, radical_inverse((uint32_t)s), N, t, b, roughness);
        const float vh = N.x * H.x + N.y * H.y + N.z * H.z;
        const float3 L = normalize3f(make_float3(2.0f * vh * H.x - N.x, 2.0f * vh * H.y - N.y, 2.0f * vh * H.z - N.z));
        const float ndl = fmaxf(N.x * L.x + N.y * L.y + N.z * L.z, 0.0f);
        if (ndl > 0.0f) {
            const float ndh = fmaxf(vh, 0.0f);
            const float den = ndh * ndh * (a2 - 1.0f) + 1.0f;
            const float D = a2 / (kPi * den * den);
            const float pdf = D * ndh / (4.0f * ndh) + 0.0001f;
            const float sa_sample = 1.0f / ((float)n_samples * pdf + 0.0001f);
            const float lod = roughness == 0.0f ? 0.0f : 0.5f * log2f(sa_sample / sa_texel);
            const float3 col = cube_sample(c, L, lod);
            acc.x += col.x * ndl; acc.y += col.y * ndl; acc.z += col.z * ndl;
            wsum += ndl;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xFFFFFFFFu, acc.x, o); acc.y += __shfl_xor_sync(0xFFFFFFFFu, acc.y, o);
        acc.z += __shfl_xor_sync(0xFFFFFFFFu, acc.z, o); wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, o);
    }
    if (lane == 0) { out[3 * (size_t)warp] = acc.x / wsum; out[3 * (size_t)warp + 1] = acc.y / wsum; out[3 * (size_t)warp + 2] = acc.z / wsum; }
}

// ---- BRDF LUT: one warp per texel (brdf.frag:69-107) -------------------------------------------------------------------------
__device__ __forceinline__ float g_schlick(float ndv, float roughness) { const float k = roughness * roughness / 2.0f; return ndv / (ndv * (1.0f - k) + k); }
__global__ void __launch_bounds__(256) brdf_lut_kernel(int w, int h, int n_samples, float *out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= w * h) return;
    const int x = warp % w, y = warp / w;
    const float ndv = ((float)x + 0.5f) / (float)w, roughness = ((float)y + 0.5f) / (float)h;
    const float3 V = make_float3(sqrtf(1.0f - ndv * ndv), 0.0f, ndv), N = make_float3(0.f, 0.f, 1.f);
    float3 t, b;
    tangent_frame(N, 0.999f, t, b);
    float A = 0.f, B = 0.f;
    for (int s = lane; s < n_samples; s += 32) {
        const float3 H = sample_ggx((float)s / (float)n_samples, radical_inverse((uint32_t)s), N, t, b, roughness);
        const float vh = V.x * H.x + V.y * H.y + V.z * H.z;
        const float3 L = normalize3f(make_float3(2.0f * vh * H.x - V.x, 2.0f * vh * H.y - V.y, 2.0f * vh * H.z - V.z));
        const float ndl = fmaxf(L.z, 0.0f), ndh = fmaxf(H.z, 0.0f), vdh = fmaxf(vh, 0.0f);
        if (ndl > 0.0f) {
            const float G = g_schlick(ndl, roughness) * g_schlick(fmaxf(ndv, 0.0f), roughness);
            const float gv = (G * vdh) / (ndh * ndv);
            const float fc = powf(1.0f - vdh, 5.0f);
            A += (1.0f - fc) * gv;
            B += fc * gv;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { A += __shfl_xor_sync(0xFFFFFFFFu, A, o); B += __shfl_xor_sync(0xFFFFFFFFu, B, o); }
    if (lane == 0) { out[2 * (size_t)warp] = A / (float)n_samples; out[2 * (size_t)warp + 1] = B / (float)n_samples; }
}

// ---- env cube -> SH: one thread per phi column / texel row, shared-memory tree like the shaders' -------------------------
template <int ORDER>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float *o) {
    o[0] = 0.282095f;
    if (ORDER >= 2) { o[1] = 0.488603f * y; o[2] = 0.488603f * z; o[3] = 0.488603f * x; }
    if (ORDER >= 3) {
        const float x2 = x * x, y2 = y * y, z2 = z * z;
        o[4] = 1.092548f * x * y; o[5] = 1.092548f * y * z; o[6] = 0.315392f * (3.0f * z2 - 1.0f); o[7] = 1.092548f * x * z; o[8] = 0.546274f * (x2 - y2);
        if (ORDER >= 4) {
            o[9] = 0.590044f * y * (3.0f * x2 - y2); o[10] = 2.890611f * x * y * z; o[11] = 0.457046f * y * (5.0f * z2 - 1.0f);
            o[12] = 0.373176f * z * (5.0f * z2 - 3.0f); o[13] = 0.457046f * x * (5.0f * z2 - 1.0f); o[14] = 1.445306f * z * (x2 - y2);
            o[15] = 0.590044f * x * (x2 - 3.0f * y2);
        }
        if (ORDER >= 5) {
            o[16] = 2.503343f * x * y * (x2 - y2); o[17] = 1.770131f * y * z * (3.0f * x2 - y2); o[18] = 0.946175f * x * y * (7.0f * z2 - 1.0f);
            o[19] = 0.669047f * y * z * (7.0f * z2 - 3.0f); o[20] = 0.105786f * (35.0f * z2 * z2 - 30.0f * z2 + 3.0f);
            o[21] = 0.669047f * x * z * (7.0f * z2 - 3.0f); o[22] = 0.473087f * (x2 - y2) * (7.0f * z2 - 1.0f);
            o[23] = 1.770131f * x * z * (x2 - 3.0f * y2); o[24] = 0.625836f * (x2 * (x2 - 3.0f * y2) - y2 * (3.0f * x2 - y2));
        }
    }
}
template <int ORDER>
__global__ void env_project_sh_kernel(CubeView c, int method, int size, const float *thetas, int n_theta, double *out) {
    constexpr int N2 = ORDER * ORDER;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= size) return;
    float part[N2][3];
#pragma unroll
    for (int k = 0; k < N2; k++) part[k][0] = part[k][1] = part[k][2] = 0.f;
    if (method == 0) {
        const float delta = 2.0f * kPi / (float)size, phi = delta * (float)tid;
        float sp, cp;
        sincosf(phi, &sp, &cp);
        for (int q = 0; q < n_theta; q++) {
            float st, ct;
            sincosf(thetas[q], &st, &ct);
            const float sx = st * cp, sy = st * sp, sz = ct;
            const float3 col = cube_sample(c, make_float3(sy, sz, sx), 0.0f);     // sampleVec.yzx
            float y[N2];
            sh_basis<ORDER>(sx, sy, sz, y);
#pragma unroll
            for (int k = 0; k < N2; k++) { part[k][0] += col.x * st * y[k]; part[k][1] += col.y * st * y[k]; part[k][2] += col.z * st * y[k]; }
        }
    } else {
        for (int f = 0; f < 6; f++)
            for (int tu = 0; tu < size; tu++) {
                const float3 p = face_dir(f, (float)tu / (float)size * 2.0f - 1.0f, (float)tid / (float)size * 2.0f - 1.0f);
                const float wx = p.z, wy = p.x, wz = p.y;                        // pos.zxy
                const float d2 = wx * wx + wy * wy + wz * wz, il = 1.0f / sqrtf(d2), wt = 1.0f / (sqrtf(d2) * d2);
                const float3 col = cube_sample(c, p, 0.0f);
                float y[N2];
                sh_basis<ORDER>(wx * il, wy * il, wz * il, y);
#pragma unroll
                for (int k = 0; k < N2; k++) { part[k][0] += col.x * wt * y[k]; part[k][1] += col.y * wt * y[k]; part[k][2] += col.z * wt * y[k]; }
            }
    }
#pragma unroll
    for (int k = 0; k < N2; k++)
        for (int ch = 0; ch < 3; ch++) atomicAdd(&out[3 * k + ch], (double)part[k][ch]);
}

}  // namespace

struct prt_env {
    prt_ctx *ctx = nullptr;
    float4 *cube = nullptr;
    int n0 = 0, levels = 0;
};

#define ENV_TRY(expr)                                                                                                   \
    do {                                                                                                                \
        cudaError_t e_ = (expr);                                                                                        \
        if (e_ != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int prt_env_create(prt_ctx *c, const float *eq, int w, int h, int cube_size, prt_env **out) {
    if (!c || !eq || !out || w < 1 || h < 1) return prt_set_error(PRT_ERR_INVALID, "prt_env_create: bad argument");
    *out = nullptr;
    if (cube_size < 1 || (cube_size & (cube_size - 1))) return prt_set_error(PRT_ERR_INVALID, "prt_env_create: cube size must be a power of two");
    ENV_TRY(cudaSetDevice(prt_ctx_device(c)));
    cudaStream_t st = prt_ctx_stream(c);
    int levels = 0;
    while ((cube_size >> levels) >= 1) levels++;
    const size_t texels = level_off(cube_size, levels);
    float4 *cube = nullptr;
    float *d_eq = (float *)prt_ctx_scratch(c, 2, sizeof(float) * 3 * (size_t)w * h);       // staging of the equirect image, reused by later calls
    if (!d_eq) return prt_set_error(PRT_ERR_NOMEM, "prt_env_create: out of device memory");
    cudaError_t e = cudaMalloc(&cube, sizeof(float4) * texels);
    if (e != cudaSuccess) return prt_set_error(PRT_ERR_NOMEM, "prt_env_create: cudaMalloc failed");
    cudaMemcpyAsync(d_eq, eq, sizeof(float) * 3 * (size_t)w * h, cudaMemcpyHostToDevice, st);
    prt_ctx_timer_begin(c, st);
    const size_t n = (size_t)6 * cube_size * cube_size;
    equirect_to_cube_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_eq, w, h, cube_size, cube);
    for (int l = 1; l < levels; l++) {
        const int np = cube_size >> (l - 1);
        const size_t m = (size_t)6 * (np >> 1) * (np >> 1);
        cube_downsample_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(cube + level_off(cube_size, l - 1), np, cube + level_off(cube_size, l));
    }
    prt_ctx_timer_end(c, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(cube); return prt_set_error(PRT_ERR_CUDA, std::string("prt_env_create: ") + cudaGetErrorString(e)); }
    prt_env *env = new prt_env();
    env->ctx = c; env->cube = cube; env->n0 = cube_size; env->levels = levels;
    *out = env;
    return PRT_OK;
}

void prt_env_destroy(prt_env *e) {
    if (!e) return;
    cudaSetDevice(prt_ctx_device(e->ctx));
    cudaFree(e->cube);
    delete e;
}

int prt_env_levels(const prt_env *e) { return e ? e->levels : 0; }

static int download_rgb(prt_env *e, const float *d_src, size_t n_floats, float *out) {
    ENV_TRY(cudaMemcpyAsync(out, d_src, sizeof(float) * n_floats, cudaMemcpyDeviceToHost, prt_ctx_stream(e->ctx)));
    ENV_TRY(cudaStreamSynchronize(prt_ctx_stream(e->ctx)));
    return PRT_OK;
}

int prt_env_get_cube(prt_env *e, int level, float *out_rgb) {
    if (!e || !out_rgb || level < 0 || level >= e->levels) return prt_set_error(PRT_ERR_INVALID, "prt_env_get_cube: bad argument");
    ENV_TRY(cudaSetDevice(prt_ctx_device(e->ctx)));
    const int n = e->n0 >> level;
    const size_t tex = (size_t)6 * n * n;
    float *tmp = (float *)prt_ctx_scratch(e->ctx, 0, sizeof(float) * 3 * tex);
    if (!tmp) return prt_set_error(PRT_ERR_NOMEM, "prt_env_get_cube: out of device memory");
    cube_to_rgb_kernel<<<(unsigned)((tex + 255) / 256), 256, 0, prt_ctx_stream(e->ctx)>>>(e->cube + level_off(e->n0, level), tex, tmp);
    return download_rgb(e, tmp, 3 * tex, out_rgb);
}

// float-accumulated loop variables of the shaders (for (x = 0; x < limit; x += step)), reproduced on the host
static std::vector<float> loop_values(float limit, float step) {
    std::vector<float> v;
    for (float x = 0.0f; x < limit; x += step) v.push_back(x);
    return v;
}

int prt_env_irradiance(prt_env *e, int n_out, float *out_rgb) {
    if (!e || n_out < 1) return prt_set_error(PRT_ERR_INVALID, "prt_env_irradiance: bad argument");
    ENV_TRY(cudaSetDevice(prt_ctx_device(e->ctx)));
    cudaStream_t st = prt_ctx_stream(e->ctx);
    const std::vector<float> phis = loop_values(2.0f * kPi, 0.025f), thetas = loop_values(0.5f * kPi, 0.025f);   // irradiance.frag:25-29
    const size_t tex = (size_t)6 * n_out * n_out;
    float *d_tab = (float *)prt_ctx_scratch(e->ctx, 1, sizeof(float) * (phis.size() + thetas.size()));
    float *d_out = (float *)prt_ctx_scratch(e->ctx, 0, sizeof(float) * 3 * tex);
    if (!d_tab || !d_out) return prt_set_error(PRT_ERR_NOMEM, "prt_env_irradiance: out of device memory");
    cudaMemcpyAsync(d_tab, phis.data(), sizeof(float) * phis.size(), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_tab + phis.size(), thetas.data(), sizeof(float) * thetas.size(), cudaMemcpyHostToDevice, st);
    CubeView cv{e->cube, e->n0, e->levels};
    prt_ctx_timer_begin(e->ctx, st);
    irradiance_kernel<<<(unsigned)tex, 128, 0, st>>>(cv, n_out, d_tab, (int)phis.size(), d_tab + phis.size(), (int)thetas.size(), d_out);
    prt_ctx_timer_end(e->ctx, st);
    if (!out_rgb) { ENV_TRY(cudaStreamSynchronize(st)); return PRT_OK; }          // result stays on the device (timing runs)
    return download_rgb(e, d_out, 3 * tex, out_rgb);
}

int prt_env_prefilter(prt_env *e, int n_out, int mips, int n_samples, float *out_rgb) {
    if (!e || n_out < 1 || mips < 2 || n_samples < 1 || (n_out >> (mips - 1)) < 1) return prt_set_error(PRT_ERR_INVALID, "prt_env_prefilter: bad argument");
    ENV_TRY(cudaSetDevice(prt_ctx_device(e->ctx)));
    cudaStream_t st = prt_ctx_stream(e->ctx);
    const size_t total = level_off(n_out, mips);
    float *d_out = (float *)prt_ctx_scratch(e->ctx, 0, sizeof(float) * 3 * total);
    if (!d_out) return prt_set_error(PRT_ERR_NOMEM, "prt_env_prefilter: out of device memory");
    CubeView cv{e->cube, e->n0, e->levels};
    prt_ctx_timer_begin(e->ctx, st);
    for (int mip = 0; mip < mips; mip++) {
        const int n = n_out >> mip;
        const float roughness = (float)mip / (float)(mips - 1);                       // gl.cpp:561
        const size_t warps = (size_t)6 * n * n;
        prefilter_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(cv, n, roughness, n_samples, d_out + 3 * level_off(n_out, mip));
    }
    prt_ctx_timer_end(e->ctx, st);
    if (!out_rgb) { ENV_TRY(cudaStreamSynchronize(st)); return PRT_OK; }
    return download_rgb(e, d_out, 3 * total, out_rgb);
}

int prt_brdf_lut(prt_ctx *c, int w, int h, int n_samples, float *out_rg) {
    if (!c || w < 1 || h < 1 || n_samples < 1) return prt_set_error(PRT_ERR_INVALID, "prt_brdf_lut: bad argument");
    ENV_TRY(cudaSetDevice(prt_ctx_device(c)));
    cudaStream_t st = prt_ctx_stream(c);
    const size_t tex = (size_t)w * h;
    float *d_out = (float *)prt_ctx_scratch(c, 0, sizeof(float) * 2 * tex);
    if (!d_out) return prt_set_error(PRT_ERR_NOMEM, "prt_brdf_lut: out of device memory");
    prt_ctx_timer_begin(c, st);
    brdf_lut_kernel<<<(unsigned)((tex * 32 + 255) / 256), 256, 0, st>>>(w, h, n_samples, d_out);
    prt_ctx_timer_end(c, st);
    cudaError_t e = out_rg ? cudaMemcpyAsync(out_rg, d_out, sizeof(float) * 2 * tex, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string("prt_brdf_lut: ") + cudaGetErrorString(e));
    return PRT_OK;
}

int prt_env_project_sh(prt_env *e, int order, int method, int size, float *out_rgb_coeffs) {
    if (!e || !out_rgb_coeffs || order < 1 || order > 5 || (method != 0 && method != 1) || size < 1 || size > 4096)
        return prt_set_error(PRT_ERR_INVALID, "prt_env_project_sh: bad argument");
    ENV_TRY(cudaSetDevice(prt_ctx_device(e->ctx)));
    cudaStream_t st = prt_ctx_stream(e->ctx);
    const int n2 = order * order;
    const float delta = 2.0f * kPi / (float)size;
    const std::vector<float> thetas = loop_values(kPi, delta);                          // projectSH.comp:74
    double *d_acc = (double *)prt_ctx_scratch(e->ctx, 0, sizeof(double) * 3 * n2);
    float *d_th = (float *)prt_ctx_scratch(e->ctx, 1, sizeof(float) * thetas.size());
    if (!d_acc || !d_th) return prt_set_error(PRT_ERR_NOMEM, "prt_env_project_sh: out of device memory");
    cudaMemsetAsync(d_acc, 0, sizeof(double) * 3 * n2, st);
    cudaMemcpyAsync(d_th, thetas.data(), sizeof(float) * thetas.size(), cudaMemcpyHostToDevice, st);
    CubeView cv{e->cube, e->n0, e->levels};
    const int block = 64, grid = (size + block - 1) / block;
    switch (order) {
    case 1: env_project_sh_kernel<1><<<grid, block, 0, st>>>(cv, method, size, d_th, (int)thetas.size(), d_acc); break;
    case 2: env_project_sh_kernel<2><<<grid, block, 0, st>>>(cv, method, size, d_th, (int)thetas.size(), d_acc); break;
    case 3: env_project_sh_kernel<3><<<grid, block, 0, st>>>(cv, method, size, d_th, (int)thetas.size(), d_acc); break;
    case 4: env_project_sh_kernel<4><<<grid, block, 0, st>>>(cv, method, size, d_th, (int)thetas.size(), d_acc); break;
    default: env_project_sh_kernel<5><<<grid, block, 0, st>>>(cv, method, size, d_th, (int)thetas.size(), d_acc); break;
    }
    std::vector<double> acc(3 * (size_t)n2);
    cudaError_t er = cudaMemcpyAsync(acc.data(), d_acc, sizeof(double) * 3 * n2, cudaMemcpyDeviceToHost, st);
    if (er == cudaSuccess) er = cudaStreamSynchronize(st);
    if (er != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string("prt_env_project_sh: ") + cudaGetErrorString(er));
    const double scale = method == 0 ? (double)delta * (double)delta : 4.0 / (double)size / (double)size;
    for (int k = 0; k < 3 * n2; k++) out_rgb_coeffs[k] = (float)(acc[k] * scale);
    return PRT_OK;
}

// Ramamoorthi-Hanrahan polynomial pack the viewer's SH_Irad() expects (precomp_projectSH.comp:23,118-139; common/SH.glsl:17-36)
int prt_sh_pack_rh(const float *L, float *out28) {
    if (!L || !out28) return prt_set_error(PRT_ERR_INVALID, "prt_sh_pack_rh: null argument");
    const float c1 = 0.429043f, c2 = 0.511664f, c3 = 0.743125f, c4 = 0.886227f, c5 = 0.247708f;
    for (int ch = 0; ch < 3; ch++) {
        float *A = out28 + 4 * ch, *B = out28 + 12 + 4 * ch;
        A[0] = 2 * c2 * L[3 * 3 + ch]; A[1] = 2 * c2 * L[3 * 1 + ch]; A[2] = 2 * c2 * L[3 * 2 + ch]; A[3] = c4 * L[ch] - c5 * L[3 * 6 + ch];
        B[0] = 2 * c1 * L[3 * 4 + ch]; B[1] = 2 * c1 * L[3 * 7 + ch]; B[2] = 2 * c1 * L[3 * 5 + ch]; B[3] = c3 * L[3 * 6 + ch];
        out28[24 + ch] = c1 * L[3 * 8 + ch];
    }
    out28[27] = 1.0f;
    return PRT_OK;
}

}  // extern "C"
