/*
 * oracle/env.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Restates the reference's image-based-lighting passes (BASELINE config 2):
 *   equirect -> cube      src/shaders/rectangle2cube.frag:7-22 driven by LightProbe::equirectangular_to_cubemap (gl.cpp:581-591)
 *   mip chain             CubeMap::generateMipmap (gl.cpp:454-460)            [driver-defined: pinned as 2x2 box]
 *   irradiance            src/shaders/irradiance.frag:9-43   (gl.cpp:569-579, 32x32 faces, app.cpp:55)
 *   GGX prefilter         src/shaders/prefilter.frag:10-107  (gl.cpp:546-567, 256x256, 5 mips, roughness = mip/4, app.cpp:58)
 *   BRDF LUT              src/shaders/brdf.frag:9-113        (app.cpp:61-63, 512x512)
 *   env SH projection     src/shaders/bak/projectSH.comp:63-151 (lat-long), bak/image_projectSH.comp:64-133 (cube texels)
 *   R-H polynomial pack   src/shaders/precomp_projectSH.comp:23,118-139 / common/SH.glsl:17-36
 *
 * PARITY UNPINNED against the GL driver (absent): hardware cube filtering, seamless edges (platform.cpp:173), mip generation
 * and RGB16F storage are driver behaviour.  Pinned here as: FP32 storage; face selection by major axis (ties x>=y>=z); GL
 * (sc,tc) table == the reference's own cubeCoordToWorld (SH_function.h:104-109); bilinear with texel centres at (i+0.5)/N;
 * taps that fall off a face are re-projected through their direction onto the neighbouring face (nearest texel); trilinear =
 * lerp of two bilinear levels, lod clamped to the chain; implicit-LOD texture() calls use level 0.
 */
#include "prt_oracle.h"
#include "sh.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PIF 3.14159265359f

static void face_dir(int f, float u, float v, float d[3]) {
    switch (f) {
    case 0: d[0] = 1.f; d[1] = -v; d[2] = -u; break;
    case 1: d[0] = -1.f; d[1] = -v; d[2] = u; break;
    case 2: d[0] = u; d[1] = 1.f; d[2] = v; break;
    case 3: d[0] = u; d[1] = -1.f; d[2] = -v; break;
    case 4: d[0] = u; d[1] = -v; d[2] = 1.f; break;
    default: d[0] = -u; d[1] = -v; d[2] = -1.f; break;
    }
}
static void dir_face(const float d[3], int *f, float *s, float *t) {
    float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]), ma, sc, tc;
    if (ax >= ay && ax >= az) { ma = ax; if (d[0] >= 0) { *f = 0; sc = -d[2]; tc = -d[1]; } else { *f = 1; sc = d[2]; tc = -d[1]; } }
    else if (ay >= az) { ma = ay; if (d[1] >= 0) { *f = 2; sc = d[0]; tc = d[2]; } else { *f = 3; sc = d[0]; tc = -d[2]; } }
    else { ma = az; if (d[2] >= 0) { *f = 4; sc = d[0]; tc = -d[1]; } else { *f = 5; sc = -d[0]; tc = -d[1]; } }
    *s = 0.5f * (sc / ma + 1.0f);
    *t = 0.5f * (tc / ma + 1.0f);
}

/* cube storage: levels concatenated; level l has 6 faces of n_l x n_l RGB float, n_l = n0 >> l */
static size_t level_offset(int n0, int l) { size_t o = 0; for (int i = 0; i < l; i++) { size_t n = (size_t)(n0 >> i); o += 6 * n * n * 3; } return o; }
size_t prt_o_cube_floats(int n0, int levels) { return level_offset(n0, levels); }
int prt_o_cube_levels(int n0) { int l = 0; while ((n0 >> l) >= 1) l++; return l; }

static const float *texel(const float *cube, int n0, int l, int f, int i, int j) {
    int n = n0 >> l;
    if (i < 0 || j < 0 || i >= n || j >= n) {
        float d[3], s, t;
        face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n - 1.0f, d);
        dir_face(d, &f, &s, &t);
        i = (int)floorf(s * (float)n); j = (int)floorf(t * (float)n);
        i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
        j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
    }
    return cube + level_offset(n0, l) + ((size_t)f * n * n + (size_t)j * n + i) * 3;
}
static void cube_bilinear(const float *cube, int n0, int l, const float d[3], float out[3]) {
    int f; float s, t;
    dir_face(d, &f, &s, &t);
    int n = n0 >> l;
    float x = s * (float)n - 0.5f, y = t * (float)n - 0.5f;
    float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
    int i = (int)x0, j = (int)y0;
    const float *a = texel(cube, n0, l, f, i, j), *b = texel(cube, n0, l, f, i + 1, j);
    const float *c = texel(cube, n0, l, f, i, j + 1), *e = texel(cube, n0, l, f, i + 1, j + 1);
    for (int k = 0; k < 3; k++) {
        float top = a[k] + fx * (b[k] - a[k]), bot = c[k] + fx * (e[k] - c[k]);
        out[k] = top + fy * (bot - top);
    }
}
void prt_o_cube_sample(const float *cube, int n0, int levels, const float d[3], float lod, float out[3]) {
    if (!(lod > 0.0f)) lod = 0.0f;
    if (lod > (float)(levels - 1)) lod = (float)(levels - 1);
    int l0 = (int)floorf(lod);
    float f = lod - (float)l0;
    cube_bilinear(cube, n0, l0, d, out);
    if (f > 0.0f && l0 + 1 < levels) {
        float o1[3];
        cube_bilinear(cube, n0, l0 + 1, d, o1);
        for (int k = 0; k < 3; k++) out[k] = out[k] + f * (o1[k] - out[k]);
    }
}

/* SampleSphericalMap (rectangle2cube.frag:7-15) of a normalised direction; acos argument clamped against rounding */
void prt_o_equirect_uv(const float v[3], float uv[2]) {
    uv[0] = atan2f(-v[0], -v[2]) * 0.1591f + 0.5f;
    uv[1] = acosf(fminf(1.0f, fmaxf(-1.0f, v[1]))) * 0.3183f;
}

/* rectangle2cube.frag:7-22 + Tex2D bilinear, CLAMP_TO_EDGE (gl.cpp:375-380); then 2x2 box mips */
void prt_o_env_equirect_to_cube(const float *eq, int w, int h, int n0, int levels, float *cube) {
    for (int f = 0; f < 6; f++)
        for (int j = 0; j < n0; j++)
            for (int i = 0; i < n0; i++) {
                float d[3];
                face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n0 - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n0 - 1.0f, d);
                float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                float vn[3] = { d[0] * inv, d[1] * inv, d[2] * inv }, uvq[2];
                prt_o_equirect_uv(vn, uvq);
                float u = uvq[0], v = uvq[1];
                float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
                float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
                int i0 = (int)x0, j0 = (int)y0, i1 = i0 + 1, j1 = j0 + 1;
                i0 = i0 < 0 ? 0 : (i0 > w - 1 ? w - 1 : i0); i1 = i1 < 0 ? 0 : (i1 > w - 1 ? w - 1 : i1);
                j0 = j0 < 0 ? 0 : (j0 > h - 1 ? h - 1 : j0); j1 = j1 < 0 ? 0 : (j1 > h - 1 ? h - 1 : j1);
                float *o = cube + ((size_t)f * n0 * n0 + (size_t)j * n0 + i) * 3;
                for (int k = 0; k < 3; k++) {
                    float a = eq[((size_t)j0 * w + i0) * 3 + k], b = eq[((size_t)j0 * w + i1) * 3 + k];
                    float c = eq[((size_t)j1 * w + i0) * 3 + k], e = eq[((size_t)j1 * w + i1) * 3 + k];
                    float top = a + fx * (b - a), bot = c + fx * (e - c);
                    o[k] = top + fy * (bot - top);
                }
            }
    for (int l = 1; l < levels; l++) {
        int n = n0 >> l, np = n0 >> (l - 1);
        const float *src = cube + level_offset(n0, l - 1);
        float *dst = cube + level_offset(n0, l);
        for (int f = 0; f < 6; f++)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++)
                    for (int k = 0; k < 3; k++) {
                        const float *p = src + ((size_t)f * np * np + (size_t)(2 * j) * np + 2 * i) * 3 + k;
                        dst[((size_t)f * n * n + (size_t)j * n + i) * 3 + k] = 0.25f * ((p[0] + p[3]) + (p[(size_t)np * 3] + p[(size_t)np * 3 + 3]));
                    }
    }
}

static void tangent_frame(const float N[3], float thr, float right[3], float up[3]) {
    float u0[3] = { 0.f, 0.f, 1.f };
    if (!(fabsf(N[2]) < thr)) { u0[0] = 1.f; u0[2] = 0.f; }
    float r[3] = { u0[1] * N[2] - u0[2] * N[1], u0[2] * N[0] - u0[0] * N[2], u0[0] * N[1] - u0[1] * N[0] };
    float inv = 1.0f / sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    right[0] = r[0] * inv; right[1] = r[1] * inv; right[2] = r[2] * inv;
    up[0] = N[1] * right[2] - N[2] * right[1]; up[1] = N[2] * right[0] - N[0] * right[2]; up[2] = N[0] * right[1] - N[1] * right[0];
}

/* irradiance.frag:9-43 for one fragment: P = CubeTexPos (un-normalised position on the cube) */
void prt_o_env_irradiance_dir(const float *cube, int n0, int levels, const float P[3], float o[3]) {
                float N[3] = { P[0], P[1], P[2] }, right[3], up[3];
                float inv = 1.0f / sqrtf(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
                N[0] *= inv; N[1] *= inv; N[2] *= inv;
                tangent_frame(N, 0.99f, right, up);
                float acc[3] = { 0, 0, 0 }, nr = 0.0f;
                for (float phi = 0.0f; phi < 2.0f * PIF; phi += 0.025f)
                    for (float theta = 0.0f; theta < 0.5f * PIF; theta += 0.025f) {
                        float st = sinf(theta), ct = cosf(theta), sp = sinf(phi), cp = cosf(phi);
                        float tx = st * cp, ty = st * sp, tz = ct, d[3], c[3];
                        for (int k = 0; k < 3; k++) d[k] = tx * right[k] + ty * up[k] + tz * N[k];
                        prt_o_cube_sample(cube, n0, levels, d, 0.0f, c);
                        for (int k = 0; k < 3; k++) acc[k] += c[k] * ct * st;
                        nr += 1.0f;
                    }
                for (int k = 0; k < 3; k++) o[k] = PIF * PIF * acc[k] * (1.0f / nr);
}
void prt_o_env_irradiance(const float *cube, int n0, int levels, int n_out, float *out) {
    for (int f = 0; f < 6; f++)
        for (int j = 0; j < n_out; j++)
            for (int i = 0; i < n_out; i++) {
                float P[3];
                face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n_out - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n_out - 1.0f, P);
                prt_o_env_irradiance_dir(cube, n0, levels, P, out + ((size_t)f * n_out * n_out + (size_t)j * n_out + i) * 3);
            }
}

static float radical_inverse(uint32_t bits) {
    bits = (bits << 16) | (bits >> 16);
    bits = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
    bits = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
    bits = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
    bits = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
    return (float)bits * 2.3283064365386963e-10f;
}
/* ImportanceSampleGGX (prefilter.frag:42-62, brdf.frag:24-44): note the 0.999 threshold */
static void sample_ggx(float xi_x, float xi_y, const float N[3], float roughness, float H[3]) {
    float a = roughness * roughness;
    float phi = 2.0f * PIF * xi_x;
    float ct = sqrtf((1.0f - xi_y) / (1.0f + (a * a - 1.0f) * xi_y));
    float st = sqrtf(1.0f - ct * ct);
    float hx = cosf(phi) * st, hy = sinf(phi) * st, hz = ct, t[3], b[3];
    tangent_frame(N, 0.999f, t, b);
    float s[3];
    for (int k = 0; k < 3; k++) s[k] = t[k] * hx + b[k] * hy + N[k] * hz;
    float inv = 1.0f / sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    H[0] = s[0] * inv; H[1] = s[1] * inv; H[2] = s[2] * inv;
}

/* prefilter.frag:64-107 for one fragment: P = CubeTexPos (un-normalised position on the cube) */
void prt_o_env_prefilter_dir(const float *cube, int n0, int levels, const float P[3], float roughness, int n_samples, float o[3]) {
                    float N[3] = { P[0], P[1], P[2] };
                    float inv = 1.0f / sqrtf(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
                    N[0] *= inv; N[1] *= inv; N[2] *= inv;
                    float acc[3] = { 0, 0, 0 }, wsum = 0.0f;
                    for (int s = 0; s < n_samples; s++) {
                        float H[3], L[3], c[3];
                        sample_ggx((float)s / (float)n_samples, radical_inverse((uint32_t)s), N, roughness, H);
                        float vh = N[0] * H[0] + N[1] * H[1] + N[2] * H[2];
                        for (int k = 0; k < 3; k++) L[k] = 2.0f * vh * H[k] - N[k];
                        float il = 1.0f / sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
                        L[0] *= il; L[1] *= il; L[2] *= il;
                        float ndl = fmaxf(N[0] * L[0] + N[1] * L[1] + N[2] * L[2], 0.0f);
                        if (ndl > 0.0f) {
                            float ndh = fmaxf(vh, 0.0f), hdv = ndh;
                            float a = roughness * roughness, a2 = a * a;
                            float den = ndh * ndh * (a2 - 1.0f) + 1.0f;
                            float D = a2 / (PIF * den * den);
                            float pdf = D * ndh / (4.0f * hdv) + 0.0001f;
                            float sa_texel = 4.0f * PIF / (6.0f * 512.0f * 512.0f);
                            float sa_sample = 1.0f / ((float)n_samples * pdf + 0.0001f);
                            float lod = roughness == 0.0f ? 0.0f : 0.5f * log2f(sa_sample / sa_texel);
                            prt_o_cube_sample(cube, n0, levels, L, lod, c);
                            for (int k = 0; k < 3; k++) acc[k] += c[k] * ndl;
                            wsum += ndl;
                        }
                    }
                    for (int k = 0; k < 3; k++) o[k] = acc[k] / wsum;
}
/* out: levels concatenated like the cube storage, n_out >> mip, roughness = mip/(mips-1) (gl.cpp:556-566) */
void prt_o_env_prefilter(const float *cube, int n0, int levels, int n_out, int mips, int n_samples, float *out) {
    for (int mip = 0; mip < mips; mip++) {
        int n = n_out >> mip;
        float roughness = (float)mip / (float)(mips - 1);
        float *dst = out + level_offset(n_out, mip);
        for (int f = 0; f < 6; f++)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++) {
                    float P[3];
                    face_dir(f, 2.0f * ((float)i + 0.5f) / (float)n - 1.0f, 2.0f * ((float)j + 0.5f) / (float)n - 1.0f, P);
                    prt_o_env_prefilter_dir(cube, n0, levels, P, roughness, n_samples, dst + ((size_t)f * n * n + (size_t)j * n + i) * 3);
                }
    }
}

/* brdf.frag:47-113: out[y][x][2] = (A,B), NdotV = (x+0.5)/w, roughness = (y+0.5)/h */
static float g_schlick(float ndv, float roughness) { float k = roughness * roughness / 2.0f; return ndv / (ndv * (1.0f - k) + k); }
void prt_o_brdf_lut(int w, int h, int n_samples, float *out) {
    const float N[3] = { 0.f, 0.f, 1.f };
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float ndv = ((float)x + 0.5f) / (float)w, roughness = ((float)y + 0.5f) / (float)h;
            float V[3] = { sqrtf(1.0f - ndv * ndv), 0.0f, ndv };
            float A = 0.0f, B = 0.0f;
            for (int s = 0; s < n_samples; s++) {
                float H[3], L[3];
                sample_ggx((float)s / (float)n_samples, radical_inverse((uint32_t)s), N, roughness, H);
                float vh = V[0] * H[0] + V[1] * H[1] + V[2] * H[2];
                for (int k = 0; k < 3; k++) L[k] = 2.0f * vh * H[k] - V[k];
                float il = 1.0f / sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
                L[0] *= il; L[1] *= il; L[2] *= il;
                float ndl = fmaxf(L[2], 0.0f), ndh = fmaxf(H[2], 0.0f), vdh = fmaxf(vh, 0.0f);
                if (ndl > 0.0f) {
                    float G = g_schlick(fmaxf(ndl, 0.0f), roughness) * g_schlick(fmaxf(ndv, 0.0f), roughness);
                    float gv = (G * vdh) / (ndh * ndv);
                    float fc = powf(1.0f - vdh, 5.0f);
                    A += (1.0f - fc) * gv;
                    B += fc * gv;
                }
            }
            out[((size_t)y * w + x) * 2] = A / (float)n_samples;
            out[((size_t)y * w + x) * 2 + 1] = B / (float)n_samples;
        }
}

/* env cube -> SH (RGB x order^2).  method 0: lat-long quadrature (bak/projectSH.comp:63-151, SIZE phi-threads, theta += 2pi/SIZE);
 * method 1: cube-texel quadrature (bak/image_projectSH.comp:64-133, SIZE^2 texel corners per face, weight |p|^-3 * 4/SIZE^2). */
void prt_o_env_project_sh(const float *cube, int n0, int levels, int order, int method, int size, float *out) {
    int n2 = order * order;
    double acc[25][3];
    memset(acc, 0, sizeof acc);
    if (method == 0) {
        float delta = 2.0f * PIF / (float)size;
        for (int xid = 0; xid < size; xid++) {
            float phi = delta * (float)xid, sp = sinf(phi), cp = cosf(phi);
            float part[25][3];
            memset(part, 0, sizeof part);
            for (float theta = 0.0f; theta < PIF; theta += delta) {
                float st = sinf(theta), ct = cosf(theta);
                float sv[3] = { st * cp, st * sp, ct };           /* sh-space ("directX") */
                float gl[3] = { sv[1], sv[2], sv[0] }, c[3], y[25];  /* sampleVec.yzx: to GL space */
                prt_o_cube_sample(cube, n0, levels, gl, 0.0f, c);
                prt_sh_eval(order, 0, sv[0], sv[1], sv[2], y);
                for (int k = 0; k < n2; k++) for (int ch = 0; ch < 3; ch++) part[k][ch] += c[ch] * st * y[k];
            }
            for (int k = 0; k < n2; k++) for (int ch = 0; ch < 3; ch++) acc[k][ch] += (double)part[k][ch];
        }
        for (int k = 0; k < n2; k++) for (int ch = 0; ch < 3; ch++) out[3 * k + ch] = (float)(acc[k][ch] * (double)delta * (double)delta);
    } else {
        for (int tv = 0; tv < size; tv++) {
            float part[25][3];
            memset(part, 0, sizeof part);
            for (int f = 0; f < 6; f++)
                for (int tu = 0; tu < size; tu++) {
                    float p[3], c[3], y[25];
                    face_dir(f, (float)tu / (float)size * 2.0f - 1.0f, (float)tv / (float)size * 2.0f - 1.0f, p);
                    float w[3] = { p[2], p[0], p[1] };            /* pos.zxy: GL -> sh-space */
                    float d2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], il = 1.0f / sqrtf(d2);
                    float wt = 1.0f / (sqrtf(d2) * d2);
                    prt_o_cube_sample(cube, n0, levels, p, 0.0f, c);
                    prt_sh_eval(order, 0, w[0] * il, w[1] * il, w[2] * il, y);
                    for (int k = 0; k < n2; k++) for (int ch = 0; ch < 3; ch++) part[k][ch] += c[ch] * wt * y[k];
                }
            for (int k = 0; k < n2; k++) for (int ch = 0; ch < 3; ch++) acc[k][ch] += (double)part[k][ch];
        }
        for (int k = 0; k < n2; k++) for (int ch = 0; ch < 3; ch++) out[3 * k + ch] = (float)(acc[k][ch] * (4.0 / (double)size / (double)size));
    }
}

/* Ramamoorthi-Hanrahan pack (precomp_projectSH.comp:23,118-139): L[9][3] -> 7 vec4:
 * Ar,Ag,Ab = (2c2 L11, 2c2 L1-1, 2c2 L10, c4 L00 - c5 L20); Br,Bg,Bb = (2c1 L2-2, 2c1 L21, 2c1 L2-1, c3 L20); C = (c1 L22 rgb, 1) */
void prt_o_sh_pack_rh(const float *L, float *out28) {
    const float c1 = 0.429043f, c2 = 0.511664f, c3 = 0.743125f, c4 = 0.886227f, c5 = 0.247708f;
    for (int ch = 0; ch < 3; ch++) {
        float *A = out28 + 4 * ch, *B = out28 + 12 + 4 * ch;
        A[0] = 2 * c2 * L[3 * 3 + ch]; A[1] = 2 * c2 * L[3 * 1 + ch]; A[2] = 2 * c2 * L[3 * 2 + ch]; A[3] = c4 * L[ch] - c5 * L[3 * 6 + ch];
        B[0] = 2 * c1 * L[3 * 4 + ch]; B[1] = 2 * c1 * L[3 * 7 + ch]; B[2] = 2 * c1 * L[3 * 5 + ch]; B[3] = c3 * L[3 * 6 + ch];
        out28[24 + ch] = c1 * L[3 * 8 + ch];
    }
    out28[27] = 1.0f;
}
