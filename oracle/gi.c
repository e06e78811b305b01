/*
 * oracle/gi.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product library).
 *
 * CPU restatement of the reference's per-frame probe pipeline:
 *   Paral_Shadow::set_dir / render   src/opengl/gl.cpp:620-648        (sky light matrix; depth map)
 *   relight.comp                     src/shaders/relight.comp:68-82   (per-surfel direct light + SH feedback)
 *   lights                           src/shaders/common/light.glsl:18-46
 *   ShadowCalculation                src/shaders/common/paral_shadow.glsl:4-36
 *   get_albedo                       src/shaders/colored_wall.glsl:3-11
 *   SH_Irad                          src/shaders/common/SH.glsl:17-36 (GL_LINEAR, CLAMP_TO_EDGE volumes, volume.cpp:33-41)
 *   transfer2volume.comp             src/shaders/transfer2volume.comp:36-147
 *
 * parity unpinned: the reference runs these as GLSL on a GL driver (FP16 volume storage, driver-defined filtering and
 * rasterisation of the depth map) and ships no golden output.  Pinned here: FP32 storage; the depth map is the closest hit
 * of one ray per texel centre (what the depth raster of gl.cpp:633-648 samples); nearest shadow lookup with a border of 1;
 * trilinear volume lookup with texel centres at (i+0.5)/res and clamped indices; every float operation explicitly rounded
 * (arith.h) so the CUDA kernels can match bit for bit.
 */
#include "arith.h"
#include "prt_oracle.h"
#include <stdlib.h>

/* Paral_Shadow::set_dir (gl.cpp:620-631): direction from (up, dir) in [0,1]^2, glm::ortho(-30,30,-30,30,0.1,60) * glm::lookAt(30 d, 0, up).
 * Column-major 4x4 like glm.  Computed in double, rounded once. */
void prt_o_paral_shadow_matrix(float up, float dir, float out_dir[3], float m[16]) {
    const double PI = 3.14159265359;                                /* util.h:6 */
    double theta = PI * (double)up, phi = 2.0 * PI * (double)dir;
    double d[3] = { sin(theta) * sin(phi), cos(theta), sin(theta) * cos(phi) };
    double upv[3] = { 0.0, 1.0, 0.0 };
    if (up < 0.1f || up > 0.9f) { upv[1] = 0.0; upv[2] = 1.0; }
    double eye[3] = { 30.0 * d[0], 30.0 * d[1], 30.0 * d[2] };
    double f[3] = { -eye[0], -eye[1], -eye[2] };
    double fl = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int k = 0; k < 3; k++) f[k] /= fl;
    double s[3] = { f[1] * upv[2] - f[2] * upv[1], f[2] * upv[0] - f[0] * upv[2], f[0] * upv[1] - f[1] * upv[0] };
    double sl = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int k = 0; k < 3; k++) s[k] /= sl;
    double u[3] = { s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0] };
    double view[4][4] = { { s[0], u[0], -f[0], 0 }, { s[1], u[1], -f[1], 0 }, { s[2], u[2], -f[2], 0 },
                          { -(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2]), -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]),
                            f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2], 1 } };            /* [column][row] */
    const double l = -30, r = 30, b = -30, t = 30, n = 0.1, fa = 60;
    double proj[4][4] = { { 2 / (r - l), 0, 0, 0 }, { 0, 2 / (t - b), 0, 0 }, { 0, 0, -2 / (fa - n), 0 },
                          { -(r + l) / (r - l), -(t + b) / (t - b), -(fa + n) / (fa - n), 1 } };
    for (int c = 0; c < 4; c++)
        for (int rr = 0; rr < 4; rr++) {
            double a = 0;
            for (int k = 0; k < 4; k++) a += proj[k][rr] * view[c][k];
            m[4 * c + rr] = (float)a;
        }
    for (int k = 0; k < 3; k++) out_dir[k] = (float)d[k];
}

/* world-space ray of NDC (x, y, z in [-1,1]) for an AFFINE light matrix: origin at z = -1, direction to z = +1 */
static int light_rays(const float m[16], double O[3], double U[3], double V[3], double D[3]) {
    if (m[3] != 0.f || m[7] != 0.f || m[11] != 0.f || m[15] != 1.f) return -1;
    double a[3][3], inv[3][3];                                          /* a[row][col] */
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a[r][c] = m[4 * c + r];
    double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                 a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (det == 0.0) return -1;
    inv[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det; inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det; inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
    inv[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det; inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det; inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
    inv[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det; inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det; inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
    double rhs[3] = { 0.0 - m[12], 0.0 - m[13], -1.0 - m[14] };
    for (int r = 0; r < 3; r++) {
        O[r] = inv[r][0] * rhs[0] + inv[r][1] * rhs[1] + inv[r][2] * rhs[2];
        U[r] = inv[r][0]; V[r] = inv[r][1]; D[r] = 2.0 * inv[r][2];
    }
    return 0;
}

/* Paral_Shadow::render (gl.cpp:633-648): depth[j*size+i] in [0,1] of the nearest surface seen through texel (i,j); 1 = nothing */
int prt_o_shadow_map(const prt_o_scene *sc, const float m[16], int size, float *depth) {
    double O[3], U[3], V[3], D[3];
    if (light_rays(m, O, U, V, D)) return -1;
    const float of[3] = { (float)O[0], (float)O[1], (float)O[2] }, uf[3] = { (float)U[0], (float)U[1], (float)U[2] };
    const float vf[3] = { (float)V[0], (float)V[1], (float)V[2] }, df[3] = { (float)D[0], (float)D[1], (float)D[2] };
    for (int j = 0; j < size; j++)
        for (int i = 0; i < size; i++) {
            const float x = ((float)i + 0.5f) / (float)size * 2.0f - 1.0f, y = ((float)j + 0.5f) / (float)size * 2.0f - 1.0f;
            float org[3];
            for (int k = 0; k < 3; k++) org[k] = fmaf(y, vf[k], fmaf(x, uf[k], of[k]));
            float t = 1.0f, ng[3]; uint32_t prim;
            if (!prt_o_closest_hit(sc, org, df, 0.0f, 1.0f, 1, &t, &prim, ng)) t = 1.0f;
            depth[(size_t)j * size + i] = t;
        }
    return 0;
}

/* colored_wall.glsl:3-11 */
static void get_albedo(const float pos[3], float a[3]) {
    a[0] = a[1] = a[2] = 0.4f;
    if (pos[0] > 5.9f) {
        a[0] = a[1] = a[2] = 0.1f;
        a[((int)(pos[1] / 6.0f + 100.0f) + (int)(pos[2] / 6.0f + 100.0f)) % 3] += 0.7f;
    }
}

/* paral_shadow.glsl:4-36 with a nearest, clamp-to-border(1.0) depth texture (gl.cpp:603-607) */
static float shadow_calc(const float m[16], const float *depth, int size, const float pos[3], const float ldir[3], const float n[3]) {
    if (!depth) return 0.0f;
    float p[3];
    for (int r = 0; r < 3; r++) p[r] = fmaf(m[r], pos[0], fmaf(m[4 + r], pos[1], fmaf(m[8 + r], pos[2], m[12 + r])));
    const float w = fmaf(m[3], pos[0], fmaf(m[7], pos[1], fmaf(m[11], pos[2], m[15])));
    for (int r = 0; r < 3; r++) p[r] = fmaf(p[r] / w, 0.5f, 0.5f);
    float closest = 1.0f;
    if (p[0] >= 0.0f && p[0] < 1.0f && p[1] >= 0.0f && p[1] < 1.0f) {
        int i = (int)(p[0] * (float)size), j = (int)(p[1] * (float)size);
        if (i > size - 1) i = size - 1;
        if (j > size - 1) j = size - 1;
        closest = depth[(size_t)j * size + i];
    }
    const float ndl = fmaf(n[2], ldir[2], fmaf(n[1], ldir[1], n[0] * ldir[0]));
    const float bias = 0.1f * fmaxf(0.05f * (1.0f - ndl), 0.005f);
    float shadow = (p[2] - bias > closest) ? 1.0f : 0.0f;
    if (p[2] > 1.0f) shadow = 0.0f;
    return shadow;
}

/* trilinear fetch of a [rz][ry][rx][7][4] volume at texture coordinate c in [0,1]^3 (GL_LINEAR, CLAMP_TO_EDGE) */
static void volume_fetch(const float *vol, const int res[3], const float c[3], float out[28]) {
    int i0[3], i1[3]; float f[3];
    for (int a = 0; a < 3; a++) {
        const float u = fmaf(c[a], (float)res[a], -0.5f), fl = floorf(u);
        f[a] = u - fl;
        int i = (int)fl;
        i0[a] = i < 0 ? 0 : (i > res[a] - 1 ? res[a] - 1 : i);
        i1[a] = i + 1 < 0 ? 0 : (i + 1 > res[a] - 1 ? res[a] - 1 : i + 1);
    }
    for (int k = 0; k < 28; k++) out[k] = 0.0f;
    for (int dz = 0; dz < 2; dz++)
        for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++) {
                const float w = ((dx ? f[0] : 1.0f - f[0]) * (dy ? f[1] : 1.0f - f[1])) * (dz ? f[2] : 1.0f - f[2]);
                const size_t vox = ((size_t)(dz ? i1[2] : i0[2]) * res[1] + (dy ? i1[1] : i0[1])) * res[0] + (dx ? i1[0] : i0[0]);
                for (int k = 0; k < 28; k++) out[k] = fmaf(w, vol[28 * vox + k], out[k]);
            }
}

/* SH.glsl:17-36 */
static void sh_irad(const float *vol, const int res[3], const float scene_size[3], const float n[3], const float pos[3], float out[3]) {
    const float N[4] = { n[2], n[0], n[1], 1.0f };                    /* N.zxyw */
    float c[3], t[28];
    for (int a = 0; a < 3; a++) c[a] = (pos[a] - (-1.0f * scene_size[a])) / (2.0f * scene_size[a]);
    volume_fetch(vol, res, c, t);
    const float BN[4] = { N[0] * N[1], N[0] * N[2], N[1] * N[2], N[2] * N[2] };
    const float cc = fmaf(N[0], N[0], -(N[1] * N[1]));
    for (int ch = 0; ch < 3; ch++) {
        const float *A = t + 4 * ch, *B = t + 4 * (3 + ch);
        const float x = fmaf(A[3], N[3], fmaf(A[2], N[2], fmaf(A[1], N[1], A[0] * N[0])));
        const float y = fmaf(B[3], BN[3], fmaf(B[2], BN[2], fmaf(B[1], BN[1], B[0] * BN[0])));
        const float z = t[24 + ch] * cc;
        out[ch] = fmaxf((x + y) + z, 0.0f);
    }
}

static float len3(const float v[3]) { return sqrtf(fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0]))); }

/* relight.comp:68-82.  surfels[n][6]; albedo (optional) [n][3], NULL = colored_wall.glsl; depth (optional) size^2; volumes
 * (optional unless multi_bounce) [rz][ry][rx][7][4]; radiance [n][4] is read and written (temporal blend). */
void prt_o_relight(const prt_o_relight_params *P, uint32_t n, const float *surfels, const float *albedo, const float *depth, int shadow_size,
                   const float *volumes, const int volume_res[3], const float scene_size[3], float *radiance) {
    for (uint32_t s = 0; s < n; s++) {
        const float *pos = surfels + 6 * (size_t)s, *N = pos + 3;
        float alb[3];
        if (albedo) { alb[0] = albedo[3 * (size_t)s]; alb[1] = albedo[3 * (size_t)s + 1]; alb[2] = albedo[3 * (size_t)s + 2]; }
        else get_albedo(pos, alb);
        const float shadow = shadow_calc(P->light_space_matrix, depth, shadow_size, pos, P->sky_direction, N);
        /* Eval_ParalLight (light.glsl:43-46) */
        const float sky_cos = fmaxf(fmaf(P->sky_direction[2], N[2], fmaf(P->sky_direction[1], N[1], P->sky_direction[0] * N[0])), 0.0f);
        /* Eval_CastLight (light.glsl:18-31) */
        float cast[3] = { 0, 0, 0 };
        {
            const float d[3] = { P->cast_position[0] - pos[0], P->cast_position[1] - pos[1], P->cast_position[2] - pos[2] };
            const float dist = len3(d), inv = 1.0f / dist;
            const float ld[3] = { d[0] * inv, d[1] * inv, d[2] * inv };
            const float nd[3] = { -P->cast_direction[0], -P->cast_direction[1], -P->cast_direction[2] };
            const float ninv = 1.0f / len3(nd);
            const float theta = fmaf(ld[2], nd[2] * ninv, fmaf(ld[1], nd[1] * ninv, ld[0] * (nd[0] * ninv)));
            if (theta > P->cast_cutoff) {
                const float icos = fmaxf(fmaf(ld[2], N[2], fmaf(ld[1], N[1], ld[0] * N[0])), 0.0f);
                const float soft = (theta - P->cast_cutoff) / (1.0f - P->cast_cutoff);
                for (int k = 0; k < 3; k++) cast[k] = ((soft * P->cast_intensity[k]) * icos) / (dist * dist);
            }
        }
        /* Eval_PointLight (light.glsl:33-41) */
        float amb[3] = { 0, 0, 0 };
        if (P->ambient_intensity[0] != 0.f || P->ambient_intensity[1] != 0.f || P->ambient_intensity[2] != 0.f) {
            const float d[3] = { P->ambient_position[0] - pos[0], P->ambient_position[1] - pos[1], P->ambient_position[2] - pos[2] };
            const float dist = len3(d), inv = 1.0f / dist;
            const float icos = fmaxf(fmaf(d[2] * inv, N[2], fmaf(d[1] * inv, N[1], (d[0] * inv) * N[0])), 0.0f);
            for (int k = 0; k < 3; k++) amb[k] = (P->ambient_intensity[k] * icos) / (dist * dist);
        }
        float irr[3] = { 0, 0, 0 };
        if (P->multi_bounce && volumes) {
            const float q[3] = { fmaf(P->sh_shift, N[0], pos[0]), fmaf(P->sh_shift, N[1], pos[1]), fmaf(P->sh_shift, N[2], pos[2]) };
            sh_irad(volumes, volume_res, scene_size, N, q, irr);
        }
        float *r = radiance + 4 * (size_t)s;
        for (int k = 0; k < 3; k++) {
            float rad = alb[k] * ((((1.0f - shadow) * (P->sky_intensity[k] * sky_cos)) + cast[k]) + amb[k]);
            if (P->multi_bounce && volumes) rad = rad + ((alb[k] * P->atten) * irr[k]) / PRT_PI_F;
            r[k] = fmaf(P->temp_weight, rad, (1.0f - P->temp_weight) * r[k]);
        }
        r[3] = 1.0f;
    }
}

/* transfer2volume.comp:36-147; probe_sh [pz][py][px][7][4], weights [vz][vy][vx][4] x 2, out [vz][vy][vx][7][4].
 * texelFetch outside the probe grid contributes 0 (robust buffer access; calculate_weight gives such corners weight 0). */
void prt_o_transfer_to_volume(const float *probe_sh, const int probe_res[3], const float *w0123, const float *w4567,
                              const int volume_res[3], float *out) {
    static const int off[8][3] = { {0,0,1}, {1,0,1}, {1,0,0}, {0,0,0}, {0,1,0}, {0,1,1}, {1,1,1}, {1,1,0} };
    for (int z = 0; z < volume_res[2]; z++)
        for (int y = 0; y < volume_res[1]; y++)
            for (int x = 0; x < volume_res[0]; x++) {
                const size_t v = ((size_t)z * volume_res[1] + y) * volume_res[0] + x;
                const int id[3] = { x, y, z };
                int anchor[3];
                for (int a = 0; a < 3; a++) {
                    const float vp = (((float)id[a] + 0.5f) / (float)volume_res[a]) * (float)probe_res[a] - 0.5f;
                    anchor[a] = (int)floorf(vp);
                }
                float acc[28];
                for (int k = 0; k < 28; k++) acc[k] = 0.0f;
                for (int c = 0; c < 8; c++) {
                    const float w = c < 4 ? w0123[4 * v + c] : w4567[4 * v + c - 4];
                    const int px = anchor[0] + off[c][0], py = anchor[1] + off[c][1], pz = anchor[2] + off[c][2];
                    if (px < 0 || py < 0 || pz < 0 || px >= probe_res[0] || py >= probe_res[1] || pz >= probe_res[2]) continue;
                    const float *src = probe_sh + 28 * (((size_t)pz * probe_res[1] + py) * probe_res[0] + px);
                    for (int k = 0; k < 28; k++) acc[k] = fmaf(w, src[k], acc[k]);
                }
                for (int k = 0; k < 28; k++) out[28 * v + k] = acc[k];
            }
}
