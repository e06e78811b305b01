/*
 * oracle/sh.h -- TEST INFRASTRUCTURE ONLY.
 * Real spherical-harmonic basis, bands l = 0..4 (25 functions), index k = l(l+1)+m
 * (sh::GetIndex, reference raytracing.cpp:334,348).
 *
 * l <= 2 restates reference src/sh/SH_function.h:7-41 with its 6-digit literals and *no*
 * Condon-Shortley sign -- the convention the viewer's SH_Irad() (src/shaders/common/SH.glsl:17-36)
 * and precomp_projectSH.comp:118-139 expect.  l = 3,4 are not in the reference; standard real-SH
 * polynomials validated by the orthonormality known-answer test.
 * cs_phase = 1 multiplies odd-|m| functions by -1, reproducing google/spherical-harmonics'
 * sh::EvalSH sign convention (the un-vendored library raytracing.cpp:226 calls).
 *
 * Argument d is the *sh-space* direction: world (x,y,z) -> (z,x,y)  (raytracing.cpp:226,
 * volume.cpp:255, SH.glsl:19 "openGL to directX").
 */
#ifndef PRT_ORACLE_SH_H
#define PRT_ORACLE_SH_H

static inline void prt_sh_eval(int order, int cs_phase, float x, float y, float z, float *out) {
    /* order = number of bands (1..5) */
    const float o = cs_phase ? -1.0f : 1.0f;
    out[0] = 0.282095f;
    if (order < 2) return;
    out[1] = o * 0.488603f * y;
    out[2] = 0.488603f * z;
    out[3] = o * 0.488603f * x;
    if (order < 3) return;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    out[4] = 1.092548f * x * y;
    out[5] = o * 1.092548f * y * z;
    out[6] = 0.315392f * (3.0f * z2 - 1.0f);
    out[7] = o * 1.092548f * x * z;
    out[8] = 0.546274f * (x2 - y2);
    if (order < 4) return;
    out[9] = o * 0.590044f * y * (3.0f * x2 - y2);
    out[10] = 2.890611f * x * y * z;
    out[11] = o * 0.457046f * y * (5.0f * z2 - 1.0f);
    out[12] = 0.373176f * z * (5.0f * z2 - 3.0f);
    out[13] = o * 0.457046f * x * (5.0f * z2 - 1.0f);
    out[14] = 1.445306f * z * (x2 - y2);
    out[15] = o * 0.590044f * x * (x2 - 3.0f * y2);
    if (order < 5) return;
    out[16] = 2.503343f * x * y * (x2 - y2);
    out[17] = o * 1.770131f * y * z * (3.0f * x2 - y2);
    out[18] = 0.946175f * x * y * (7.0f * z2 - 1.0f);
    out[19] = o * 0.669047f * y * z * (7.0f * z2 - 3.0f);
    out[20] = 0.105786f * (35.0f * z2 * z2 - 30.0f * z2 + 3.0f);
    out[21] = o * 0.669047f * x * z * (7.0f * z2 - 3.0f);
    out[22] = 0.473087f * (x2 - y2) * (7.0f * z2 - 1.0f);
    out[23] = o * 1.770131f * x * z * (x2 - 3.0f * y2);
    out[24] = 0.625836f * (x2 * (x2 - 3.0f * y2) - y2 * (3.0f * x2 - y2));
}

/* Clamped-cosine zonal factors A_l/pi, l=0..4 (Ramamoorthi-Hanrahan): the analytic
 * "unshadowed" transfer the reference only hints at (scene/model.cpp:29-31 rotate_cos_lobe*INV_PI). */
static const float PRT_COS_LOBE[5] = { 1.0f, 2.0f / 3.0f, 0.25f, 0.0f, -1.0f / 24.0f };

#endif
