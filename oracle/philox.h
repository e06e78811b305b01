/*
 * oracle/philox.h -- TEST INFRASTRUCTURE ONLY.
 * Philox4x32-10 counter-based RNG (Salmon et al., SC'11; Random123 constants).  Replaces the
 * reference's irreproducible thread_local std::mt19937 (raytracing.cpp:14-18): see SURVEY
 * section 7 "Sample-identical parity".  Integer-only, so CPU and GPU agree exactly.
 *   stream 0: strata jitter     ctr = (sample s, 0, 0, 0)          -> (xi1, xi2)
 *   stream 1: bounce randoms    ctr = (vertex, sample, bounce, 1)  -> (u, v)
 *   stream 2: preview tracer    ctr = (pixel, frame, bounce, 2)    -> (u, v)
 *   key = (seed, 0x50525421)
 * uniform float = (x >> 8) * 2^-24 in [0,1)  (same range as uniform_real_distribution<float>(0,1))
 */
#ifndef PRT_ORACLE_PHILOX_H
#define PRT_ORACLE_PHILOX_H
#include <stdint.h>

static inline void prt_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define PRT_PHILOX_KEY1 0x50525421u

static inline float prt_u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

static inline void prt_rand2(uint32_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t stream, float *u, float *v) {
    uint32_t ctr[4] = { a, b, c, stream }, key[2] = { seed, PRT_PHILOX_KEY1 }, o[4];
    prt_philox4x32_10(ctr, key, o);
    *u = prt_u01(o[0]);
    *v = prt_u01(o[1]);
}
#endif
