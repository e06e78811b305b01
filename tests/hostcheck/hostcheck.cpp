// tests/hostcheck/hostcheck.cpp -- TEST TOOLING ONLY (never linked into libprt_b200.so).
// Compiles the product's BVH8 builder and the traversal header as plain C++ so that the CPU test-suite can
// check, without a GPU, that the 8-wide compressed BVH never culls a triangle the pinned test accepts
// (compared with the oracle's brute force) and that any-hit / closest-hit answers equal the oracle's.
#include "../../prt_b200/csrc/bvh8.h"
#include "../../prt_b200/csrc/traverse.cuh"
#include <cstdio>
#include <cstdlib>

using namespace prt;

extern "C" {
void *hc_build(const float *pos, size_t stride, uint32_t nv, const uint32_t *idx, uint32_t nt) {
    HostBVH8 *b = new HostBVH8();
    char err[256];
    if (build_bvh8(pos, stride, nv, idx, nt, b, err, sizeof err) != 0) { fprintf(stderr, "hc_build: %s\n", err); delete b; return nullptr; }
    return b;
}
void hc_free(void *h) { HostBVH8 *b = (HostBVH8 *)h; if (b) { free_bvh8(b); delete b; } }
void hc_info(void *h, uint32_t *out) { HostBVH8 *b = (HostBVH8 *)h; out[0] = b->n_nodes; out[1] = b->n_tris; out[2] = b->max_depth; }
// rays: n x 8 floats (org, tnear, dir, tfar)
void hc_any_hit(void *h, const float *rays, uint32_t n, uint8_t *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    for (uint32_t i = 0; i < n; i++) {
        const float *r = rays + 8 * (size_t)i;
        Trav t; t.reset_counters(); t.init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]);
        out[i] = t.run<true>(b->nodes, b->tris, 0, false) == TRAV_HIT;
    }
}
void hc_closest_hit(void *h, const float *rays, uint32_t n, float *out_t, uint32_t *out_prim, float *out_ng) {
    HostBVH8 *b = (HostBVH8 *)h;
    for (uint32_t i = 0; i < n; i++) {
        const float *r = rays + 8 * (size_t)i;
        Trav t; t.reset_counters(); t.init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]);
        int rc = t.run<false>(b->nodes, b->tris, 0, false);
        if (rc == TRAV_HIT) { out_t[i] = t.best_t; out_prim[i] = t.best_prim; f3 g = t.hit_ng(b->tris); out_ng[3 * i] = g.x; out_ng[3 * i + 1] = g.y; out_ng[3 * i + 2] = g.z; }
        else { out_t[i] = INFINITY; out_prim[i] = 0xFFFFFFFFu; out_ng[3 * i] = out_ng[3 * i + 1] = out_ng[3 * i + 2] = 0.f; }
    }
}
}

// work counters of the host traversal (node visits, triangle tests) summed over a ray batch
extern "C" void hc_count_work(void *h, const float *rays, uint32_t n, uint64_t *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    uint64_t nv = 0, nt = 0;
    for (uint32_t i = 0; i < n; i++) {
        const float *r = rays + 8 * (size_t)i;
        Trav t; t.reset_counters(); t.init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]);
        t.run<true>(b->nodes, b->tris, 0, false);
        nv += t.n_node_visits; nt += t.n_tri_tests;
    }
    out[0] = nv; out[1] = nt;
}
