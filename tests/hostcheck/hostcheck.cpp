// tests/hostcheck/hostcheck.cpp -- TEST TOOLING ONLY (never linked into libprt_b200.so).
// Compiles the product's BVH8 builder and the traversal header as plain C++ so that the CPU test-suite can
// check, without a GPU, that the 8-wide compressed BVH never culls a triangle the pinned test accepts
// (compared with the oracle's brute force) and that any-hit / closest-hit answers equal the oracle's.
#include "../../prt_b200/csrc/bvh8.h"
#include "../../prt_b200/csrc/traverse.cuh"
#include "../../prt_b200/csrc/horizon_math.cuh"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>

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
        Trav t; t.reset_counters(); t.init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]); t.start_root();
        out[i] = t.run<true>(b->nodes, b->tris, 0, false) == TRAV_HIT;
    }
}
void hc_closest_hit(void *h, const float *rays, uint32_t n, float *out_t, uint32_t *out_prim, float *out_ng) {
    HostBVH8 *b = (HostBVH8 *)h;
    for (uint32_t i = 0; i < n; i++) {
        const float *r = rays + 8 * (size_t)i;
        Trav t; t.reset_counters(); t.init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]); t.start_root();
        t.run<false>(b->nodes, b->tris, 0, false);
        if (t.best_prim != 0xFFFFFFFFu) { out_t[i] = t.best_t; out_prim[i] = t.best_prim; f3 g = t.hit_ng(b->tris); out_ng[3 * i] = g.x; out_ng[3 * i + 1] = g.y; out_ng[3 * i + 2] = g.z; }
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
        Trav t; t.reset_counters(); t.init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]); t.start_root();
        t.run<true>(b->nodes, b->tris, 0, false);
        nv += t.n_node_visits; nt += t.n_tri_tests;
    }
    out[0] = nv; out[1] = nt;
}

// Experiment (design study): cost of traversing bundles of 32 same-origin rays as ONE packet with a
// shared stack (a node is visited when any live ray hits its box) versus per-ray traversal.
// out[0] = packet node visits, out[1] = packet triangle tests (one per triangle per packet), out[2] = sum over rays of
// per-ray node visits, out[3] = sum of per-ray tri tests, out[4] = lane-slots doing useful tri tests in packet mode
extern "C" void hc_packet_stats(void *h, const float *rays, uint32_t n, uint64_t *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    uint64_t pv = 0, pt = 0, rv = 0, rt = 0, useful = 0;
    for (uint32_t base = 0; base + 32 <= n; base += 32) {
        Trav tr[32]; bool live[32];
        for (int l = 0; l < 32; l++) {
            const float *r = rays + 8 * (size_t)(base + l);
            tr[l].reset_counters(); tr[l].init(mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7]); tr[l].start_root();
            Trav t = tr[l]; t.run<true>(b->nodes, b->tris, 0, false); rv += t.n_node_visits; rt += t.n_tri_tests;
            live[l] = true;
        }
        uint32_t stack[256]; int sp = 0; stack[sp++] = 0; int nlive = 32;
        while (sp && nlive) {
            uint32_t node = stack[--sp];
            pv++;
            const Node8 &nd = b->nodes[node];
            uint32_t inner_slots = 0, tri_bits = 0; uint32_t lane_tri[32];
            for (int l = 0; l < 32; l++) {
                lane_tri[l] = 0;
                if (!live[l]) continue;
                tr[l].visit_node(b->nodes, node);
                uint32_t hits = tr[l].ng.y >> 24, oi = tr[l].octinv4 & 7u;
                for (int k = 0; k < 8; k++) if (hits & (1u << k)) inner_slots |= 1u << (k ^ oi);
                lane_tri[l] = tr[l].tg.y; tri_bits |= tr[l].tg.y;
            }
            for (int k = 0; k < 24 && nlive; k++) if (tri_bits & (1u << k)) {
                pt++;
                for (int l = 0; l < 32; l++) if (live[l] && (lane_tri[l] & (1u << k))) {
                    useful++;
                    float t; uint32_t prim;
                    if (tr[l].tri_test(b->tris, nd.tri_base + k, false, t, prim)) { live[l] = false; nlive--; }
                }
            }
            for (int s = 7; s >= 0; s--) if (inner_slots & (1u << s)) {
                uint32_t rel = (uint32_t)__builtin_popcount(nd.imask & ((1u << s) - 1u));
                if (sp < 256) stack[sp++] = nd.child_base + rel;
            }
        }
    }
    out[0] = pv; out[1] = pt; out[2] = rv; out[3] = rt; out[4] = useful;
}

// Experiment: per-origin entry list.  Descend the nodes whose box contains the origin; children that do not contain it
// and are not entirely below the tangent plane become candidates (subtree roots or leaves).  Reports, for rays of one
// origin: number of candidates, candidate boxes hit per ray, node visits and triangle tests left per ray.
struct Cand { float lo[3], hi[3]; uint32_t node; uint32_t tri0, ntri; };
static void decode_child(const Node8 &nd, int s, float lo[3], float hi[3]) {
    const float sc[3] = { PRT_U2F((uint32_t)nd.ex << 23), PRT_U2F((uint32_t)nd.ey << 23), PRT_U2F((uint32_t)nd.ez << 23) };
    const float p[3] = { nd.px, nd.py, nd.pz };
    const uint8_t *ql[3] = { nd.qlox, nd.qloy, nd.qloz }, *qh[3] = { nd.qhix, nd.qhiy, nd.qhiz };
    for (int a = 0; a < 3; a++) { lo[a] = p[a] + ql[a][s] * sc[a]; hi[a] = p[a] + qh[a][s] * sc[a]; }
}
extern "C" void hc_entry_stats(void *h, const float *org, const float *nrm, const float *dirs, uint32_t nrays, int max_cands, double *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    std::vector<Cand> cands; std::vector<uint32_t> queue; queue.push_back(0);
    int expanded = 0;
    while (!queue.empty()) {
        uint32_t x = queue.back(); queue.pop_back(); expanded++;
        const Node8 &nd = b->nodes[x];
        for (int s = 0; s < 8; s++) {
            if (!nd.meta[s]) continue;
            Cand c; decode_child(nd, s, c.lo, c.hi);
            float mx = 0.f, far2 = 0.f; bool inside = true;
            for (int a = 0; a < 3; a++) {
                float l = c.lo[a] - org[a], hh = c.hi[a] - org[a];
                mx += std::max(nrm[a] * l, nrm[a] * hh);
                far2 = std::max(far2, std::max(std::fabs(l), std::fabs(hh)));
                if (org[a] < c.lo[a] || org[a] > c.hi[a]) inside = false;
            }
            if (mx < -1e-5f * far2) continue;   // wholly below the tangent plane
            bool inner = (nd.imask >> s) & 1;
            if (inner) {
                uint32_t child = nd.child_base + __builtin_popcount(nd.imask & ((1u << s) - 1u));
                if (inside && (int)(cands.size() + queue.size()) < max_cands) { queue.push_back(child); continue; }
                c.node = child; c.ntri = 0; c.tri0 = 0;
            } else {
                c.node = 0xFFFFFFFFu; c.tri0 = nd.tri_base + (nd.meta[s] & 31); c.ntri = __builtin_popcount(nd.meta[s] >> 5);
            }
            cands.push_back(c);
        }
    }
    uint64_t box_hits = 0, nv = 0, nt = 0, occl = 0;
    for (uint32_t i = 0; i < nrays; i++) {
        f3 o = mk3(org[0], org[1], org[2]), d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        Trav t; t.reset_counters(); t.init(o, d, 0.f, INFINITY);
        bool hit = false;
        for (size_t k = 0; k < cands.size() && !hit; k++) {
            const Cand &c = cands[k];
            float t0 = 0.f, t1 = INFINITY; const float id[3] = { t.idx, t.idy, t.idz }; const float oo[3] = { o.x, o.y, o.z };
            for (int a = 0; a < 3; a++) { float ta = (c.lo[a] - oo[a]) * id[a], tb = (c.hi[a] - oo[a]) * id[a]; t0 = std::max(t0, std::min(ta, tb)); t1 = std::min(t1, std::max(ta, tb)); }
            if (!(t0 <= t1)) continue;
            box_hits++;
            if (c.node == 0xFFFFFFFFu) {
                for (uint32_t j = 0; j < c.ntri && !hit; j++) { float tt; uint32_t pr; nt++; hit = t.tri_test(b->tris, c.tri0 + j, false, tt, pr); }
            } else {
                t.ng.x = c.node; t.ng.y = 0x80000000u; t.tg.y = 0; t.sp = 0; t.n_node_visits = 0; t.n_tri_tests = 0;
                hit = t.run<true>(b->nodes, b->tris, 0, false) == TRAV_HIT;
                nv += t.n_node_visits; nt += t.n_tri_tests;
            }
        }
        occl += hit;
    }
    out[0] = (double)cands.size(); out[1] = (double)box_hits / nrays; out[2] = (double)nv / nrays; out[3] = (double)nt / nrays;
    out[4] = (double)occl / nrays; out[5] = expanded;
}

// Experiment: entry list expanded further by apparent size.  After the chain of nodes containing the origin, the subtree
// candidate with the largest (radius / distance)^2 is replaced by its children until `total_max` candidates exist or no
// candidate exceeds `min_ratio`.  Same outputs as hc_entry_stats.
extern "C" void hc_entry_stats2(void *h, const float *org, const float *nrm, const float *dirs, uint32_t nrays, int chain_max, int total_max,
                                float min_ratio, double *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    std::vector<Cand> cands; std::vector<uint32_t> queue; queue.push_back(0);
    int expanded = 0;
    auto push_children = [&](uint32_t x, bool chain) {
        const Node8 &nd = b->nodes[x];
        for (int s = 0; s < 8; s++) {
            if (!nd.meta[s]) continue;
            Cand c; decode_child(nd, s, c.lo, c.hi);
            float mx = 0.f, far2 = 0.f; bool inside = true;
            for (int a = 0; a < 3; a++) {
                float l = c.lo[a] - org[a], hh = c.hi[a] - org[a];
                mx += std::max(nrm[a] * l, nrm[a] * hh);
                far2 = std::max(far2, std::max(std::fabs(l), std::fabs(hh)));
                if (org[a] < c.lo[a] || org[a] > c.hi[a]) inside = false;
            }
            if (mx < -1e-5f * far2) continue;
            bool inner = (nd.imask >> s) & 1;
            if (inner) {
                uint32_t child = nd.child_base + __builtin_popcount(nd.imask & ((1u << s) - 1u));
                if (chain && inside && (int)(cands.size() + queue.size()) < chain_max) { queue.push_back(child); continue; }
                c.node = child; c.ntri = 0; c.tri0 = 0;
            } else {
                c.node = 0xFFFFFFFFu; c.tri0 = nd.tri_base + (nd.meta[s] & 31); c.ntri = __builtin_popcount(nd.meta[s] >> 5);
            }
            cands.push_back(c);
        }
    };
    while (!queue.empty()) { uint32_t x = queue.back(); queue.pop_back(); expanded++; push_children(x, true); }
    for (;;) {
        int best = -1; float best_ratio = min_ratio;
        for (size_t k = 0; k < cands.size(); k++) {
            if (cands[k].node == 0xFFFFFFFFu) continue;
            float r2 = 0.f, d2 = 0.f;
            for (int a = 0; a < 3; a++) { float e = 0.5f * (cands[k].hi[a] - cands[k].lo[a]), c = 0.5f * (cands[k].hi[a] + cands[k].lo[a]) - org[a]; r2 += e * e; d2 += c * c; }
            float ratio = r2 / std::max(d2, 1e-30f);
            if (ratio > best_ratio) { best_ratio = ratio; best = (int)k; }
        }
        if (best < 0 || (int)cands.size() + 7 > total_max) break;
        uint32_t node = cands[best].node;
        cands.erase(cands.begin() + best);
        expanded++;
        push_children(node, false);
    }
    uint64_t box_hits = 0, nv = 0, nt = 0, occl = 0;
    for (uint32_t i = 0; i < nrays; i++) {
        f3 o = mk3(org[0], org[1], org[2]), d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        Trav t; t.reset_counters(); t.init(o, d, 0.f, INFINITY);
        bool hit = false;
        for (size_t k = 0; k < cands.size() && !hit; k++) {
            const Cand &c = cands[k];
            float t0 = 0.f, t1 = INFINITY; const float id[3] = { t.idx, t.idy, t.idz }; const float oo[3] = { o.x, o.y, o.z };
            for (int a = 0; a < 3; a++) { float ta = (c.lo[a] - oo[a]) * id[a], tb = (c.hi[a] - oo[a]) * id[a]; t0 = std::max(t0, std::min(ta, tb)); t1 = std::min(t1, std::max(ta, tb)); }
            if (!(t0 <= t1)) continue;
            box_hits++;
            if (c.node == 0xFFFFFFFFu) {
                for (uint32_t j = 0; j < c.ntri && !hit; j++) { float tt; uint32_t pr; nt++; hit = t.tri_test(b->tris, c.tri0 + j, false, tt, pr); }
            } else {
                t.ng.x = c.node; t.ng.y = 0x80000000u; t.tg.y = 0; t.sp = 0; t.n_node_visits = 0; t.n_tri_tests = 0;
                hit = t.run<true>(b->nodes, b->tris, 0, false) == TRAV_HIT;
                nv += t.n_node_visits; nt += t.n_tri_tests;
            }
        }
        occl += hit;
    }
    out[0] = (double)cands.size(); out[1] = (double)box_hits / nrays; out[2] = (double)nv / nrays; out[3] = (double)nt / nrays;
    out[4] = (double)occl / nrays; out[5] = expanded;
}

// Experiment: directional binning of entry-list candidates.  Each candidate box is bounded by a cone (bounding sphere seen
// from the origin) and registered in the cells of a G x G grid over the unit disk (the projected hemisphere, local frame)
// that the cone can touch.  Reports the mean number of candidates a ray must test: (a) its own cell, (b) the union of the
// cells of its 32-ray bundle (rays given in processing order).
extern "C" void hc_grid_stats(void *h, const float *org, const float *nrm, const float *local_dirs, uint32_t nrays, int G, double *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    std::vector<Cand> cands; std::vector<uint32_t> queue; queue.push_back(0);
    while (!queue.empty()) {
        uint32_t x = queue.back(); queue.pop_back();
        const Node8 &nd = b->nodes[x];
        for (int s = 0; s < 8; s++) {
            if (!nd.meta[s]) continue;
            Cand c; decode_child(nd, s, c.lo, c.hi);
            float mx = 0.f, far2 = 0.f; bool inside = true;
            for (int a = 0; a < 3; a++) {
                float l = c.lo[a] - org[a], hh = c.hi[a] - org[a];
                mx += std::max(nrm[a] * l, nrm[a] * hh);
                far2 = std::max(far2, std::max(std::fabs(l), std::fabs(hh)));
                if (org[a] < c.lo[a] || org[a] > c.hi[a]) inside = false;
            }
            if (mx < -1e-5f * far2) continue;
            bool inner = (nd.imask >> s) & 1;
            if (inner && inside) { queue.push_back(nd.child_base + __builtin_popcount(nd.imask & ((1u << s) - 1u))); continue; }
            cands.push_back(c);
        }
    }
    // local frame
    f3 N = mk3(nrm[0], nrm[1], nrm[2]); Frame fr = make_frame(N);
    std::vector<std::vector<uint64_t>> cell(G * G, std::vector<uint64_t>(2, 0));
    int n_all = 0;
    for (size_t k = 0; k < cands.size() && k < 128; k++) {
        const Cand &c = cands[k];
        float cc[3], r2 = 0.f, d2 = 0.f;
        for (int a = 0; a < 3; a++) { cc[a] = 0.5f * (c.lo[a] + c.hi[a]) - org[a]; float e = 0.5f * (c.hi[a] - c.lo[a]); r2 += e * e; d2 += cc[a] * cc[a]; }
        int x0 = 0, x1 = G - 1, y0 = 0, y1 = G - 1;
        if (d2 > r2 * 1.0001f) {
            float d = std::sqrt(d2), sina = std::sqrt(r2) / d, cosa = std::sqrt(std::max(0.f, 1.f - sina * sina));
            float rho = std::sqrt(std::max(0.f, 2.f - 2.f * cosa)) * 1.001f + 1e-4f;
            float ax = (fr.right.x * cc[0] + fr.right.y * cc[1] + fr.right.z * cc[2]) / d, ay = (fr.up.x * cc[0] + fr.up.y * cc[1] + fr.up.z * cc[2]) / d;
            x0 = std::max(0, (int)std::floor((ax - rho + 1.f) * 0.5f * G)); x1 = std::min(G - 1, (int)std::floor((ax + rho + 1.f) * 0.5f * G));
            y0 = std::max(0, (int)std::floor((ay - rho + 1.f) * 0.5f * G)); y1 = std::min(G - 1, (int)std::floor((ay + rho + 1.f) * 0.5f * G));
        } else n_all++;
        for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) cell[y * G + x][k >> 6] |= 1ull << (k & 63);
    }
    double own = 0, uni = 0; uint64_t missed = 0;
    for (uint32_t base = 0; base + 32 <= nrays; base += 32) {
        uint64_t u[2] = {0, 0};
        for (int l = 0; l < 32; l++) {
            const float *d = local_dirs + 3 * (size_t)(base + l);
            int cx = std::min(G - 1, std::max(0, (int)std::floor((d[0] + 1.f) * 0.5f * G))), cy = std::min(G - 1, std::max(0, (int)std::floor((d[1] + 1.f) * 0.5f * G)));
            const auto &m = cell[cy * G + cx];
            own += __builtin_popcountll(m[0]) + __builtin_popcountll(m[1]);
            u[0] |= m[0]; u[1] |= m[1];
            // verify conservativeness: every candidate the ray's box test hits must be in its cell mask
            f3 wd = to_world(fr, mk3(d[0], d[1], d[2]));
            float id[3] = { safe_rcp(wd.x), safe_rcp(wd.y), safe_rcp(wd.z) };
            for (size_t k = 0; k < cands.size() && k < 128; k++) {
                float t0 = 0.f, t1 = INFINITY;
                for (int a = 0; a < 3; a++) { float ta = (cands[k].lo[a] - org[a]) * id[a], tb = (cands[k].hi[a] - org[a]) * id[a]; t0 = std::max(t0, std::min(ta, tb)); t1 = std::min(t1, std::max(ta, tb)); }
                if (t0 <= t1 && !((m[k >> 6] >> (k & 63)) & 1)) missed++;
            }
        }
        uni += 32.0 * (__builtin_popcountll(u[0]) + __builtin_popcountll(u[1]));
    }
    out[0] = (double)cands.size(); out[1] = own / nrays; out[2] = uni / nrays; out[3] = (double)missed; out[4] = n_all;
}

// ---- per-item bounds of the horizon map (prt_b200/csrc/horizon_math.cuh compiled as plain C++) ----------------------------------
// out[0] = first bin, out[1] = last bin (unwrapped), value = bound of sin(elevation) incl. margin (<= 0: the item is empty)
extern "C" void hc_frame(const float *n, float *out9) {
    const Frame f = make_frame(mk3(n[0], n[1], n[2]));
    out9[0] = f.right.x; out9[1] = f.right.y; out9[2] = f.right.z; out9[3] = f.up.x; out9[4] = f.up.y; out9[5] = f.up.z;
    out9[6] = f.n.x; out9[7] = f.n.y; out9[8] = f.n.z;
}
extern "C" float hc_hz_pang(float x, float y) { return hz_pang(x, y); }
// triangle given in the LOCAL frame of the origin (z = height above the tangent plane)
extern "C" void hc_hz_triangle(const float *q, uint32_t n, int *bins, float *val) {
    for (uint32_t i = 0; i < n; i++) {
        const float *t = q + 9 * (size_t)i;
        const HzItem it = hz_triangle(mk3(t[0], t[1], t[2]), mk3(t[3], t[4], t[5]), mk3(t[6], t[7], t[8]));
        bins[2 * i] = it.b0; bins[2 * i + 1] = it.b1; val[i] = it.v;
    }
}
// boxes in WORLD space relative to the origin: centre c, half extents e; n = vertex normal (unit).  which: 0 cheap (cone clamped
// by the slab bound), 1 exact box bound, 2 cone alone (only defined for d2 > 1.05 r2; otherwise the unbounded item)
extern "C" void hc_hz_box(const float *c, const float *e, const float *nrm, uint32_t n, int which, int *bins, float *val) {
    for (uint32_t i = 0; i < n; i++) {
        const f3 cc = mk3(c[3 * i], c[3 * i + 1], c[3 * i + 2]), ee = mk3(e[3 * i], e[3 * i + 1], e[3 * i + 2]);
        const Frame fr = make_frame(mk3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        const float r2 = ee.x * ee.x + ee.y * ee.y + ee.z * ee.z, d2 = cc.x * cc.x + cc.y * cc.y + cc.z * cc.z;
        HzItem it;
        if (which == 0) it = hz_cheap_box(cc, ee, r2, d2, fr);
        else if (which == 1) it = hz_box(cc, ee, fr);
        else it = d2 > 1.05f * r2 ? hz_sphere(cc, r2, d2, fr) : hz_item(0.f, 0.f, true, 1.0f);
        bins[2 * i] = it.b0; bins[2 * i + 1] = it.b1; val[i] = it.v;
    }
}

// box cut by an oriented slab (hz_slab_value, horizon_math.cuh): boxes as in hc_hz_box; slab[5i..] = m (3), L0, U0 relative to the origin
extern "C" void hc_hz_slab(const float *c, const float *e, const float *nrm, const float *slab, uint32_t n, float *val) {
    for (uint32_t i = 0; i < n; i++) {
        const f3 cc = mk3(c[3 * i], c[3 * i + 1], c[3 * i + 2]), ee = mk3(e[3 * i], e[3 * i + 1], e[3 * i + 2]);
        const Frame fr = make_frame(mk3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        const float *s = slab + 5 * (size_t)i;
        val[i] = hz_slab_value(cc, ee, fr.n, mk3(s[0], s[1], s[2]), s[3], s[4]);
    }
}
// the builder's slabs: out[8 * node ..] = m (3), d0, d1, 0, 0, 0
extern "C" uint32_t hc_slabs(void *h, float *out, uint32_t cap) {
    HostBVH8 *b = (HostBVH8 *)h;
    const uint32_t n = std::min(cap, b->n_nodes);
    if (out && b->slabs) std::memcpy(out, b->slabs, (size_t)n * sizeof(Slab32));
    return b->slabs ? b->n_nodes : 0u;
}
// triangle range below every node: out[2 * node] = first triangle (index into the emitted Tri48 array), out[2 * node + 1] = count
extern "C" void hc_node_tri_ranges(void *h, uint32_t *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    // depth-first emission: children are emitted after their parent; a reverse sweep accumulates the ranges
    std::vector<uint32_t> lo(b->n_nodes, 0xFFFFFFFFu), hi(b->n_nodes, 0u);
    for (uint32_t x = b->n_nodes; x-- > 0;) {
        const Node8 &nd = b->nodes[x];
        uint32_t rank = 0;
        for (int s = 0; s < 8; s++) {
            if (!nd.meta[s]) continue;
            if ((nd.imask >> s) & 1) { const uint32_t ch = nd.child_base + rank++; lo[x] = std::min(lo[x], lo[ch]); hi[x] = std::max(hi[x], hi[ch]); }
            else {
                const uint32_t t0 = nd.tri_base + (nd.meta[s] & 31u), cnt = (uint32_t)__builtin_popcount(nd.meta[s] >> 5);
                lo[x] = std::min(lo[x], t0); hi[x] = std::max(hi[x], t0 + cnt);
            }
        }
    }
    for (uint32_t x = 0; x < b->n_nodes; x++) { out[2 * x] = lo[x]; out[2 * x + 1] = hi[x] > lo[x] ? hi[x] - lo[x] : 0u; }
}
// emitted triangles: out[9 * t ..] = v0, v0 + e1, v0 + e2
extern "C" void hc_tris(void *h, float *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    for (uint32_t t = 0; t < b->n_tris; t++) {
        const Tri48 &T = b->tris[t];
        float *o = out + 9 * (size_t)t;
        o[0] = T.v0x; o[1] = T.v0y; o[2] = T.v0z; o[3] = T.v0x + T.e1x; o[4] = T.v0y + T.e1y; o[5] = T.v0z + T.e1z;
        o[6] = T.v0x + T.e2x; o[7] = T.v0y + T.e2y; o[8] = T.v0z + T.e2z;
    }
}

// the builder's fourth slab axis: out[12 * node ..] = M (3), D0, then qlo[8], qhi[8] as floats would not fit: raw 32-byte records instead
extern "C" uint32_t hc_dops(void *h, uint8_t *out, uint32_t cap) {
    HostBVH8 *b = (HostBVH8 *)h;
    const uint32_t n = std::min(cap, b->n_nodes);
    if (out && b->dops) std::memcpy(out, b->dops, (size_t)n * sizeof(Dop32));
    return b->dops ? b->n_nodes : 0u;
}
// per node and slot: the triangle range of the child (first, count; count 0 = empty slot): out[16 * node + 2 * slot ..]
extern "C" void hc_child_tri_ranges(void *h, uint32_t *out) {
    HostBVH8 *b = (HostBVH8 *)h;
    std::vector<uint32_t> lo(b->n_nodes, 0xFFFFFFFFu), hi(b->n_nodes, 0u);
    for (uint32_t x = b->n_nodes; x-- > 0;) {
        const Node8 &nd = b->nodes[x];
        uint32_t rank = 0;
        for (int s = 0; s < 8; s++) {
            out[16 * (size_t)x + 2 * s] = 0; out[16 * (size_t)x + 2 * s + 1] = 0;
            if (!nd.meta[s]) continue;
            uint32_t t0, t1;
            if ((nd.imask >> s) & 1) { const uint32_t ch = nd.child_base + rank++; t0 = lo[ch]; t1 = hi[ch]; }
            else { t0 = nd.tri_base + (nd.meta[s] & 31u); t1 = t0 + (uint32_t)__builtin_popcount(nd.meta[s] >> 5); }
            out[16 * (size_t)x + 2 * s] = t0; out[16 * (size_t)x + 2 * s + 1] = t1 - t0;
            lo[x] = std::min(lo[x], t0); hi[x] = std::max(hi[x], t1);
        }
    }
}
