/*
 * oracle/bvh.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for what the reference delegates to Intel Embree 3 (not in /root/reference):
 * RTScene (raytracing.cpp:58-99, light_probe.cpp:44-93), rtcIntersect1 (raytracing.cpp:254,
 * light_probe.cpp:119) and rtcOccluded1 (light_probe.cpp:128).  A binary binned-SAH BVH over
 * padded triangle boxes with scalar stack traversal; the ray/triangle decision itself is the
 * pinned test in arith.h, so results do not depend on the tree (checked against brute force).
 * Deliberately a different tree from the product's 8-wide compressed BVH.
 */
#include "prt_oracle.h"
#include "arith.h"
#include <stdlib.h>
#include <float.h>

typedef struct { v3 v0, e1, e2; uint32_t prim; } otri;
typedef struct {
    float lo[3], hi[3];
    uint32_t left;   /* internal: index of left child (right = left+1); leaf: first triangle */
    uint32_t count;  /* 0 = internal */
} onode;

struct prt_o_scene {
    uint32_t n_tris, n_nodes;
    otri *tris;   /* leaf order */
    onode *nodes;
    float pad;
};

typedef struct { float lo[3], hi[3], c[3]; uint32_t id; } oref;

#define NBINS 16
#define LEAF_MAX 4

static inline float box_area(const float lo[3], const float hi[3]) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

static void build(struct prt_o_scene *sc, oref *refs, uint32_t n) {
    /* explicit work stack: (node index, first, count) */
    typedef struct { uint32_t node, first, count; } job;
    size_t cap = 64;
    job *stack = (job *)malloc(cap * sizeof(job));
    size_t sp = 0;
    sc->nodes = (onode *)malloc(sizeof(onode) * (size_t)(2 * (size_t)n + 1));
    sc->n_nodes = 1;
    stack[sp++] = (job){ 0, 0, n };
    while (sp) {
        job j = stack[--sp];
        onode *nd = &sc->nodes[j.node];
        float clo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, chi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        for (int a = 0; a < 3; a++) { nd->lo[a] = FLT_MAX; nd->hi[a] = -FLT_MAX; }
        for (uint32_t i = j.first; i < j.first + j.count; i++) {
            const oref *r = &refs[i];
            for (int a = 0; a < 3; a++) {
                if (r->lo[a] < nd->lo[a]) nd->lo[a] = r->lo[a];
                if (r->hi[a] > nd->hi[a]) nd->hi[a] = r->hi[a];
                if (r->c[a] < clo[a]) clo[a] = r->c[a];
                if (r->c[a] > chi[a]) chi[a] = r->c[a];
            }
        }
        if (j.count <= LEAF_MAX) { nd->left = j.first; nd->count = j.count; continue; }
        /* binned SAH over the centroid box */
        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        for (int a = 0; a < 3; a++) {
            float ext = chi[a] - clo[a];
            if (!(ext > 0.0f)) continue;
            float scale = (float)NBINS / ext;
            uint32_t cnt[NBINS] = { 0 };
            float blo[NBINS][3], bhi[NBINS][3];
            for (int b = 0; b < NBINS; b++)
                for (int k = 0; k < 3; k++) { blo[b][k] = FLT_MAX; bhi[b][k] = -FLT_MAX; }
            for (uint32_t i = j.first; i < j.first + j.count; i++) {
                const oref *r = &refs[i];
                int b = (int)((r->c[a] - clo[a]) * scale);
                if (b >= NBINS) b = NBINS - 1;
                if (b < 0) b = 0;
                cnt[b]++;
                for (int k = 0; k < 3; k++) {
                    if (r->lo[k] < blo[b][k]) blo[b][k] = r->lo[k];
                    if (r->hi[k] > bhi[b][k]) bhi[b][k] = r->hi[k];
                }
            }
            float la[NBINS], ra[NBINS];
            uint32_t lc[NBINS], rc[NBINS];
            float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
            uint32_t c = 0;
            for (int b = 0; b < NBINS; b++) {
                for (int k = 0; k < 3; k++) { if (blo[b][k] < lo[k]) lo[k] = blo[b][k]; if (bhi[b][k] > hi[k]) hi[k] = bhi[b][k]; }
                c += cnt[b]; lc[b] = c; la[b] = c ? box_area(lo, hi) : 0.0f;
            }
            for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
            c = 0;
            for (int b = NBINS - 1; b >= 0; b--) {
                for (int k = 0; k < 3; k++) { if (blo[b][k] < lo[k]) lo[k] = blo[b][k]; if (bhi[b][k] > hi[k]) hi[k] = bhi[b][k]; }
                c += cnt[b]; rc[b] = c; ra[b] = c ? box_area(lo, hi) : 0.0f;
            }
            for (int b = 0; b < NBINS - 1; b++) {
                if (!lc[b] || !rc[b + 1]) continue;
                float cost = la[b] * (float)lc[b] + ra[b + 1] * (float)rc[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        uint32_t mid;
        if (best_axis < 0) {
            mid = j.first + j.count / 2; /* all centroids coincide: split the list */
        } else {
            int a = best_axis;
            float scale = (float)NBINS / (chi[a] - clo[a]);
            uint32_t i = j.first, k = j.first + j.count;
            while (i < k) {
                int b = (int)((refs[i].c[a] - clo[a]) * scale);
                if (b >= NBINS) b = NBINS - 1;
                if (b < 0) b = 0;
                if (b <= best_bin) i++;
                else { oref t = refs[i]; refs[i] = refs[--k]; refs[k] = t; }
            }
            mid = i;
            if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
        }
        uint32_t l = sc->n_nodes; sc->n_nodes += 2;
        nd->left = l; nd->count = 0;
        if (sp + 2 > cap) { cap *= 2; stack = (job *)realloc(stack, cap * sizeof(job)); }
        stack[sp++] = (job){ l + 1, mid, j.first + j.count - mid };
        stack[sp++] = (job){ l, j.first, mid - j.first };
    }
    free(stack);
}

prt_o_scene *prt_o_scene_create(const float *pos, size_t stride, uint32_t nv, const uint32_t *idx, uint32_t nt) {
    if (!pos || !idx || nt == 0) return NULL;
    if (stride == 0) stride = 12;
    struct prt_o_scene *sc = (struct prt_o_scene *)calloc(1, sizeof(*sc));
    sc->n_tris = nt;
    oref *refs = (oref *)malloc(sizeof(oref) * (size_t)nt);
    otri *src = (otri *)malloc(sizeof(otri) * (size_t)nt);
    float amax = 0.0f;
    for (uint32_t i = 0; i < nv; i++) {
        const float *p = (const float *)((const char *)pos + (size_t)i * stride);
        for (int a = 0; a < 3; a++) { float f = fabsf(p[a]); if (f > amax) amax = f; }
    }
    /* conservative padding so that the box test can never cull what the pinned triangle test accepts */
    sc->pad = amax * 1.6e-5f + 1e-30f;
    for (uint32_t t = 0; t < nt; t++) {
        v3 p[3];
        for (int k = 0; k < 3; k++) {
            uint32_t vi = idx[3 * (size_t)t + k];
            if (vi >= nv) { free(refs); free(src); free(sc); return NULL; }
            const float *q = (const float *)((const char *)pos + (size_t)vi * stride);
            p[k] = v3_make(q[0], q[1], q[2]);
        }
        src[t].v0 = p[0]; src[t].e1 = v3_sub(p[1], p[0]); src[t].e2 = v3_sub(p[2], p[0]); src[t].prim = t;
        oref *r = &refs[t];
        r->id = t;
        const float *c0 = &p[0].x, *c1 = &p[1].x, *c2 = &p[2].x;
        for (int a = 0; a < 3; a++) {
            float lo = fminf(c0[a], fminf(c1[a], c2[a])), hi = fmaxf(c0[a], fmaxf(c1[a], c2[a]));
            r->lo[a] = lo - sc->pad; r->hi[a] = hi + sc->pad;
            r->c[a] = 0.5f * (lo + hi);
        }
    }
    build(sc, refs, nt);
    sc->tris = (otri *)malloc(sizeof(otri) * (size_t)nt);
    for (uint32_t i = 0; i < nt; i++) sc->tris[i] = src[refs[i].id];
    free(src); free(refs);
    return sc;
}

void prt_o_scene_destroy(prt_o_scene *sc) {
    if (!sc) return;
    free(sc->tris); free(sc->nodes); free(sc);
}
uint32_t prt_o_scene_ntris(const prt_o_scene *sc) { return sc ? sc->n_tris : 0; }

static inline float safe_inv(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    return 1.0f / d;
}

/* slab test with slack; never used for the decision, only for culling */
static inline int box_hit(const onode *n, const float o[3], const float id[3], float tnear, float tfar, float *tmin_out) {
    float t0 = tnear, t1 = tfar;
    for (int a = 0; a < 3; a++) {
        float ta = (n->lo[a] - o[a]) * id[a], tb = (n->hi[a] - o[a]) * id[a];
        float tn = fminf(ta, tb), tf = fmaxf(ta, tb);
        if (tn > t0) t0 = tn;
        if (tf < t1) t1 = tf;
    }
    *tmin_out = t0;
    return t0 <= t1 * 1.0000005f + 1e-30f;
}

int prt_o_any_hit(const prt_o_scene *sc, const float org[3], const float dir[3], float tnear, float tfar, int use_bvh) {
    v3 O = v3_make(org[0], org[1], org[2]), D = v3_make(dir[0], dir[1], dir[2]);
    if (!use_bvh) {
        for (uint32_t i = 0; i < sc->n_tris; i++)
            if (prt_tri_test(O, D, tnear, tfar, sc->tris[i].v0, sc->tris[i].e1, sc->tris[i].e2, NULL)) return 1;
        return 0;
    }
    float id[3] = { safe_inv(dir[0]), safe_inv(dir[1]), safe_inv(dir[2]) };
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const onode *n = &sc->nodes[stack[--sp]];
        float tm;
        if (!box_hit(n, org, id, tnear, tfar, &tm)) continue;
        if (n->count) {
            for (uint32_t i = n->left; i < n->left + n->count; i++)
                if (prt_tri_test(O, D, tnear, tfar, sc->tris[i].v0, sc->tris[i].e1, sc->tris[i].e2, NULL)) return 1;
        } else {
            /* nearer child first: occluders close to the origin are the likely ones */
            const onode *l = &sc->nodes[n->left];
            int ax = 0; float e = -1.f;
            for (int a = 0; a < 3; a++) { float d = n->hi[a] - n->lo[a]; if (d > e) { e = d; ax = a; } }
            float cl = l->lo[ax] + l->hi[ax], cr = l[1].lo[ax] + l[1].hi[ax];
            int left_first = (dir[ax] >= 0.0f) ? (cl <= cr) : (cl >= cr);
            if (sp + 2 > 128) return -1;
            if (left_first) { stack[sp++] = n->left + 1; stack[sp++] = n->left; }
            else { stack[sp++] = n->left; stack[sp++] = n->left + 1; }
        }
    }
    return 0;
}

int prt_o_closest_hit(const prt_o_scene *sc, const float org[3], const float dir[3], float tnear, float tfar,
                      int use_bvh, float *t_out, uint32_t *prim_out, float ng[3]) {
    v3 O = v3_make(org[0], org[1], org[2]), D = v3_make(dir[0], dir[1], dir[2]);
    float best_t = INFINITY; uint32_t best_prim = 0xFFFFFFFFu; int found = 0; uint32_t best_i = 0;
    /* a candidate is valid w.r.t. the *initial* interval; the winner is min (t, prim): order independent */
    if (!use_bvh) {
        for (uint32_t i = 0; i < sc->n_tris; i++) {
            float t;
            if (prt_tri_test(O, D, tnear, tfar, sc->tris[i].v0, sc->tris[i].e1, sc->tris[i].e2, &t)) {
                uint32_t p = sc->tris[i].prim;
                if (!found || t < best_t || (t == best_t && p < best_prim)) { best_t = t; best_prim = p; best_i = i; found = 1; }
            }
        }
    } else {
        float id[3] = { safe_inv(dir[0]), safe_inv(dir[1]), safe_inv(dir[2]) };
        uint32_t stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const onode *n = &sc->nodes[stack[--sp]];
            float tm;
            float lim = found ? best_t : tfar;
            if (!box_hit(n, org, id, tnear, lim, &tm)) continue;
            if (n->count) {
                for (uint32_t i = n->left; i < n->left + n->count; i++) {
                    float t;
                    if (prt_tri_test(O, D, tnear, tfar, sc->tris[i].v0, sc->tris[i].e1, sc->tris[i].e2, &t)) {
                        uint32_t p = sc->tris[i].prim;
                        if (!found || t < best_t || (t == best_t && p < best_prim)) { best_t = t; best_prim = p; best_i = i; found = 1; }
                    }
                }
            } else {
                if (sp + 2 > 128) return -1;
                const onode *l = &sc->nodes[n->left];
                int ax = 0; float e = -1.f;
                for (int a = 0; a < 3; a++) { float d = n->hi[a] - n->lo[a]; if (d > e) { e = d; ax = a; } }
                float cl = l->lo[ax] + l->hi[ax], cr = l[1].lo[ax] + l[1].hi[ax];
                int left_first = (dir[ax] >= 0.0f) ? (cl <= cr) : (cl >= cr);
                if (left_first) { stack[sp++] = n->left + 1; stack[sp++] = n->left; }
                else { stack[sp++] = n->left; stack[sp++] = n->left + 1; }
            }
        }
    }
    if (!found) return 0;
    if (t_out) *t_out = best_t;
    if (prim_out) *prim_out = best_prim;
    if (ng) { v3 g = v3_cross(sc->tris[best_i].e1, sc->tris[best_i].e2); ng[0] = g.x; ng[1] = g.y; ng[2] = g.z; }
    return 1;
}
