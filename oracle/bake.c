/*
 * oracle/bake.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Restates the reference's per-vertex diffuse SH transfer bake:
 *   bake_SH        src/raytracing/raytracing.cpp:320-360  (estimator, origin offset, 1/S normalisation)
 *   renderSH       src/raytracing/raytracing.cpp:228-278  (path logic, cut-offs, bounce offsets)
 *   sampling       src/raytracing/raytracing.cpp:101-107,130-160
 *   lightSH        src/raytracing/raytracing.cpp:224-227  (sh-space permutation (z,x,y))
 * Documented departures: one stratified (u,v) table per run shared by every vertex and
 * coefficient (Philox stream 0) instead of a fresh mt19937 jitter per (vertex, coefficient, sample);
 * counter-based bounce randoms keyed (vertex, sample, bounce); one trace per sample shared by all
 * coefficients (faithful=1 re-traces per coefficient, identical results); accumulation in double
 * (the reference accumulates in float, raytracing.cpp:348).
 */
#include "prt_oracle.h"
#include "arith.h"
#include "philox.h"
#include "sh.h"
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

void prt_o_sh_eval(int order, int cs_phase, const float d[3], float *out) { prt_sh_eval(order, cs_phase, d[0], d[1], d[2], out); }
void prt_o_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { prt_philox4x32_10(ctr, key, out); }
void prt_o_sincos2pi(float v, float *s, float *c) { prt_sincos2pi(v, s, c); }
void prt_o_cosine_world(float u, float v, const float N[3], float local[3], float world[3], float frame9[9], float *pdf) {
    v3 l = prt_cosine_local(u, v);
    frame3 f = prt_frame(v3_make(N[0], N[1], N[2]));
    v3 w = prt_to_world(&f, l);
    local[0] = l.x; local[1] = l.y; local[2] = l.z;
    world[0] = w.x; world[1] = w.y; world[2] = w.z;
    frame9[0] = f.right.x; frame9[1] = f.right.y; frame9[2] = f.right.z; frame9[3] = f.up.x; frame9[4] = f.up.y; frame9[5] = f.up.z;
    frame9[6] = f.n.x; frame9[7] = f.n.y; frame9[8] = f.n.z;
    *pdf = l.z / PRT_PI_F;
}
int prt_o_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

void prt_o_sample_table(const prt_o_bake_params *p, float *uv, float *dirs) {
    int Ru = p->samples_u, Rv = p->samples_v;
    for (int i = 0; i < Ru; i++)
        for (int j = 0; j < Rv; j++) {
            int s = i * Rv + j;
            float x1 = 0.5f, x2 = 0.5f;
            if (p->jitter) prt_rand2(p->seed, (uint32_t)s, 0u, 0u, 0u, &x1, &x2);
            /* raytracing.cpp:338-339: x = (i+random())/res -> radius^2, y = (j+random())/res -> angle */
            float u = ((float)i + x1) / (float)Ru;
            float v = ((float)j + x2) / (float)Rv;
            v3 l = prt_cosine_local(u, v);
            if (uv) { uv[2 * s] = u; uv[2 * s + 1] = v; }
            if (dirs) { dirs[3 * s] = l.x; dirs[3 * s + 1] = l.y; dirs[3 * s + 2] = l.z; }
        }
}

typedef struct {
    const prt_o_scene *sc;
    const char *pos, *nrm;
    size_t stride;
    uint32_t n, base;
    const prt_o_bake_params *p;
    const float *dirs;
    float *out;
    uint32_t *vis;
    int faithful;
    volatile uint32_t *next;
    uint64_t rays, segs;
} job_t;

/* One path (renderSH).  Returns the weight Lw.x carried to the environment and the final
 * direction (0 weight when the path is absorbed). *primary_visible = 1 iff the first segment escaped. */
static float trace_path(const job_t *J, v3 pos, v3 dir, int depth, uint32_t vid, uint32_t s,
                        v3 *final_dir, int *primary_visible, uint64_t *segs) {
    const prt_o_bake_params *p = J->p;
    float Lw[3] = { 1.f, 1.f, 1.f };
    float tnear = 0.0f;
    *primary_visible = 0;
    for (int i = 0; i < depth; i++) {
        if (fmaxf(Lw[0], fmaxf(Lw[1], Lw[2])) < 0.01f) break;          /* :249 */
        float o[3] = { pos.x, pos.y, pos.z }, d[3] = { dir.x, dir.y, dir.z };
        (*segs)++;
        if (i == depth - 1) {
            /* last segment: only hit/miss is observable (a hit can never reach the environment) */
            if (!prt_o_any_hit(J->sc, o, d, tnear, INFINITY, 1)) { if (i == 0) *primary_visible = 1; *final_dir = dir; return Lw[0]; }
            return 0.0f;
        }
        float t; uint32_t prim; float ng[3];
        if (!prt_o_closest_hit(J->sc, o, d, tnear, INFINITY, 1, &t, &prim, ng)) {   /* :257-261 */
            if (i == 0) *primary_visible = 1;
            *final_dir = dir; return Lw[0];
        }
        v3 n = v3_normalize(v3_make(ng[0], ng[1], ng[2]));                  /* :263-264 */
        if (v3_dot(dir, n) >= -1e-4f) return 0.0f;                          /* :265 */
        pos = v3_madd(pos, t, dir);                                         /* :266 */
        float u, v;
        prt_rand2(p->seed, vid, s, (uint32_t)i, 1u, &u, &v);                /* :267 random(), random() */
        v3 l = prt_cosine_local(u, v);
        float pdf = l.z / PRT_PI_F;
        frame3 f = prt_frame(n);
        dir = prt_to_world(&f, l);
        if (pdf <= 1e-4f) return 0.0f;                                      /* :269 */
        Lw[0] *= p->albedo[0]; Lw[1] *= p->albedo[1]; Lw[2] *= p->albedo[2]; /* :271 */
        float sign = v3_dot(dir, n) < 0.0f ? -1.0f : 1.0f;                  /* :273 */
        pos = v3_madd(pos, sign * p->bounce_eps, dir);                      /* :274 */
        tnear = p->bounce_eps;                                              /* :275 */
    }
    return 0.0f;
}

static void bake_vertex(job_t *J, uint32_t i) {
    const prt_o_bake_params *p = J->p;
    const float *P = (const float *)(J->pos + (size_t)i * J->stride);
    const float *Nn = (const float *)(J->nrm + (size_t)i * J->stride);
    int n2 = p->order * p->order;
    int S = p->samples_u * p->samples_v;
    int words = (S + 31) / 32;
    float *out = J->out + (size_t)i * n2;
    uint32_t vid = J->base + i;
    v3 N = v3_make(Nn[0], Nn[1], Nn[2]);
    if (p->mode == PRT_O_UNSHADOWED_ANALYTIC) {
        /* scene/model.cpp:29-31 hint: rotate_cos_lobe(norm) * INV_PI */
        float y[25];
        prt_sh_eval(p->order, p->cs_phase, N.z, N.x, N.y, y);
        for (int l = 0, k = 0; l < p->order; l++)
            for (int m = -l; m <= l; m++, k++) out[k] = PRT_COS_LOBE[l] * y[k];
        return;
    }
    frame3 f = prt_frame(N);
    v3 org = v3_madd(v3_make(P[0], P[1], P[2]), p->origin_eps, N);          /* :343 */
    int depth = p->mode == PRT_O_INTERREFLECT ? p->bounces + 1 : 1;          /* :345 */
    double acc[25] = { 0 };
    uint32_t *vis = J->vis ? J->vis + (size_t)i * words : NULL;
    if (vis) for (int w = 0; w < words; w++) vis[w] = 0;
    int passes = J->faithful ? n2 : 1;
    for (int pass = 0; pass < passes; pass++) {
        for (int s = 0; s < S; s++) {
            v3 l = v3_make(J->dirs[3 * s], J->dirs[3 * s + 1], J->dirs[3 * s + 2]);
            v3 wi = prt_to_world(&f, l);                                    /* :340 */
            float w; v3 fd = wi; int pv = 0;
            if (p->mode == PRT_O_UNSHADOWED) { w = 1.0f; pv = 1; }
            else { J->rays++; w = trace_path(J, org, wi, depth, vid, (uint32_t)s, &fd, &pv, &J->segs); }
            if (vis && pass == 0 && pv) vis[s >> 5] |= 1u << (s & 31);
            if (w != 0.0f) {
                float y[25];
                prt_sh_eval(p->order, p->cs_phase, fd.z, fd.x, fd.y, y);    /* :226 */
                if (J->faithful) acc[pass] += (double)(w * y[pass]);
                else for (int k = 0; k < n2; k++) acc[k] += (double)(w * y[k]);
            }
        }
    }
    for (int k = 0; k < n2; k++) out[k] = (float)(acc[k] / (double)S);      /* :350 */
}

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    for (;;) {
        uint32_t b = __sync_fetch_and_add(J->next, 16u);
        if (b >= J->n) break;
        uint32_t e = b + 16u < J->n ? b + 16u : J->n;
        for (uint32_t i = b; i < e; i++) bake_vertex(J, i);
    }
    return NULL;
}

int prt_o_bake_transfer(const prt_o_scene *sc, const float *pos, const float *nrm, size_t stride, uint32_t n,
                        uint32_t vertex_id_base, const prt_o_bake_params *p, float *out, uint32_t *vis,
                        int n_threads, int faithful, uint64_t *counters) {
    if (!p || !pos || !nrm || !out) return -1;
    if (p->order < 1 || p->order > 5 || p->samples_u < 1 || p->samples_v < 1) return -2;
    if (!sc && (p->mode == PRT_O_SHADOWED || p->mode == PRT_O_INTERREFLECT)) return -3;
    if (stride == 0) stride = 12;
    int S = p->samples_u * p->samples_v;
    float *dirs = (float *)malloc(sizeof(float) * 3 * (size_t)S);
    prt_o_sample_table(p, NULL, dirs);
    if (n_threads <= 0) n_threads = prt_o_hw_threads();
    if ((uint32_t)n_threads > n) n_threads = n ? (int)n : 1;
    volatile uint32_t next = 0;
    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (job_t){ sc, (const char *)pos, (const char *)nrm, stride, n, vertex_id_base, p, dirs, out, vis, faithful, &next, 0, 0 };
        if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    uint64_t rays = jobs[0].rays, segs = jobs[0].segs;
    for (int t = 1; t < n_threads; t++) { pthread_join(th[t], NULL); rays += jobs[t].rays; segs += jobs[t].segs; }
    if (counters) { counters[0] = rays; counters[1] = segs; }
    free(th); free(jobs); free(dirs);
    return 0;
}

/* ---- bake_SH in the reference's own order, fed with the reference's own random sequence ---------------------------------------
 * raytracing.cpp:320-360 + renderSH :228-278 literally: vertices in order, one pass per coefficient (l, m), every pass re-jittered
 * (x = (i + random()) / res, y = (j + random()) / res), a closest-hit query per path segment, two more random() per front-face
 * hit -- also on the last segment, whose bounce direction is sampled before the loop ends --, float accumulation, division by
 * res * res.  `rnd` is the sequence of random() values (the test replays std::mt19937 + uniform_real_distribution<float>, the
 * generator of raytracing.cpp:14-18; single-threaded, so the consumption order is the program order); u_first: which of the two
 * random() calls in `cosineSampleHemisphere(random(), random(), normal)` (raytracing.cpp:267; unspecified by C++) lands in u.
 * Used only to pin this oracle against the reference's bake_SH compiled here with Embree replaced by this oracle's tracer
 * (oracle/ref_bake.cpp); shares frame / sampling / SH / tracer with prt_o_bake_transfer.  Returns random() values consumed.
 * rnd == NULL: the same literal loop, but drawing what prt_o_bake_transfer draws (Philox: one jitter pair per stratum shared by all
 * passes, bounce pairs keyed (vertex, sample, segment), seed / vertex base as given) -- the link between this literal restatement
 * and the production oracle, which must then agree up to the float accumulation. */
uint64_t prt_o_bake_transfer_ref_order(const prt_o_scene *sc, const float *pos, const float *nrm, size_t stride, uint32_t n, int order, int res,
                                       int max_path_length, const float albedo[3], int cs_phase, const float *rnd, uint64_t n_rnd, int u_first,
                                       uint32_t seed, uint32_t vid_base, float *out) {
    if (stride == 0) stride = 12;
    const int n2 = order * order, depth = max_path_length - 1;
    uint64_t at = 0;
#define NEXT_RND() (at < n_rnd ? rnd[at++] : (at++, 0.5f))
    for (uint32_t vi = 0; vi < n; vi++) {
        const float *P = (const float *)((const char *)pos + (size_t)vi * stride), *Nn = (const float *)((const char *)nrm + (size_t)vi * stride);
        const v3 N = v3_make(Nn[0], Nn[1], Nn[2]);
        const frame3 f = prt_frame(N);
        const v3 org = v3_madd(v3_make(P[0], P[1], P[2]), 1e-4f, N);                                   /* :343 */
        for (int k = 0; k < n2; k++) {
            float acc = 0.0f;
            for (int i = 0; i < res; i++)
                for (int j = 0; j < res; j++) {
                    float j1, j2;
                    if (rnd) { j1 = NEXT_RND(); j2 = NEXT_RND(); }
                    else prt_rand2(seed, (uint32_t)(i * res + j), 0u, 0u, 0u, &j1, &j2);
                    const float x = ((float)i + j1) / (float)res, y = ((float)j + j2) / (float)res;                /* :338-339 */
                    v3 dir = prt_to_world(&f, prt_cosine_local(x, y));
                    v3 p = org;
                    float Lw[3] = { 1.f, 1.f, 1.f }, tnear = 0.0f, L = 0.0f;
                    for (int s = 0; s < depth; s++) {
                        if (fmaxf(Lw[0], fmaxf(Lw[1], Lw[2])) < 0.01f) break;                          /* :249 */
                        float o[3] = { p.x, p.y, p.z }, d[3] = { dir.x, dir.y, dir.z }, t, ng[3]; uint32_t prim;
                        if (!prt_o_closest_hit(sc, o, d, tnear, INFINITY, 1, &t, &prim, ng)) {         /* :257-261 */
                            float yv[25];
                            prt_sh_eval(order, cs_phase, dir.z, dir.x, dir.y, yv);
                            L = Lw[0] * yv[k];
                            break;
                        }
                        const v3 nn = v3_normalize(v3_make(ng[0], ng[1], ng[2]));
                        if (v3_dot(dir, nn) >= -1e-4f) break;                                          /* :265 */
                        p = v3_madd(p, t, dir);                                                        /* :266 */
                        float r0, r1;                                                                  /* :267 */
                        if (rnd) { r0 = NEXT_RND(); r1 = NEXT_RND(); }
                        else { prt_rand2(seed, vid_base + vi, (uint32_t)(i * res + j), (uint32_t)s, 1u, &r0, &r1); u_first = 1; }
                        const v3 l = prt_cosine_local(u_first ? r0 : r1, u_first ? r1 : r0);
                        const frame3 fb = prt_frame(nn);
                        dir = prt_to_world(&fb, l);
                        if (l.z / PRT_PI_F <= 1e-4f) break;                                            /* :269 */
                        Lw[0] *= albedo[0]; Lw[1] *= albedo[1]; Lw[2] *= albedo[2];
                        const float sign = v3_dot(dir, nn) < 0.0f ? -1.0f : 1.0f;
                        p = v3_madd(p, sign * 1e-5f, dir);                                             /* :273-274 */
                        tnear = 1e-5f;
                    }
                    acc += L;                                                                          /* :348 */
                }
            out[(size_t)vi * n2 + k] = acc / (float)(res * res);                                       /* :350 */
        }
    }
#undef NEXT_RND
    return at;
}
