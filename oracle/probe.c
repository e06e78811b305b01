/*
 * oracle/probe.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Restates the reference's per-probe radiance-transfer capture and projection (BASELINE config 3):
 *   probe positions         src/sh/volume.cpp:83-90        (-size + 2*size/res*(idx+0.5), x fastest)
 *   capture                 src/sh/volume.cpp:229-271      (per texel: sky / back-face skips, cube_coord, solid angle,
 *                                                           sh-space direction, surfel cluster id, transfer += SH9*dOmega)
 *   surfel clustering       src/sh/volume.cpp:205-223      (floor(pos), signed principal axis of the normal)
 *   surfel table            src/sh/volume.cpp:301-312      (mean position, normalised mean normal)
 *   projection + pack       src/shaders/precomp_projectSH.comp:32-143
 *   Fibonacci directions    src/raytracing/light_probe.cpp:137-152 (get_dirs)
 * Departures: the reference rasterises a 64x64x6 G-buffer per probe with OpenGL and reads it back
 * (FP16 positions); here each direction is a closest-hit ray in FP32, the hit normal is the normalised geometric normal, and
 * surfel ids are the rank of the cluster key (x,y,z,direction lexicographic, i.e. std::map<std::array<int,4>> order) instead of
 * first-seen numbering -- parity with the reference's ids is defined up to that permutation.
 */
#include "prt_oracle.h"
#include "arith.h"
#include "sh.h"
#include <stdlib.h>
#include <string.h>

/* light_probe.cpp:137-152, computed in double, stored float */
void prt_o_fibonacci_dirs(int n, float *dirs) {
    double pi = 2 * acos(0.0), gold = 3 - sqrt(5.0);
    for (int i = 0; i < n; i++) {
        double z = 1 - ((double)i / (double)(n - 1)) * 2, theta = pi * i * gold;
        double r = sqrt(1 - z * z);
        dirs[3 * i] = (float)(cos(theta) * r); dirs[3 * i + 1] = (float)(sin(theta) * r); dirs[3 * i + 2] = (float)z;
    }
}
/* SH_function.h:96-112 generalised to res x res texels per face: direction through the texel centre + the reference's
 * solid angle 4/res^2/|c|^3 (volume.cpp:251-254) */
void prt_o_cube_dirs(int res, float *dirs, float *weights) {
    int k = 0;
    for (int f = 0; f < 6; f++)
        for (int y = 0; y < res; y++)
            for (int x = 0; x < res; x++, k++) {
                float u = (float)(((double)x + 0.5) / (double)res) * 2.0f - 1.0f, v = (float)(((double)y + 0.5) / (double)res) * 2.0f - 1.0f;
                float d[3];
                switch (f) {
                case 0: d[0] = 1.f; d[1] = -v; d[2] = -u; break;
                case 1: d[0] = -1.f; d[1] = -v; d[2] = u; break;
                case 2: d[0] = u; d[1] = 1.f; d[2] = v; break;
                case 3: d[0] = u; d[1] = -1.f; d[2] = -v; break;
                case 4: d[0] = u; d[1] = -v; d[2] = 1.f; break;
                default: d[0] = -u; d[1] = -v; d[2] = -1.f; break;
                }
                float w = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                w *= sqrtf(w);
                dirs[3 * k] = d[0]; dirs[3 * k + 1] = d[1]; dirs[3 * k + 2] = d[2];
                weights[k] = 4.0f / (float)res / (float)res / w;
            }
}

void prt_o_probe_positions(const int res[3], const float size[3], float *pos) {
    int k = 0;
    for (int z = 0; z < res[2]; z++)
        for (int y = 0; y < res[1]; y++)
            for (int x = 0; x < res[0]; x++, k++) {
                int id[3] = { x, y, z };
                for (int a = 0; a < 3; a++) {
                    float ds = 2.0f / (float)res[a] * size[a];
                    pos[3 * k + a] = -size[a] + ds * (0.5f + (float)id[a]);
                }
            }
}

/* volume.cpp:205-223 packed so that integer order == lexicographic (x,y,z,direction) order */
static int cluster_key(v3 pos, v3 n, uint64_t *key) {
    int dir = 0;
    float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
    if (ax > ay && ax > az) dir = n.x > 0 ? 0 : 1;
    if (ay > ax && ay > az) dir = n.y > 0 ? 2 : 3;
    if (az > ax && az > ay) dir = n.z > 0 ? 4 : 5;
    float fx = floorf(pos.x), fy = floorf(pos.y), fz = floorf(pos.z);
    if (!(fabsf(fx) < 32768.f && fabsf(fy) < 32768.f && fabsf(fz) < 32768.f)) return 0;
    uint64_t x = (uint64_t)((int)fx + 32768), y = (uint64_t)((int)fy + 32768), z = (uint64_t)((int)fz + 32768);
    *key = (x << 35) | (y << 19) | (z << 3) | (uint64_t)dir;
    return 1;
}

typedef struct { uint64_t key; uint32_t ray; float t; v3 n; v3 pos; } hitrec;
static int cmp_hit(const void *a, const void *b) {
    const hitrec *x = (const hitrec *)a, *y = (const hitrec *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->ray < y->ray ? -1 : (x->ray > y->ray);
}
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : (x > y); }

struct prt_o_csr {
    uint32_t n_probes, nnz, n_prim;
    uint32_t *range;     /* [n_probes][2] */
    uint32_t *ids;       /* [nnz] */
    float *transfer;     /* [nnz][9] */
    float *surfels;      /* [n_prim][6] mean position, normalised mean normal */
    uint64_t *keys;      /* [n_prim] sorted cluster keys */
    double *sums;        /* [n_prim][7] sum of hit positions, sum of hit normals, hit count (merging partial captures) */
};

prt_o_csr *prt_o_probe_capture(const prt_o_scene *sc, const float *probe_pos, uint32_t n_probes, const float *dirs,
                               const float *weights, uint32_t n_dirs) {
    struct prt_o_csr *c = (struct prt_o_csr *)calloc(1, sizeof(*c));
    c->n_probes = n_probes;
    c->range = (uint32_t *)calloc((size_t)n_probes * 2, 4);
    size_t cap = 1024, nnz = 0;
    uint64_t *ekey = (uint64_t *)malloc(cap * 8);
    float *etr = (float *)malloc(cap * 36);
    double *eacc = (double *)malloc(cap * 7 * 8);          /* per entry: sum pos, sum normal, count */
    hitrec *h = (hitrec *)malloc(sizeof(hitrec) * n_dirs);
    for (uint32_t p = 0; p < n_probes; p++) {
        v3 P = v3_make(probe_pos[3 * p], probe_pos[3 * p + 1], probe_pos[3 * p + 2]);
        uint32_t nh = 0;
        for (uint32_t r = 0; r < n_dirs; r++) {
            float t; uint32_t prim; float ng[3];
            if (!prt_o_closest_hit(sc, &probe_pos[3 * p], &dirs[3 * r], 0.0f, INFINITY, 1, &t, &prim, ng)) continue;   /* sky: volume.cpp:246 */
            v3 n = v3_normalize(v3_make(ng[0], ng[1], ng[2]));
            v3 D = v3_make(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
            v3 pos = v3_madd(P, t, D);
            v3 tp = v3_sub(pos, P);
            if (v3_dot(tp, n) > 0.0f) continue;                                                                     /* back face: :249 */
            uint64_t key;
            if (!cluster_key(pos, n, &key)) continue;
            h[nh].key = key; h[nh].ray = r; h[nh].t = t; h[nh].n = n; h[nh].pos = pos; nh++;
        }
        qsort(h, nh, sizeof(hitrec), cmp_hit);
        c->range[2 * p] = (uint32_t)nnz;
        for (uint32_t i = 0; i < nh;) {
            uint32_t j = i;
            float tr[9] = { 0 };
            double acc[7] = { 0 };
            while (j < nh && h[j].key == h[i].key) {
                v3 tp = v3_sub(h[j].pos, P);
                float m = fmaxf(fabsf(tp.x), fmaxf(fabsf(tp.y), fabsf(tp.z)));
                v3 cc = v3_make(tp.x / m, tp.y / m, tp.z / m);                                                       /* :250 */
                v3 d = v3_normalize(v3_make(cc.z, cc.x, cc.y));                                                      /* :255-256 */
                float y[9];
                prt_sh_eval(3, 0, d.x, d.y, d.z, y);
                float w = weights[h[j].ray];
                for (int k = 0; k < 9; k++) tr[k] += y[k] * w;                                                       /* :260 */
                acc[0] += h[j].pos.x; acc[1] += h[j].pos.y; acc[2] += h[j].pos.z;
                acc[3] += h[j].n.x; acc[4] += h[j].n.y; acc[5] += h[j].n.z; acc[6] += 1.0;
                j++;
            }
            if (nnz == cap) { cap *= 2; ekey = (uint64_t *)realloc(ekey, cap * 8); etr = (float *)realloc(etr, cap * 36); eacc = (double *)realloc(eacc, cap * 56); }
            ekey[nnz] = h[i].key; memcpy(etr + 9 * nnz, tr, 36); memcpy(eacc + 7 * nnz, acc, 56); nnz++;
            i = j;
        }
        c->range[2 * p + 1] = (uint32_t)nnz;
    }
    free(h);
    /* global ids = rank of the key among all distinct keys */
    uint64_t *sorted = (uint64_t *)malloc((nnz ? nnz : 1) * 8);
    memcpy(sorted, ekey, nnz * 8);
    qsort(sorted, nnz, 8, cmp_u64);
    uint32_t np = 0;
    for (size_t i = 0; i < nnz; i++) if (i == 0 || sorted[i] != sorted[i - 1]) sorted[np++] = sorted[i];
    c->nnz = (uint32_t)nnz; c->n_prim = np; c->keys = sorted;
    c->ids = (uint32_t *)malloc((nnz ? nnz : 1) * 4);
    c->transfer = etr;
    double *sacc = (double *)calloc((size_t)(np ? np : 1) * 7, 8);
    for (size_t i = 0; i < nnz; i++) {
        size_t lo = 0, hi = np;
        while (lo + 1 < hi) { size_t mid = (lo + hi) / 2; if (sorted[mid] <= ekey[i]) lo = mid; else hi = mid; }
        c->ids[i] = (uint32_t)lo;
        for (int k = 0; k < 7; k++) sacc[7 * lo + k] += eacc[7 * i + k];
    }
    c->surfels = (float *)malloc((size_t)(np ? np : 1) * 24);
    for (uint32_t s = 0; s < np; s++) {
        double cnt = sacc[7 * s + 6], nx = sacc[7 * s + 3] / cnt, ny = sacc[7 * s + 4] / cnt, nz = sacc[7 * s + 5] / cnt;
        double il = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
        c->surfels[6 * s] = (float)(sacc[7 * s] / cnt); c->surfels[6 * s + 1] = (float)(sacc[7 * s + 1] / cnt); c->surfels[6 * s + 2] = (float)(sacc[7 * s + 2] / cnt);
        c->surfels[6 * s + 3] = (float)(nx * il); c->surfels[6 * s + 4] = (float)(ny * il); c->surfels[6 * s + 5] = (float)(nz * il);
    }
    c->sums = sacc; free(eacc); free(ekey);
    return c;
}

void prt_o_csr_sizes(const prt_o_csr *c, uint32_t *nnz, uint32_t *n_prim) { *nnz = c->nnz; *n_prim = c->n_prim; }
void prt_o_csr_get(const prt_o_csr *c, uint32_t *range, uint32_t *ids, float *transfer, float *surfels, uint64_t *keys) {
    if (range) memcpy(range, c->range, (size_t)c->n_probes * 8);
    if (ids) memcpy(ids, c->ids, (size_t)c->nnz * 4);
    if (transfer) memcpy(transfer, c->transfer, (size_t)c->nnz * 36);
    if (surfels) memcpy(surfels, c->surfels, (size_t)c->n_prim * 24);
    if (keys) memcpy(keys, c->keys, (size_t)c->n_prim * 8);
}
void prt_o_csr_get_sums(const prt_o_csr *c, double *sums) { memcpy(sums, c->sums, (size_t)c->n_prim * 56); }
void prt_o_csr_destroy(prt_o_csr *c) {
    if (!c) return;
    free(c->sums); free(c->range); free(c->ids); free(c->transfer); free(c->surfels); free(c->keys); free(c);
}

/* precomp_projectSH.comp:51-139: L_k = sum_i transfer[9i+k] * radiance[ID[i]].rgb; window; R-H pack -> out[n_probes][7][4] */
void prt_o_project_arrays(const uint32_t *range, uint32_t n_probes, const uint32_t *ids, const float *transfer, const float *radiance_rgba, float *out) {
    const float PI = 3.14159265359f;
    const float w1 = 3.f / PI * sinf(PI / 3), w2 = 3.f / 2 / PI * sinf(2 * PI / 3);                    /* :104-113 */
    for (uint32_t p = 0; p < n_probes; p++) {
        float L[27] = { 0 };
        for (uint32_t i = range[2 * p]; i < range[2 * p + 1]; i++) {
            const float *rad = radiance_rgba + 4 * (size_t)ids[i];
            for (int k = 0; k < 9; k++) for (int ch = 0; ch < 3; ch++) L[3 * k + ch] += transfer[9 * (size_t)i + k] * rad[ch];
        }
        for (int k = 1; k < 4; k++) for (int ch = 0; ch < 3; ch++) L[3 * k + ch] *= w1;
        for (int k = 4; k < 9; k++) for (int ch = 0; ch < 3; ch++) L[3 * k + ch] *= w2;
        prt_o_sh_pack_rh(L, out + 28 * (size_t)p);
    }
}
void prt_o_probe_project(const prt_o_csr *c, const float *radiance_rgba, float *out) {
    prt_o_project_arrays(c->range, c->n_probes, c->ids, c->transfer, radiance_rgba, out);
}

/* ---- calculate_weight (src/raytracing/light_probe.cpp:156-367): voxel -> 8-probe trilinear weights masked by visibility ----
 * pass 1 (:207-229): 100 Fibonacci closest-hit rays per voxel; inside score = #(dot(dir, Ng) > 0.01) / #hits (0/0 = NaN, as in
 *                    the reference; Ng is the UNNORMALISED geometric normal); score[n_voxels] = 999 for out-of-grid neighbours.
 * pass 2 (:261-361): trilinear weights of the 8 surrounding probes (corner order of the diagram at :269-294); voxels with
 *                    score > 0.2 are moved to the 3x3x3 neighbour with the smallest score (:320-332); each probe is tested with
 *                    a segment any-hit ray (org = voxel, dir = probe - voxel, tfar = 1; :250-251); weights are masked and
 *                    renormalised, or all zero when nothing is visible (:341-357).                                          */
static void grid_pos(const int res[3], const float size[3], const int id[3], float out[3]) {
    for (int a = 0; a < 3; a++) {
        float ds = 2.f / (float)res[a] * size[a];
        out[a] = -size[a] + ds * (0.5f + (float)id[a]);
    }
}
void prt_o_volume_weights(const prt_o_scene *sc, const int probe_res[3], const int volume_res[3], const float scene_size[3],
                          float *w0123, float *w4567, float *score_out) {
    const int rx = volume_res[0], ry = volume_res[1], rz = volume_res[2];
    const size_t nvox = (size_t)rx * ry * rz;
    float *score = (float *)malloc(sizeof(float) * (nvox + 1));
    float dirs[300];
    prt_o_fibonacci_dirs(100, dirs);
    for (int z = 0; z < rz; z++) for (int y = 0; y < ry; y++) for (int x = 0; x < rx; x++) {
        int id[3] = { x, y, z }; float pos[3];
        grid_pos(volume_res, scene_size, id, pos);
        int hits = 0; float inside = 0.f;
        for (int r = 0; r < 100; r++) {
            float t, ng[3]; uint32_t prim;
            if (prt_o_closest_hit(sc, pos, dirs + 3 * r, 0.f, INFINITY, 1, &t, &prim, ng)) {
                hits++;
                if (v3_dot(v3_make(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]), v3_make(ng[0], ng[1], ng[2])) > 0.01f) inside += 1.f;
            }
        }
        score[((size_t)z * ry + y) * rx + x] = inside / (float)hits;
    }
    score[nvox] = 999.f;
    if (score_out) memcpy(score_out, score, sizeof(float) * nvox);
    for (int z = 0; z < rz; z++) for (int y = 0; y < ry; y++) for (int x = 0; x < rx; x++) {
        const size_t index = ((size_t)z * ry + y) * rx + x;
        int vid[3] = { x, y, z }, anchor[3]; float vpos[3], apos[3], fr[3];
        for (int a = 0; a < 3; a++) {
            float tc = ((float)vid[a] + 0.5f) / (float)volume_res[a];
            tc = tc * (float)probe_res[a] - 0.5f;
            anchor[a] = (int)floorf(tc);
        }
        grid_pos(volume_res, scene_size, vid, vpos);
        grid_pos(probe_res, scene_size, anchor, apos);
        for (int a = 0; a < 3; a++) fr[a] = (vpos[a] - apos[a]) * (float)probe_res[a] / (2.f * scene_size[a]);
        static const int off[8][3] = { {0,0,1}, {1,0,1}, {1,0,0}, {0,0,0}, {0,1,0}, {0,1,1}, {1,1,1}, {1,1,0} };
        float w[8] = { (1 - fr[0]) * (1 - fr[1]) * fr[2], fr[0] * (1 - fr[1]) * fr[2], fr[0] * (1 - fr[1]) * (1 - fr[2]), (1 - fr[0]) * (1 - fr[1]) * (1 - fr[2]),
                       (1 - fr[0]) * fr[1] * (1 - fr[2]), (1 - fr[0]) * fr[1] * fr[2], fr[0] * fr[1] * fr[2], fr[0] * fr[1] * (1 - fr[2]) };
        float vp[3] = { vpos[0], vpos[1], vpos[2] };
        if (score[index] > 0.2f) {
            float min_score = score[index];
            for (int nx = -1; nx <= 1; nx++) for (int ny = -1; ny <= 1; ny++) for (int nz = -1; nz <= 1; nz++) {
                int q[3] = { x + nx, y + ny, z + nz };
                size_t ni = (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= rx || q[1] >= ry || q[2] >= rz) ? nvox : ((size_t)q[2] * ry + q[1]) * rx + q[0];
                if (score[ni] < min_score) { grid_pos(volume_res, scene_size, q, vp); min_score = score[ni]; }
            }
        }
        float vis[8], valid = 0.f;
        for (int i = 0; i < 8; i++) {
            int pid[3] = { anchor[0] + off[i][0], anchor[1] + off[i][1], anchor[2] + off[i][2] };
            vis[i] = 0.f;
            if (pid[0] >= 0 && pid[1] >= 0 && pid[2] >= 0 && pid[0] < probe_res[0] && pid[1] < probe_res[1] && pid[2] < probe_res[2]) {
                float pp[3], d[3];
                grid_pos(probe_res, scene_size, pid, pp);
                for (int a = 0; a < 3; a++) d[a] = pp[a] - vp[a];
                vis[i] = prt_o_any_hit(sc, vp, d, 0.f, 1.f, 1) ? 0.f : 1.f;
            }
            valid += vis[i];
        }
        if (valid > 0.f) {
            float sum = 0.f;
            for (int i = 0; i < 8; i++) { w[i] *= vis[i]; }
            for (int i = 0; i < 8; i++) sum += w[i];
            for (int i = 0; i < 8; i++) w[i] /= sum;
        } else for (int i = 0; i < 8; i++) w[i] = 0.f;
        memcpy(w0123 + 4 * index, w, 16); memcpy(w4567 + 4 * index, w + 4, 16);
    }
    free(score);
}
