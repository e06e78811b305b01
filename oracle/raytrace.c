/*
 * oracle/raytrace.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product library).
 *
 * CPU restatement of the reference's progressive preview tracer:
 *   raytrace      src/raytracing/raytracing.cpp:280-317   camera rays, running mean, gamma, RGBA8 pack
 *   renderAO      src/raytracing/raytracing.cpp:177-222   ambient-occlusion path (white environment, albedo per bounce)
 *   renderNormal  src/raytracing/raytracing.cpp:162-175   0.5 * (normalize(Ng) + 1)
 *
 * parity unpinned: the reference draws its bounce randoms from a thread_local mt19937 (raytracing.cpp:14-18), so its images are
 * not reproducible and it ships none.  Pinned here: Philox stream 2 keyed (pixel, frame, bounce); PRT-ARITH v1 float ops.
 */
#include "arith.h"
#include "philox.h"
#include "prt_oracle.h"

static void camera_basis(const prt_o_camera *c, int w, int h, v3 *bl, v3 *Up, v3 *Right) {
    const double PI = 3.14159265358979323846;
    const float height = (float)(2.0 * tan((double)c->zoom_deg * PI / 180.0 / 2.0));    /* :293 */
    const float width = height * (float)w / (float)h;                                   /* :294 */
    *Up = v3_make(height * c->up[0], height * c->up[1], height * c->up[2]);
    *Right = v3_make(width * c->right[0], width * c->right[1], width * c->right[2]);
    *bl = v3_make((c->front[0] - 0.5f * Up->x) - 0.5f * Right->x, (c->front[1] - 0.5f * Up->y) - 0.5f * Right->y,
                  (c->front[2] - 0.5f * Up->z) - 0.5f * Right->z);                      /* :297 */
}

/* renderAO: returns L.x,y,z */
static void render_ao(const prt_o_scene *sc, v3 pos, v3 dir, int depth, const float albedo[3], uint32_t seed, uint32_t pixel, uint32_t frame,
                      float L[3]) {
    const float eps = 1e-5f;
    float Lw[3] = { 1.f, 1.f, 1.f }, tnear = 0.0f;
    L[0] = L[1] = L[2] = 0.0f;
    for (int i = 0; i < depth; i++) {
        if (fmaxf(Lw[0], fmaxf(Lw[1], Lw[2])) < 0.01f) break;                            /* :196 */
        float o[3] = { pos.x, pos.y, pos.z }, d[3] = { dir.x, dir.y, dir.z }, t, ng[3];
        uint32_t prim;
        if (!prt_o_closest_hit(sc, o, d, tnear, INFINITY, 1, &t, &prim, ng)) { L[0] = Lw[0]; L[1] = Lw[1]; L[2] = Lw[2]; break; }   /* :204-207 */
        v3 n = v3_normalize(v3_make(ng[0], ng[1], ng[2]));                               /* :209-210 */
        pos = v3_madd(pos, t, dir);                                                      /* :211 */
        float u, v;
        prt_rand2(seed, pixel, frame, (uint32_t)i, 2u, &u, &v);                          /* :212 */
        v3 l = prt_cosine_local(u, v);
        float pdf = l.z / PRT_PI_F;
        frame3 f = prt_frame(n);
        dir = prt_to_world(&f, l);
        if (pdf <= 1e-4f) break;                                                         /* :214 */
        Lw[0] *= albedo[0]; Lw[1] *= albedo[1]; Lw[2] *= albedo[2];                      /* :216 */
        float sign = v3_dot(dir, n) < 0.0f ? -1.0f : 1.0f;                               /* :218 */
        pos = v3_madd(pos, sign * eps, dir);                                             /* :219 */
        tnear = eps;                                                                     /* :220 */
    }
}

void prt_o_raytrace(const prt_o_scene *sc, const prt_o_camera *cam, int w, int h, int max_path_length, const float albedo[3], int gamma,
                    int mode, uint32_t seed, uint32_t frame, float *accum, uint8_t *pixels) {
    v3 bl, Up, Right;
    camera_basis(cam, w, h, &bl, &Up, &Right);
    const v3 P = v3_make(cam->position[0], cam->position[1], cam->position[2]);
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const float fy = (float)j / (float)h, fx = (float)i / (float)w;
            const v3 dir = v3_make(fmaf(fx, Right.x, fmaf(fy, Up.x, bl.x)), fmaf(fx, Right.y, fmaf(fy, Up.y, bl.y)), fmaf(fx, Right.z, fmaf(fy, Up.z, bl.z)));   /* :299 */
            float L[3] = { 0, 0, 0 };
            if (mode == 1) {                                                             /* renderNormal */
                float o[3] = { P.x, P.y, P.z }, d[3] = { dir.x, dir.y, dir.z }, t, ng[3]; uint32_t prim;
                if (prt_o_closest_hit(sc, o, d, 0.0f, INFINITY, 1, &t, &prim, ng)) {
                    v3 n = v3_normalize(v3_make(ng[0], ng[1], ng[2]));
                    L[0] = 0.5f * (n.x + 1.0f); L[1] = 0.5f * (n.y + 1.0f); L[2] = 0.5f * (n.z + 1.0f);
                }
            } else render_ao(sc, P, dir, max_path_length, albedo, seed, (uint32_t)(j * w + i), frame, L);
            float *a = accum + 4 * ((size_t)j * w + i);
            a[0] += L[0]; a[1] += L[1]; a[2] += L[2]; a[3] += 1.0f;                      /* :303-304 */
            const float wgt = 1.0f / a[3];                                               /* :305 */
            uint8_t *px = pixels + 4 * ((size_t)j * w + i);
            for (int k = 0; k < 3; k++) {
                float c = fminf(fmaxf(a[k] * wgt, 0.0f), 1.0f);                          /* :306 */
                if (gamma) c = powf(c, (float)(1 / 2.2));                                /* :307-308 */
                px[k] = (uint8_t)(255.0f * c);                                           /* :309-311 */
            }
            px[3] = 255;
        }
}
