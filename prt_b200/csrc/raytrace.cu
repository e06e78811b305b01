// raytrace.cu -- sm_100a kernel + C ABI for the reference's progressive preview tracer:
//
//   raytrace      src/raytracing/raytracing.cpp:280-317   camera rays, running mean, gamma, RGBA8 pack
//   renderAO      src/raytracing/raytracing.cpp:177-222   ambient-occlusion path (white environment, albedo per bounce)
//   renderNormal  src/raytracing/raytracing.cpp:162-175   0.5 * (normalize(Ng) + 1)
//
// One thread per pixel (a 8x4 pixel tile per warp so neighbouring rays share BVH nodes), closest-hit traversal of the 8-wide BVH,
// the bounce step of the bake's interreflection path.  The accumulation buffer (sum rgb, count) stays in HBM between frames; bounce
// randoms are Philox stream 2 keyed (pixel, frame, bounce) instead of the reference's thread-local mt19937, so frames are reproducible.
#include "../../include/prt_b200.h"
#include "abi_internal.h"
#include "traverse.cuh"

#include <cuda_runtime.h>
#include <math.h>
#include <string>

using namespace prt;

struct prt_film {
    prt_ctx *ctx = nullptr;
    int w = 0, h = 0;
    float4 *accum = nullptr;
    uchar4 *pixels = nullptr;
    uint32_t frame = 0;
};

namespace {

struct RayArgs {
    const Node8 *nodes; const Tri48 *tris;
    f3 P, bl, Up, Right;
    int w, h, depth, gamma, mode;
    float albedo[3];
    uint32_t seed, frame;
    float4 *accum; uchar4 *pixels;
};

__global__ void __launch_bounds__(128) raytrace_kernel(const RayArgs A) {
    // 8x4 tiles: lane -> (lane & 7, lane >> 3)
    const int tiles_x = (A.w + 7) >> 3;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int i = (warp % tiles_x) * 8 + (lane & 7), j = (warp / tiles_x) * 4 + (lane >> 3);
    if (i >= A.w || j >= A.h) return;
    const uint32_t pixel = (uint32_t)(j * A.w + i);
    const float fy = PRT_DIV((float)j, (float)A.h), fx = PRT_DIV((float)i, (float)A.w);
    f3 dir = mk3(PRT_FMA(fx, A.Right.x, PRT_FMA(fy, A.Up.x, A.bl.x)), PRT_FMA(fx, A.Right.y, PRT_FMA(fy, A.Up.y, A.bl.y)),
                 PRT_FMA(fx, A.Right.z, PRT_FMA(fy, A.Up.z, A.bl.z)));                      // raytracing.cpp:299
    f3 pos = A.P;
    float L0 = 0.f, L1 = 0.f, L2 = 0.f;
    Trav tr;
    tr.reset_counters();
    if (A.mode == 1) {                                                                      // renderNormal
        tr.init(pos, dir, 0.0f, INFINITY);
        tr.start_root();
        tr.run<false>(A.nodes, A.tris, 0, false);
        if (tr.best_prim != 0xFFFFFFFFu) {
            const f3 n = normalize3(tr.hit_ng(A.tris));
            L0 = PRT_MUL(0.5f, PRT_ADD(n.x, 1.0f)); L1 = PRT_MUL(0.5f, PRT_ADD(n.y, 1.0f)); L2 = PRT_MUL(0.5f, PRT_ADD(n.z, 1.0f));
        }
    } else {                                                                                // renderAO
        const float eps = 1e-5f;
        float Lw0 = 1.f, Lw1 = 1.f, Lw2 = 1.f, tnear = 0.0f;
        for (int seg = 0; seg < A.depth; seg++) {
            if (fmaxf(Lw0, fmaxf(Lw1, Lw2)) < 0.01f) break;                                 // :196
            tr.init(pos, dir, tnear, INFINITY);
            tr.start_root();
            tr.run<false>(A.nodes, A.tris, 0, false);
            if (tr.best_prim == 0xFFFFFFFFu) { L0 = Lw0; L1 = Lw1; L2 = Lw2; break; }       // :204-207
            const f3 n = normalize3(tr.hit_ng(A.tris));                                     // :209-210
            pos = madd3(pos, tr.best_t, dir);                                               // :211
            float u, v;
            rand2(A.seed, pixel, A.frame, (uint32_t)seg, 2u, u, v);                         // :212
            const f3 l = cosine_local(u, v);
            const float pdf = PRT_DIV(l.z, kPiF);
            const Frame fb = make_frame(n);
            dir = to_world(fb, l);
            if (pdf <= 1e-4f) break;                                                        // :214
            Lw0 = PRT_MUL(Lw0, A.albedo[0]); Lw1 = PRT_MUL(Lw1, A.albedo[1]); Lw2 = PRT_MUL(Lw2, A.albedo[2]);   // :216
            const float sg = dot3(dir, n) < 0.0f ? -1.0f : 1.0f;                            // :218
            pos = madd3(pos, PRT_MUL(sg, eps), dir);                                        // :219
            tnear = eps;                                                                    // :220
        }
    }
    float4 a = A.accum[pixel];
    a.x = PRT_ADD(a.x, L0); a.y = PRT_ADD(a.y, L1); a.z = PRT_ADD(a.z, L2); a.w = PRT_ADD(a.w, 1.0f);        // :303-304
    A.accum[pixel] = a;
    const float wgt = PRT_DIV(1.0f, a.w);                                                   // :305
    float c[3] = {PRT_MUL(a.x, wgt), PRT_MUL(a.y, wgt), PRT_MUL(a.z, wgt)};
    unsigned char q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float x = fminf(fmaxf(c[k], 0.0f), 1.0f);                                           // :306
        if (A.gamma) x = powf(x, (float)(1 / 2.2));                                         // :307-308
        q[k] = (unsigned char)PRT_MUL(255.0f, x);                                           // :309-311
    }
    A.pixels[pixel] = make_uchar4(q[0], q[1], q[2], 255);
}

}  // namespace

#define RT_TRY(expr)                                                                                                    \
    do {                                                                                                                \
        cudaError_t e_ = (expr);                                                                                        \
        if (e_ != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

void prt_film_destroy(prt_film *f) {
    if (!f) return;
    cudaSetDevice(prt_ctx_device(f->ctx));
    cudaFree(f->accum); cudaFree(f->pixels);
    delete f;
}

int prt_film_create(prt_ctx *ctx, int32_t width, int32_t height, prt_film **out) {
    if (!ctx || !out || width <= 0 || height <= 0 || (long long)width * height > (1ll << 28)) return prt_set_error(PRT_ERR_INVALID, "prt_film_create: bad argument");
    *out = nullptr;
    RT_TRY(cudaSetDevice(prt_ctx_device(ctx)));
    prt_film *f = new prt_film();
    f->ctx = ctx; f->w = width; f->h = height;
    const size_t n = (size_t)width * height;
    cudaError_t e = cudaMalloc(&f->accum, 16 * n);
    if (e == cudaSuccess) e = cudaMalloc(&f->pixels, 4 * n);
    if (e == cudaSuccess) e = cudaMemset(f->accum, 0, 16 * n);
    if (e == cudaSuccess) e = cudaMemset(f->pixels, 0, 4 * n);
    if (e != cudaSuccess) { prt_film_destroy(f); return prt_set_error(PRT_ERR_CUDA, std::string("prt_film_create: ") + cudaGetErrorString(e)); }
    *out = f;
    return PRT_OK;
}

int prt_film_reset(prt_film *f) {                                                          // camera.dirty (raytracing.cpp:282-285)
    if (!f) return prt_set_error(PRT_ERR_INVALID, "prt_film_reset: null argument");
    RT_TRY(cudaSetDevice(prt_ctx_device(f->ctx)));
    RT_TRY(cudaMemsetAsync(f->accum, 0, 16 * (size_t)f->w * f->h, prt_ctx_stream(f->ctx)));
    f->frame = 0;
    return PRT_OK;
}

int prt_raytrace(prt_scene *scene, prt_film *f, const prt_camera *cam, int32_t max_path_length, const float albedo[3], int32_t gamma,
                 int32_t mode, uint32_t seed, int32_t n_frames) {
    if (!scene || !f || !cam || !albedo || max_path_length < 0 || n_frames < 0 || (mode != PRT_RAYTRACE_AO && mode != PRT_RAYTRACE_NORMAL))
        return prt_set_error(PRT_ERR_INVALID, "prt_raytrace: bad argument");
    const prt_scene_view sv = prt_scene_get_view(scene);
    if (sv.ctx != f->ctx) return prt_set_error(PRT_ERR_INVALID, "prt_raytrace: scene and film belong to different contexts");
    RT_TRY(cudaSetDevice(prt_ctx_device(sv.ctx)));
    cudaStream_t st = prt_ctx_stream(sv.ctx);
    RayArgs A{};
    A.nodes = sv.nodes; A.tris = sv.tris;
    const double PI = 3.14159265358979323846;
    const float height = (float)(2.0 * tan((double)cam->zoom_deg * PI / 180.0 / 2.0));      // :293
    const float width = height * (float)f->w / (float)f->h;                                  // :294
    A.Up = mk3(height * cam->up[0], height * cam->up[1], height * cam->up[2]);
    A.Right = mk3(width * cam->right[0], width * cam->right[1], width * cam->right[2]);
    A.bl = mk3((cam->front[0] - 0.5f * A.Up.x) - 0.5f * A.Right.x, (cam->front[1] - 0.5f * A.Up.y) - 0.5f * A.Right.y,
               (cam->front[2] - 0.5f * A.Up.z) - 0.5f * A.Right.z);                          // :297
    A.P = mk3(cam->position[0], cam->position[1], cam->position[2]);
    A.w = f->w; A.h = f->h; A.depth = max_path_length; A.gamma = gamma; A.mode = mode;
    for (int k = 0; k < 3; k++) A.albedo[k] = albedo[k];
    A.seed = seed; A.accum = f->accum; A.pixels = f->pixels;
    const long long warps = (long long)((f->w + 7) / 8) * ((f->h + 3) / 4);
    for (int32_t it = 0; it < n_frames; it++) {
        A.frame = f->frame++;
        raytrace_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(A);
    }
    RT_TRY(cudaGetLastError());
    RT_TRY(cudaStreamSynchronize(st));
    return PRT_OK;
}

int prt_film_download(const prt_film *f, float *accum, uint8_t *pixels) {
    if (!f) return prt_set_error(PRT_ERR_INVALID, "prt_film_download: null argument");
    RT_TRY(cudaSetDevice(prt_ctx_device(f->ctx)));
    RT_TRY(cudaStreamSynchronize(prt_ctx_stream(f->ctx)));
    const size_t n = (size_t)f->w * f->h;
    if (accum) RT_TRY(cudaMemcpy(accum, f->accum, 16 * n, cudaMemcpyDeviceToHost));
    if (pixels) RT_TRY(cudaMemcpy(pixels, f->pixels, 4 * n, cudaMemcpyDeviceToHost));
    return PRT_OK;
}

}  // extern "C"
