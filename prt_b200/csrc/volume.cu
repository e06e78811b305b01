// volume.cu -- sm_100a kernels + C ABI for calculate_weight (reference src/raytracing/light_probe.cpp:156-367): the
// visibility-masked trilinear weights that blend the 8 surrounding probes into every voxel of the SH volume
// (consumed by SH_volume::set_visibility, src/sh/volume.cpp:318-333 -> volume_weight0123 / volume_weight4567).
//
//   pass 1 (light_probe.cpp:207-229)  one warp per voxel: 100 Fibonacci closest-hit rays, inside score = #(dot(dir,Ng) > 0.01) / #hits
//   pass 2 (light_probe.cpp:261-361)  one thread per voxel: trilinear weights, relocation to the least-inside 3x3x3 neighbour,
//                                     8 segment any-hit rays (tfar = 1), masking + renormalisation
#include "../../include/prt_b200.h"
#include "abi_internal.h"
#include "traverse.cuh"

#include <cuda_runtime.h>
#include <math.h>

using namespace prt;

namespace {

struct Grid { int rx, ry, rz; float sx, sy, sz; };
// cal_probe_pos / cal_volume_pos (light_probe.cpp:176-185): -size + (2/res*size) * (0.5 + id), explicitly rounded ops
__device__ __forceinline__ f3 grid_pos(const Grid &g, int x, int y, int z) {
    const float dx = PRT_MUL(PRT_DIV(2.f, (float)g.rx), g.sx), dy = PRT_MUL(PRT_DIV(2.f, (float)g.ry), g.sy), dz = PRT_MUL(PRT_DIV(2.f, (float)g.rz), g.sz);
    return mk3(PRT_ADD(-g.sx, PRT_MUL(dx, PRT_ADD(0.5f, (float)x))), PRT_ADD(-g.sy, PRT_MUL(dy, PRT_ADD(0.5f, (float)y))),
               PRT_ADD(-g.sz, PRT_MUL(dz, PRT_ADD(0.5f, (float)z))));
}

__global__ void __launch_bounds__(256) volume_score_kernel(const Node8 *nodes, const Tri48 *tris, Grid vol, const float *dirs, int n_dirs, float *score) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t nvox = (size_t)vol.rx * vol.ry * vol.rz;
    if (warp >= nvox) return;
    const int x = (int)(warp % vol.rx), y = (int)((warp / vol.rx) % vol.ry), z = (int)(warp / ((size_t)vol.rx * vol.ry));
    const f3 pos = grid_pos(vol, x, y, z);
    int hits = 0, inside = 0;
    for (int r = lane; r < n_dirs; r += 32) {
        const f3 d = mk3(dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]);
        Trav tr;
        tr.reset_counters();
        tr.init(pos, d, 0.f, INFINITY);
        tr.start_root();
        tr.run<false>(nodes, tris, 0, false);
        if (tr.best_prim != 0xFFFFFFFFu) {
            hits++;
            if (dot3(d, tr.hit_ng(tris)) > 0.01f) inside++;        // light_probe.cpp:221, unnormalised Ng
        }
    }
    hits = __reduce_add_sync(0xFFFFFFFFu, hits);
    inside = __reduce_add_sync(0xFFFFFFFFu, inside);
    if (lane == 0) score[warp] = PRT_DIV((float)inside, (float)hits);   // 0/0 = NaN like the reference
}

__global__ void __launch_bounds__(128) volume_weight_kernel(const Node8 *nodes, const Tri48 *tris, Grid vol, Grid prb, const float *score, float4 *w0123, float4 *w4567) {
    const size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nvox = (size_t)vol.rx * vol.ry * vol.rz;
    if (index >= nvox) return;
    const int x = (int)(index % vol.rx), y = (int)((index / vol.rx) % vol.ry), z = (int)(index / ((size_t)vol.rx * vol.ry));
    const int vid[3] = {x, y, z}, vres[3] = {vol.rx, vol.ry, vol.rz}, pres[3] = {prb.rx, prb.ry, prb.rz};
    const float size[3] = {vol.sx, vol.sy, vol.sz};
    int anchor[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float tc = PRT_DIV(PRT_ADD((float)vid[a], 0.5f), (float)vres[a]);
        tc = PRT_SUB(PRT_MUL(tc, (float)pres[a]), 0.5f);
        anchor[a] = (int)floorf(tc);
    }
    const f3 vpos = grid_pos(vol, x, y, z), apos = grid_pos(prb, anchor[0], anchor[1], anchor[2]);
    const float vp0[3] = {vpos.x, vpos.y, vpos.z}, ap[3] = {apos.x, apos.y, apos.z};
    float fr[3];
#pragma unroll
    for (int a = 0; a < 3; a++) fr[a] = PRT_DIV(PRT_MUL(PRT_SUB(vp0[a], ap[a]), (float)pres[a]), PRT_MUL(2.f, size[a]));
    float w[8] = {(1 - fr[0]) * (1 - fr[1]) * fr[2], fr[0] * (1 - fr[1]) * fr[2], fr[0] * (1 - fr[1]) * (1 - fr[2]), (1 - fr[0]) * (1 - fr[1]) * (1 - fr[2]),
                  (1 - fr[0]) * fr[1] * (1 - fr[2]), (1 - fr[0]) * fr[1] * fr[2], fr[0] * fr[1] * fr[2], fr[0] * fr[1] * (1 - fr[2])};
    f3 vp = vpos;
    const float s0 = score[index];
    if (s0 > 0.2f) {                                                                    // light_probe.cpp:320-332
        float min_score = s0;
        for (int nx = -1; nx <= 1; nx++) for (int ny = -1; ny <= 1; ny++) for (int nz = -1; nz <= 1; nz++) {
            const int qx = x + nx, qy = y + ny, qz = z + nz;
            const bool oob = qx < 0 || qy < 0 || qz < 0 || qx >= vol.rx || qy >= vol.ry || qz >= vol.rz;
            const float ns = oob ? 999.f : score[((size_t)qz * vol.ry + qy) * vol.rx + qx];
            if (ns < min_score) { vp = grid_pos(vol, qx, qy, qz); min_score = ns; }
        }
    }
    const int off[8][3] = {{0, 0, 1}, {1, 0, 1}, {1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}};   // diagram :269-294
    float vis[8], valid = 0.f;
    for (int i = 0; i < 8; i++) {
        const int px = anchor[0] + off[i][0], py = anchor[1] + off[i][1], pz = anchor[2] + off[i][2];
        vis[i] = 0.f;
        if (px >= 0 && py >= 0 && pz >= 0 && px < prb.rx && py < prb.ry && pz < prb.rz) {
            const f3 pp = grid_pos(prb, px, py, pz);
            Trav tr;
            tr.reset_counters();
            tr.init(vp, sub3(pp, vp), 0.f, 1.f);                                       // segment test, light_probe.cpp:250
            tr.start_root();
            vis[i] = tr.run<true>(nodes, tris, 0, false) == TRAV_HIT ? 0.f : 1.f;
        }
        valid += vis[i];
    }
    if (valid > 0.f) {
        float sum = 0.f;
        for (int i = 0; i < 8; i++) w[i] *= vis[i];
        for (int i = 0; i < 8; i++) sum += w[i];
        for (int i = 0; i < 8; i++) w[i] = w[i] / sum;
    } else {
        for (int i = 0; i < 8; i++) w[i] = 0.f;
    }
    w0123[index] = make_float4(w[0], w[1], w[2], w[3]);
    w4567[index] = make_float4(w[4], w[5], w[6], w[7]);
}

}  // namespace

extern "C" int prt_volume_weights(prt_scene *scene, const int32_t probe_res[3], const int32_t volume_res[3], const float scene_size[3],
                                  float *w0123, float *w4567, float *inside_score) {
    if (!scene || !probe_res || !volume_res || !scene_size || (!w0123) != (!w4567)) return prt_set_error(PRT_ERR_INVALID, "prt_volume_weights: null argument");
    for (int a = 0; a < 3; a++)
        if (probe_res[a] < 1 || volume_res[a] < 1 || !(scene_size[a] > 0.f)) return prt_set_error(PRT_ERR_INVALID, "prt_volume_weights: bad grid");
    const prt_scene_view sv = prt_scene_get_view(scene);
    cudaError_t e = cudaSetDevice(prt_ctx_device(sv.ctx));
    if (e != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, cudaGetErrorString(e));
    cudaStream_t st = prt_ctx_stream(sv.ctx);
    const Grid vol{volume_res[0], volume_res[1], volume_res[2], scene_size[0], scene_size[1], scene_size[2]};
    const Grid prb{probe_res[0], probe_res[1], probe_res[2], scene_size[0], scene_size[1], scene_size[2]};
    const size_t nvox = (size_t)vol.rx * vol.ry * vol.rz;
    float dirs[300];
    prt_fibonacci_dirs(100, dirs);                                                     // static out_dirs = get_dirs(), light_probe.cpp:187
    // scratch owned by the context (no cudaMalloc / cudaFree per call); w0123 == w4567 == NULL keeps the weights on the device
    float *d_dirs = (float *)prt_ctx_scratch(sv.ctx, 3, sizeof dirs), *d_score = (float *)prt_ctx_scratch(sv.ctx, 4, 4 * nvox);
    float4 *d_w0 = (float4 *)prt_ctx_scratch(sv.ctx, 5, 16 * nvox), *d_w1 = (float4 *)prt_ctx_scratch(sv.ctx, 6, 16 * nvox);
    if (!d_dirs || !d_score || !d_w0 || !d_w1) return prt_set_error(PRT_ERR_NOMEM, "prt_volume_weights: out of device memory");
    cudaMemcpyAsync(d_dirs, dirs, sizeof dirs, cudaMemcpyHostToDevice, st);
    prt_ctx_timer_begin(sv.ctx, st);
    volume_score_kernel<<<(unsigned)((nvox * 32 + 255) / 256), 256, 0, st>>>(sv.nodes, sv.tris, vol, d_dirs, 100, d_score);
    volume_weight_kernel<<<(unsigned)((nvox + 127) / 128), 128, 0, st>>>(sv.nodes, sv.tris, vol, prb, d_score, d_w0, d_w1);
    prt_ctx_timer_end(sv.ctx, st);
    if (w0123) {
        cudaMemcpyAsync(w0123, d_w0, 16 * nvox, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(w4567, d_w1, 16 * nvox, cudaMemcpyDeviceToHost, st);
        if (inside_score) cudaMemcpyAsync(inside_score, d_score, 4 * nvox, cudaMemcpyDeviceToHost, st);
    }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return prt_set_error(PRT_ERR_CUDA, std::string("prt_volume_weights: ") + cudaGetErrorString(e));
    return PRT_OK;
}
