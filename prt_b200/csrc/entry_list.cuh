// entry_list.cuh -- per-origin entry lists (device only), shared by the bake kernels.
#pragma once
#include "traverse.cuh"
#include <cuda_runtime.h>

namespace prt {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- per-origin entry list ---------------------------------------------------------------------------------------
// All S rays of a vertex share one origin, so every one of them would repeat the same descent through the chain of
// nodes whose boxes contain that origin.  Once per vertex the warp flattens that chain: it expands (breadth of 8
// lanes = 8 children) every node containing the origin and records each remaining child -- subtree root or leaf --
// that is not wholly below the tangent plane as a *candidate* box in shared memory (centre/half-extent relative to
// the origin).  Each ray then tests the candidate list in lockstep (no divergence, no node decode) and traverses only
// the subtrees whose candidate box it hits.  Decisions are still made by the pinned triangle test, so results are
// unchanged; the list only removes work.
constexpr int kMaxCand = 96;
struct EntryList {
    float4 ca[kMaxCand];       // centre - origin (xyz), half extent x
    float4 cb[kMaxCand];       // half extent y, z, group x, group y (see Trav::start_group)
    uint32_t queue[kMaxCand];
};

__device__ __forceinline__ int build_entry_list(const Node8 *nodes, const f3 O, const f3 N, EntryList &W, const int lane) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int n = 0, qn = 1;
    if (lane == 0) W.queue[0] = 0u;
    __syncwarp();
    while (qn > 0) {
        const uint32_t x = W.queue[qn - 1];
        qn--;
        __syncwarp();
        if (n + qn + 8 > kMaxCand) {
            // out of room: keep the node itself as an always-hit candidate (graceful fallback to plain traversal)
            if (lane == 0) {
                W.ca[n] = make_float4(0.f, 0.f, 0.f, INFINITY);
                W.cb[n] = make_float4(INFINITY, INFINITY, __uint_as_float(x), __uint_as_float(0x80000000u));
            }
            n++;
            continue;
        }
        const char *np = reinterpret_cast<const char *>(nodes + x);
        const u4 n0 = ld16(np), n1 = ld16(np + 16), n2 = ld16(np + 32), n3 = ld16(np + 48), n4 = ld16(np + 64);
        bool keep = false, expand = false;
        float cx = 0.f, cy = 0.f, cz = 0.f, ex = 0.f, ey = 0.f, ez = 0.f;
        uint32_t gx = 0u, gy = 0u;
        if (lane < 8) {
            const int h = lane >> 2, sh = 8 * (lane & 3);
            const uint32_t meta = ((h ? n1.w : n1.z) >> sh) & 0xFFu;
            if (meta) {
                const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
                            sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
                const float lox = __uint_as_float(n0.x) + (float)(((h ? n2.y : n2.x) >> sh) & 0xFFu) * sx;
                const float loy = __uint_as_float(n0.y) + (float)(((h ? n2.w : n2.z) >> sh) & 0xFFu) * sy;
                const float loz = __uint_as_float(n0.z) + (float)(((h ? n3.y : n3.x) >> sh) & 0xFFu) * sz;
                const float hix = __uint_as_float(n0.x) + (float)(((h ? n3.w : n3.z) >> sh) & 0xFFu) * sx;
                const float hiy = __uint_as_float(n0.y) + (float)(((h ? n4.y : n4.x) >> sh) & 0xFFu) * sy;
                const float hiz = __uint_as_float(n0.z) + (float)(((h ? n4.w : n4.z) >> sh) & 0xFFu) * sz;
                cx = 0.5f * (lox + hix) - O.x; cy = 0.5f * (loy + hiy) - O.y; cz = 0.5f * (loz + hiz) - O.z;
                ex = 0.5f * (hix - lox); ey = 0.5f * (hiy - loy); ez = 0.5f * (hiz - loz);
                // absorb the rounding of the centre/half-extent form (boxes are already padded by the builder)
                ex += 4e-7f * (fabsf(cx) + ex); ey += 4e-7f * (fabsf(cy) + ey); ez += 4e-7f * (fabsf(cz) + ez);
                const float top = N.x * cx + N.y * cy + N.z * cz + fabsf(N.x) * ex + fabsf(N.y) * ey + fabsf(N.z) * ez;
                const float far = fmaxf(fabsf(cx) + ex, fmaxf(fabsf(cy) + ey, fabsf(cz) + ez));
                keep = !(top < -1e-5f * far);   // something of the box lies above the tangent plane
                const bool inside = fabsf(cx) <= ex && fabsf(cy) <= ey && fabsf(cz) <= ez;
                const bool inner = (n0.w >> (24 + lane)) & 1u;
                if (inner) {
                    gx = n1.x + __popc((n0.w >> 24) & lt_mask);
                    gy = 0x80000000u;
                    expand = keep && inside;
                } else {
                    gx = n1.y + (meta & 31u);
                    gy = meta >> 5;          // unary count: 1, 3, 7 = one bit per triangle
                }
            }
        }
        const unsigned keep_b = __ballot_sync(kFull, keep), exp_b = __ballot_sync(kFull, expand), cand_b = keep_b & ~exp_b;
        if (keep && !expand) {
            const int slot = n + __popc(cand_b & lt_mask);
            W.ca[slot] = make_float4(cx, cy, cz, ex);
            W.cb[slot] = make_float4(ey, ez, __uint_as_float(gx), __uint_as_float(gy));
        }
        if (expand) W.queue[qn + __popc(exp_b & lt_mask)] = gx;
        n += __popc(cand_b);
        qn += __popc(exp_b);
        __syncwarp();
    }
    return n;
}

// Bit k of the result words = candidate k is a leaf (triangle group) rather than a subtree root.
__device__ __forceinline__ void entry_list_leaf_mask(const EntryList &W, const int n, const int lane, uint32_t lm[3]) {
#pragma unroll
    for (int w = 0; w < 3; w++) {
        const int k = 32 * w + lane;
        const bool leaf = k < n && __float_as_uint(W.cb[k].w) <= 0x00FFFFFFu;
        lm[w] = __ballot_sync(kFull, leaf);
    }
}

// approximate reciprocal for box tests only (never used by the pinned triangle test)
__device__ __forceinline__ float rcp_box(float d) {
    if (fabsf(d) < 1e-18f) d = copysignf(1e-18f, d);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}

// exclusive prefix sum over the warp of two 16-bit counters packed in one word; total returned through `total`
__device__ __forceinline__ uint32_t warp_excl_scan_packed(uint32_t v, const int lane, uint32_t &total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(kFull, incl, 31);
    return incl - v;
}

// Tests the lane's ray (origin = list origin, interval [0, inf)) against all candidate boxes; one bit per candidate.
__device__ __forceinline__ void scan_entry_list(const EntryList &W, const int n, const float idx, const float idy, const float idz, uint32_t m[3]) {
    const float aix = fabsf(idx), aiy = fabsf(idy), aiz = fabsf(idz);
#pragma unroll
    for (int w = 0; w < 3; w++) {
        uint32_t bits = 0u, bit = 1u;
        const int k1 = min(n, 32 * (w + 1));
#pragma unroll 4
        for (int k = 32 * w; k < k1; k++) {
            const float4 a = W.ca[k], b = W.cb[k];
            const float tx = a.x * idx, ty = a.y * idy, tz = a.z * idz;
            const float tmin = fmaxf(fmaxf(fmaf(-a.w, aix, tx), fmaf(-b.x, aiy, ty)), fmaf(-b.y, aiz, tz));
            const float tmax = fminf(fminf(fmaf(a.w, aix, tx), fmaf(b.x, aiy, ty)), fmaf(b.y, aiz, tz));
            if (tmin <= tmax && tmax >= 0.0f) bits |= bit;
            bit += bit;
        }
        m[w] = bits;
    }
}

}  // namespace prt
