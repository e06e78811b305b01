// bvh8.h -- flattened 8-wide compressed BVH layout shared by the host builder and the sm_100a kernels.
//
// Replaces the acceleration structure the reference obtains from Intel Embree 3 behind RTScene
// (reference src/raytracing/raytracing.cpp:58-99, light_probe.cpp:44-93).  Layout follows the compressed
// wide BVH of Ylitie, Karras & Laine (HPG 2017): one 80-byte node = five 16-byte vector loads, child
// boxes quantised to 8 bits per plane relative to the node origin, children stored in an octant-sorted
// slot order so traversal needs no per-ray distance sort.  Triangles are 48 bytes = three 16-byte loads.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace prt {

struct alignas(16) Node8 {
    float px, py, pz;       // node origin (min corner)
    uint8_t ex, ey, ez;     // per-axis power-of-two scale, biased exponent byte: scale = as_float(e << 23)
    uint8_t imask;          // bit s set: slot s holds an internal child
    uint32_t child_base;    // index of first internal child node (children contiguous in slot order)
    uint32_t tri_base;      // index of first triangle referenced by this node's leaf slots
    uint8_t meta[8];        // internal: 0x20 | (24+slot); leaf: (unary tri count << 5) | tri offset; empty: 0
    uint8_t qlox[8], qloy[8], qloz[8];
    uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

// v0 + edges for the pinned Moeller-Trumbore test; w of the first vector carries the
// caller's primitive index.
struct alignas(16) Tri48 {
    float v0x, v0y, v0z; uint32_t prim;
    float e1x, e1y, e1z; uint32_t pad1;
    float e2x, e2y, e2z; uint32_t pad2;
};
static_assert(sizeof(Tri48) == 48, "Tri48 must be 48 bytes");

constexpr int kStackEntries = 48;   // traversal stack capacity (node groups); builder enforces max depth

// Oriented slab of one 8-wide node (horizon pass only, entry_list.cuh): every triangle below the node lies between the planes
// m . x = d0 and m . x = d1, m = the (unit) area-weighted mean normal of those triangles.  A surface patch is a thin sheet inside its
// fat axis-aligned box; the slab lets the horizon builder bound the height of the sheet above a tangent plane instead of the box's.
// m = 0: no slab (degenerate mean normal): d0 = -3e38, d1 = 3e38.
// (A ray / slab test in the TRAVERSAL pass -- filtering (ray, node) items before their node is opened -- removes 40 % of the node tests on
// the bench mesh but was 20 % slower on the GPU: the extra steps cost as many instructions as they save and one more dependent fetch
// each; measured and removed, profiles/r2_slab_filter_rejected_ncu_summary.txt.)
struct alignas(16) Slab32 {
    float mx, my, mz, d0;
    float d1, pad0, pad1, pad2;
};
static_assert(sizeof(Slab32) == 32, "Slab32 must be 32 bytes");

// Fourth slab axis of the node test (traversal pass, traverse.cuh): the node's mean normal m (the Slab32's), scaled so that
// s(x) = M . x - D0 runs over [0, 255] across the node's own extent along m, and the extent of every CHILD's triangles along it, quantised
// outward to those 255 steps plus one step of padding either side (covers the float evaluation of s on the device).  A child is culled
// when the ray's interval inside its box does not meet its [qlo, qhi] along m: a surface patch is a thin sheet in its box, and a ray that
// grazes the surface passes through many such boxes beside their sheets.  M = 0: no axis (qlo = 0, qhi = 255: never culls).
struct alignas(16) Dop32 {
    float Mx, My, Mz, D0;
    uint8_t qlo[8], qhi[8];
};
static_assert(sizeof(Dop32) == 32, "Dop32 must be 32 bytes");

struct HostBVH8 {
    Node8 *nodes = nullptr;
    Tri48 *tris = nullptr;
    Slab32 *slabs = nullptr;        // [n_nodes]
    Dop32 *dops = nullptr;          // [n_nodes]
    uint32_t n_nodes = 0, n_tris = 0, max_depth = 0;
    double sah_cost = 0.0, build_seconds = 0.0;
    float pad = 0.f;
};

// Builds on the host with all hardware threads. Returns 0 on success, <0 on error (message in err).
int build_bvh8(const float *pos, size_t stride_bytes, uint32_t n_verts, const uint32_t *tri_idx, uint32_t n_tris,
               HostBVH8 *out, char *err, size_t err_len);
void free_bvh8(HostBVH8 *);

}  // namespace prt
