// bvh_build.cpp -- host builder for the 8-wide compressed BVH (layout in bvh8.h).
//
// Takes the place of rtcCommitScene() inside the reference's RTScene constructors
// (reference src/raytracing/raytracing.cpp:58-94, light_probe.cpp:44-87).  Three stages:
//   1. multi-threaded binned-SAH binary build over padded triangle boxes (leaves <= 3 triangles),
//   2. SAH-optimal collapse of the binary tree into 8-wide nodes (dynamic programming over slot budgets),
//   3. octant-ordered slot assignment, conservative 8-bit quantisation and depth-first emission so that
//      the internal children of a node and the triangles of its leaf slots are contiguous.
#include "bvh8.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace prt {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = 3.0e38f; hi[a] = -3.0e38f; } }
    void grow(const Box &b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct BNode {
    Box box;
    uint32_t left;   // internal: left child (right = left + 1); leaf: first index into ids
    uint32_t count;  // 0 = internal
};

constexpr int kBins = 64;
constexpr uint32_t kLeafMax = 3;
constexpr int kMedianDepth = 36;   // beyond this binary depth fall back to median splits (bounds the stack)

struct Builder {
    const Box *tb = nullptr;        // padded triangle boxes
    const float *cent = nullptr;    // centroids, 3 per triangle
    uint32_t *ids = nullptr;
    BNode *nodes = nullptr;
    std::atomic<uint32_t> n_nodes{0};

    struct Job { uint32_t node, first, count; int depth; };
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Job> queue;
    int busy = 0;
    bool done = false;

    // splits [first, first+count); returns mid or 0 when it must become a leaf
    uint32_t split(uint32_t first, uint32_t count, int depth, Box &nb) {
        Box cb; cb.reset(); nb.reset();
        for (uint32_t i = first; i < first + count; i++) {
            uint32_t t = ids[i];
            nb.grow(tb[t]);
            for (int a = 0; a < 3; a++) { float c = cent[3 * (size_t)t + a]; cb.lo[a] = std::min(cb.lo[a], c); cb.hi[a] = std::max(cb.hi[a], c); }
        }
        if (count <= kLeafMax) return 0;
        int best_axis = -1, best_bin = -1;
        float best_cost = 3.0e38f;
        const int nbin = (int)std::min<uint32_t>((uint32_t)kBins, std::max<uint32_t>(4u, 2u * count));     // few primitives need few bins
        if (depth < kMedianDepth) {
            for (int a = 0; a < 3; a++) {
                float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.f)) continue;
                float scale = (float)nbin / ext;
                uint32_t cnt[kBins];
                for (int b = 0; b < nbin; b++) cnt[b] = 0;
                Box bb[kBins];
                for (int b = 0; b < nbin; b++) bb[b].reset();
                for (uint32_t i = first; i < first + count; i++) {
                    uint32_t t = ids[i];
                    int b = std::min(nbin - 1, std::max(0, (int)((cent[3 * (size_t)t + a] - cb.lo[a]) * scale)));
                    cnt[b]++; bb[b].grow(tb[t]);
                }
                float ra[kBins]; uint32_t rc[kBins];
                Box acc; acc.reset(); uint32_t c = 0;
                for (int b = nbin - 1; b >= 0; b--) { acc.grow(bb[b]); c += cnt[b]; rc[b] = c; ra[b] = c ? acc.area() : 0.f; }
                acc.reset(); c = 0;
                for (int b = 0; b < nbin - 1; b++) {
                    acc.grow(bb[b]); c += cnt[b];
                    if (!c || !rc[b + 1]) continue;
                    float cost = acc.area() * (float)c + ra[b + 1] * (float)rc[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
                }
            }
        }
        uint32_t mid;
        if (best_axis >= 0) {
            int a = best_axis;
            float scale = (float)nbin / (cb.hi[a] - cb.lo[a]);
            uint32_t i = first, k = first + count;
            while (i < k) {
                int b = std::min(nbin - 1, std::max(0, (int)((cent[3 * (size_t)ids[i] + a] - cb.lo[a]) * scale)));
                if (b <= best_bin) i++; else std::swap(ids[i], ids[--k]);
            }
            mid = i;
            if (mid != first && mid != first + count) return mid;
        }
        // median split along the widest centroid axis (also the deep-tree fallback)
        int a = 0; float e = -1.f;
        for (int k = 0; k < 3; k++) { float d = cb.hi[k] - cb.lo[k]; if (d > e) { e = d; a = k; } }
        mid = first + count / 2;
        std::nth_element(ids + first, ids + mid, ids + first + count,
                         [&](uint32_t x, uint32_t y) { return cent[3 * (size_t)x + a] < cent[3 * (size_t)y + a]; });
        return mid;
    }

    void subtree(Job j, std::vector<Job> &local, bool share) {
        local.clear(); local.push_back(j);
        while (!local.empty()) {
            Job w = local.back(); local.pop_back();
            BNode &nd = nodes[w.node];
            uint32_t mid = split(w.first, w.count, w.depth, nd.box);
            if (!mid) { nd.left = w.first; nd.count = w.count; continue; }
            uint32_t l = n_nodes.fetch_add(2);
            nd.left = l; nd.count = 0;
            Job jl{l, w.first, mid - w.first, w.depth + 1}, jr{l + 1, mid, w.first + w.count - mid, w.depth + 1};
            if (share && w.count > 32768) {
                std::lock_guard<std::mutex> g(mu);
                queue.push_back(jl); queue.push_back(jr);
                cv.notify_all();
            } else { local.push_back(jr); local.push_back(jl); }
        }
    }

    void worker() {
        std::vector<Job> local;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty() || (busy == 0) || done; });
                if (queue.empty()) { if (busy == 0 || done) { done = true; cv.notify_all(); return; } continue; }
                j = queue.back(); queue.pop_back(); busy++;
            }
            subtree(j, local, true);
            {
                std::lock_guard<std::mutex> g(mu);
                busy--;
                if (busy == 0 && queue.empty()) done = true;
                cv.notify_all();
            }
        }
    }
};

inline uint8_t exp_byte(float ext) {
    // smallest e with 255 * 2^e >= ext
    if (!(ext > 0.f)) return 1;
    int e = (int)std::ceil(std::log2((double)ext / 255.0));
    while (std::ldexp(255.0, e) < (double)ext) e++;
    while (e > -126 && std::ldexp(255.0, e - 1) >= (double)ext) e--;
    int b = e + 127;
    return (uint8_t)std::min(254, std::max(1, b));
}

}  // namespace

int build_bvh8(const float *pos, size_t stride, uint32_t nv, const uint32_t *idx, uint32_t nt, HostBVH8 *out,
               char *err, size_t err_len) {
    auto fail = [&](const char *m) { if (err && err_len) snprintf(err, err_len, "%s", m); return -1; };
    if (!pos || !idx || !out || nt == 0 || nv == 0) return fail("build_bvh8: empty mesh");
    if (nt > 0x7FFFFFF0u / 3) return fail("build_bvh8: too many triangles");
    if (stride == 0) stride = 12;
    auto t0 = std::chrono::steady_clock::now();
    auto P = [&](uint32_t i) { return (const float *)((const char *)pos + (size_t)i * stride); };

    float amax = 0.f;
    for (uint32_t i = 0; i < nv; i++) { const float *p = P(i); for (int a = 0; a < 3; a++) { if (!std::isfinite(p[a])) return fail("build_bvh8: non-finite vertex"); amax = std::max(amax, std::fabs(p[a])); } }
    // Boxes are padded so that box culling (float slab test on quantised planes) can never reject a
    // triangle the pinned Moeller-Trumbore test accepts (error analysis).
    const float pad = amax * 1.6e-5f + 1e-30f;

    std::vector<Box> tb(nt);
    std::vector<float> cent(3 * (size_t)nt);
    std::vector<uint32_t> ids(nt);
    for (uint32_t t = 0; t < nt; t++) {
        uint32_t i0 = idx[3 * (size_t)t], i1 = idx[3 * (size_t)t + 1], i2 = idx[3 * (size_t)t + 2];
        if (i0 >= nv || i1 >= nv || i2 >= nv) return fail("build_bvh8: triangle index out of range");
        const float *a = P(i0), *b = P(i1), *c = P(i2);
        for (int k = 0; k < 3; k++) {
            float lo = std::min(a[k], std::min(b[k], c[k])), hi = std::max(a[k], std::max(b[k], c[k]));
            tb[t].lo[k] = lo - pad; tb[t].hi[k] = hi + pad; cent[3 * (size_t)t + k] = 0.5f * (lo + hi);
        }
        ids[t] = t;
    }

    // ---- stage 1: binary build ---------------------------------------------------------------------
    std::vector<BNode> bn(2 * (size_t)nt + 2);
    Builder B;
    B.tb = tb.data(); B.cent = cent.data(); B.ids = ids.data(); B.nodes = bn.data(); B.n_nodes = 1;
    B.queue.push_back({0, 0, nt, 0});
    unsigned nth = std::max(1u, std::thread::hardware_concurrency());
    if (nt < 65536) nth = 1;
    if (nth == 1) { std::vector<Builder::Job> local; Builder::Job j = B.queue.back(); B.queue.clear(); B.subtree(j, local, false); }
    else {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < nth; i++) th.emplace_back([&] { B.worker(); });
        for (auto &t : th) t.join();
    }

    // ---- stage 2+3: collapse to 8-wide, order slots, quantise, emit depth-first ----------------------
    const uint32_t n_bin = B.n_nodes.load();
    // SAH-optimal collapse (Ylitie, Karras & Laine 2017, section 4.1) by dynamic programming over the binary tree:
    //   C(n, i)  = least cost of representing the subtree of n with at most i slots of one 8-wide node
    //   C(n, 1)  = n is a leaf slot (binary leaves only: they already hold <= 3 triangles), or an 8-wide node of its own:
    //              area(n) * c_node + D(n, 8)
    //   D(n, j)  = min over k of C(left, k) + C(right, j - k)                       ("distribute j slots over the children")
    //   C(n, i)  = min(D(n, i), C(n, i - 1))
    // Every triangle ends up in its binary leaf whatever the cut, so the leaf terms are the same for all cuts and the recurrence
    // minimises the area-weighted number of 8-wide nodes: bottom nodes come out full instead of the 2-child nodes a top-down
    // greedy cut leaves behind (42 % of all nodes on the bench mesh).  Children are allocated after their parent, so one
    // reverse sweep visits children first.
    std::vector<float> dpC(7 * (size_t)n_bin, 0.f);          // C(n, 1..7); binary leaves: 0 (their cost is cut-independent)
    std::vector<uint8_t> dpK(8 * (size_t)n_bin, 0);          // dpK[8n + j - 1], j = 2..8: slots given to the left child by D(n, j); 0 = "use C(n, j-1)"
    for (uint32_t n = n_bin; n-- > 0;) {
        if (bn[n].count) continue;
        const uint32_t l = bn[n].left, r = l + 1;
        const float *Cl = &dpC[7 * (size_t)l], *Cr = &dpC[7 * (size_t)r];
        float D[9];
        for (int j = 2; j <= 8; j++) {
            float best = 3.0e38f; int bk = 1;
            for (int k = 1; k < j; k++) {
                if (k > 7 || j - k > 7) continue;
                const float c = Cl[k - 1] + Cr[j - k - 1];
                if (c < best) { best = c; bk = k; }
            }
            D[j] = best; dpK[8 * (size_t)n + j - 1] = (uint8_t)bk;
        }
        float *Cn = &dpC[7 * (size_t)n];
        Cn[0] = bn[n].box.area() + D[8];                     // c_node = 1
        for (int i = 2; i <= 7; i++) {
            if (D[i] < Cn[i - 2]) Cn[i - 1] = D[i];
            else { Cn[i - 1] = Cn[i - 2]; dpK[8 * (size_t)n + i - 1] = 0; }
        }
    }
    std::vector<Node8> wn; wn.reserve(n_bin / 4 + 16);
    Tri48 *tris = (Tri48 *)std::malloc(sizeof(Tri48) * (size_t)nt);
    if (!tris) return fail("build_bvh8: out of memory");
    uint32_t n_tri_out = 0, max_depth = 0;
    double sah = 0.0;
    const double root_area = std::max(1e-30f, bn[0].box.area());

    // triangle range ids[first, first + count) of every binary node (children are allocated after their parent)
    std::vector<uint32_t> bfirst(n_bin), bcount(n_bin);
    for (uint32_t n = n_bin; n-- > 0;) {
        if (bn[n].count) { bfirst[n] = bn[n].left; bcount[n] = bn[n].count; }
        else { bfirst[n] = bfirst[bn[n].left]; bcount[n] = bcount[bn[n].left] + bcount[bn[n].left + 1]; }
    }
    std::vector<uint32_t> wide_bnode; wide_bnode.reserve(n_bin / 4 + 16);      // binary node behind every 8-wide node
    wide_bnode.push_back(0);
    std::vector<uint32_t> wide_child; wide_child.reserve(8 * (n_bin / 4 + 16));  // [8 * node + slot] binary node of the child, 0xFFFFFFFF = empty
    wide_child.resize(8, 0xFFFFFFFFu);

    struct WJob { uint32_t bnode, wnode, depth; };
    std::vector<WJob> stack;
    wn.emplace_back(); std::memset(&wn[0], 0, sizeof(Node8));
    stack.push_back({0, 0, 1});
    while (!stack.empty()) {
        WJob j = stack.back(); stack.pop_back();
        max_depth = std::max(max_depth, j.depth);
        const BNode &root = bn[j.bnode];
        uint32_t ch[8]; int nch = 0;
        if (root.count) ch[nch++] = j.bnode;
        else { ch[nch++] = root.left; ch[nch++] = root.left + 1; }
        if (!root.count) {
            // unfold the optimal cut: (binary node, slot budget) pairs; budget 1 = the node takes one slot
            nch = 0;
            struct Cut { uint32_t node; int budget; };
            Cut cs[16]; int ncs = 0;
            const int k8 = dpK[8 * (size_t)j.bnode + 7];
            cs[ncs++] = {root.left + 1, 8 - k8};
            cs[ncs++] = {root.left, k8};
            while (ncs) {
                const Cut c = cs[--ncs];
                int i = c.budget;
                if (!bn[c.node].count) while (i > 1 && dpK[8 * (size_t)c.node + i - 1] == 0) i--;      // "use fewer slots"
                if (bn[c.node].count || i == 1) { ch[nch++] = c.node; continue; }
                const int k = dpK[8 * (size_t)c.node + i - 1];
                cs[ncs++] = {bn[c.node].left + 1, i - k};
                cs[ncs++] = {bn[c.node].left, k};
            }
        }
        // slot assignment: slot s should hold the child lying farthest along d_s = (s&4?+:-, s&2?+:-, s&1?+:-),
        // because a ray with sign octant o visits slots in decreasing (s ^ (7-o)) order.
        float cx = 0.5f * (root.box.lo[0] + root.box.hi[0]), cy = 0.5f * (root.box.lo[1] + root.box.hi[1]), cz = 0.5f * (root.box.lo[2] + root.box.hi[2]);
        float cost[8][8];
        for (int i = 0; i < nch; i++) {
            const Box &b = bn[ch[i]].box;
            float ox = 0.5f * (b.lo[0] + b.hi[0]) - cx, oy = 0.5f * (b.lo[1] + b.hi[1]) - cy, oz = 0.5f * (b.lo[2] + b.hi[2]) - cz;
            for (int s = 0; s < 8; s++) cost[i][s] = ((s & 4) ? ox : -ox) + ((s & 2) ? oy : -oy) + ((s & 1) ? oz : -oz);
        }
        int slot_child[8]; for (int s = 0; s < 8; s++) slot_child[s] = -1;
        bool used_c[8] = {false};
        for (int it = 0; it < nch; it++) {
            int bi = -1, bs = -1; float bc = -3.0e38f;
            for (int i = 0; i < nch; i++) if (!used_c[i])
                for (int s = 0; s < 8; s++) if (slot_child[s] < 0 && cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
            used_c[bi] = true; slot_child[bs] = bi;
        }
        Node8 nd; std::memset(&nd, 0, sizeof(nd));
        nd.px = root.box.lo[0]; nd.py = root.box.lo[1]; nd.pz = root.box.lo[2];
        nd.ex = exp_byte(root.box.hi[0] - root.box.lo[0]);
        nd.ey = exp_byte(root.box.hi[1] - root.box.lo[1]);
        nd.ez = exp_byte(root.box.hi[2] - root.box.lo[2]);
        const double sc[3] = { std::ldexp(1.0, (int)nd.ex - 127), std::ldexp(1.0, (int)nd.ey - 127), std::ldexp(1.0, (int)nd.ez - 127) };
        int n_internal = 0;
        for (int s = 0; s < 8; s++) if (slot_child[s] >= 0 && !bn[ch[slot_child[s]]].count) n_internal++;
        nd.child_base = (uint32_t)wn.size();
        nd.tri_base = n_tri_out;
        if (n_internal) { wn.resize(wn.size() + n_internal); wide_bnode.resize(wn.size()); wide_child.resize(8 * wn.size(), 0xFFFFFFFFu); }
        for (int s = 0; s < 8; s++) wide_child[8 * (size_t)j.wnode + s] = slot_child[s] < 0 ? 0xFFFFFFFFu : ch[slot_child[s]];
        uint32_t rank = 0, toff = 0;
        uint8_t *qlo[3] = { nd.qlox, nd.qloy, nd.qloz }, *qhi[3] = { nd.qhix, nd.qhiy, nd.qhiz };
        for (int s = 0; s < 8; s++) {
            if (slot_child[s] < 0) { for (int a = 0; a < 3; a++) { qlo[a][s] = 255; qhi[a][s] = 0; } nd.meta[s] = 0; continue; }
            const BNode &c = bn[ch[slot_child[s]]];
            const float plo[3] = { nd.px, nd.py, nd.pz };
            for (int a = 0; a < 3; a++) {
                double l = std::floor(((double)c.box.lo[a] - (double)plo[a]) / sc[a]);
                double h = std::ceil(((double)c.box.hi[a] - (double)plo[a]) / sc[a]);
                qlo[a][s] = (uint8_t)std::min(255.0, std::max(0.0, l));
                qhi[a][s] = (uint8_t)std::min(255.0, std::max(0.0, h));
            }
            sah += (double)c.box.area() / root_area * (c.count ? (double)c.count : 1.0);
            if (!c.count) {
                nd.imask |= (uint8_t)(1u << s);
                nd.meta[s] = (uint8_t)(0x20 | (24 + s));
                stack.push_back({ch[slot_child[s]], nd.child_base + rank, j.depth + 1});
                wide_bnode[nd.child_base + rank] = ch[slot_child[s]];
                rank++;
            } else {
                uint32_t unary = c.count == 1 ? 1u : (c.count == 2 ? 3u : 7u);
                nd.meta[s] = (uint8_t)((unary << 5) | toff);
                for (uint32_t k = 0; k < c.count; k++) {
                    uint32_t t = ids[c.left + k];
                    const float *a = P(idx[3 * (size_t)t]), *b = P(idx[3 * (size_t)t + 1]), *cc = P(idx[3 * (size_t)t + 2]);
                    Tri48 &o = tris[n_tri_out++];
                    o.v0x = a[0]; o.v0y = a[1]; o.v0z = a[2]; o.prim = t;
                    o.e1x = b[0] - a[0]; o.e1y = b[1] - a[1]; o.e1z = b[2] - a[2]; o.pad1 = 0;
                    o.e2x = cc[0] - a[0]; o.e2y = cc[1] - a[1]; o.e2z = cc[2] - a[2]; o.pad2 = 0;
                }
                toff += c.count;
            }
        }
        wn[j.wnode] = nd;
    }
    if (n_tri_out != nt) { std::free(tris); return fail("build_bvh8: internal error (triangle count)"); }
    if (max_depth + 2 > (uint32_t)kStackEntries) { std::free(tris); return fail("build_bvh8: tree too deep for the traversal stack"); }

    out->n_nodes = (uint32_t)wn.size();
    out->nodes = (Node8 *)std::malloc(sizeof(Node8) * wn.size());
    if (!out->nodes) { std::free(tris); return fail("build_bvh8: out of memory"); }
    std::memcpy(out->nodes, wn.data(), sizeof(Node8) * wn.size());
    // ---- oriented slabs (bvh8.h): mean normal of the triangles below every 8-wide node and their extent along it -----------------
    out->slabs = (Slab32 *)std::malloc(sizeof(Slab32) * wn.size());
    out->dops = (Dop32 *)std::malloc(sizeof(Dop32) * wn.size());
    if (!out->slabs || !out->dops) { std::free(tris); std::free(out->nodes); std::free(out->slabs); std::free(out->dops); out->nodes = nullptr; out->slabs = nullptr; out->dops = nullptr; return fail("build_bvh8: out of memory"); }
    {
        const uint32_t n_wide = (uint32_t)wn.size();
        std::atomic<uint32_t> next{0};
        auto slab_worker = [&]() {
            for (;;) {
                const uint32_t w0 = next.fetch_add(64);
                if (w0 >= n_wide) return;
                for (uint32_t w = w0; w < std::min(n_wide, w0 + 64); w++) {
                    const uint32_t b = wide_bnode[w], first = bfirst[b], count = bcount[b];
                    Slab32 sl; std::memset(&sl, 0, sizeof(sl));
                    sl.d0 = -3.0e38f; sl.d1 = 3.0e38f;
                    double sx = 0, sy = 0, sz = 0;
                    for (uint32_t i = first; i < first + count; i++) {
                        const uint32_t t = ids[i];
                        const float *a = P(idx[3 * (size_t)t]), *bb = P(idx[3 * (size_t)t + 1]), *c = P(idx[3 * (size_t)t + 2]);
                        const double e1[3] = {(double)bb[0] - a[0], (double)bb[1] - a[1], (double)bb[2] - a[2]};
                        const double e2[3] = {(double)c[0] - a[0], (double)c[1] - a[1], (double)c[2] - a[2]};
                        sx += e1[1] * e2[2] - e1[2] * e2[1]; sy += e1[2] * e2[0] - e1[0] * e2[2]; sz += e1[0] * e2[1] - e1[1] * e2[0];
                    }
                    const double len = std::sqrt(sx * sx + sy * sy + sz * sz);
                    if (len > 0.0 && std::isfinite(len)) {
                        // the slab is stated for the float vector the kernels will use
                        const float mx = (float)(sx / len), my = (float)(sy / len), mz = (float)(sz / len);
                        double lo = 3.0e38, hi = -3.0e38;
                        for (uint32_t i = first; i < first + count; i++) {
                            const uint32_t t = ids[i];
                            for (int k = 0; k < 3; k++) {
                                const float *q = P(idx[3 * (size_t)t + k]);
                                const double d = (double)mx * q[0] + (double)my * q[1] + (double)mz * q[2];
                                lo = std::min(lo, d); hi = std::max(hi, d);
                            }
                        }
                        // same padding as the triangle boxes: absorbs the float evaluation of m . x on the device
                        sl.mx = mx; sl.my = my; sl.mz = mz;
                        sl.d0 = std::nextafter((float)(lo - (double)pad), -3.0e38f); sl.d1 = std::nextafter((float)(hi + (double)pad), 3.0e38f);
                    }
                    out->slabs[w] = sl;
                    // fourth slab axis (Dop32): extents of the children along m on a 255-step grid over the node's own extent, which is
                    // floored at 2^-9 of the scene size so that the float evaluation of M . o stays far below one step
                    Dop32 dp; std::memset(&dp, 0, sizeof(dp));
                    for (int s = 0; s < 8; s++) { dp.qlo[s] = 0; dp.qhi[s] = 255; }
                    if (sl.mx != 0.f || sl.my != 0.f || sl.mz != 0.f) {
                        const double T = std::max((double)sl.d1 - (double)sl.d0, (double)amax * (1.0 / 512.0));
                        const double scale = 255.0 / T, base = 0.5 * ((double)sl.d0 + (double)sl.d1) - 0.5 * T;
                        dp.Mx = (float)(sl.mx * scale); dp.My = (float)(sl.my * scale); dp.Mz = (float)(sl.mz * scale);
                        dp.D0 = (float)(base * scale);
                        for (int s = 0; s < 8; s++) {
                            const uint32_t cb = wide_child[8 * (size_t)w + s];
                            if (cb == 0xFFFFFFFFu) continue;
                            double lo = 3.0e38, hi = -3.0e38;
                            for (uint32_t i = bfirst[cb]; i < bfirst[cb] + bcount[cb]; i++) {
                                const uint32_t t = ids[i];
                                for (int k = 0; k < 3; k++) {
                                    const float *q = P(idx[3 * (size_t)t + k]);
                                    // evaluated with the float vector the kernels use
                                    const double d = (double)dp.Mx * q[0] + (double)dp.My * q[1] + (double)dp.Mz * q[2] - (double)dp.D0;
                                    lo = std::min(lo, d); hi = std::max(hi, d);
                                }
                            }
                            const double padq = (double)pad * scale + 1.0;            // the triangle-box padding, in steps, + one step
                            dp.qlo[s] = (uint8_t)std::min(255.0, std::max(0.0, std::floor(lo - padq)));
                            dp.qhi[s] = (uint8_t)std::min(255.0, std::max(0.0, std::ceil(hi + padq)));
                        }
                    }
                    out->dops[w] = dp;
                }
            }
        };
        if (nth == 1) slab_worker();
        else {
            std::vector<std::thread> th;
            for (unsigned i = 0; i < nth; i++) th.emplace_back(slab_worker);
            for (auto &t : th) t.join();
        }
    }
    out->tris = tris; out->n_tris = nt; out->max_depth = max_depth; out->sah_cost = sah; out->pad = pad;
    out->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

void free_bvh8(HostBVH8 *b) {
    if (!b) return;
    std::free(b->nodes); std::free(b->tris); std::free(b->slabs); std::free(b->dops);
    b->nodes = nullptr; b->tris = nullptr; b->slabs = nullptr; b->dops = nullptr;
}

}  // namespace prt
