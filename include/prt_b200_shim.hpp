// prt_b200_shim.hpp -- header-only C++ shim that re-creates the reference's own entry points on top of the C ABI
// (include/prt_b200.h), so the reference's call sites compile unchanged:
//
//   void bake_SH(Mesh& gl_mesh);                    reference src/raytracing/raytracing.h:15, raytracing.cpp:320-360
//   class RTScene { RTScene(Mesh&); RTScene(Model&); ~RTScene(); }   raytracing.h:6-12, light_probe.h:6-12
//   struct Ray { first_hit, any_hit, hit_normal }   light_probe.cpp:95-133
//
// It is templated on the reference's Mesh / Model types instead of including "opengl/gl.h", so it also compiles in
// this repository's tests against a stand-in Mesh with the same members:
//   Mesh::Vert { glm::vec3 pos; glm::vec3 norm; float sh_coeff[9]; }   (60 bytes, gl.h:76-80)
//   verts(), edit_verts() -> std::vector<Vert>&, indices() -> std::vector<uint32_t>&
// Parameters that the reference reads from the App singleton (albedo, sh_resolution, max_path_length; app.h:55,70-71)
// are passed through prt_shim::Settings, defaulting to the reference's defaults.
#pragma once
#include "prt_b200.h"

#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <vector>

namespace prt_shim {

struct Settings {
    float albedo[3] = {1.f, 1.f, 1.f};   // App::albedo, app.h:55
    int sh_resolution = 32;              // App::sh_resolution, app.h:71
    int max_path_length = 2;             // App::max_path_length, app.h:70  (depth = max_path_length - 1)
    int order = 2;                       // file-scope `int order = 2` (raytracing.cpp:320): bands 0..order
    uint32_t seed = 0x50525400u;
    int device = -1;
};

inline void check(int rc, const char *what) {
    if (rc != PRT_OK) throw std::runtime_error(std::string(what) + ": " + prt_last_error());
}

// process-wide context, like the reference's global `RTCDevice device = initializeDevice();` (raytracing.cpp:52)
inline prt_ctx *device(int id = -1) {
    static prt_ctx *ctx = nullptr;
    if (!ctx) check(prt_ctx_create(id, &ctx), "prt_ctx_create");
    return ctx;
}

// RTScene(Mesh&) / RTScene(Model&): copy positions + index triples, build the BVH (raytracing.cpp:58-94).
class RTScene {
public:
    template <class MeshT>
    explicit RTScene(MeshT &gl_mesh) {
        const auto &v = gl_mesh.verts();
        const auto &i = gl_mesh.indices();
        using Vert = typename std::decay<decltype(v[0])>::type;
        check(prt_scene_create(device(), reinterpret_cast<const float *>(v.data()), sizeof(Vert), (uint32_t)v.size(),
                               reinterpret_cast<const uint32_t *>(i.data()), (uint32_t)(i.size() / 3), &scene),
              "prt_scene_create");
    }
    // RTScene(Model&): the reference only handles single-mesh models correctly (light_probe.cpp:72-83 writes every mesh
    // from index 0); this overload concatenates meshes properly.
    template <class ModelT>
    static RTScene from_model(ModelT &gl_model) {
        std::vector<float> pos;
        std::vector<uint32_t> idx;
        for (auto &m : gl_model.meshes) {
            const uint32_t base = (uint32_t)(pos.size() / 3);
            for (const auto &v : m.verts()) { pos.push_back(v.pos.x); pos.push_back(v.pos.y); pos.push_back(v.pos.z); }
            for (auto i : m.indices()) idx.push_back(base + (uint32_t)i);
        }
        RTScene s;
        check(prt_scene_create(device(), pos.data(), 12, (uint32_t)(pos.size() / 3), idx.data(), (uint32_t)(idx.size() / 3), &s.scene),
              "prt_scene_create");
        return s;
    }
    RTScene(RTScene &&o) noexcept : scene(o.scene) { o.scene = nullptr; }
    RTScene(const RTScene &) = delete;
    ~RTScene() { prt_scene_destroy(scene); }
    prt_scene *scene = nullptr;

private:
    RTScene() = default;
};

// struct Ray (light_probe.cpp:95-133): single-ray convenience on top of the batched ABI.
struct Ray {
    float r[8];
    float t = std::numeric_limits<float>::infinity();
    float ng[3] = {0, 0, 0};
    uint32_t prim = 0xFFFFFFFFu;
    Ray(const float org[3], const float dir[3], float tnear = 0.f, float tfar = std::numeric_limits<float>::infinity()) {
        r[0] = org[0]; r[1] = org[1]; r[2] = org[2]; r[3] = tnear; r[4] = dir[0]; r[5] = dir[1]; r[6] = dir[2]; r[7] = tfar;
    }
    bool first_hit(RTScene &s) { check(prt_trace_closest_hit(s.scene, r, 1, &t, &prim, ng), "prt_trace_closest_hit"); return prim != 0xFFFFFFFFu; }
    bool any_hit(RTScene &s) { uint8_t h = 0; check(prt_trace_any_hit(s.scene, r, 1, &h), "prt_trace_any_hit"); return h != 0; }
    const float *hit_normal() const { return ng; }   // unnormalised Ng, Embree convention (light_probe.cpp:123-125)
};

// void bake_SH(Mesh& gl_mesh): reads verts()[i].pos/.norm and indices(), writes edit_verts()[i].sh_coeff[k] in place.
template <class MeshT>
void bake_SH(MeshT &gl_mesh, const Settings &app = Settings()) {
    RTScene rtscene{gl_mesh};
    auto &verts = gl_mesh.edit_verts();   // sets Mesh::dirty like the reference (gl.cpp:280-283)
    using Vert = typename std::decay<decltype(verts[0])>::type;
    static_assert(sizeof(Vert) % 4 == 0, "Mesh::Vert must be float-aligned");
    prt_bake_params p;
    prt_bake_params_default(&p);
    p.order = app.order + 1;
    p.samples_u = p.samples_v = app.sh_resolution;
    p.seed = app.seed;
    p.bounces = app.max_path_length - 2;
    p.mode = p.bounces > 0 ? PRT_INTERREFLECT : PRT_SHADOWED;
    for (int c = 0; c < 3; c++) p.albedo[c] = app.albedo[c];
    const int n2 = p.order * p.order;
    std::vector<float> rows((size_t)verts.size() * n2);
    const float *base = reinterpret_cast<const float *>(verts.data());
    check(prt_bake_transfer(device(), rtscene.scene, base + offsetof(Vert, pos) / 4, base + offsetof(Vert, norm) / 4, sizeof(Vert),
                            (uint32_t)verts.size(), 0, &p, rows.data(), nullptr),
          "prt_bake_transfer");
    if (p.order >= 3) check(prt_scatter_sh9(rows.data(), p.order, (uint32_t)verts.size(), verts.data(), sizeof(Vert), offsetof(Vert, sh_coeff)), "prt_scatter_sh9");
}

}  // namespace prt_shim
