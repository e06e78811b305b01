// prt_b200_shim.hpp -- header-only C++ shim that re-creates the reference's own entry points on top of the C ABI
// (include/prt_b200.h), so the reference's call sites compile unchanged:
//
//   void bake_SH(Mesh& gl_mesh);                    reference src/raytracing/raytracing.h:15, raytracing.cpp:320-360
//   class RTScene { RTScene(Mesh&); RTScene(Model&); ~RTScene(); }   raytracing.h:6-12, light_probe.h:6-12
//   struct Ray { first_hit, any_hit, hit_normal }   light_probe.cpp:95-133
//   Volume_weight calculate_weight(Model&, ivec3 probe_res, ivec3 volume_res, vec3 scene_size)   light_probe.h:14-18
//   SH_volume::precompute / project_sh              src/sh/volume.h:14,55-56 (volume.cpp:149-316,388-452)   -> prt_shim::ProbeBake
//   LightProbe::equirectangular_to_cubemap / irradiance / prefilter, the BRDF LUT pass   src/opengl/gl.h:273-298, app.cpp:61-63
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
#include <utility>
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
    std::vector<int> devices;            // more than one entry: bake_SH shards the vertices over these GPUs (prt_group_*)
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
template <class MeshT> void bake_SH_multi(MeshT &gl_mesh, const Settings &app);
template <class MeshT>
void bake_SH(MeshT &gl_mesh, const Settings &app = Settings()) {
    if (app.devices.size() > 1) { bake_SH_multi(gl_mesh, app); return; }
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

// bake_SH over several GPUs of one box (Settings::devices): same contract, vertices sharded by the library's multi-GPU driver.
template <class MeshT>
void bake_SH_multi(MeshT &gl_mesh, const Settings &app) {
    prt_group *g = nullptr;
    check(prt_group_create(app.devices.data(), (int)app.devices.size(), &g), "prt_group_create");
    prt_group_scene *gs = nullptr;
    const auto &v = gl_mesh.verts();
    const auto &idx = gl_mesh.indices();
    using Vert = typename std::decay<decltype(v[0])>::type;
    int rc = prt_group_scene_create(g, reinterpret_cast<const float *>(v.data()), sizeof(Vert), (uint32_t)v.size(),
                                    reinterpret_cast<const uint32_t *>(idx.data()), (uint32_t)(idx.size() / 3), &gs);
    if (rc != PRT_OK) { prt_group_destroy(g); check(rc, "prt_group_scene_create"); }
    auto &verts = gl_mesh.edit_verts();
    prt_bake_params p;
    prt_bake_params_default(&p);
    p.order = app.order + 1; p.samples_u = p.samples_v = app.sh_resolution; p.seed = app.seed; p.bounces = app.max_path_length - 2;
    p.mode = p.bounces > 0 ? PRT_INTERREFLECT : PRT_SHADOWED;
    for (int c = 0; c < 3; c++) p.albedo[c] = app.albedo[c];
    std::vector<float> rows((size_t)verts.size() * p.order * p.order);
    const float *base = reinterpret_cast<const float *>(verts.data());
    rc = prt_group_bake_transfer(g, gs, base + offsetof(Vert, pos) / 4, base + offsetof(Vert, norm) / 4, sizeof(Vert), (uint32_t)verts.size(), &p,
                                 rows.data(), PRT_GATHER_AUTO, nullptr);
    prt_group_scene_destroy(gs); prt_group_destroy(g);
    check(rc, "prt_group_bake_transfer");
    if (p.order >= 3) check(prt_scatter_sh9(rows.data(), p.order, (uint32_t)verts.size(), verts.data(), sizeof(Vert), offsetof(Vert, sh_coeff)), "prt_scatter_sh9");
}

// Volume_weight calculate_weight(Model&, glm::ivec3 probe_res, glm::ivec3 volume_res, glm::vec3 scene_size) (light_probe.h:14-18,
// light_probe.cpp:156-367).  Vec4 = glm::vec4 in the reference; any 16-byte 4-float type works.
template <class Vec4>
struct Volume_weight_t {
    std::vector<Vec4> weight0123;
    std::vector<Vec4> weight4567;
};
template <class Vec4, class ModelT, class IVec3, class Vec3>
Volume_weight_t<Vec4> calculate_weight(ModelT &gl_model, IVec3 probe_res, IVec3 volume_res, Vec3 scene_size) {
    static_assert(sizeof(Vec4) == 16, "Volume_weight holds glm::vec4");
    RTScene rtscene = RTScene::from_model(gl_model);
    const int32_t pr[3] = {(int32_t)probe_res.x, (int32_t)probe_res.y, (int32_t)probe_res.z};
    const int32_t vr[3] = {(int32_t)volume_res.x, (int32_t)volume_res.y, (int32_t)volume_res.z};
    const float sz[3] = {(float)scene_size.x, (float)scene_size.y, (float)scene_size.z};
    Volume_weight_t<Vec4> result;
    const size_t n = (size_t)vr[0] * vr[1] * vr[2];
    result.weight0123.resize(n); result.weight4567.resize(n);
    check(prt_volume_weights(rtscene.scene, pr, vr, sz, reinterpret_cast<float *>(result.weight0123.data()),
                             reinterpret_cast<float *>(result.weight4567.data()), nullptr), "prt_volume_weights");
    return result;
}

// SH_volume::precompute (volume.cpp:149-316): what the reference uploads afterwards -- probe_range (:286-295), ID_buffer (:275-276),
// transfer_buffer (:273-274, 9 floats per entry), primitive_buffer (:314-315, 2 x vec3 per surfel), num_primitive -- and
// SH_volume::project_sh (volume.cpp:388-452) on the result.  Probe positions follow SH_volume::init (volume.cpp:83-90); the capture
// directions are the reference's 64 x 64 x 6 cube texels with its 4 / res^2 / |c|^3 solid angles (volume.cpp:251-254).
struct ProbeBake {
    std::vector<uint32_t> range;         // [n_probes][2] (start, end)
    std::vector<uint32_t> ids;           // [nnz] surfel id
    std::vector<float> transfer;         // [nnz][9]
    std::vector<float> surfels;          // [num_primitive][6] mean position, normalised mean normal
    uint32_t n_probes = 0, num_primitive = 0;
    prt_csr *csr = nullptr;              // device-resident CSR for project_sh / prt_gi_create
    ProbeBake() = default;
    ProbeBake(const ProbeBake &) = delete;
    ProbeBake(ProbeBake &&o) noexcept { *this = std::move(o); }
    ProbeBake &operator=(ProbeBake &&o) noexcept {
        range.swap(o.range); ids.swap(o.ids); transfer.swap(o.transfer); surfels.swap(o.surfels);
        n_probes = o.n_probes; num_primitive = o.num_primitive; std::swap(csr, o.csr);
        return *this;
    }
    ~ProbeBake() { prt_csr_destroy(csr); }
    // [n_probes][7][4] packed volumes Ar Ag Ab Br Bg Bb C (precomp_projectSH.comp:118-139) from per-surfel radiance [num_primitive][4]
    std::vector<float> project_sh(const std::vector<float> &radiance_rgba) const {
        if (radiance_rgba.size() != (size_t)num_primitive * 4) throw std::runtime_error("project_sh: radiance must be [num_primitive][4]");
        std::vector<float> out((size_t)n_probes * 28);
        check(prt_probe_project(csr, radiance_rgba.data(), out.data()), "prt_probe_project");
        return out;
    }
};
inline ProbeBake precompute(RTScene &scene, const int32_t probe_res[3], const float scene_size[3], int cubemap_res = 64) {
    ProbeBake b;
    b.n_probes = (uint32_t)(probe_res[0] * probe_res[1] * probe_res[2]);
    std::vector<float> probes((size_t)b.n_probes * 3), dirs((size_t)6 * cubemap_res * cubemap_res * 3), w((size_t)6 * cubemap_res * cubemap_res);
    check(prt_probe_positions(probe_res, scene_size, probes.data()), "prt_probe_positions");
    check(prt_cube_dirs(cubemap_res, dirs.data(), w.data()), "prt_cube_dirs");
    // the capture kernel keeps a probe's rays in shared memory: <= 4096 directions per probe.  The reference's 64^2 x 6 = 24 576 texels
    // do not fit; the largest cube-texel set that does is 26^2 x 6 = 4056 (same directions rule, same solid-angle formula)
    if (w.size() > 4096) {
        cubemap_res = 26;
        dirs.resize((size_t)6 * cubemap_res * cubemap_res * 3); w.resize((size_t)6 * cubemap_res * cubemap_res);
        check(prt_cube_dirs(cubemap_res, dirs.data(), w.data()), "prt_cube_dirs");
    }
    check(prt_probe_capture(scene.scene, probes.data(), b.n_probes, dirs.data(), w.data(), (uint32_t)w.size(), &b.csr), "prt_probe_capture");
    uint64_t nnz = 0;
    check(prt_csr_sizes(b.csr, nullptr, &nnz, &b.num_primitive, nullptr), "prt_csr_sizes");
    b.range.resize((size_t)b.n_probes * 2); b.ids.resize(nnz); b.transfer.resize(nnz * 9); b.surfels.resize((size_t)b.num_primitive * 6);
    check(prt_csr_download(b.csr, b.range.data(), b.ids.data(), b.transfer.data(), b.surfels.data(), nullptr), "prt_csr_download");
    return b;
}

// LightProbe passes (gl.h:273-298, gl.cpp:546-591) + the BRDF LUT (app.cpp:61-63) with host images in and out.
class LightProbe {
public:
    // load_hdr + equirectangular_to_cubemap + generateMipmap (util.cpp:6-24, gl.cpp:581-591,454-460); cube_size 512 in app.cpp:44
    LightProbe(const float *equirect_rgb, int w, int h, int cube_size = 512) : n0(cube_size) {
        check(prt_env_create(device(), equirect_rgb, w, h, cube_size, &env), "prt_env_create");
    }
    LightProbe(const LightProbe &) = delete;
    ~LightProbe() { prt_env_destroy(env); }
    std::vector<float> cube(int level = 0) {                                     // [6][n][n][3]
        const int n = n0 >> level;
        std::vector<float> out((size_t)6 * n * n * 3);
        check(prt_env_get_cube(env, level, out.data()), "prt_env_get_cube");
        return out;
    }
    std::vector<float> irradiance(int n_out = 32) {                              // app.cpp:55
        std::vector<float> out((size_t)6 * n_out * n_out * 3);
        check(prt_env_irradiance(env, n_out, out.data()), "prt_env_irradiance");
        return out;
    }
    std::vector<float> prefilter(int n_out = 256, int mips = 5, int n_samples = 1024) {     // app.cpp:58; mips one after another
        size_t total = 0;
        for (int m = 0; m < mips; m++) total += (size_t)6 * (n_out >> m) * (n_out >> m) * 3;
        std::vector<float> out(total);
        check(prt_env_prefilter(env, n_out, mips, n_samples, out.data()), "prt_env_prefilter");
        return out;
    }
    std::vector<float> project_sh(int order = 3, int method = 0, int size = 256) {           // [order^2][3]
        std::vector<float> out((size_t)order * order * 3);
        check(prt_env_project_sh(env, order, method, size, out.data()), "prt_env_project_sh");
        return out;
    }
    prt_env *env = nullptr;
private:
    int n0;
};
inline std::vector<float> brdf_lut(int w = 512, int h = 512, int n_samples = 1024) {          // [h][w][2] = (A, B)
    std::vector<float> out((size_t)w * h * 2);
    check(prt_brdf_lut(device(), w, h, n_samples, out.data()), "prt_brdf_lut");
    return out;
}

}  // namespace prt_shim
