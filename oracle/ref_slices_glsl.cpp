// oracle/ref_slices_glsl.cpp -- TEST INFRASTRUCTURE ONLY.  The GLSL half of ref_slices.cpp (see its header): the reference's
// shaders compiled as C++ through its vendored glm, with -fsingle-precision-constant so that literals are float as in GLSL.
#include <cstdio>
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <glm/glm.hpp>
#include "prt_oracle.h"

// ---- GLSL slices ----------------------------------------------------------------------------------------------------
namespace glsl_env {
using namespace glm;
typedef unsigned int uint;
struct texel_t { vec3 rgb; };
struct samplerCube { const float *cube; int n0, levels; };
static texel_t textureLod(const samplerCube &s, vec3 d, float lod) {
    float dd[3] = { d.x, d.y, d.z }, o[3];
    prt_o_cube_sample(s.cube, s.n0, s.levels, dd, lod, o);
    return texel_t{ vec3(o[0], o[1], o[2]) };
}
static texel_t texture(const samplerCube &s, vec3 d) { return textureLod(s, d, 0.0f); }   // implicit LOD pinned to level 0 (env.c)
}  // namespace glsl_env

namespace ref_brdf {
using namespace glm;
typedef unsigned int uint;
#include "_ref/slice_brdf.inc"
}
namespace ref_prefilter {
using namespace glsl_env;
static vec4 FragColor; static vec3 CubeTexPos; static samplerCube environment; static float roughness;
#define main shader_main
#include "_ref/slice_prefilter.inc"
#undef main
}
namespace ref_irradiance {
using namespace glsl_env;
static vec4 FragColor; static vec3 CubeTexPos; static samplerCube environment;
#define main shader_main
#include "_ref/slice_irradiance.inc"
#undef main
}
namespace ref_rect2cube {
using namespace glm;
#include "_ref/slice_rect2cube.inc"
}


// ---- precomp_projectSH.comp (the live per-probe projection kernel, volume.cpp:388-416): lines 22-23 (constants) and 32-143 (main) ----
// One work group = 128 invocations run as 128 host threads; barrier() is a std::barrier, `shared` a static array, the buffer / image
// bindings are plain arrays (texelFetch / imageStore below).  RGBA16F image stores are kept in float (storage format = driver side).
#include <barrier>
#include <thread>
#include <vector>
namespace ref_project {
using namespace glm;
typedef unsigned int uint;
#define Thread_Size 128
#define THIRD_BAND
#define NUM_COEFF 9
struct samplerBuffer { const float *p; int comps; };
struct usamplerBuffer { const unsigned *p; };
struct usampler3D { const unsigned *p; };                       // probe id = (x, 0, 0): one row of probes
struct image3D { float *p; };
struct ftexel { float x; vec3 rgb; };
struct utexel { unsigned x; };
struct u2texel { uvec2 xy; };
static ftexel texelFetch(const samplerBuffer &s, int i) { ftexel t; t.x = s.p[(size_t)i * s.comps]; t.rgb = s.comps >= 3 ? vec3(s.p[(size_t)i * s.comps], s.p[(size_t)i * s.comps + 1], s.p[(size_t)i * s.comps + 2]) : vec3(t.x); return t; }
static utexel texelFetch(const usamplerBuffer &s, int i) { return utexel{ s.p[i] }; }
static u2texel texelFetch(const usampler3D &s, ivec3 id, int) { return u2texel{ uvec2(s.p[2 * id.x], s.p[2 * id.x + 1]) }; }
static void imageStore(image3D &im, ivec3 id, vec4 v) { float *o = im.p + 4 * (size_t)id.x; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
static samplerBuffer radiance, transfer;
static usamplerBuffer range_ID;
static usampler3D probe_range;
static image3D SH_Ar, SH_Ag, SH_Ab, SH_Br, SH_Bg, SH_Bb, SH_C;
static vec3 surfel_SH[NUM_COEFF][Thread_Size];
static thread_local uvec3 gl_WorkGroupID, gl_LocalInvocationID;
static inline vec3 operator/(const vec3 &v, int s) { return v / float(s); }        // GLSL converts the int operand implicitly; glm's template cannot
static std::barrier<> *g_bar = nullptr;
static void barrier() { g_bar->arrive_and_wait(); }
#define main shader_main
#include "_ref/slice_project_const.inc"
#include "_ref/slice_project_main.inc"
#undef main
}  // namespace ref_project

// out: [n_probes][7][4] = Ar Ag Ab Br Bg Bb C
void ref_glsl_project(const unsigned *range2, int n_probes, const unsigned *ids, const float *transfer9, const float *radiance_rgba, float *out) {
    using namespace ref_project;
    std::vector<float> img[7];
    for (auto &v : img) v.assign((size_t)n_probes * 4, 0.f);
    ref_project::radiance = samplerBuffer{ radiance_rgba, 4 }; ref_project::transfer = samplerBuffer{ transfer9, 1 };
    range_ID = usamplerBuffer{ ids }; probe_range = usampler3D{ range2 };
    SH_Ar.p = img[0].data(); SH_Ag.p = img[1].data(); SH_Ab.p = img[2].data(); SH_Br.p = img[3].data(); SH_Bg.p = img[4].data(); SH_Bb.p = img[5].data(); SH_C.p = img[6].data();
    for (int p = 0; p < n_probes; p++) {                                   // glDispatchCompute(probe_res.x, y, z): one group per probe
        std::barrier<> bar(Thread_Size);
        g_bar = &bar;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < Thread_Size; t++)
            th.emplace_back([p, t]() { gl_WorkGroupID = glm::uvec3(p, 0, 0); gl_LocalInvocationID = glm::uvec3(t, 0, 0); shader_main(); });
        for (auto &x : th) x.join();
    }
    for (int p = 0; p < n_probes; p++)
        for (int k = 0; k < 7; k++)
            for (int c = 0; c < 4; c++) out[((size_t)p * 7 + k) * 4 + c] = img[k][(size_t)p * 4 + c];
}

void ref_glsl_brdf(float ndotv, float roughness, float out[2]) {
    glm::vec2 r = ref_brdf::IntegrateBRDF(ndotv, roughness);
    out[0] = r.x; out[1] = r.y;
}
void ref_glsl_prefilter(const float *cube, int n0, int levels, const float P[3], float roughness, float out[3]) {
    ref_prefilter::environment = glsl_env::samplerCube{ cube, n0, levels };
    ref_prefilter::CubeTexPos = glm::vec3(P[0], P[1], P[2]);
    ref_prefilter::roughness = roughness;
    ref_prefilter::shader_main();
    out[0] = ref_prefilter::FragColor.x; out[1] = ref_prefilter::FragColor.y; out[2] = ref_prefilter::FragColor.z;
}
void ref_glsl_irradiance(const float *cube, int n0, int levels, const float P[3], float out[3]) {
    ref_irradiance::environment = glsl_env::samplerCube{ cube, n0, levels };
    ref_irradiance::CubeTexPos = glm::vec3(P[0], P[1], P[2]);
    ref_irradiance::shader_main();
    out[0] = ref_irradiance::FragColor.x; out[1] = ref_irradiance::FragColor.y; out[2] = ref_irradiance::FragColor.z;
}
void ref_glsl_rect2cube(const float v[3], float uv[2]) {
    glm::vec2 r = ref_rect2cube::SampleSphericalMap(glm::vec3(v[0], v[1], v[2]));
    uv[0] = r.x; uv[1] = r.y;
}
