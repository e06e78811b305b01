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
