// oracle/ref_slices.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the pieces of the reference's OWN hot-path source that need neither Embree nor a GL context, straight from
// /root/reference: `make -C oracle ref` cuts the line ranges below out of the reference files with sed into oracle/_ref/*.inc
// (git-ignored build output, never committed) and this driver #includes them unmodified.  Its outputs are the golden fixtures
// tests/golden/ref_*.txt (tests/golden/make_golden.py) the oracle -- and through it the CUDA kernels -- are pinned against.
//
//   slice_util_pi.inc        src/util/util.h:6-7                       PI, INV_PI (float)
//   slice_frame.inc          src/raytracing/raytracing.cpp:101-107     frame(N)
//   slice_sampling.inc       src/raytracing/raytracing.cpp:130-160     UniformSampleDisk, cosineSampleHemisphere{,PDF}, (u,v,N) overload
//   slice_get_dirs.inc       src/raytracing/light_probe.cpp:136-152    get_dirs (Fibonacci sphere)
//   slice_brdf.inc           src/shaders/brdf.frag:5-107               Hammersley, ImportanceSampleGGX, GeometrySmith, IntegrateBRDF
//   slice_prefilter.inc      src/shaders/prefilter.frag:8-107          DistributionGGX, ..., main()
//   slice_irradiance.inc     src/shaders/irradiance.frag:7-43          main()
//   slice_rect2cube.inc      src/shaders/rectangle2cube.frag:7-15      SampleSphericalMap
//   slice_project_*.inc      src/shaders/precomp_projectSH.comp:22-23,32-143   the live per-probe projection kernel: CSR SpMV over 128
//                                                                      invocations, shared-memory tree reduction, window, R-H pack
//
// The GLSL is compiled as C++ through the reference's vendored glm (`using namespace glm`) in a translation unit of its own
// (ref_slices_glsl.cpp) with -fsingle-precision-constant, so that literals are float as in GLSL; the C++ slices are compiled
// without that flag (get_dirs computes in double).  samplerCube / texture() / textureLod() are bound to the oracle's PINNED sampler
// (oracle/env.c prt_o_cube_sample: the GL driver's filtering is not in the reference tree, see env.c's header) -- so the
// shaders' own arithmetic (sample generation, pdf, mip choice, weights, normalisation) is what these fixtures pin.
// <stdlib.h>/<math.h> are included so that the unqualified abs(N.z) of raytracing.cpp:103,155 binds to the float overload,
// as it does under the reference's MSVC toolchain (with <cmath> alone glibc would pick int abs(int)).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <algorithm>
#include <utility>
#include <vector>
#include <glm/glm.hpp>
#include "prt_oracle.h"

// ---- C++ slices -----------------------------------------------------------------------------------------------------
namespace ref_rt {
#include "_ref/slice_util_pi.inc"
#include "_ref/slice_frame.inc"
#include "_ref/slice_sampling.inc"
}
namespace ref_lp {
#include "_ref/slice_get_dirs.inc"
}

// ---- GLSL slices: compiled in ref_slices_glsl.cpp with -fsingle-precision-constant (GLSL literals are float); the C++ slices above
// keep C++ literal semantics (double), as under the reference's own compiler
void ref_glsl_brdf(float ndotv, float roughness, float out[2]);
void ref_glsl_prefilter(const float *cube, int n0, int levels, const float P[3], float roughness, float out[3]);
void ref_glsl_irradiance(const float *cube, int n0, int levels, const float P[3], float out[3]);
void ref_glsl_rect2cube(const float v[3], float uv[2]);
void ref_glsl_project(const unsigned *range2, int n_probes, const unsigned *ids, const float *transfer9, const float *radiance_rgba, float *out);

static std::vector<float> read_floats(const char *path) {
    std::vector<float> v;
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    v.resize((size_t)n / 4);
    if (fread(v.data(), 4, v.size(), f) != v.size()) { fprintf(stderr, "short read\n"); exit(2); }
    fclose(f);
    return v;
}

// deterministic directions / parameters shared with the tests (which re-create them from the fixture's own columns)
static void fib(int i, int n, float d[3]) {
    double z = 1 - (double(i) + 0.5) / n * 2, th = 2.39996322972865332 * i, r = std::sqrt(1 - z * z);
    d[0] = float(std::cos(th) * r); d[1] = float(std::sin(th) * r); d[2] = float(z);
}

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "sampling")) {
        // frame(N) and cosineSampleHemisphere(u, v, N): 40 normals (incl. the |N.z| >= 0.99 branch and the axes) x 12 (u, v)
        const int nn = 40;
        for (int i = 0; i < nn; i++) {
            float n[3];
            fib(i, nn - 4, n);
            if (i == nn - 4) { n[0] = 0; n[1] = 0; n[2] = 1; }
            if (i == nn - 3) { n[0] = 0; n[1] = 0; n[2] = -1; }
            if (i == nn - 2) { n[0] = 0.1f; n[1] = 0.05f; n[2] = 0.99373f; float l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]); n[0] /= l; n[1] /= l; n[2] /= l; }
            if (i == nn - 1) { n[0] = 1; n[1] = 0; n[2] = 0; }
            glm::vec3 N(n[0], n[1], n[2]);
            glm::mat3 F = ref_rt::frame(N);
            printf("frame %.9g %.9g %.9g : %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", N.x, N.y, N.z, F[0].x, F[0].y, F[0].z, F[1].x,
                   F[1].y, F[1].z, F[2].x, F[2].y, F[2].z);
            for (int k = 0; k < 12; k++) {
                float u = (float(k % 4) + 0.37f) / 4.0f, v = (float(k) + 0.61f) / 12.0f;
                if (k == 0) { u = 0.0f; v = 0.0f; }
                if (k == 11) { u = 0.999999f; v = 0.999999f; }
                auto s = ref_rt::cosineSampleHemisphere(u, v, N);
                glm::vec3 l = ref_rt::cosineSampleHemisphere(u, v);
                printf("sample %.9g %.9g %.9g %.9g %.9g : %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", N.x, N.y, N.z, u, v, l.x, l.y, l.z, s.first.x,
                       s.first.y, s.first.z, s.second);
            }
        }
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "get_dirs")) {
        auto d = ref_lp::get_dirs(atoi(argv[2]));
        for (auto &v : d) printf("dir %.9g %.9g %.9g\n", v.x, v.y, v.z);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "brdf")) {
        // IntegrateBRDF at the texel centres of a w x h LUT (TexCoords of the full-screen quad, gl.h:204-209)
        const int w = atoi(argv[2]), h = atoi(argv[3]);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                float r[2];
                ref_glsl_brdf((float(x) + 0.5f) / float(w), (float(y) + 0.5f) / float(h), r);
                printf("lut %d %d : %.9g %.9g\n", x, y, r[0], r[1]);
            }
        return 0;
    }
    if (argc >= 5 && (!strcmp(argv[1], "prefilter") || !strcmp(argv[1], "irradiance"))) {
        // cube file: raw float32 levels as laid out by oracle/env.c (prt_o_env_equirect_to_cube); n0 = face size; n_out = size of
        // the rendered cube.  Fragments = texel centres; CubeTexPos is the un-normalised position on the unit cube (skybox.vert
        // passes the cube vertex through), from the GL cube-map face table (OpenGL 4.5 spec table 8.19).
        std::vector<float> cube = read_floats(argv[2]);
        const int n0 = atoi(argv[3]), n_out = atoi(argv[4]);
        const int levels = prt_o_cube_levels(n0);
        const bool pre = !strcmp(argv[1], "prefilter");
        for (int mip = 0; mip < (pre ? 5 : 1); mip++) {
            const int n = n_out >> mip;
            for (int f = 0; f < 6; f++) {
                const int ti[4] = { 0, n - 1, n / 2, n / 3 }, tj[4] = { 0, 0, n / 3, n - 1 };
                for (int k = 0; k < (n <= 2 ? n * n : 4); k++) {
                    const int i = n <= 2 ? k % n : ti[k], j = n <= 2 ? k / n : tj[k];
                    const float u = 2.0f * (float(i) + 0.5f) / float(n) - 1.0f, v = 2.0f * (float(j) + 0.5f) / float(n) - 1.0f;
                    const float T[6][3] = { { 1, -v, -u }, { -1, -v, u }, { u, 1, v }, { u, -1, -v }, { u, -v, 1 }, { -u, -v, -1 } };
                    float o[3];
                    if (pre) {
                        ref_glsl_prefilter(cube.data(), n0, levels, T[f], float(mip) / 4.0f, o);   // roughness = mip / 4: gl.cpp:556-566
                        printf("prefilter %d %d %d %d : %.9g %.9g %.9g\n", mip, f, i, j, o[0], o[1], o[2]);
                    } else {
                        ref_glsl_irradiance(cube.data(), n0, levels, T[f], o);
                        printf("irradiance %d %d %d : %.9g %.9g %.9g\n", f, i, j, o[0], o[1], o[2]);
                    }
                }
            }
        }
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "project")) {
        // csr.bin: uint32 n_probes, nnz, n_surfels; range[n_probes][2]; ids[nnz]; float transfer[nnz][9]; float radiance[n_surfels][4]
        FILE *f = fopen(argv[2], "rb");
        if (!f) { perror(argv[2]); return 2; }
        unsigned hdr[3];
        if (fread(hdr, 4, 3, f) != 3) return 2;
        std::vector<unsigned> range(2 * (size_t)hdr[0]), ids(hdr[1]);
        std::vector<float> tr(9 * (size_t)hdr[1]), rad(4 * (size_t)hdr[2]), out(28 * (size_t)hdr[0]);
        if (fread(range.data(), 8, hdr[0], f) != hdr[0] || fread(ids.data(), 4, hdr[1], f) != hdr[1] || fread(tr.data(), 36, hdr[1], f) != hdr[1] ||
            fread(rad.data(), 16, hdr[2], f) != hdr[2]) return 2;
        fclose(f);
        ref_glsl_project(range.data(), (int)hdr[0], ids.data(), tr.data(), rad.data(), out.data());
        for (unsigned p = 0; p < hdr[0]; p++) {
            printf("probe %u :", p);
            for (int k = 0; k < 28; k++) printf(" %.9g", out[28 * (size_t)p + k]);
            printf("\n");
        }
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "rect2cube")) {
        const int n = atoi(argv[2]);
        for (int i = 0; i < n; i++) {
            float d[3];
            fib(i, n, d);
            float uv[2];
            ref_glsl_rect2cube(d, uv);
            printf("uv %.9g %.9g %.9g : %.9g %.9g\n", d[0], d[1], d[2], uv[0], uv[1]);
        }
        return 0;
    }
    fprintf(stderr, "usage: ref_slices sampling | get_dirs n | brdf w h | prefilter cube.f32 n0 n_out | irradiance cube.f32 n0 n_out | rect2cube n | project csr.bin\n");
    return 2;
}
