// oracle/ref_bake.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the reference's OWN per-vertex bake -- /root/reference/src/raytracing/raytracing.cpp, the whole file, included UNMODIFIED
// from where it lies -- with its external dependencies replaced by stand-ins (oracle/ref_stub/): Embree 3 by the CPU oracle's tracer,
// google/spherical-harmonics by its published polynomials, the GL / platform headers by the members the file reads.  bake_SH(Mesh&)
// then executes the reference's own scene set-up (RTScene), frame / cosine sampling, renderSH path logic and estimator.  Its output
// is the fixture tests/golden/ref_bake_SH_*.txt the oracle is pinned against (tests/test_oracle_pinned.py), fed with the same
// std::mt19937 sequence the reference's random() draws (raytracing.cpp:14-18; PSTL without TBB runs the vertex loop serially).
//
// Two accommodations, neither touching the source: glibc declares `long random(void)`, which collides with the reference's own
// `inline float random()` (MSVC has no such function) -- the name is re-spelled by a macro after the system headers are in; and
// <stdlib.h> / <math.h> are included so that the unqualified abs(N.z) (raytracing.cpp:103,155) is the float overload, as under MSVC.
//
// usage: ref_bake mesh.bin out.txt sh_resolution max_path_length albedo
//   mesh.bin: uint32 n_verts, n_tris; float pos[n_verts][3]; float nrm[n_verts][3]; uint32 tri[n_tris][3]
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <execution>
#include <algorithm>
#include <random>
#include <mutex>
#include <vector>
#include <string>
#include <limits>
#include <tuple>
#define random prt_reference_random
#include "raytracing/raytracing.cpp"
#undef random

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: ref_bake mesh.bin out.txt sh_resolution max_path_length albedo\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint32_t nv = 0, nt = 0;
    if (fread(&nv, 4, 1, f) != 1 || fread(&nt, 4, 1, f) != 1) return 2;
    std::vector<float> pos(3 * (size_t)nv), nrm(3 * (size_t)nv);
    std::vector<Mesh::Index> idx(3 * (size_t)nt);
    if (fread(pos.data(), 12, nv, f) != nv || fread(nrm.data(), 12, nv, f) != nv || fread(idx.data(), 12, nt, f) != nt) return 2;
    fclose(f);
    std::vector<Mesh::Vert> verts(nv);
    for (uint32_t i = 0; i < nv; i++) {
        verts[i].pos = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        verts[i].norm = glm::vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
    }
    Mesh mesh(std::move(verts), std::move(idx));
    App &app = App::get();
    app.sh_resolution = atoi(argv[3]);
    app.max_path_length = atoi(argv[4]);
    app.albedo[0] = app.albedo[1] = app.albedo[2] = (float)atof(argv[5]);
    FILE *keep = stdout;
    (void)keep;
    bake_SH(mesh);                                   // prints "\r baking... i/n" to stdout
    FILE *o = fopen(argv[2], "w");
    if (!o) { perror(argv[2]); return 2; }
    for (const auto &v : mesh.verts()) {
        for (int k = 0; k < 9; k++) fprintf(o, "%.9g%c", v.sh_coeff[k], k == 8 ? '\n' : ' ');
    }
    fclose(o);
    return 0;
}
