// oracle/ref_weight.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the reference's OWN calculate_weight -- /root/reference/src/raytracing/light_probe.cpp, the whole file, included UNMODIFIED --
// with Embree 3 replaced by the CPU oracle's tracer (oracle/ref_stub/embree3/rtcore.h) and Model reduced to its `meshes` member.
// Output: the fixture tests/golden/ref_volume_weight.txt (weight0123 / weight4567 per voxel) the oracle's prt_o_volume_weights is
// pinned against.
// usage: ref_weight mesh.bin out.txt probe_res volume_res scene_size
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <vector>
#include <limits>
#include "raytracing/light_probe.cpp"

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: ref_weight mesh.bin out.txt probe_res volume_res scene_size\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint32_t nv = 0, nt = 0;
    if (fread(&nv, 4, 1, f) != 1 || fread(&nt, 4, 1, f) != 1) return 2;
    std::vector<float> pos(3 * (size_t)nv), nrm(3 * (size_t)nv);
    std::vector<Mesh::Index> idx(3 * (size_t)nt);
    if (fread(pos.data(), 12, nv, f) != nv || fread(nrm.data(), 12, nv, f) != nv || fread(idx.data(), 12, nt, f) != nt) return 2;
    fclose(f);
    std::vector<Mesh::Vert> verts(nv);
    for (uint32_t i = 0; i < nv; i++) verts[i].pos = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    Model model;
    model.meshes.emplace_back(std::move(verts), std::move(idx));
    const int pr = atoi(argv[3]), vr = atoi(argv[4]);
    const float sz = (float)atof(argv[5]);
    Volume_weight w = calculate_weight(model, glm::ivec3(pr), glm::ivec3(vr), glm::vec3(sz));
    FILE *o = fopen(argv[2], "w");
    if (!o) { perror(argv[2]); return 2; }
    for (size_t i = 0; i < w.weight0123.size(); i++)
        fprintf(o, "%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", w.weight0123[i].x, w.weight0123[i].y, w.weight0123[i].z, w.weight0123[i].w,
                w.weight4567[i].x, w.weight4567[i].y, w.weight4567[i].z, w.weight4567[i].w);
    fclose(o);
    return 0;
}
