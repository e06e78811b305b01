// oracle/ref_probe.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the CPU half of the reference's OWN per-probe capture -- SH_volume::precompute, /root/reference/src/sh/volume.cpp:185-315, cut
// out with sed at build time (oracle/Makefile `ref`) and compiled unmodified as the body of a member function: surfel clustering
// (get_surfel_ID, first-seen numbering), the per-texel loop (sky / back-face skips, cube_coord, solid angle, sh-space direction,
// transfer_weight[ID] += SH9{d} * solid_angle), CSR emission in std::map order, probe ranges, surfel averaging -- with the reference's
// own SH9 / cubeCoordToWorld (src/sh/SH_function.h, included from where it lies).
// What is NOT the reference's: the G-buffer.  The reference rasterises a 64 x 64 x 6 cubemap with OpenGL and reads it back with
// glGetTexImage (volume.cpp:152-183,241-244); here capture_GBuffer() fills the same arrays by one closest-hit ray of the CPU oracle
// through every texel centre (direction = the reference's cubeCoordToWorld), FP32 positions, normalised geometric normals (SURVEY.md
// section 7: the departures of the ray-cast capture).  The GL upload calls of the slice are recorded instead of executed.
// Output: the fixture tests/golden/ref_probe_capture.txt the oracle's prt_o_probe_capture is pinned against (up to the permutation
// of surfel ids: first-seen numbering here, rank of the cluster key there).
//
// usage: ref_probe mesh.bin out.txt probe_res scene_size n_probes_max
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#define FMT_HEADER_ONLY
#include <fmt/core.h>
#include <fmt/format.h>
#include <glm/glm.hpp>
typedef float GLfloat;
typedef unsigned int GLuint;
typedef int GLsizei;
typedef int GLint;
#include "sh/SH_function.h"
#include "prt_oracle.h"

// ---- the GL calls of the slice, recorded -------------------------------------------------------------------------------------
enum { GL_TEXTURE_CUBE_MAP = 1, GL_TEXTURE_CUBE_MAP_POSITIVE_X = 100, GL_RGB = 2, GL_FLOAT = 3, GL_TEXTURE_BUFFER = 4, GL_STATIC_DRAW = 5, GL_DYNAMIC_COPY = 6,
       GL_TEXTURE_3D = 7, GL_RG_INTEGER = 8, GL_UNSIGNED_INT = 9 };
static GLuint g_bound_tex = 0, g_bound_buf = 0;
static std::map<GLuint, std::vector<unsigned char>> g_buffers;
static std::vector<glm::uvec2> g_ranges;
static std::vector<glm::vec3> g_face_pos[6], g_face_norm[6];
enum { TEX_POS = 11, TEX_NORM = 12, TEX_RANGE = 13, BUF_TRANSFER = 21, BUF_ID = 22, BUF_RAD = 23, BUF_PRIM = 24 };
static void glBindTexture(int, GLuint t) { g_bound_tex = t; }
static void glGetTexImage(int face_target, int, int, int, void *out) {
    const std::vector<glm::vec3> &src = (g_bound_tex == TEX_POS ? g_face_pos : g_face_norm)[face_target - GL_TEXTURE_CUBE_MAP_POSITIVE_X];
    std::memcpy(out, src.data(), src.size() * sizeof(glm::vec3));
}
static void glBindBuffer(int, GLuint b) { g_bound_buf = b; }
static void glBufferData(int, size_t size, const void *data, int) {
    std::vector<unsigned char> &b = g_buffers[g_bound_buf];
    b.assign(size, 0);
    if (data) std::memcpy(b.data(), data, size);
}
static void glTexSubImage3D(int, int, int, int, int, int, int, int, int, int, const void *data) { g_ranges.push_back(*(const glm::uvec2 *)data); }

struct Precompute {
    // the members of SH_volume the slice reads (src/sh/volume.h:28-52)
    glm::ivec3 probe_res{1};
    std::vector<glm::vec3> probe_positions;
    int num_primitive = 0;
    static const GLsizei cubemap_res = 64;
    GLuint GBuffer_pos = TEX_POS, GBuffer_norm = TEX_NORM, probe_range = TEX_RANGE;
    GLuint transfer_buffer = BUF_TRANSFER, ID_buffer = BUF_ID, rad_buffer = BUF_RAD, primitive_buffer = BUF_PRIM;
    const prt_o_scene *scene = nullptr;

    // stands in for the G-buffer raster (volume.cpp:152-183): one closest-hit ray per texel centre
    void capture_GBuffer(glm::vec3 pos) {
        for (int face = 0; face < 6; face++) {
            g_face_pos[face].assign(cubemap_res * cubemap_res, glm::vec3(0));
            g_face_norm[face].assign(cubemap_res * cubemap_res, glm::vec3(0));
            for (int y = 0; y < cubemap_res; y++)
                for (int x = 0; x < cubemap_res; x++) {
                    const glm::vec3 d = cubeCoordToWorld(x, y, face);
                    float o[3] = { pos.x, pos.y, pos.z }, dd[3] = { d.x, d.y, d.z }, t, ng[3];
                    uint32_t prim;
                    if (!prt_o_closest_hit(scene, o, dd, 0.0f, INFINITY, 1, &t, &prim, ng)) continue;          // sky: normal stays 0
                    const float inv = 1.0f / sqrtf(fmaf(ng[2], ng[2], fmaf(ng[1], ng[1], ng[0] * ng[0])));
                    g_face_norm[face][y * cubemap_res + x] = glm::vec3(ng[0] * inv, ng[1] * inv, ng[2] * inv);
                    g_face_pos[face][y * cubemap_res + x] = glm::vec3(fmaf(t, d.x, pos.x), fmaf(t, d.y, pos.y), fmaf(t, d.z, pos.z));
                }
        }
    }
    void run() {
#include "_ref/slice_precompute.inc"
    }
};

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: ref_probe mesh.bin out.txt probe_res scene_size n_probes_max\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint32_t nv = 0, nt = 0;
    if (fread(&nv, 4, 1, f) != 1 || fread(&nt, 4, 1, f) != 1) return 2;
    std::vector<float> pos(3 * (size_t)nv), nrm(3 * (size_t)nv);
    std::vector<uint32_t> idx(3 * (size_t)nt);
    if (fread(pos.data(), 12, nv, f) != nv || fread(nrm.data(), 12, nv, f) != nv || fread(idx.data(), 12, nt, f) != nt) return 2;
    fclose(f);
    Precompute P;
    P.scene = prt_o_scene_create(pos.data(), 12, nv, idx.data(), nt);
    const int res = atoi(argv[3]);
    const float size = (float)atof(argv[4]);
    const int n_max = atoi(argv[5]);
    // SH_volume::init, volume.cpp:83-90 (restated: it sits between GL calls), x fastest; the first n_max probes are baked
    const glm::vec3 scene_size(size), ds = 2.f / glm::vec3(res) * scene_size;
    for (int z = 0; z < res; z++) for (int y = 0; y < res; y++) for (int x = 0; x < res; x++)
        if ((int)P.probe_positions.size() < n_max) P.probe_positions.push_back(-scene_size + ds * (glm::vec3(0.5) + glm::vec3(x, y, z)));
    P.probe_res = glm::ivec3((int)P.probe_positions.size(), 1, 1);
    P.run();
    FILE *o = fopen(argv[2], "w");
    if (!o) { perror(argv[2]); return 2; }
    const std::vector<unsigned char> &tr = g_buffers[BUF_TRANSFER], &id = g_buffers[BUF_ID], &pr = g_buffers[BUF_PRIM];
    const size_t nnz = id.size() / 4;
    fprintf(o, "sizes %zu %zu %d\n", P.probe_positions.size(), nnz, P.num_primitive);
    for (size_t p = 0; p < P.probe_positions.size(); p++)
        fprintf(o, "probe %.9g %.9g %.9g : %u %u\n", P.probe_positions[p].x, P.probe_positions[p].y, P.probe_positions[p].z, g_ranges[p].x, g_ranges[p].y);
    for (size_t i = 0; i < nnz; i++) {
        fprintf(o, "entry %u :", ((const GLuint *)id.data())[i]);
        for (int k = 0; k < 9; k++) fprintf(o, " %.9g", ((const float *)tr.data())[9 * i + k]);
        fprintf(o, "\n");
    }
    for (int s = 0; s < P.num_primitive; s++) {
        const float *q = (const float *)pr.data() + 6 * s;
        fprintf(o, "surfel %d : %.9g %.9g %.9g %.9g %.9g %.9g\n", s, q[0], q[1], q[2], q[3], q[4], q[5]);
    }
    fclose(o);
    return 0;
}
