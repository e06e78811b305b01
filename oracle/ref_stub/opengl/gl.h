// oracle/ref_stub/opengl/gl.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's src/opengl/gl.h (which needs glad and a GL
// context) exposing what raytracing.cpp / light_probe.cpp / util.h use of it: the GL scalar typedefs, class Mesh with the reference's
// Vert layout and accessors (gl.h:72-110, gl.cpp:270-300), a Tex2D declaration, glm and fmt.
#pragma once
#include <string>
#include <vector>
#include <cassert>
#ifndef FMT_HEADER_ONLY
#define FMT_HEADER_ONLY
#endif
#include <fmt/core.h>
#include <fmt/format.h>
#include <glm/glm.hpp>
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>
#include <glm/mat4x4.hpp>

typedef unsigned int GLuint;
typedef int GLint;
typedef float GLfloat;

class Tex2D;
class Shader;

class Mesh {
public:
    typedef GLuint Index;
    struct Vert {
        glm::vec3 pos;
        glm::vec3 norm;
        GLfloat sh_coeff[9];
    };
    glm::mat4 Mat_model = glm::mat4(1);
    Mesh() {}
    Mesh(std::vector<Vert> &&vertices, std::vector<Index> &&indices) : _verts(std::move(vertices)), _idxs(std::move(indices)) {}
    std::vector<Vert> &edit_verts() { dirty = true; return _verts; }
    std::vector<Index> &edit_indices() { dirty = true; return _idxs; }
    const std::vector<Vert> &verts() const { return _verts; }
    const std::vector<Index> &indices() const { return _idxs; }
    bool dirty = true;
private:
    std::vector<Vert> _verts;
    std::vector<Index> _idxs;
};
static_assert(sizeof(Mesh::Vert) == 60, "Mesh::Vert is 60 bytes in the reference (gl.h:76-80)");
