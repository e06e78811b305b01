// oracle/ref_stub/scene/model.h -- TEST INFRASTRUCTURE ONLY.  The reference's Model (src/scene/model.h) without assimp: just the
// `meshes` member RTScene(Model&) and calculate_weight read.
#pragma once
#include "opengl/gl.h"
class Model {
public:
    std::vector<Mesh> meshes;
    glm::mat4 Mat_model = glm::mat4(1);
};
