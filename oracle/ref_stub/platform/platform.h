// oracle/ref_stub/platform/platform.h -- TEST INFRASTRUCTURE ONLY.  The members of the reference's App singleton / Platform / Camera
// that raytracing.cpp reads (app.h:31-71, platform.h:27-40, util/camera.h), with the reference's default values; no GLFW / ImGui.
#pragma once
#include "opengl/gl.h"
struct Camera {
    glm::vec3 Position{0.f, 0.f, 3.f}, Front{0.f, 0.f, -1.f}, Up{0.f, 1.f, 0.f}, Right{1.f, 0.f, 0.f};
    float Zoom = 45.f;
    bool dirty = true;
};
class Platform {
public:
    int SCR_WIDTH = 64, SCR_HEIGHT = 48;
};
class App {
public:
    struct Pixel { unsigned char r, g, b, a; };
    struct Pixel_accum { glm::vec3 color; float count; };
    static App &get();
    Platform plt_storage;
    Platform &plt = plt_storage;
    std::vector<Pixel> pixels;
    std::vector<Pixel_accum> pixels_w;
    Camera camera;
    bool gamma = true;
    float albedo[3] = {1.f, 1.f, 1.f};     // app.h:55
    int max_path_length = 2;               // app.h:70
    int sh_resolution = 32;                // app.h:71
};
