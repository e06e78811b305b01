// oracle/ref_shfun.cpp -- TEST INFRASTRUCTURE ONLY.
// Driver that compiles the reference's OWN header src/sh/SH_function.h in place (nothing is copied)
// and prints SH9 / cubeCoordToWorld values; tests compare oracle/sh.h and the CUDA basis against it.
// Usage: ref_shfun sh9 x y z   |   ref_shfun cube x y face   |   ref_shfun golden  (fixed vector set)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
typedef float GLfloat;
#include <glm/glm.hpp>
#include "sh/SH_function.h"

static void print_sh9(float x, float y, float z) {
    SH9 s{ vec3(x, y, z) };
    for (int i = 0; i < 9; i++) printf("%.9g%c", s.data[i], i == 8 ? '\n' : ' ');
}
int main(int argc, char **argv) {
    if (argc >= 5 && !strcmp(argv[1], "sh9")) { print_sh9(atof(argv[2]), atof(argv[3]), atof(argv[4])); return 0; }
    if (argc >= 5 && !strcmp(argv[1], "cube")) {
        vec3 w = cubeCoordToWorld(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
        printf("%.9g %.9g %.9g\n", w.x, w.y, w.z); return 0;
    }
    if (argc >= 2 && !strcmp(argv[1], "golden")) {
        // 64 Fibonacci directions, then 6 faces x 4 texels of cubeCoordToWorld
        const int n = 64;
        for (int i = 0; i < n; i++) {
            double z = 1 - (double(i) + 0.5) / n * 2, th = 2.39996322972865332 * i, r = std::sqrt(1 - z * z);
            float x = float(std::cos(th) * r), y = float(std::sin(th) * r), zz = float(z);
            printf("sh9 %.9g %.9g %.9g :", x, y, zz); printf(" "); print_sh9(x, y, zz);
        }
        int tx[4] = { 0, 17, 40, 63 }, ty[4] = { 0, 50, 3, 63 };
        for (int f = 0; f < 6; f++) for (int k = 0; k < 4; k++) {
            vec3 w = cubeCoordToWorld(tx[k], ty[k], f);
            printf("cube %d %d %d : %.9g %.9g %.9g\n", tx[k], ty[k], f, w.x, w.y, w.z);
        }
        return 0;
    }
    fprintf(stderr, "usage: ref_shfun sh9 x y z | cube x y face | golden\n");
    return 2;
}
