// oracle/ref_stub/sh/spherical_harmonics.h -- TEST INFRASTRUCTURE ONLY.
// raytracing.cpp:10 includes "sh/spherical_harmonics.h" -- google/spherical-harmonics (+ Eigen), which is in neither the reference
// tree nor this image (parity unpinned for this dependency).  This stand-in restates the two calls the reference
// makes (raytracing.cpp:226,334,348) from the library's published hard-coded polynomials for l <= 4, in double, with ITS sign
// convention: the Condon-Shortley phase (-1)^m is included (e.g. Y_1^-1 = -0.488603 y), unlike the reference's own SH_function.h.
#pragma once
namespace Eigen {
struct Vector3d {
    double v[3];
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
};
}  // namespace Eigen
namespace sh {
inline int GetIndex(int l, int m) { return l * (l + 1) + m; }
inline double EvalSH(int l, int m, const Eigen::Vector3d &d) {
    const double x = d.x(), y = d.y(), z = d.z(), x2 = x * x, y2 = y * y, z2 = z * z;
    switch (GetIndex(l, m)) {
    case 0: return 0.282095;
    case 1: return -0.488603 * y;
    case 2: return 0.488603 * z;
    case 3: return -0.488603 * x;
    case 4: return 1.092548 * x * y;
    case 5: return -1.092548 * y * z;
    case 6: return 0.315392 * (-x2 - y2 + 2.0 * z2);
    case 7: return -1.092548 * x * z;
    case 8: return 0.546274 * (x2 - y2);
    case 9: return -0.590044 * y * (3.0 * x2 - y2);
    case 10: return 2.890611 * x * y * z;
    case 11: return -0.457046 * y * (4.0 * z2 - x2 - y2);
    case 12: return 0.373176 * z * (2.0 * z2 - 3.0 * x2 - 3.0 * y2);
    case 13: return -0.457046 * x * (4.0 * z2 - x2 - y2);
    case 14: return 1.445306 * z * (x2 - y2);
    case 15: return -0.590044 * x * (x2 - 3.0 * y2);
    default: return 0.0;      // the reference bakes bands 0..2 (file-scope `int order = 2`, raytracing.cpp:320)
    }
}
}  // namespace sh
