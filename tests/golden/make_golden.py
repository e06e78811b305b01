"""Regenerates the fixtures that pin the oracle to the reference's OWN source, compiled in place by `make -C oracle ref`
(oracle/_ref/ref_shfun: src/sh/SH_function.h; oracle/_ref/ref_slices: the line ranges listed in oracle/ref_slices.cpp).
Run in the build container (needs /root/reference); the fixtures are committed, the binaries are not.

  ref_shfun.txt       SH9 / cubeCoordToWorld                      (src/sh/SH_function.h)
  ref_sampling.txt    frame(N), cosineSampleHemisphere(u, v, N)   (src/raytracing/raytracing.cpp:101-107,130-160)
  ref_get_dirs.txt    get_dirs(100), get_dirs(4096)               (src/raytracing/light_probe.cpp:136-152)
  ref_brdf.txt        IntegrateBRDF on a 32x32 LUT                (src/shaders/brdf.frag:5-107)
  ref_prefilter.txt   prefilter.frag main() at texel centres of a 16^2 cube, 5 mips (roughness mip/4), environment = ENV_CUBE below
  ref_irradiance.txt  irradiance.frag main() at texel centres of an 8^2 cube, same environment
  ref_rect2cube.txt   SampleSphericalMap at 64 directions          (src/shaders/rectangle2cube.frag:7-15)
"""
import os
import subprocess
import sys
import tempfile

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(os.path.dirname(here))
sys.path.insert(0, root)

PREF_OUT, IRR_OUT = 16, 8                  # sizes of the rendered prefilter (5 mips) / irradiance cubes the fixtures sample texels of
ENV_W, ENV_H, ENV_CUBE = 256, 128, 64      # synthetic_env(ENV_W, ENV_H) -> oracle cube of ENV_CUBE^2 faces (tests rebuild the same cube)


def env_cube():
    from oracle import pyoracle
    from prt_b200 import hdr
    return pyoracle.EnvCube(hdr.synthetic_env(ENV_W, ENV_H), ENV_CUBE)


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle"), "libprt_oracle.so", "ref"])
    ref = os.path.join(root, "oracle", "_ref")

    def emit(name, exe, *args):
        out = subprocess.check_output([os.path.join(ref, exe), *[str(a) for a in args]], text=True)
        with open(os.path.join(here, name), "w") as f:
            f.write(out)
        print(f"{name}: {len(out.splitlines())} lines")

    emit("ref_shfun.txt", "ref_shfun", "golden")
    emit("ref_sampling.txt", "ref_slices", "sampling")
    out = "".join(subprocess.check_output([os.path.join(ref, "ref_slices"), "get_dirs", str(n)], text=True) for n in (100, 4096))
    with open(os.path.join(here, "ref_get_dirs.txt"), "w") as f:
        f.write(out)
    emit("ref_brdf.txt", "ref_slices", "brdf", 32, 32)
    emit("ref_rect2cube.txt", "ref_slices", "rect2cube", 64)
    cube = env_cube()
    with tempfile.NamedTemporaryFile(suffix=".f32") as tf:
        cube.data.tofile(tf.name)
        emit("ref_prefilter.txt", "ref_slices", "prefilter", tf.name, ENV_CUBE, PREF_OUT)
        emit("ref_irradiance.txt", "ref_slices", "irradiance", tf.name, ENV_CUBE, IRR_OUT)


if __name__ == "__main__":
    main()
