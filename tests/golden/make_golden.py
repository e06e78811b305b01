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
  ref_project.txt     precomp_projectSH.comp main() (the live per-probe projection kernel: 128 invocations as host threads, barrier() =
                      std::barrier) on the random CSR of project_case() below: 7 probes with 0 .. 700 entries
and, from the reference's WHOLE files compiled unmodified with Embree replaced by the oracle's tracer (oracle/ref_bake.cpp,
oracle/ref_weight.cpp, stand-in headers under oracle/ref_stub/):
  ref_bake_SH_*.txt       bake_SH(Mesh&) (src/raytracing/raytracing.cpp:320-360 + renderSH :228-278) on bumpy_torus meshes: sh_coeff[9] of
                          every vertex; cases: shadow (32x24 mesh, sh_resolution 16, max_path_length 2), bounce (same mesh, 8, 4, albedo 0.5),
                          default (16x12 mesh, the reference's defaults 32 / 2 / 1.0)
  ref_probe_capture.txt   the CPU half of SH_volume::precompute (src/sh/volume.cpp:185-315: clustering, per-texel accumulation, CSR emission,
                          surfel averaging; oracle/ref_probe.cpp) for the first 3 probes of a 4^3 grid over data/cube.obj + a torus, 64^2 x 6
                          texels per probe, G-buffer = one oracle closest-hit ray per texel centre
  ref_volume_weight.txt   calculate_weight(Model&, 4^3 probes, 12^3 voxels, scene_size 6.18) (src/raytracing/light_probe.cpp:156-367) on
                          data/cube.obj + a torus: weight0123 | weight4567 of every voxel
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


    import numpy as np
    rng, ids, tr, rad = project_case()
    with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
        with open(tf.name, "wb") as f:
            f.write(np.array([len(rng), len(ids), len(rad)], np.uint32).tobytes() + rng.tobytes() + ids.tobytes() + tr.tobytes() + rad.tobytes())
        emit("ref_project.txt", "ref_slices", "project", tf.name)
    ref_bake_and_weight(ref)


def project_case():
    """CSR + surfel radiance fed to the reference's projection kernel; ranges of 0, 1, 127, 128, 129, 700 and 33 entries straddle its
    128-invocation work group"""
    import numpy as np
    rs = np.random.RandomState(4)
    lens = [0, 1, 127, 128, 129, 700, 33]
    n_prim = 57
    rng = np.zeros((len(lens), 2), np.uint32)
    o = 0
    for i, n in enumerate(lens):
        rng[i] = (o, o + n)
        o += n
    ids = rs.randint(0, n_prim, o).astype(np.uint32)
    tr = (rs.randn(o, 9) * 0.01).astype(np.float32)
    rad = (rs.rand(n_prim, 4) * 3).astype(np.float32)
    return rng, ids, tr, rad


BAKE_CASES = {"shadow": (32, 24, 16, 2, 1.0), "bounce": (32, 24, 8, 4, 0.5), "default": (16, 12, 32, 2, 1.0)}   # nu, nv, sh_resolution, max_path_length, albedo
WEIGHT_CASE = (4, 12, 6.18)                                                                                      # probe_res, volume_res, scene_size


def weight_scene():
    import numpy as np
    from prt_b200 import meshes
    pos, _, tri = meshes.load_obj_assimp(os.path.join(here, "cube.obj"))
    tp, _, tt = meshes.bumpy_torus(40, 28)
    return (np.concatenate([pos, tp * np.float32(1.3)]).astype(np.float32),
            np.concatenate([tri, tt + np.uint32(len(pos))]).astype(np.uint32))


def write_mesh(path, pos, nrm, tri):
    import numpy as np
    with open(path, "wb") as f:
        f.write(np.array([len(pos), len(tri)], np.uint32).tobytes())
        f.write(np.ascontiguousarray(pos, np.float32).tobytes())
        f.write(np.ascontiguousarray(nrm, np.float32).tobytes())
        f.write(np.ascontiguousarray(tri, np.uint32).tobytes())


def ref_bake_and_weight(ref):
    import numpy as np
    from prt_b200 import meshes
    with tempfile.TemporaryDirectory() as td:
        for name, (nu, nv, res, mpl, alb) in BAKE_CASES.items():
            pos, nrm, tri = meshes.bumpy_torus(nu, nv)
            write_mesh(os.path.join(td, "m.bin"), pos, nrm, tri)
            out = os.path.join(here, f"ref_bake_SH_{name}.txt")
            subprocess.check_call([os.path.join(ref, "ref_bake"), os.path.join(td, "m.bin"), out, str(res), str(mpl), str(alb)], stdout=subprocess.DEVNULL)
            print(f"ref_bake_SH_{name}.txt: {len(open(out).readlines())} lines")
        pos, tri = weight_scene()
        write_mesh(os.path.join(td, "s.bin"), pos, np.zeros_like(pos), tri)
        out = os.path.join(here, "ref_volume_weight.txt")
        subprocess.check_call([os.path.join(ref, "ref_weight"), os.path.join(td, "s.bin"), out] + [str(x) for x in WEIGHT_CASE])
        print(f"ref_volume_weight.txt: {len(open(out).readlines())} lines")
        out = os.path.join(here, "ref_probe_capture.txt")
        subprocess.check_call([os.path.join(ref, "ref_probe"), os.path.join(td, "s.bin"), out, "4", "6.18", "3"], stdout=subprocess.DEVNULL)
        print(f"ref_probe_capture.txt: {len(open(out).readlines())} lines")


if __name__ == "__main__":
    main()
