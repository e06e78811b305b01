"""prt_b200 -- B200-native PRT precomputation hot path (host-side mirror of the reference interface).

The product is the C-ABI library ``prt_b200/csrc/libprt_b200.so`` (``include/prt_b200.h``); this package is
the thin Python binding used by the tests and by ``bench.py``.  It mirrors the reference's entry points:

* ``RTScene``            -- reference ``RTScene(Mesh&)`` / ``RTScene(Model&)`` (src/raytracing/raytracing.cpp:58-99,
  src/raytracing/light_probe.cpp:44-93) with ``any_hit`` / ``first_hit`` (light_probe.cpp:95-133)
* ``bake_SH``            -- reference ``bake_SH(Mesh&)`` (src/raytracing/raytracing.cpp:320-360)
* ``ProbeTransfer`` / ``calculate_weight`` / ``SHVolume`` -- reference ``SH_volume::precompute / set_visibility / relight /
  project_sh`` (src/sh/volume.cpp:149-452)
* ``LightProbe`` / ``brdf_lut`` -- reference ``LightProbe`` passes (src/opengl/gl.cpp:546-591)

There is no CPU fallback: importing works anywhere, but creating a context without the built CUDA library or
without a B200-class GPU raises.
"""
from .api import (DeviceBuffer, Group, GroupStats, GATHER_NONE, GATHER_NCCL, GATHER_P2P, GATHER_AUTO, mesh_hash, hash_arrays, cache_save_transfer, cache_load_transfer, cache_save_csr, cache_load_csr, Film, Camera, raytrace, AO, NORMAL, SHVolume, RelightParams, paral_shadow_matrix, shadow_map, calculate_weight, ProbeTransfer, probe_positions, fibonacci_dirs, cube_dirs, LightProbe, brdf_lut, sh_pack_rh, BakeParams, Context, PRTError, RTScene, bake_SH, bake_transfer, lib_path, load_library,  # noqa: F401
                  SHADOWED, UNSHADOWED, INTERREFLECT, UNSHADOWED_ANALYTIC)

__all__ = ["DeviceBuffer", "Group", "GroupStats", "GATHER_NONE", "GATHER_NCCL", "GATHER_P2P", "GATHER_AUTO", "mesh_hash", "hash_arrays", "cache_save_transfer", "cache_load_transfer", "cache_save_csr", "cache_load_csr", "Film", "Camera", "raytrace", "AO", "NORMAL", "SHVolume", "RelightParams", "paral_shadow_matrix", "shadow_map", "calculate_weight", "ProbeTransfer", "probe_positions", "fibonacci_dirs", "cube_dirs", "LightProbe", "brdf_lut", "sh_pack_rh", "BakeParams", "Context", "PRTError", "RTScene", "bake_SH", "bake_transfer", "lib_path", "load_library",
           "SHADOWED", "UNSHADOWED", "INTERREFLECT", "UNSHADOWED_ANALYTIC"]
