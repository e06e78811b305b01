#!/usr/bin/env python
"""tools/tsan_emulated.py -- data-race check of the product's warp-cooperative code WITHOUT a GPU: the horizon builder, the traversal
pass and the interreflection pass run on the warp emulator (one thread per lane, tests/hostcheck) under ThreadSanitizer, which sees
every shared-memory access that is not ordered by a warp collective -- the CPU counterpart of compute-sanitizer's racecheck.
Run through tools/tsan_emulated.sh (builds the instrumented library and preloads libtsan)."""
import sys, os, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import conftest
# load the TSan build instead of the regular one
L = C.CDLL(os.environ.get('PRT_HOSTCHECK_TSAN', '/tmp/libhostcheck_tsan.so'))
ref = conftest.load_hostcheck()
for name in ('hc_build','hc_free','hc_horizon_maps','hc_bake_wave','hc_bake_inter','hc_use_slabs','hc_horizon_mid','hc_wave_dop'):
    getattr(L,name).argtypes = getattr(ref,name).argtypes; getattr(L,name).restype = getattr(ref,name).restype
from oracle import pyoracle as oracle
from prt_b200 import meshes
from test_horizon_math import _maps
from test_wave_emulated import processing_table, run_wave, run_inter
pos, nrm, tri = meshes.bumpy_torus(96, 64)
h = L.hc_build(pos.ctypes.data, 12, len(pos), tri.ctypes.data, len(tri))
sel = np.arange(5, len(pos), 211)[:10]
op = oracle.make_params(order=3, samples_u=16, samples_v=16)
tab, bins = processing_table(oracle, op)
hz, _ = _maps(L, h, pos[sel], nrm[sel])
need = ~(tab[None, :, 2] > hz[:, bins])
nw = np.ascontiguousarray(np.packbits(need, axis=1, bitorder="little")).view(np.uint32).copy()
got, vis = run_wave(L, h, pos[sel], nrm[sel], tab, 3, need=nw)
got2, vis2 = run_wave(L, h, pos[sel], nrm[sel], tab, 3)
opi = oracle.make_params(mode=oracle.INTERREFLECT, order=4, samples_u=16, samples_v=16, bounces=3, albedo=(0.5,0.5,0.5))
tabi, _ = processing_table(oracle, opi)
gi, vi, _ = run_inter(L, h, pos[sel], nrm[sel], tabi, 4, nw, 3, (0.5,0.5,0.5), opi.seed)
print("tsan run finished", float(got.sum()), float(gi.sum()))
L.hc_free(h)
