"""GPU parity tests for the preview tracer (raytrace / renderAO / renderNormal) against the oracle."""
import numpy as np
import pytest

from prt_b200 import meshes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["ao", "normal"])
def test_raytrace_matches_oracle(prt, oracle, mode):
    tp, _, tt = meshes.bumpy_torus(64, 40)                   # open scene: torus resting over a floor quad
    fp = np.array([[-8, -1.2, -8], [8, -1.2, -8], [8, -1.2, 8], [-8, -1.2, 8]], np.float32)
    ft = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)
    rot = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], np.float32)              # torus axis -> y
    pos = np.concatenate([tp @ rot.T, fp]).astype(np.float32)
    tri = np.concatenate([tt, ft + np.uint32(len(tp))]).astype(np.uint32)
    gs, os_ = prt.RTScene(pos, tri), oracle.Scene(pos, tri)
    cam = prt.Camera.look_at((4.5, 3.0, 4.0), (0, -0.3, 0), zoom_deg=50)
    w, h = 83, 47                                            # not multiples of the 8x4 tile
    m = prt.AO if mode == "ao" else prt.NORMAL
    film = prt.Film(w, h, gs.ctx)
    o_acc = None
    for frame in range(3):
        prt.raytrace(gs, film, cam, max_path_length=4, albedo=(0.7, 0.6, 0.5), mode=m, n_frames=1)
        o_acc, o_px = oracle.raytrace(os_, cam, w, h, accum=o_acc, max_path_length=4, albedo=(0.7, 0.6, 0.5), mode=m, frame=frame)
        g_acc, g_px = film.download()
        assert np.array_equal(g_acc, o_acc)                                       # paths are bit-identical (pinned arithmetic, Philox)
        assert np.abs(g_px.astype(int) - o_px.astype(int)).max() <= 1               # powf differs in the last ulp
    if mode == "ao":
        lum = g_acc[..., 0] / g_acc[..., 3]
        assert 0.0 < lum.mean() < 1.0 and len(np.unique(lum)) > 20                # a real image: sky, floor, torus, contact shadows
    film.reset()
    prt.raytrace(gs, film, cam, max_path_length=4, albedo=(0.7, 0.6, 0.5), mode=m, n_frames=3)      # 3 frames in one call == 3 calls
    assert np.array_equal(film.download()[0], g_acc)


def test_raytrace_convex_white_is_one(prt):
    pos, nrm, tri = meshes.icosphere(4)
    gs = prt.RTScene(pos, tri)
    film = prt.Film(64, 64, gs.ctx)
    prt.raytrace(gs, film, prt.Camera.look_at((0, 0, 4), (0, 0, 0)), max_path_length=3, gamma=False, n_frames=2)
    acc, px = film.download()
    assert np.allclose(acc[..., :3], 2.0) and (acc[..., 3] == 2).all() and (px == 255).all()
    with pytest.raises(prt.PRTError):
        prt.raytrace(gs, film, prt.Camera.look_at((0, 0, 4), (0, 0, 0)), mode=7)
