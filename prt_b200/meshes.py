"""Synthetic meshes and a minimal OBJ reader for tests and bench.

The reference's named asset data/buddha.obj is a missing blob (reference .MISSING_LARGE_BLOBS:1-6), so the
bench generates meshes of the named sizes instead:

* ``icosphere(k)``      -- convex; V = 10*4^k + 2, F = 20*4^k  (reference data/sphere.obj is k = 4)
* ``bumpy_torus(nu,nv)``-- self-occluding; V = nu*nv, F = 2*nu*nv; 737x737 is "buddha-scale"
  (543 169 V / 1 086 338 F; Stanford Happy Buddha: 543 652 V / 1 087 716 F)

A vertex is (pos, normal) as in the reference's Mesh::Vert (src/opengl/gl.h:76-80).  The synthetic meshes carry smooth
area-weighted vertex normals (an OBJ with normals passes through assimp unchanged); ``load_obj_assimp`` reproduces what the
reference's import flags (src/scene/model.cpp:72) make of an OBJ without normals: flat normals, vertices split per face.
"""
from __future__ import annotations

import numpy as np


def vertex_normals(pos: np.ndarray, tri: np.ndarray) -> np.ndarray:
    """Area-weighted smooth vertex normals (float32)."""
    p = pos.astype(np.float64)
    fn = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    n = np.zeros_like(p)
    for a in range(3):
        for k in range(3):
            n[:, a] += np.bincount(tri[:, k], weights=fn[:, a], minlength=len(p))
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    return (n / ln).astype(np.float32)


def icosphere(k: int):
    """Unit icosphere with k subdivisions; normals = positions."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(k):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * (len(v) + 1) + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // (len(v) + 1), uniq % (len(v) + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        nf = len(f)
        m01, m12, m20 = base + inv[:nf], base + inv[nf:2 * nf], base + inv[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    pos = v.astype(np.float32)
    nrm = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    return pos, nrm.astype(np.float32), f.astype(np.uint32)


def bumpy_torus(nu: int, nv: int, seed: int = 7, R: float = 2.0, r: float = 0.8, amp: float = 0.15, fscale: int = 1):
    """Torus (major R, minor r) with a radial displacement amp*sum_j a_j sin(f_j u+phi_j) sin(g_j v+psi_j).

    Vertex (i, j) -> index i*nv + j, u = 2 pi i/nu (around the axis), v = 2 pi j/nv (around the tube).
    Parameters come from numpy's MT19937 (RandomState(seed)); outward-facing CCW triangles.  ``fscale`` multiplies the bump
    frequencies: deeper, narrower folds at the same amplitude (the "heavily self-occluding" bench workload uses amp 0.25, fscale 3: 45 % of the rays occluded).
    """
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.3, 1.0, 4)
    fu = rs.randint(1, 7, 4)
    gv = rs.randint(1, 9, 4)
    phi = rs.uniform(0, 2 * np.pi, 4)
    psi = rs.uniform(0, 2 * np.pi, 4)
    u = (np.arange(nu, dtype=np.float64) * (2 * np.pi / nu))[:, None]
    v = (np.arange(nv, dtype=np.float64) * (2 * np.pi / nv))[None, :]
    d = np.zeros((nu, nv))
    for j in range(4):
        d += a[j] * np.sin(fscale * fu[j] * u + phi[j]) * np.sin(fscale * gv[j] * v + psi[j])
    rr = r + amp * d
    x = (R + rr * np.cos(v)) * np.cos(u)
    y = rr * np.sin(v) + 0.0 * u
    z = (R + rr * np.cos(v)) * np.sin(u)
    pos = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = np.arange(nu, dtype=np.int64)[:, None]
    j = np.arange(nv, dtype=np.int64)[None, :]
    i1, j1 = (i + 1) % nu, (j + 1) % nv
    v00, v10, v01, v11 = i * nv + j, i1 * nv + j, i * nv + j1, i1 * nv + j1
    t0 = np.stack([v00, v01, v10], -1).reshape(-1, 3)
    t1 = np.stack([v10, v01, v11], -1).reshape(-1, 3)
    tri = np.concatenate([t0, t1], axis=0).astype(np.uint32)
    nrm = vertex_normals(pos, tri.astype(np.int64))
    return pos, nrm, tri


def load_obj(path: str):
    """Minimal OBJ reader: v / vn / f (polygons fan-triangulated).  If the file has no normals, smooth
    vertex normals are generated; with normals, (pos, normal) pairs are de-duplicated like assimp's
    JoinIdenticalVertices (reference src/scene/model.cpp:72)."""
    vs, vns, faces = [], [], []
    with open(path) as fh:
        for line in fh:
            s = line.split()
            if not s:
                continue
            if s[0] == "v":
                vs.append([float(s[1]), float(s[2]), float(s[3])])
            elif s[0] == "vn":
                vns.append([float(s[1]), float(s[2]), float(s[3])])
            elif s[0] == "f":
                corners = []
                for c in s[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    corners.append((vi - 1 if vi > 0 else len(vs) + vi, ni - 1 if ni > 0 else (len(vns) + ni if ni < 0 else -1)))
                for k in range(1, len(corners) - 1):
                    faces.append([corners[0], corners[k], corners[k + 1]])
    vs = np.asarray(vs, dtype=np.float32)
    if not vns:
        tri = np.asarray([[c[0] for c in f] for f in faces], dtype=np.uint32)
        return vs, vertex_normals(vs, tri.astype(np.int64)), tri
    vns = np.asarray(vns, dtype=np.float32)
    lut, pos, nrm, tri = {}, [], [], []
    for f in faces:
        t = []
        for c in f:
            if c not in lut:
                lut[c] = len(pos)
                pos.append(vs[c[0]])
                nrm.append(vns[c[1]] if c[1] >= 0 else np.zeros(3, np.float32))
            t.append(lut[c])
        tri.append(t)
    return np.asarray(pos, np.float32), np.asarray(nrm, np.float32), np.asarray(tri, np.uint32)


def load_obj_assimp(path: str):
    """What a vertex IS in the reference: ``Model(path)`` imports with ``aiProcess_Triangulate | aiProcess_GenNormals |
    aiProcess_JoinIdenticalVertices`` (reference src/scene/model.cpp:72) and ``processMesh`` copies ``mVertices`` / ``mNormals`` into
    ``Mesh::Vert`` (model.cpp:14-47).  assimp's OBJ importer emits one vertex per face corner; for a file WITHOUT normals (the
    reference's data/cube.obj and data/sphere.obj) GenNormals gives every corner the FLAT normal of its face, normalize((v1-v0)x(v2-v0))
    in float, and JoinIdenticalVertices then merges corners whose position, normal and texture coordinate are all identical, keeping
    first occurrences in order -- so coplanar neighbours share vertices and everything else is split per face.  A file WITH normals
    keeps them and only the join applies.  Returns (pos [V,3] f32, nrm [V,3] f32, tri [F,3] u32); polygons are fan-triangulated
    (assimp's ear clipping differs only for non-convex polygons; the reference's assets are triangles)."""
    vs, vns, vts, faces = [], [], [], []
    with open(path) as fh:
        for line in fh:
            s = line.split()
            if not s:
                continue
            if s[0] == "v":
                vs.append([float(x) for x in s[1:4]])
            elif s[0] == "vn":
                vns.append([float(x) for x in s[1:4]])
            elif s[0] == "vt":
                vts.append([float(x) for x in (s[1:3] + ["0"])[:2]])
            elif s[0] == "f":
                c = []
                for tok in s[1:]:
                    q = (tok.split("/") + ["", ""])[:3]
                    idx = [int(x) if x else 0 for x in q]
                    c.append(tuple(i - 1 if i > 0 else (n + i if i < 0 else -1) for i, n in zip(idx, (len(vs), len(vts), len(vns)))))
                for k in range(1, len(c) - 1):
                    faces.append((c[0], c[k], c[k + 1]))
    vs = np.asarray(vs, np.float32)
    vns = np.asarray(vns, np.float32).reshape(-1, 3)
    vts = np.asarray(vts, np.float32).reshape(-1, 2)
    fv = np.asarray([[c[0] for c in f] for f in faces], np.int64)
    corner_pos = vs[fv]                                                           # [F,3,3]
    if len(vns):
        fn = np.asarray([[c[2] for c in f] for f in faces], np.int64)
        corner_nrm = np.where((fn >= 0)[..., None], vns[np.maximum(fn, 0)], np.float32(0))
    else:
        e1, e2 = corner_pos[:, 1] - corner_pos[:, 0], corner_pos[:, 2] - corner_pos[:, 0]
        n = np.cross(e1, e2).astype(np.float32)
        ln = np.sqrt((n * n).sum(1, keepdims=True, dtype=np.float32))
        n = np.where(ln > 0, n / np.where(ln > 0, ln, 1), n).astype(np.float32)   # NormalizeSafe
        corner_nrm = np.repeat(n[:, None, :], 3, axis=1)
    if len(vts):
        ft = np.asarray([[c[1] for c in f] for f in faces], np.int64)
        corner_uv = np.where((ft >= 0)[..., None], vts[np.maximum(ft, 0)], np.float32(0))
    else:
        corner_uv = np.zeros(corner_pos.shape[:2] + (2,), np.float32)
    rec = np.concatenate([corner_pos, corner_nrm, corner_uv], -1).reshape(-1, 8).astype(np.float32) + np.float32(0)   # -0 -> +0
    # (assimp joins with an epsilon of 1e-5 on every component; exact equality is used here -- it differs only for corners whose
    # face normals agree to 1e-5 without being equal)
    _, first, inv = np.unique(rec.view(np.uint32), axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                                      # first occurrences, in file order
    rank = np.empty(len(order), np.int64)
    rank[order] = np.arange(len(order))
    keep = first[order]
    return rec[keep, 0:3].copy(), rec[keep, 3:6].copy(), rank[inv.reshape(-1)].reshape(-1, 3).astype(np.uint32)


def morton_order(pos: np.ndarray) -> np.ndarray:
    """Permutation sorting vertices along a 30-bit Morton curve (locality for the traversal kernel)."""
    p = pos.astype(np.float64)
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-30) * 1023.0), 0, 1023).astype(np.uint64)

    def spread(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x

    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return np.argsort(code, kind="stable")
