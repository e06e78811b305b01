"""Environment inputs for BASELINE config 2: a Radiance .hdr (RGBE) reader equivalent to the reference's
``stbi_loadf`` call (reference src/util/util.cpp:12) and a synthetic HDR environment for machines without the asset."""
from __future__ import annotations

import numpy as np


def load_hdr(path: str) -> np.ndarray:
    """Reads a Radiance RGBE image (-Y h +X w, flat or new-style RLE) -> [h, w, 3] float32, top row first.
    Decoding rule of stb_image's stbi__hdr_convert: value = mantissa * 2^(exponent - 136), zero when exponent == 0."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    header_ok = False
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if line.startswith(b"#?"):
            header_ok = True
        if line == b"":
            break
    if not header_ok:
        raise ValueError("not a Radiance HDR file")
    end = data.index(b"\n", pos)
    res = data[pos:end].split()
    pos = end + 1
    if res[0] != b"-Y" or res[2] != b"+X":
        raise ValueError("unsupported HDR orientation")
    h, w = int(res[1]), int(res[3])
    buf = np.frombuffer(data, np.uint8, offset=pos)
    rgbe = np.zeros((h, w, 4), np.uint8)
    p = 0
    for y in range(h):
        if w < 8 or w >= 32768 or buf[p] != 2 or buf[p + 1] != 2 or (buf[p + 2] & 0x80):
            rgbe[y:] = buf[p:p + (h - y) * w * 4].reshape(h - y, w, 4)   # flat
            break
        p += 4
        for c in range(4):
            x = 0
            while x < w:
                n = int(buf[p]); p += 1
                if n > 128:
                    n -= 128
                    rgbe[y, x:x + n, c] = buf[p]; p += 1
                else:
                    rgbe[y, x:x + n, c] = buf[p:p + n]; p += n
                x += n
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * scale[..., None]).astype(np.float32)


def synthetic_env(w: int = 1600, h: int = 800, seed: int = 3) -> np.ndarray:
    """Deterministic HDR equirect (sky gradient + ground + a few bright lobes up to ~16, like newport_loft's range)."""
    rs = np.random.RandomState(seed)
    v = (np.arange(h, dtype=np.float64) + 0.5) / h * np.pi            # polar angle from +Y
    u = ((np.arange(w, dtype=np.float64) + 0.5) / w - 0.5) * 2 * np.pi
    st, ct = np.sin(v)[:, None], np.cos(v)[:, None]
    d = np.stack([-np.sin(u)[None, :] * st, np.broadcast_to(ct, (h, w)), -np.cos(u)[None, :] * st], -1)
    sky = np.array([0.25, 0.4, 0.8]) * np.clip(d[..., 1:2], 0, 1) ** 0.5 + np.array([0.15, 0.12, 0.1]) * np.clip(-d[..., 1:2], 0, 1)
    img = sky + 0.05
    for _ in range(6):
        c = rs.normal(size=3); c /= np.linalg.norm(c)
        power, sharp = rs.uniform(2, 16), rs.uniform(20, 400)
        col = rs.uniform(0.6, 1.0, 3)
        img = img + power * col * np.exp(sharp * (d @ c - 1.0))[..., None]
    img *= 1.0 + 0.1 * np.sin(7 * u)[None, :, None] * np.cos(5 * v)[:, None, None]
    return img.astype(np.float32)
