"""Seeded synthetic inputs shared by the tests, the golden generator and bench.py.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- pure numpy, deterministic across machines.
"""
import numpy as np

JOINTS = 21


def frame_c2(i, size=368):
    """C2 / C5 frame i: uniform-noise BGR image."""
    return np.random.default_rng(1000 + i).integers(0, 256, (size, size, 3), dtype=np.uint8)


def stream_frame(stream, k, size=368):
    """C4: frame k of synthetic stream `stream`: a fixed noise image circularly shifted by k pixels (moving argmax)."""
    base = np.random.default_rng(2000 + stream).integers(0, 256, (size, size, 3), dtype=np.uint8)
    return np.ascontiguousarray(np.roll(base, shift=(k, 2 * k), axis=(0, 1)))


def synthetic_maps(seed, n_scales, hs=46, border_joints=True):
    """Heat-maps with one dominant blob per joint (+ noise) and smooth location maps, NHWC float32 [n,hs,hs,21] x 4.

    With ``border_joints`` the first joints get their peak on the map border / corners, which is where the reference's
    x8 upsample produces exact ties.  Scale i>0 maps are a shrunken copy of the scale-0 content
    plus independent noise, like a real pyramid.
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:hs, 0:hs].astype(np.float64)
    centers = rng.uniform(3, hs - 4, size=(JOINTS, 2))
    if border_joints:
        forced = [(0, 0), (0, hs - 1), (hs - 1, 0), (hs - 1, hs - 1), (0, 17), (hs - 1, 30), (21, 0), (9, hs - 1)]
        for j, c in enumerate(forced):
            centers[j] = c
    sig = rng.uniform(1.0, 2.5, size=JOINTS)
    maps = [np.zeros((n_scales, hs, hs, JOINTS), np.float32) for _ in range(4)]
    scales = [1.0, 0.7, 0.85, 0.6][:n_scales]
    for i, s in enumerate(scales):
        for j in range(JOINTS):
            cy = (centers[j, 0] - hs / 2) * s + hs / 2
            cx = (centers[j, 1] - hs / 2) * s + hs / 2
            blob = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * (sig[j] * s) ** 2))
            maps[0][i, :, :, j] = (blob + 0.02 * rng.standard_normal((hs, hs))).astype(np.float32)
        for m in maps[1:]:
            nc = -(-hs // 8)  # 6 for the 46 x 46 maps the committed fixtures were minted with
            coarse = rng.uniform(-5, 5, size=(nc, nc, JOINTS))
            fine = np.kron(coarse, np.ones((8, 8, 1)))[:hs, :hs, :]
            m[i] = (fine + 0.05 * rng.standard_normal((hs, hs, JOINTS))).astype(np.float32)
    return tuple(maps)
