"""Synthetic PIV frame pairs with a known Poiseuille displacement: the benchmark / smoke-test input
generator.  Host-side numpy, not part of the compute path."""
import numpy as np


def poiseuille_truth(H, W):
    """u(y) = -4 (1 - ((y - (H-1)/2) / (H/2))^2), v = 0 -- the profile of the reference's bundled test pair."""
    y = np.arange(H, dtype=np.float64)
    u = -4.0 * (1.0 - ((y - (H - 1) / 2.0) / (H / 2.0)) ** 2)
    return np.repeat(u[:, None], W, axis=1).astype(np.float32), np.zeros((H, W), dtype=np.float32)


def synthetic_piv_pair(H, W, seed=0):
    """Two 8-bit particle images as float32 (H, W): 6 particles per 256 px, Gaussian blobs (sigma 0.75 px) with
    light-sheet intensity 255 exp(-z^2/2), rendered at x -/+ u(y)/2 for frame 0 / 1."""
    rng = np.random.default_rng(seed)
    n = (H * W * 6) // 256
    px = rng.uniform(-8.0, W + 8.0, n)
    py = rng.uniform(-8.0, H + 8.0, n)
    peak = 255.0 * np.exp(-0.5 * rng.standard_normal(n) ** 2)
    uy = -4.0 * (1.0 - ((py - (H - 1) / 2.0) / (H / 2.0)) ** 2)
    frames = []
    offs = np.arange(-4, 5)
    for sgn in (-0.5, 0.5):
        img = np.zeros((H, W), dtype=np.float64)
        x = px + sgn * uy
        x0 = np.rint(x).astype(np.int64)
        y0 = np.rint(py).astype(np.int64)
        for dy in offs:
            yy = y0 + dy
            wy = np.exp(-((yy - py) ** 2) / (2 * 0.75 ** 2))
            oky = (yy >= 0) & (yy < H)
            for dx in offs:
                xx = x0 + dx
                ok = oky & (xx >= 0) & (xx < W)
                val = peak * wy * np.exp(-((xx - x) ** 2) / (2 * 0.75 ** 2))
                np.add.at(img, (yy[ok], xx[ok]), val[ok])
        frames.append(np.clip(np.rint(img), 0, 255).astype(np.uint8).astype(np.float32))
    return frames[0], frames[1]
