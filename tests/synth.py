"""Seeded synthetic RGB-D frames with the statistics of the reference's captures
(SURVEY.md 8d / Appendix C): piecewise-smooth planes 400-3000 mm with 2 mm noise, ~5 % far
background, ~42 % holes laid out as horizontal runs, a few saturated 65535 pixels."""
import numpy as np


def synth_depth(h, w, seed, hole_frac=0.42):
    rng = np.random.default_rng(seed)
    v, u = np.mgrid[0:h, 0:w].astype(np.float32)
    a, b = rng.uniform(-0.6, 0.6, 2)
    base = rng.uniform(600, 2200) + a * (u - w / 2) + b * (v - h / 2)
    step = (u > rng.uniform(0.3, 0.7) * w) * rng.uniform(-300, 600)
    z = base + step + rng.normal(0, 2.0, (h, w))
    far = rng.random((h, w)) < 0.05
    z = np.where(far, rng.uniform(3000, 10000, (h, w)), z)
    z = np.clip(z, 150, 65000)
    # run-coherent holes: threshold a field that is smooth along rows
    fld = rng.normal(size=(h, w // 16 + 2))
    fld = np.repeat(fld, 16, axis=1)[:, :w] + 0.35 * rng.normal(size=(h, w))
    thr = np.quantile(fld, hole_frac)
    z = np.where(fld < thr, 0, z)
    out = z.astype(np.uint16)
    sat = rng.random((h, w)) < 0.0003
    out[sat] = 65535
    return out


def synth_color(h, w, seed):
    rng = np.random.default_rng(seed + 1000)
    v, u = np.mgrid[0:h, 0:w]
    g = ((u * 255 // max(w - 1, 1))[..., None] + (v * 255 // max(h - 1, 1))[..., None] * np.array([1, 2, 3])) % 256
    n = rng.integers(0, 256, (h, w, 3))
    return ((g + n) // 2).astype(np.uint8)


def synth_mask(h, w, seed, cover=0.15):
    rng = np.random.default_rng(seed + 2000)
    v, u = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), bool)
    for _ in range(3):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        rx, ry = rng.uniform(0.1, 0.25) * w, rng.uniform(0.1, 0.25) * h
        m |= ((u - cx) / rx) ** 2 + ((v - cy) / ry) ** 2 < 1
    return (m * 255).astype(np.uint8)


def synth_batch(b, h, w, seed0=0):
    d = np.stack([synth_depth(h, w, seed0 + i) for i in range(b)])
    c = np.stack([synth_color(h, w, seed0 + i) for i in range(b)])
    return d, c


def bumpy_surface(rng, n, noise=0.0):
    """A curved, non-symmetric surface patch (what a depth camera sees of an object): constrains all six degrees of freedom."""
    u = rng.uniform(-0.15, 0.15, n)
    v = rng.uniform(-0.10, 0.10, n)
    z = 0.6 + 0.35 * u * u - 0.5 * u * v + 0.04 * np.sin(25.0 * u) * np.cos(18.0 * v) + 0.6 * v * v * v
    P = np.stack([u, v, z], axis=1)
    return P + rng.normal(size=P.shape) * noise if noise else P
