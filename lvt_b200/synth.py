"""Deterministic synthetic stereo / RGB-D streams.

canvas(seed, H, Wc): float32 canvas filled with 128, H*Wc/density rectangles painted in draw
order, Gaussian blur sigma 1, additive N(0, noise) baked into the canvas, rounded to u8.
Frame t: left = canvas[:, s*t : s*t+W], right = canvas[:, s*t+d : s*t+d+W] -- a fronto-parallel
plane at Z = fx*b/d seen by a camera translating +x by s*Z/fx per frame (ground truth R = I).

Only elementwise float32 operations in a fixed order are used, so the bytes do not depend on
the BLAS / SIMD build of numpy.
"""
import numpy as np

# name -> (W, H, d, s, density, noise) ; calibration and tuning overrides live in configs.py
_BLUR = np.exp(-0.5 * np.arange(-3, 4, dtype=np.float64) ** 2)
_BLUR = (_BLUR / _BLUR.sum()).astype(np.float32)


def _blur_axis(a, axis):
    pad = [(0, 0), (0, 0)]
    pad[axis] = (3, 3)
    p = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    n = a.shape[axis]
    for k in range(7):
        sl = [slice(None), slice(None)]
        sl[axis] = slice(k, k + n)
        out += _BLUR[k] * p[tuple(sl)]
    return out


def canvas(seed, H, Wc, density=180, noise=2.0):
    rng = np.random.default_rng(seed)
    img = np.full((H, Wc), 128.0, np.float32)
    n = (H * Wc) // density
    xs = rng.integers(0, Wc, n)
    ys = rng.integers(0, H, n)
    ws = rng.integers(4, 40, n)
    hs = rng.integers(4, 40, n)
    vs = rng.integers(0, 256, n)
    for x, y, w, h, v in zip(xs, ys, ws, hs, vs):
        img[y:y + h, x:x + w] = v
    img = _blur_axis(_blur_axis(img, 1), 0)
    img = img + rng.normal(0.0, noise, img.shape).astype(np.float32)
    return np.clip(img + 0.5, 0, 255).astype(np.uint8)


class StereoStream:
    """Frames are views into one canvas; `frame(t)` returns contiguous (left, right) copies."""

    def __init__(self, W, H, n_frames, seed=0, disparity=20, step=16, density=180, noise=2.0):
        self.W, self.H, self.n_frames = W, H, n_frames
        self.d, self.s = disparity, step
        self.canvas = canvas(seed, H, W + disparity + step * (n_frames - 1), density, noise)

    def frame(self, t):
        o = self.s * t
        left = np.ascontiguousarray(self.canvas[:, o:o + self.W])
        right = np.ascontiguousarray(self.canvas[:, o + self.d:o + self.d + self.W])
        return left, right

    def ground_truth_t(self, t, fx, baseline):
        """camera position at frame t (R = I): x advances by s*Z/fx per frame, Z = fx*b/d."""
        Z = fx * baseline / self.d
        return np.array([t * self.s * Z / fx, 0.0, 0.0])


class RgbdStream:
    def __init__(self, W, H, n_frames, seed=1, step=8, depth=2.0, density=180, noise=2.0):
        self.W, self.H, self.n_frames, self.s, self.depth = W, H, n_frames, step, depth
        self.canvas = canvas(seed, H, W + step * (n_frames - 1), density, noise)

    def frame(self, t):
        o = self.s * t
        gray = np.ascontiguousarray(self.canvas[:, o:o + self.W])
        return gray, np.full((self.H, self.W), self.depth, np.float32)
