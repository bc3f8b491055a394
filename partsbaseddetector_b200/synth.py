"""Synthetic inputs for parity tests and the benchmark:
frames = low-pass noise + 40% white noise + a few filled shapes, BGR u8."""
import numpy as np


def _box_blur(a, k):
    # separable box filter with edge replication, numpy only
    pad = k // 2
    out = a.astype(np.float32)
    for axis in (0, 1):
        p = np.pad(out, [(pad, pad) if i == axis else (0, 0) for i in range(out.ndim)], mode="edge")
        c = np.cumsum(p, axis=axis, dtype=np.float64)
        z = np.zeros_like(np.take(c, [0], axis=axis))
        c = np.concatenate([z, c], axis=axis)
        hi = np.take(c, np.arange(k, k + out.shape[axis]), axis=axis)
        lo = np.take(c, np.arange(0, out.shape[axis]), axis=axis)
        out = ((hi - lo) / k).astype(np.float32)
    return out


def synth_frame(idx, h=480, w=640):
    rng = np.random.default_rng(1234 + idx)
    base = _box_blur(rng.integers(0, 256, (h, w, 3)).astype(np.float32), 9)
    base = (base - base.mean()) * 3.0 + 128.0
    img = 0.6 * base + 0.4 * rng.integers(0, 256, (h, w, 3)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(int(rng.integers(3, 6))):
        col = rng.integers(0, 256, 3).astype(np.float32)
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        rx, ry = rng.integers(w // 16, w // 4), rng.integers(h // 16, h // 4)
        if rng.integers(0, 2):
            mask = (np.abs(xx - cx) < rx) & (np.abs(yy - cy) < ry)
        else:
            mask = ((xx - cx) / float(rx)) ** 2 + ((yy - cy) / float(ry)) ** 2 < 1.0
        img[mask] = 0.5 * img[mask] + 0.5 * col
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_frames(n, h=480, w=640, start=0):
    return np.stack([synth_frame(start + i, h, w) for i in range(n)])


def synth_score_map(idx, h, w):
    return np.random.default_rng(4242 + idx).standard_normal((h, w)).astype(np.float32)
