"""Seeded synthetic inputs for the hot path (numpy only; shared by tests and bench.py).

Images follow SURVEY.md §8(d): a scene of ~400 random filled rectangles / discs (uniform grey, 8-80 px)
seen through a window that drifts a few pixels per frame, plus per-frame Gaussian noise (sigma 4) and a
sigma-1 blur, so that every pyramid level of a 640x480 frame has a few thousand FAST-20 corners and
consecutive frames share structure (matches exist).  Frame 0 of camera 1 can be replaced by a flat
image to exercise the minThFAST fallback and the empty-level path.
"""
import numpy as np


def make_scene(seed, h, w, n_shapes=400):
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128, np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(n_shapes):
        g = float(rng.integers(0, 256))
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        sx, sy = int(rng.integers(8, 81)), int(rng.integers(8, 81))
        x0, x1 = max(cx - sx // 2, 0), min(cx + sx // 2 + 1, w)
        y0, y1 = max(cy - sy // 2, 0), min(cy + sy // 2 + 1, h)
        if rng.random() < 0.5:
            img[y0:y1, x0:x1] = g
        else:
            r = sx // 2
            m = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
            img[y0:y1, x0:x1][m] = g
    return img


def _blur_sigma1(img):
    k = np.exp(-0.5 * np.arange(-3, 4, dtype=np.float32) ** 2)
    k /= k.sum()
    p = np.pad(img, 3, mode="reflect")
    t = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(7))
    return sum(k[i] * t[i:i + img.shape[0], :] for i in range(7))


def render_frame(scene, ox, oy, W, H, noise_seed):
    rng = np.random.default_rng(noise_seed)
    crop = scene[oy:oy + H, ox:ox + W]
    f = _blur_sigma1(crop + rng.normal(0.0, 4.0, crop.shape).astype(np.float32))
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def dual_sequence(seed, n_frames, W=640, H=480, cams=2, flat_first=False, max_shift=8):
    """uint8 [n_frames][cams][H][W]: one drifting window per camera over its own scene."""
    out = np.empty((n_frames, cams, H, W), np.uint8)
    margin = 64
    for c in range(cams):
        scene = make_scene(seed * 131 + c, H + 2 * margin, W + 2 * margin)
        rng = np.random.default_rng(seed * 977 + c + 17)
        ox, oy = margin, margin
        for k in range(n_frames):
            out[k, c] = render_frame(scene, ox, oy, W, H, (seed * 1000003 + k) * 2 + c)
            ox = int(np.clip(ox + rng.integers(-max_shift, max_shift + 1), 0, 2 * margin))
            oy = int(np.clip(oy + rng.integers(-max_shift, max_shift + 1), 0, 2 * margin))
    if flat_first and cams > 1:
        out[0, 1] = 97
    return out


def tiled_batch(seed, n_frames, W=640, H=480, cams=2, unique=16):
    """Cheap large batch for benchmarking: `unique` rendered dual-frames, the rest are cyclic shifts of them
    (every frame distinct as a byte string, same corner statistics)."""
    base = dual_sequence(seed, min(unique, n_frames), W, H, cams)
    out = np.empty((n_frames, cams, H, W), np.uint8)
    for k in range(n_frames):
        b = base[k % base.shape[0]]
        r = k // base.shape[0]
        out[k] = np.roll(b, shift=(3 * r, 5 * r), axis=(1, 2)) if r else b
    return out


def random_descriptors(seed, n, p_flip=None, base=None):
    """n x 32 uint8 descriptors; with `base` given, a noisy copy (each bit flipped with prob p_flip)."""
    rng = np.random.default_rng(seed)
    if base is None:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    flips = np.packbits(rng.random((base.shape[0], 256)) < p_flip, axis=1)
    return base ^ flips
