"""Deterministic synthetic camera rigs -- workload DATA for tests and bench.py.

No tracking arithmetic lives here: frames are textured backgrounds with a sub-pixel global shift and
independently moving textured rectangles (the "targets"), shaped like PETS2009 pedestrians
(w in [30,55]*(W/768), h = 2.4*w).  seed = 1234 + 100*cam + frame, numpy default_rng.
"""
import cv2
import numpy as np


def texture(h, w, rng, low_texture=False):
    """Band-limited random texture: coarse uniform grid upsampled bicubically + blurred noise."""
    gh, gw = h // 4 + 8, w // 4 + 8
    grid = rng.integers(0, 256, size=(gh, gw)).astype(np.float32)
    img = cv2.resize(grid, (w + 32, h + 32), interpolation=cv2.INTER_CUBIC)
    img = cv2.GaussianBlur(img, (0, 0), 1.5)
    noise = rng.normal(0.0, 12.0, size=img.shape).astype(np.float32)
    img += cv2.GaussianBlur(noise, (0, 0), 0.8)
    if low_texture:
        # flat and heavily blurred regions so that some windows fail the min-eigenvalue gate
        for _ in range(6):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            ww, hh = int(rng.integers(w // 8, w // 3)), int(rng.integers(h // 8, h // 3))
            if rng.random() < 0.5:
                img[y:y + hh, x:x + ww] = 128.0
            else:
                img[y:y + hh, x:x + ww] = cv2.GaussianBlur(img[y:y + hh, x:x + ww], (0, 0), 6.0)
    return img  # float32 (h+32, w+32)


def _shifted(img, dx, dy, w, h, off=16):
    """Crop (h, w) out of img at offset `off` after translating the content by (dx, dy), bicubic."""
    m = np.float32([[1, 0, dx - off], [0, 1, dy - off]])
    return cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT_101)


class Camera:
    """One synthetic camera: frame(t) -> u8 gray [h, w]; boxes(t) -> float64 [n_targets, 4] (x,y,w,h)."""

    def __init__(self, width, height, n_targets, cam=0, low_texture=False, static_frac=0.10,
                 leaving_frac=0.05, max_speed=4.0):
        self.w, self.h, self.cam = width, height, cam
        rng = np.random.default_rng(1234 + 100 * cam)
        self.bg = texture(height, width, rng, low_texture)
        s = width / 768.0
        self.targets = []
        for i in range(n_targets):
            tw = float(rng.uniform(30, 55)) * s
            th = 2.4 * tw
            x = float(rng.uniform(8, max(9.0, width - tw - 8)))
            y = float(rng.uniform(8, max(9.0, height - th - 8)))
            u = rng.random()
            if u < static_frac:
                v = (0.0, 0.0)
            elif u < static_frac + leaving_frac:
                x = width - tw * 0.6
                v = (float(max_speed * s * 4), 0.0)
            else:
                v = (float(rng.uniform(-max_speed, max_speed)) * s, float(rng.uniform(-max_speed, max_speed)) * s)
            tex = texture(int(np.ceil(th)) + 2, int(np.ceil(tw)) + 2, np.random.default_rng(99991 * (cam + 1) + i))
            self.targets.append(dict(x=x, y=y, w=tw, h=th, v=v, tex=tex))

    def boxes(self, t):
        return np.array([[tg["x"] + tg["v"][0] * t, tg["y"] + tg["v"][1] * t, tg["w"], tg["h"]]
                         for tg in self.targets], np.float64).reshape(-1, 4)

    def frame(self, t):
        rng = np.random.default_rng(1234 + 100 * self.cam + t)
        gdx, gdy = rng.uniform(-0.5, 0.5, size=2) if t > 0 else (0.0, 0.0)
        img = _shifted(self.bg, float(gdx), float(gdy), self.w, self.h)
        for tg, box in zip(self.targets, self.boxes(t)):
            x, y, tw, th = box
            ix, iy = int(np.floor(x)), int(np.floor(y))
            pw, ph = int(np.ceil(tw)), int(np.ceil(th))
            patch = _shifted(tg["tex"], x - ix, y - iy, pw, ph)
            x0, y0, x1, y1 = max(ix, 0), max(iy, 0), min(ix + pw, self.w), min(iy + ph, self.h)
            if x1 <= x0 or y1 <= y0:
                continue
            img[y0:y1, x0:x1] = patch[y0 - iy:y1 - iy, x0 - ix:x1 - ix]
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def seeds_in_boxes(boxes, n_per_target, width, height, rng):
    """Uniform random seed points inside each (cropped) box -> float32 [n_targets*n_per_target, 2], target ids."""
    pts, tid = [], []
    for i, (x, y, w, h) in enumerate(boxes):
        x0, y0 = max(x, 2.0), max(y, 2.0)
        x1, y1 = min(x + w, width - 3.0), min(y + h, height - 3.0)
        if x1 <= x0:
            x0, x1 = 2.0, width - 3.0
        if y1 <= y0:
            y0, y1 = 2.0, height - 3.0
        px = rng.uniform(x0, x1, size=n_per_target)
        py = rng.uniform(y0, y1, size=n_per_target)
        pts.append(np.stack([px, py], 1))
        tid.append(np.full(n_per_target, i, np.int32))
    if not pts:
        return np.zeros((0, 2), np.float32), np.zeros(0, np.int32)
    return np.concatenate(pts).astype(np.float32), np.concatenate(tid)


def grid_seeds(width, height, stride=4, start=2):
    """Config C4's dense grid: x,y in {2,6,10,...}."""
    xs = np.arange(start, width, stride, dtype=np.float32)
    ys = np.arange(start, height, stride, dtype=np.float32)
    gx, gy = np.meshgrid(xs, ys)
    return np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)
