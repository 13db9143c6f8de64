"""fov_masked: the static corner mask of FovRenderer (envs/fov.py:46-68).

Four opaque black triangles with legs m = int(size * mask_frac) are drawn into the corners of the
observation with pygame.draw.polygon and blitted over every composed frame (fov.py:96-99).  The mask is
static, so the host rasterises it once and the engine keeps it device resident (cbev_upload_fov_mask).
The scan-line fill restates pygame 2.6.1 draw.c:draw_fillpoly (edge pixels included; parity unpinned,
see DESIGN.md -- for these 45-degree triangles every intersection is an integer, so the rounding rule of
the fill does not matter)."""
from __future__ import annotations

import math

import numpy as np


def fill_polygon(mask: np.ndarray, points) -> None:
    """Set mask[y, x] = 1 inside the polygon (scan-line fill, both end points of every span included)."""
    h, w = mask.shape
    xs = [int(p[0]) for p in points]
    ys = [int(p[1]) for p in points]
    n = len(points)
    miny, maxy = min(ys), max(ys)

    def hline(y, x1, x2):
        if 0 <= y < h:
            lo, hi = max(min(x1, x2), 0), min(max(x1, x2), w - 1)
            if hi >= lo:
                mask[y, lo:hi + 1] = 1

    if miny == maxy:
        hline(miny, min(xs), max(xs))
        return
    for y in range(miny, maxy + 1):
        inter = []
        for i in range(n):
            ip = i - 1 if i else n - 1
            y1, y2 = ys[ip], ys[i]
            if y1 < y2:
                x1, x2 = xs[ip], xs[i]
            elif y1 > y2:
                y2, y1 = ys[ip], ys[i]
                x2, x1 = xs[ip], xs[i]
            else:
                continue
            if (y1 <= y < y2) or (y == maxy and y2 == maxy):
                v = np.float32((y - y1) * (x2 - x1)) / np.float32(y2 - y1)
                v = math.floor(v) if len(inter) % 2 == 0 else math.ceil(v)
                inter.append(int(v) + x1)
        inter.sort()
        for i in range(0, len(inter) - 1, 2):
            hline(y, inter[i], inter[i + 1])
    for i in range(n):
        ip = i - 1 if i else n - 1
        if miny < ys[i] < maxy and ys[ip] == ys[i]:
            hline(ys[i], xs[i], xs[ip])


def corner_mask(size: int = 128, mask_frac: float = 0.5) -> np.ndarray:
    """uint8 [size, size], 1 where the observation is painted black (fov.py:46-68)."""
    m = int(size * mask_frac)
    s = size
    mask = np.zeros((s, s), dtype=np.uint8)
    for pts in ([(0, 0), (m, 0), (0, m)], [(s, 0), (s - m, 0), (s, m)], [(0, s), (0, s - m), (m, s)],
                [(s, s), (s - m, s), (s, s - m)]):
        fill_polygon(mask, pts)
    return mask
