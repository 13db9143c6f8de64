"""Oracle: BEV rendering, resize, semantic masks, grayscale.  TEST INFRASTRUCTURE ONLY.

Restates reference rows a6, a7, a14-a18 as a *direct per-output-pixel*
formulation: every pixel of the 128x128 field of view is inverse-mapped through
pygame.transform.rotate's 16.16 fixed-point walk into the padded scene, whose colour is
resolved from the draw list (last drawn wins) or the class map.  The literal
surface-by-surface formulation lives in oracle/shims/pygame (used only to run the
unmodified reference); tests check that both agree pixel-for-pixel.

pygame / gymnasium arithmetic is third-party: PARITY UNPINNED beyond the reference's own
contracts (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np

# palette codes shared with the CUDA engine (include/cbev.h)
PAL_NON_DRIVABLE, PAL_DRIVABLE, PAL_SIDEWALK, PAL_VEHICLE, PAL_PEDESTRIAN, PAL_ROUTE, PAL_TL_RED, PAL_TL_YELLOW, \
    PAL_BLACK, PAL_TL_OFF = range(10)
PALETTE = np.array([
    (150, 150, 150),  # NON_DRIVABLE  semantics.py:20
    (255, 255, 255),  # DRIVABLE
    (220, 220, 220),  # SIDEWALK
    (0, 7, 175),      # VEHICLE
    (255, 0, 0),      # PEDESTRIAN
    (0, 255, 0),      # ROUTE (targets, green traffic light)
    (255, 64, 64),    # TRAFFIC_LIGHT_RED
    (255, 255, 0),    # yellow traffic light (traffic_light.py:50)
    (0, 0, 0),        # ego square / composition background (actor_manager.py:45, fov.py:92)
    (100, 100, 100),  # unknown traffic-light state (traffic_light.py:54)
], dtype=np.uint8)

MASK_CHANNELS = {  # wrappers/rgb_to_semantic.py:6-42
    "binary": ("drivable",),
    "2-class": ("drivable", "route"),
    "4-class": ("drivable", "vehicle", "pedestrian", "route"),
    "5-class": ("drivable", "sidewalk", "vehicle", "pedestrian", "route"),
    "6-class": ("non_drivable", "drivable", "sidewalk", "vehicle", "pedestrian", "route"),
    "7-class": ("non_drivable", "drivable", "sidewalk", "vehicle", "pedestrian", "route", "traffic_light_red"),
}


class FovGeometry:
    """FovRenderer constants, fov.py:30-44; padding = crop_size (world.py:69-78)."""

    def __init__(self, size=128, ax_frac=0.5, ay_frac=0.5):
        self.size = int(size)
        m = self.size - 1
        ax = max(0, min(m, int(round(m * ax_frac))))
        ay = max(0, min(m, int(round(m * ay_frac))))
        self.anchor = (ax, ay)
        radius = math.hypot(max(ax, m - ax), max(ay, m - ay))
        self.crop = max(self.size, int(math.ceil(2.0 * radius)))
        self.pad = self.crop


def crop_origin(x, y, geom: FovGeometry, map_w, map_h):
    """Follow.scroll + compute_crop_rect: camera.py:39-42, world.py:105-111, fov.py:70-79."""
    crop, pad = geom.crop, geom.pad
    off_x = int((float(pad) + float(x) * 1.0) + (-crop / 2))
    off_y = int((float(pad) + float(y) * 1.0) + (-crop / 2))
    cx = int(round(float(off_x) + crop / 2.0))
    cy = int(round(float(off_y) + crop / 2.0))
    xmin = max(0, min(max(0, map_w + 2 * pad - crop), cx - crop // 2))
    ymin = max(0, min(max(0, map_h + 2 * pad - crop), cy - crop // 2))
    return xmin, ymin


def rotate_params(yaw, crop):
    """pygame.transform.rotate set-up for angle = degrees(yaw)+90 (fov.py:84-88).

    Returns dict(mode=0 turns=k) for the exact 90-degree path or
    dict(mode=1, nx, ny, isin, icos, ax, ay, xd, yd, cy) for the fixed-point walk."""
    angle = float(np.float32(math.degrees(yaw) + 90))
    if math.fmod(angle, 90.0) == 0.0:
        a = int(angle)
        q = abs(a) // 90
        if a < 0:
            q = -q
        turns = int(math.fmod(q, 4))
        if turns < 0:
            turns += 4
        return dict(mode=0, turns=turns, nx=crop, ny=crop)
    rad = angle * 0.01745329251994329
    s, c = math.sin(rad), math.cos(rad)
    w = h = crop
    cxx, cyy, sx, sy = c * w, c * h, s * w, s * h
    nx = int(max(abs(cxx + sy), abs(cxx - sy), abs(-cxx + sy), abs(-cxx - sy)))
    ny = int(max(abs(sx + cyy), abs(sx - cyy), abs(-sx + cyy), abs(-sx - cyy)))
    return dict(mode=1, nx=nx, ny=ny, isin=int(s * 65536), icos=int(c * 65536),
                ax=(nx << 15) - int(c * ((nx - 1) << 15)), ay=(ny << 15) - int(s * ((nx - 1) << 15)),
                xd=(w - nx) << 15, yd=(h - ny) << 15, cy=ny // 2)


def draw_list(sim, with_actors=True):
    """Rectangles drawn on the padded scene in draw order (actor_manager.py:121-132):
    vehicles, pedestrians, visible targets, traffic lights.  Each is (x, y, w, h, palette)."""
    rects = []
    if not with_actors:
        return rects
    from .sim import KIND_PEDESTRIAN, KIND_VEHICLE, rect_left

    pad = sim.pad
    for kind, pal in ((KIND_VEHICLE, PAL_VEHICLE), (KIND_PEDESTRIAN, PAL_PEDESTRIAN)):
        for a in sim.actors:
            if a.kind == kind:
                rects.append((rect_left(a.x, pad, a.size), rect_left(a.y, pad, a.size), a.size, a.size, pal))
    n = len(sim.tgt_x)
    for i in range(n):
        if sim.tgt_visible_at_draw[i]:
            size = 4 if i == n - 1 else 2
            rects.append((rect_left(sim.tgt_x[i], pad, size), rect_left(sim.tgt_y[i], pad, size), size, size, PAL_ROUTE))
    tl = sim.scene.get("tl_rect")
    if tl is not None:
        for (x, y, w, h), col in zip(tl, sim.scene["tl_color"]):
            rects.append((int(x), int(y), int(w), int(h), int(col)))  # drawn WITHOUT the padding offset (quirk C-4)
    return rects


def _fill_polygon(mask, points):
    """pygame 2.6.1 draw.c:draw_fillpoly restated (scan-line fill incl. edge pixels; parity unpinned)."""
    h, w = mask.shape
    xs = [int(p[0]) for p in points]
    ys = [int(p[1]) for p in points]
    n = len(points)
    miny, maxy = min(ys), max(ys)

    def hline(yy, x1, x2):
        if 0 <= yy < h:
            lo, hi = max(min(x1, x2), 0), min(max(x1, x2), w - 1)
            if hi >= lo:
                mask[yy, lo:hi + 1] = True

    for yy in range(miny, maxy + 1):
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
            if (y1 <= yy < y2) or (yy == maxy and y2 == maxy):
                v = np.float32((yy - y1) * (x2 - x1)) / np.float32(y2 - y1)
                v = math.floor(v) if len(inter) % 2 == 0 else math.ceil(v)
                inter.append(int(v) + x1)
        inter.sort()
        for i in range(0, len(inter) - 1, 2):
            hline(yy, inter[i], inter[i + 1])
    for i in range(n):
        ip = i - 1 if i else n - 1
        if miny < ys[i] < maxy and ys[ip] == ys[i]:
            hline(ys[i], xs[i], xs[ip])


def corner_mask(size=128, mask_frac=0.5):
    """FovRenderer._build_mask_surface, fov.py:46-68: True where apply_mask paints the frame black."""
    m = int(size * mask_frac)
    s = size
    mask = np.zeros((s, s), dtype=bool)
    for pts in ([(0, 0), (m, 0), (0, m)], [(s, 0), (s - m, 0), (s, m)], [(0, s), (0, s - m), (m, s)],
                [(s, s), (s - m, s), (s, s - m)]):
        _fill_polygon(mask, pts)
    return mask


def render_fov(cls_map, geom: FovGeometry, x, y, theta, rects, fov_mask=None):
    """128x128 palette-index image of the ego-centred, ego-aligned field of view.

    world.py:137-157 (draw_fov) = crop (fov.py:70-82) -> rotate (fov.py:84-88) -> compose on
    black centred at the anchor (fov.py:90-94) -> ego 4x4 black square (hero.py:26-32)."""
    H, W = cls_map.shape
    crop, pad, size = geom.crop, geom.pad, geom.size
    xmin, ymin = crop_origin(x, y, geom, W, H)
    # padded scene crop as palette indices: map (NON_DRIVABLE outside) + rects in order
    ys = np.arange(ymin, ymin + crop) - pad
    xs = np.arange(xmin, xmin + crop) - pad
    inside = ((ys >= 0) & (ys < H))[:, None] & ((xs >= 0) & (xs < W))[None, :]
    tile = np.where(inside, cls_map[np.clip(ys, 0, H - 1)[:, None], np.clip(xs, 0, W - 1)[None, :]], PAL_NON_DRIVABLE)
    tile = tile.astype(np.uint8)
    sw, sh = W + 2 * pad, H + 2 * pad
    for rx, ry, rw, rh, pal in rects:
        x0, y0 = max(rx, 0), max(ry, 0)            # clip to the scene surface (draw.c)
        x1, y1 = min(rx + rw, sw), min(ry + rh, sh)
        x0, y0, x1, y1 = x0 - xmin, y0 - ymin, x1 - xmin, y1 - ymin
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, crop), min(y1, crop)
        if x1 > x0 and y1 > y0:
            tile[y0:y1, x0:x1] = pal
    rp = rotate_params(theta, crop)
    nx, ny = rp["nx"], rp["ny"]
    left = geom.anchor[0] - (nx >> 1)              # rotated.get_rect(center=anchor)
    top = geom.anchor[1] - (ny >> 1)
    oy, ox = np.mgrid[0:size, 0:size]
    rxp, ryp = ox - left, oy - top
    in_rot = (rxp >= 0) & (rxp < nx) & (ryp >= 0) & (ryp < ny)
    if rp["mode"] == 0:
        t = rp["turns"]
        if t == 0:
            sx, sy = rxp, ryp
        elif t == 1:
            sx, sy = crop - 1 - ryp, rxp
        elif t == 2:
            sx, sy = crop - 1 - rxp, crop - 1 - ryp
        else:
            sx, sy = ryp, crop - 1 - rxp
        oob = np.zeros_like(in_rot)
    else:
        dx = rp["ax"] + rp["isin"] * (rp["cy"] - ryp) + rp["xd"] + rxp * rp["icos"]
        dy = rp["ay"] - rp["icos"] * (rp["cy"] - ryp) + rp["yd"] + rxp * rp["isin"]
        lim = (crop << 16) - 1
        oob = (dx < 0) | (dy < 0) | (dx > lim) | (dy > lim)
        sx, sy = dx >> 16, dy >> 16
    val = tile[np.clip(sy, 0, crop - 1), np.clip(sx, 0, crop - 1)]
    val = np.where(oob, tile[0, 0], val)           # background = crop's first pixel (transform.c)
    out = np.where(in_rot, val, PAL_BLACK).astype(np.uint8)
    if fov_mask is not None:                       # apply_mask before the ego is drawn (world.py:150-156)
        out[fov_mask] = PAL_BLACK
    ax, ay = geom.anchor
    hw = int(32 / int(1024 / size))                # hero.py:14-17: 4 px at the 128 scale
    out[max(ay - (hw >> 1), 0):ay - (hw >> 1) + hw, max(ax - (hw >> 1), 0):ax - (hw >> 1) + hw] = PAL_BLACK   # ego square, rect centred on the anchor
    return out


def fov_rgb(idx_img):
    """CarlaBEV.render, carlabev.py:233-236: (H, W, 3) uint8."""
    return PALETTE[idx_img]


def _area_table(ssize, dsize):
    """OpenCV computeResizeAreaTab for one axis."""
    scale = ssize / dsize
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _linear_area_tab(ssize, dsize):
    """One axis of cv::resize's bilinear set-up in `area_mode` (OpenCV imgproc/resize.cpp, the dx loop of
    cv::resize): source index and the two 11-bit fixed-point taps (INTER_RESIZE_COEF_BITS = 11)."""
    scale = ssize / dsize
    inv = 1.0 / scale
    ofs = np.zeros(dsize, np.int32)
    taps = np.zeros((dsize, 2), np.int32)
    for dx in range(dsize):
        sx = math.floor(dx * scale)
        fx = np.float32((dx + 1) - (sx + 1) * inv)
        fx = np.float32(0.0) if fx <= 0 else np.float32(fx - np.floor(fx))
        if sx >= ssize - 1:
            fx, sx = np.float32(0.0), ssize - 1
        ofs[dx] = sx
        taps[dx, 0] = int(np.rint(np.float32((np.float32(1.0) - fx) * np.float32(2048))))
        taps[dx, 1] = int(np.rint(np.float32(fx * np.float32(2048))))
    return ofs, taps


def resize_linear_area_mode(img, out_hw):
    """cv2.resize(..., INTER_AREA) when an axis ENLARGES (EnvConfig.size = 64 with obs_size 96):
    OpenCV only implements true area interpolation for scale >= 1 on both axes and otherwise emulates it with its 8-bit
    bilinear kernel on area-mode coefficients -- HResizeLinear (int32 rows = S0 * a0 + S1 * a1) then VResizeLinear
    ((b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2) >> 2.  Pinned on cv2 itself in tests/test_oracle_contracts.py."""
    sh, sw = img.shape[:2]
    dh, dw = out_hw
    xo, xa = _linear_area_tab(sw, dw)
    yo, ya = _linear_area_tab(sh, dh)
    tail = (1,) * (img.ndim - 2)
    src = img.astype(np.int32)
    rows = src[:, xo] * xa[:, 0].reshape((1, dw) + tail) + src[:, np.minimum(xo + 1, sw - 1)] * xa[:, 1].reshape((1, dw) + tail)
    b0, b1 = ya[:, 0].reshape((dh, 1) + tail), ya[:, 1].reshape((dh, 1) + tail)
    out = (((b0 * (rows[yo] >> 4)) >> 16) + ((b1 * (rows[np.minimum(yo + 1, sh - 1)] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_area(img, out_hw):
    """gymnasium ResizeObservation -> cv2.resize(..., INTER_AREA) when shrinking (envs/__init__.py:62):
    equal size -> copy; exactly half on both axes -> OpenCV's 2x2 integer path (a + b + c + d + 2) >> 2;
    everything else (integer or fractional scale) -> area tables, float32 accumulate, round-half-even.
    Pinned on cv2 itself for 20 sizes in tests/test_oracle_contracts.py."""
    sh, sw = img.shape[:2]
    dh, dw = out_hw
    if (dh, dw) == (sh, sw):
        return img.copy()
    if dh > sh or dw > sw:
        return resize_linear_area_mode(img, out_hw)
    if (2 * dh, 2 * dw) == (sh, sw):
        s = img.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xt, yt = _area_table(sw, dw), _area_table(sh, dh)
    src = img.astype(np.float32)
    tmp = np.zeros((sh, dw) + img.shape[2:], dtype=np.float32)
    for dx, sx, a in xt:
        tmp[:, dx] += src[:, sx] * a
    out = np.zeros((dh, dw) + img.shape[2:], dtype=np.float32)
    for dy, sy, b in yt:
        out[dy] += tmp[sy] * b
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def semantic_masks(rgb, mode="6-class"):
    """rgb_to_semantic_mask, wrappers/rgb_to_semantic.py:65-142."""
    eq = lambda pal: np.all(rgb == PALETTE[pal], axis=-1)  # noqa: E731
    planes = {
        "non_drivable": eq(PAL_NON_DRIVABLE),
        "drivable": eq(PAL_DRIVABLE) | eq(PAL_ROUTE),
        "sidewalk": eq(PAL_SIDEWALK),
        "vehicle": eq(PAL_VEHICLE),
        "pedestrian": eq(PAL_PEDESTRIAN),
        "route": eq(PAL_ROUTE),
        "traffic_light_red": eq(PAL_TL_RED),
    }
    return np.stack([planes[c] for c in MASK_CHANNELS[mode]]).astype(np.float32)


def grayscale(rgb):
    """gymnasium GrayscaleObservation (envs/__init__.py:70)."""
    return np.sum(np.multiply(rgb, np.array([0.2125, 0.7154, 0.0721])), axis=-1).astype(np.uint8)


VEHICLE_CHANNEL = {m: ch.index("vehicle") for m, ch in MASK_CHANNELS.items() if "vehicle" in ch}


def fuse_vehicle_temporal(stacked, mode, history_frames=3):
    """fuse_vehicle_temporal_channels, wrappers/rgb_to_semantic.py:152-168; stacked is (F, C, H, W)."""
    v = VEHICLE_CHANNEL[mode]
    history = stacked[-history_frames:]
    current = history[-1]
    return np.concatenate([np.delete(current, v, axis=0), history[::-1, v]], axis=0).astype(np.float32)


def fuse_vehicle_weighted(stacked, mode, weights=(1.0, 0.5, 0.25)):
    """fuse_weighted_vehicle_history, wrappers/rgb_to_semantic.py:171-193."""
    v = VEHICLE_CHANNEL[mode]
    history = stacked[-len(weights):][::-1]
    current = history[0]
    acc = np.zeros_like(current[v], dtype=np.float32)
    for frame, w in zip(history, weights):
        acc += w * frame[v]
    acc = np.clip(acc, 0.0, 1.0)
    return np.concatenate([np.delete(current, v, axis=0), acc[None]], axis=0).astype(np.float32)
