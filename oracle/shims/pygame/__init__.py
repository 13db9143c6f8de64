"""pygame-lite: TEST INFRASTRUCTURE ONLY (never imported by the product path).

A small pure-Python/NumPy restatement of the part of pygame 2.6.1 (pinned by the
reference's uv.lock:792-793) that the reference's hot path calls, so that the
UNMODIFIED reference under /root/reference can be imported and stepped in this
container (pygame itself is not installed, there are no SDL libraries).  It is
used by oracle/gen_golden.py to produce the fixtures in tests/golden/.

The arithmetic is restated from upstream pygame 2.6.1 (src_c/rect.c, draw.c,
surface.c, transform.c).  PARITY UNPINNED beyond the reference's own contracts
(anchor pixel == hero colour, crop alignment; validate_simulator_semantics.py:366-414,
tests/test_seeded_scene_consistency.py:128-138): real pygame cannot run here.

Surfaces are (h, w, 3) uint8 arrays (32-bpp XRGB without alpha behaves the same
for fill/blit/draw.rect/rotate); SRCALPHA surfaces carry an extra alpha plane.
"""
from __future__ import annotations

import math as _math

import numpy as _np

SRCALPHA = 0x00010000


def init():
    return (0, 0)


def quit():
    return None


def _trunc(v):
    # pg_IntFromObj: floats are truncated toward zero ((int)double)
    return int(v)


class Rect:
    __slots__ = ("x", "y", "w", "h")

    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            if isinstance(a, Rect):
                args = (a.x, a.y, a.w, a.h)
            else:
                args = tuple(a)
        if len(args) == 2:
            (x, y), (w, h) = args
        elif len(args) == 4:
            x, y, w, h = args
        else:
            raise TypeError("Argument must be rect style object")
        self.x, self.y, self.w, self.h = _trunc(x), _trunc(y), _trunc(w), _trunc(h)

    # --- geometry ---------------------------------------------------------
    width = property(lambda s: s.w, lambda s, v: setattr(s, "w", _trunc(v)))
    height = property(lambda s: s.h, lambda s, v: setattr(s, "h", _trunc(v)))
    left = property(lambda s: s.x, lambda s, v: setattr(s, "x", _trunc(v)))
    top = property(lambda s: s.y, lambda s, v: setattr(s, "y", _trunc(v)))
    right = property(lambda s: s.x + s.w)
    bottom = property(lambda s: s.y + s.h)
    size = property(lambda s: (s.w, s.h))
    topleft = property(lambda s: (s.x, s.y))
    centerx = property(lambda s: s.x + (s.w >> 1))
    centery = property(lambda s: s.y + (s.h >> 1))

    @property
    def center(self):
        return (self.x + (self.w >> 1), self.y + (self.h >> 1))

    @center.setter
    def center(self, value):
        cx, cy = value
        self.x = _trunc(cx) - (self.w >> 1)
        self.y = _trunc(cy) - (self.h >> 1)

    def copy(self):
        return Rect(self.x, self.y, self.w, self.h)

    def colliderect(self, other):
        o = other if isinstance(other, Rect) else Rect(other)
        if self.w == 0 or self.h == 0 or o.w == 0 or o.h == 0:
            return False
        return (
            min(self.x, self.x + self.w) < max(o.x, o.x + o.w)
            and min(self.y, self.y + self.h) < max(o.y, o.y + o.h)
            and max(self.x + self.w, self.x) > min(o.x, o.x + o.w)
            and max(self.y + self.h, self.y) > min(o.y, o.y + o.h)
        )

    def collidepoint(self, *p):
        if len(p) == 1:
            p = p[0]
        px, py = p
        return self.x <= px < self.x + self.w and self.y <= py < self.y + self.h

    def clip(self, other):
        o = other if isinstance(other, Rect) else Rect(other)
        x0, y0 = max(self.x, o.x), max(self.y, o.y)
        x1, y1 = min(self.x + self.w, o.x + o.w), min(self.y + self.h, o.y + o.h)
        if x1 <= x0 or y1 <= y0:
            return Rect(self.x, self.y, 0, 0)
        return Rect(x0, y0, x1 - x0, y1 - y0)

    def __iter__(self):
        return iter((self.x, self.y, self.w, self.h))

    def __getitem__(self, i):
        return (self.x, self.y, self.w, self.h)[i]

    def __len__(self):
        return 4

    def __eq__(self, other):
        try:
            return tuple(self) == tuple(Rect(other))
        except Exception:
            return False

    def __repr__(self):
        return f"<rect({self.x}, {self.y}, {self.w}, {self.h})>"


class _Vec:
    _n = 2

    def __init__(self, *args):
        if len(args) == 0:
            vals = [0.0] * self._n
        elif len(args) == 1 and not isinstance(args[0], (int, float, _np.floating, _np.integer)):
            vals = [float(v) for v in args[0]]
        elif len(args) == 1:
            vals = [float(args[0])] * self._n
        else:
            vals = [float(v) for v in args]
        if len(vals) != self._n:
            raise ValueError("wrong vector length")
        self._v = vals

    x = property(lambda s: s._v[0], lambda s, v: s._v.__setitem__(0, float(v)))
    y = property(lambda s: s._v[1], lambda s, v: s._v.__setitem__(1, float(v)))

    def __iter__(self):
        return iter(self._v)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._v[i]

    def __setitem__(self, i, v):
        self._v[i] = float(v)

    def _coerce(self, o):
        o = list(o)
        if len(o) != self._n:
            raise TypeError("vector length mismatch")
        return [float(v) for v in o]

    def __add__(self, o):
        return type(self)([a + b for a, b in zip(self._v, self._coerce(o))])

    __radd__ = __add__

    def __sub__(self, o):
        return type(self)([a - b for a, b in zip(self._v, self._coerce(o))])

    def __rsub__(self, o):
        return type(self)([b - a for a, b in zip(self._v, self._coerce(o))])

    def __mul__(self, k):
        return type(self)([a * float(k) for a in self._v])

    __rmul__ = __mul__

    def __truediv__(self, k):
        return type(self)([a / float(k) for a in self._v])

    def __neg__(self):
        return type(self)([-a for a in self._v])

    def __array__(self, dtype=None, copy=None):
        return _np.array(self._v, dtype=dtype or float)

    def length(self):
        return _math.sqrt(sum(a * a for a in self._v))

    def __repr__(self):
        return f"<Vector{self._n}({', '.join(repr(a) for a in self._v)})>"


class Vector2(_Vec):
    _n = 2


class Vector3(_Vec):
    _n = 3
    z = property(lambda s: s._v[2], lambda s, v: s._v.__setitem__(2, float(v)))


class _MathModule:
    Vector2 = Vector2
    Vector3 = Vector3


math = _MathModule()


def _color3(color):
    c = tuple(int(v) for v in color)
    return c[:3], (c[3] if len(c) > 3 else 255)


class Surface:
    def __init__(self, size, flags=0, depth=0, masks=None, _array=None, _alpha=None):
        if _array is not None:
            self._a = _array
            self._alpha = _alpha
            return
        w, h = int(size[0]), int(size[1])
        self._a = _np.zeros((h, w, 3), dtype=_np.uint8)
        self._alpha = _np.zeros((h, w), dtype=_np.uint8) if (flags & SRCALPHA) else None

    # --- info --------------------------------------------------------------
    def get_size(self):
        return (self._a.shape[1], self._a.shape[0])

    def get_width(self):
        return self._a.shape[1]

    def get_height(self):
        return self._a.shape[0]

    def get_rect(self, **kwargs):
        r = Rect(0, 0, self._a.shape[1], self._a.shape[0])
        for k, v in kwargs.items():
            setattr(r, k, v)
        return r

    def convert(self, *a, **k):
        return self

    convert_alpha = convert

    def copy(self):
        return Surface(None, _array=self._a.copy(), _alpha=None if self._alpha is None else self._alpha.copy())

    def get_at(self, pos):
        x, y = pos
        r, g, b = (int(v) for v in self._a[y, x])
        return (r, g, b, 255 if self._alpha is None else int(self._alpha[y, x]))

    # --- drawing -----------------------------------------------------------
    def fill(self, color, rect=None):
        rgb, a = _color3(color)
        if rect is None:
            self._a[:, :] = rgb
            if self._alpha is not None:
                self._alpha[:, :] = a
            return self.get_rect()
        r = Rect(rect).clip(self.get_rect())
        if r.w > 0 and r.h > 0:
            self._a[r.y : r.y + r.h, r.x : r.x + r.w] = rgb
            if self._alpha is not None:
                self._alpha[r.y : r.y + r.h, r.x : r.x + r.w] = a
        return r

    def blit(self, source, dest, area=None, special_flags=0):
        if isinstance(dest, Rect):
            dx, dy = dest.x, dest.y
        else:
            dx, dy = _trunc(dest[0]), _trunc(dest[1])
        sh, sw = source._a.shape[:2]
        dh, dw = self._a.shape[:2]
        x0, y0 = max(dx, 0), max(dy, 0)
        x1, y1 = min(dx + sw, dw), min(dy + sh, dh)
        if x1 <= x0 or y1 <= y0:
            return Rect(dx, dy, 0, 0)
        src = source._a[y0 - dy : y1 - dy, x0 - dx : x1 - dx]
        if source._alpha is None:
            self._a[y0:y1, x0:x1] = src
        else:
            # per-pixel alpha blit; only alpha 0 / 255 occur on the reference path
            al = source._alpha[y0 - dy : y1 - dy, x0 - dx : x1 - dx]
            opaque = al == 255
            mixed = (al > 0) & ~opaque
            if mixed.any():
                raise NotImplementedError("partial alpha blit is not restated")
            self._a[y0:y1, x0:x1][opaque] = src[opaque]
        return Rect(x0, y0, x1 - x0, y1 - y0)

    def subsurface(self, *rect):
        r = Rect(*rect)
        dh, dw = self._a.shape[:2]
        if r.x < 0 or r.y < 0 or r.x + r.w > dw or r.y + r.h > dh:
            raise ValueError("subsurface rectangle outside surface area")
        return Surface(
            None,
            _array=self._a[r.y : r.y + r.h, r.x : r.x + r.w],
            _alpha=None if self._alpha is None else self._alpha[r.y : r.y + r.h, r.x : r.x + r.w],
        )


class _DrawModule:
    @staticmethod
    def rect(surface, color, rect, width=0, *a, **k):
        if width != 0:
            raise NotImplementedError("outlined rects are not on the reference hot path")
        r = Rect(rect)
        if r.w < 0:
            r.x += r.w
            r.w = -r.w
        if r.h < 0:
            r.y += r.h
            r.h = -r.h
        return surface.fill(color, r)

    @staticmethod
    def polygon(surface, color, points, width=0):
        """Filled polygon, scan-line even-odd fill including edge pixels
        (restated from pygame 2.6.1 draw.c:draw_fillpoly)."""
        if width != 0:
            raise NotImplementedError
        rgb, al = _color3(color)
        xs = [int(p[0]) for p in points]
        ys = [int(p[1]) for p in points]
        n = len(points)
        h, w = surface._a.shape[:2]
        miny, maxy = min(ys), max(ys)

        def hline(y, x1, x2):
            if y < 0 or y >= h:
                return
            if x1 > x2:
                x1, x2 = x2, x1
            x1, x2 = max(x1, 0), min(x2, w - 1)
            if x2 < x1:
                return
            surface._a[y, x1 : x2 + 1] = rgb
            if surface._alpha is not None:
                surface._alpha[y, x1 : x2 + 1] = al

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
                if (y >= y1 and y < y2) or (y == maxy and y2 == maxy):
                    v = _np.float32((y - y1) * (x2 - x1)) / _np.float32(y2 - y1)
                    v = _math.floor(v) if len(inter) % 2 == 0 else _math.ceil(v)
                    inter.append(int(v) + x1)
            inter.sort()
            for i in range(0, len(inter) - 1, 2):
                hline(y, inter[i], inter[i + 1])
        # horizontal border edges
        for i in range(n):
            ip = i - 1 if i else n - 1
            if ys[ip] == ys[i] and miny < ys[i] < maxy:
                hline(ys[i], xs[ip], xs[i])


draw = _DrawModule()


def _rotate90(src: _np.ndarray, angle_int: int) -> _np.ndarray:
    # C: numturns = (angle / 90) % 4 with truncating division / remainder
    q = abs(angle_int) // 90
    if angle_int < 0:
        q = -q
    numturns = int(_math.fmod(q, 4))
    if numturns < 0:
        numturns += 4
    sh, sw = src.shape[:2]
    if numturns == 0:
        return src.copy()
    if numturns == 1:
        # dst(x, y) = src(sw-1-y, x); dst is sh wide, sw high
        return _np.ascontiguousarray(_np.transpose(src, (1, 0, 2))[::-1, :])
    if numturns == 2:
        return _np.ascontiguousarray(src[::-1, ::-1])
    # numturns == 3: dst(x, y) = src(y, sh-1-x)
    return _np.ascontiguousarray(_np.transpose(src, (1, 0, 2))[:, ::-1])


def _c_int(v: float) -> int:
    return int(v)  # (int)double truncates toward zero


class _TransformModule:
    @staticmethod
    def rotate(surface, angle):
        """pygame 2.6.1 transform.c:surf_rotate + rotate() (see module docstring)."""
        src = surface._a
        sh, sw = src.shape[:2]
        angle = float(_np.float32(angle))  # parsed with "f" -> C float
        if _math.fmod(angle, 90.0) == 0.0:
            return Surface(None, _array=_rotate90(src, _c_int(angle)))
        rad = angle * 0.01745329251994329
        sangle = _math.sin(rad)
        cangle = _math.cos(rad)
        cx, cy_ = cangle * sw, cangle * sh
        sx, sy = sangle * sw, sangle * sh
        nxmax = _c_int(max(abs(cx + sy), abs(cx - sy), abs(-cx + sy), abs(-cx - sy)))
        nymax = _c_int(max(abs(sx + cy_), abs(sx - cy_), abs(-sx + cy_), abs(-sx - cy_)))
        bg = src[0, 0].copy()
        dw, dh = nxmax, nymax
        cy = dh // 2
        xd = (sw - dw) << 15
        yd = (sh - dh) << 15
        isin = _c_int(sangle * 65536)
        icos = _c_int(cangle * 65536)
        ax = (dw << 15) - _c_int(cangle * ((dw - 1) << 15))
        ay = (dh << 15) - _c_int(sangle * ((dw - 1) << 15))
        xmaxval = (sw << 16) - 1
        ymaxval = (sh << 16) - 1
        ys = _np.arange(dh, dtype=_np.int64)[:, None]
        xs = _np.arange(dw, dtype=_np.int64)[None, :]
        dx = ax + isin * (cy - ys) + xd + xs * icos
        dy = ay - icos * (cy - ys) + yd + xs * isin
        # the C code runs in 32-bit ints; the magnitudes here never overflow
        assert _np.abs(dx).max() < 2**31 and _np.abs(dy).max() < 2**31
        oob = (dx < 0) | (dy < 0) | (dx > xmaxval) | (dy > ymaxval)
        sxi = _np.clip(dx >> 16, 0, sw - 1)
        syi = _np.clip(dy >> 16, 0, sh - 1)
        out = src[syi, sxi]
        out[oob] = bg
        return Surface(None, _array=_np.ascontiguousarray(out))


transform = _TransformModule()


class _SurfarrayModule:
    @staticmethod
    def pixels3d(surface):
        return _np.transpose(surface._a, (1, 0, 2))

    array3d = pixels3d


surfarray = _SurfarrayModule()


class _ImageModule:
    @staticmethod
    def load(path):
        from PIL import Image

        arr = _np.array(Image.open(path).convert("RGB"), dtype=_np.uint8)
        return Surface(None, _array=arr)


image = _ImageModule()


class _Sprite:
    def __init__(self, *groups):
        pass


class _SpriteModule:
    Sprite = _Sprite


sprite = _SpriteModule()


class _Anything:
    """Permissive stand-in for pygame sub-APIs that only the GUI / display use."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


display = _Anything()
time = _Anything()
event = _Anything()
font = _Anything()
key = _Anything()
mouse = _Anything()


def __getattr__(name):  # constants such as K_a, QUIT, MOUSEBUTTONDOWN used by the GUI
    if name.startswith("__"):
        raise AttributeError(name)
    return 0
