"""Contract of oracle/shims against the REAL third-party libraries: the goldens are recorded from the
unmodified reference running on our pygame-lite / gymnasium-lite, so the bit-exact frame and mask claim is exact
relative to the shims.  Neither pygame 2.6 nor gymnasium 1.x is installable in the build image or on the GPU box (no
wheels, no network): every test here SKIPS there.  On any machine where they import, run

    python -m pytest tests/test_shim_contracts.py -q

first, and re-record the goldens (oracle/gen_golden.py) if it fails.  The shim is loaded under a private module name
so that it never shadows the real package."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_shim(pkg):
    path = os.path.join(ROOT, "oracle", "shims", pkg, "__init__.py")
    spec = importlib.util.spec_from_file_location(f"_cbev_shim_{pkg}", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def _real(pkg):
    """The real package, or skip: a module that resolves into oracle/shims is our own stand-in."""
    shims = os.path.join(ROOT, "oracle", "shims")
    saved = [p for p in sys.path if os.path.abspath(p) == shims]
    for p in saved:
        sys.path.remove(p)
    try:
        mod = pytest.importorskip(pkg)
    finally:
        sys.path[:0] = saved
    if os.path.abspath(getattr(mod, "__file__", "") or "").startswith(shims):
        pytest.skip(f"only the shim of {pkg} is importable here")
    return mod


def test_rect_semantics_match_pygame():
    pg, sh = _real("pygame"), _load_shim("pygame")
    rng = np.random.default_rng(0)
    for _ in range(2000):
        a = rng.uniform(-50, 400, 4) * np.array([1, 1, 0.1, 0.1])
        b = rng.integers(-50, 400, 4) // np.array([1, 1, 10, 10])
        ra, sa = pg.Rect(*map(float, a)), sh.Rect(*map(float, a))       # float -> int truncation (traffic_light.py:65-69)
        rb, sb = pg.Rect(*map(int, b)), sh.Rect(*map(int, b))
        assert tuple(ra) == tuple(sa) and tuple(rb) == tuple(sb)
        assert bool(ra.colliderect(rb)) == bool(sa.colliderect(sb))
        c = tuple(int(v) for v in rng.integers(-20, 500, 2))
        ra.center, sa.center = c, c                                      # Rect.center setter (transforms.py:46-51)
        assert tuple(ra) == tuple(sa) and tuple(ra.center) == tuple(sa.center)


def test_draw_rect_fill_blit_subsurface_match_pygame():
    pg, sh = _real("pygame"), _load_shim("pygame")
    rng = np.random.default_rng(1)
    for _ in range(200):
        w, h = (int(v) for v in rng.integers(8, 64, 2))
        s1, s2 = pg.Surface((w, h)), sh.Surface((w, h))
        base = tuple(int(v) for v in rng.integers(0, 256, 3))
        s1.fill(base), s2.fill(base)
        for _ in range(6):
            r = tuple(int(v) for v in rng.integers(-8, 70, 2)) + tuple(int(v) for v in rng.integers(0, 12, 2))
            col = tuple(int(v) for v in rng.integers(0, 256, 3))
            pg.draw.rect(s1, col, pg.Rect(*r)), sh.draw.rect(s2, col, sh.Rect(*r))
        a1 = np.array(pg.surfarray.pixels3d(s1))
        a2 = np.array(sh.surfarray.pixels3d(s2))
        assert np.array_equal(a1, a2)
        dst1, dst2 = pg.Surface((40, 40)), sh.Surface((40, 40))
        at = tuple(int(v) for v in rng.integers(-30, 40, 2))
        dst1.blit(s1, at), dst2.blit(s2, at)                             # clipped blit (fov.py:82-94)
        assert np.array_equal(np.array(pg.surfarray.pixels3d(dst1)), np.array(sh.surfarray.pixels3d(dst2)))


@pytest.mark.parametrize("size", [(182, 182), (230, 230), (33, 57)])
def test_transform_rotate_matches_pygame(size):
    pg, sh = _real("pygame"), _load_shim("pygame")
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size + (3,), dtype=np.uint8)
    s1, s2 = pg.Surface(size), sh.Surface(size)
    pg.surfarray.pixels3d(s1)[...] = img
    sh.surfarray.pixels3d(s2)[...] = img
    angles = list(rng.uniform(-720, 720, 200)) + [0.0, 90.0, 180.0, 270.0, -90.0, 45.0, 89.99999, 90.00001, 1e-7]
    angles += [float(np.degrees(y) + 90.0) for y in rng.uniform(-np.pi, np.pi, 100)]   # fov.py:86: degrees(yaw) + 90
    for ang in angles:
        r1, r2 = pg.transform.rotate(s1, ang), sh.transform.rotate(s2, ang)
        assert r1.get_size() == r2.get_size(), ang
        assert np.array_equal(np.array(pg.surfarray.pixels3d(r1)), np.array(sh.surfarray.pixels3d(r2))), ang


def test_gymnasium_wrappers_match():
    gym = _real("gymnasium")
    sh = _load_shim("gymnasium")
    import gymnasium.wrappers as W

    shw = importlib.import_module("_cbev_shim_gymnasium.wrappers")

    class Toy(gym.Env):
        observation_space = gym.spaces.Box(0, 255, (128, 128, 3), np.uint8)
        action_space = gym.spaces.Discrete(3)

        def __init__(self):
            self.rng = np.random.default_rng(0)

        def _obs(self):
            return self.rng.integers(0, 256, (128, 128, 3), dtype=np.uint8)

        def reset(self, *, seed=None, options=None):
            return self._obs(), {}

        def step(self, a):
            return self._obs(), 1.0, False, False, {}

    class ToyShim(Toy, sh.Env):
        observation_space = sh.spaces.Box(0, 255, (128, 128, 3), np.uint8)
        action_space = sh.spaces.Discrete(3)

    e1 = W.FrameStackObservation(W.GrayscaleObservation(W.ResizeObservation(Toy(), (96, 96))), 4)
    e2 = shw.FrameStackObservation(shw.GrayscaleObservation(shw.ResizeObservation(ToyShim(), (96, 96))), 4)
    o1, _ = e1.reset()
    o2, _ = e2.reset()
    assert np.array_equal(np.asarray(o1), np.asarray(o2))
    for _ in range(9):
        o1, *_ = e1.step(0)
        o2, *_ = e2.step(0)
        assert np.array_equal(np.asarray(o1), np.asarray(o2))
