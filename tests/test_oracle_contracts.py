"""Third-party raster / wrapper arithmetic: what CAN be pinned here.

* cv2.resize(INTER_AREA) is installed: the oracle's area resize must match it bit-for-bit.
* The reference's own contracts for the pygame boundary (tools/validate_simulator_semantics.py:60-89, 366-414;
  tests/test_seeded_scene_consistency.py:128-138) are replayed against the oracle's direct raster.
* The literal surface formulation (oracle/shims/pygame, used to run the unmodified reference) and the direct
  per-pixel formulation (oracle/raster.py, the one the CUDA kernel follows) must agree pixel-for-pixel.
Exact pygame / gymnasium values remain "parity unpinned" (see oracle/__init__.py)."""
import math
import os
import sys

import numpy as np
import pytest

from golden_util import ROOT, load_map
from oracle import raster, sim


def test_area_resize_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(12):
        img = raster.PALETTE[rng.integers(0, 9, (128, 128))]
        # blocky structure so that every tap combination (uniform and mixed) appears
        img = np.repeat(np.repeat(img[::8, ::8], 8, axis=0), 8, axis=1)
        img = np.roll(img, rng.integers(0, 8), axis=rng.integers(0, 2))
        assert np.array_equal(raster.resize_area(img, (96, 96)), cv2.resize(img, (96, 96), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("size", [(128, 128), (64, 64), (32, 32), (16, 16), (84, 84), (100, 100), (112, 112), (72, 72),
                                  (48, 48), (80, 80), (84, 96), (64, 96), (96, 64), (64, 128), (128, 64), (120, 120),
                                  (127, 127), (90, 60), (42, 42), (8, 8)])
def test_area_resize_any_obs_size_matches_cv2(size):
    """EnvConfig.obs_size is free in the reference: copy / 2x2 integer path / area tables, each against cv2."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(size[0] * 131 + size[1])
    for k in range(4):
        if k % 2 == 0:
            img = raster.PALETTE[rng.integers(0, 10, (128, 128))]
        else:
            idx = np.ones((128, 128), dtype=np.int64)
            for _ in range(40):
                x, y, w, h = rng.integers(0, 120), rng.integers(0, 120), rng.integers(1, 40), rng.integers(1, 40)
                idx[y:y + h, x:x + w] = rng.integers(0, 10)
            img = raster.PALETTE[idx]
        want = cv2.resize(img, (size[1], size[0]), interpolation=cv2.INTER_AREA)
        assert np.array_equal(raster.resize_area(img, size), want)


@pytest.mark.parametrize("view,size", [(256, (96, 96)), (256, (128, 128)), (256, (84, 84)), (256, (64, 96)), (256, (100, 60)),
                                       (256, (256, 256)), (256, (200, 200)), (256, (64, 64)), (64, (24, 24)),
                                       (64, (96, 96)), (64, (84, 84)), (64, (128, 128)), (64, (64, 64)), (64, (32, 32)),
                                       (64, (48, 96)), (64, (96, 48)), (64, (72, 100)), (64, (65, 65))])
def test_resize_from_other_view_sizes_matches_cv2(view, size):
    """EnvConfig.size 256 shrinks by 8/3 (area tables); size 64 ENLARGES, where cv2's INTER_AREA is its
    8-bit bilinear kernel on area-mode coefficients (raster.resize_linear_area_mode), also when only one axis enlarges."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(view * 7 + size[0] * 131 + size[1])
    for k in range(4):
        if k % 2 == 0:
            img = raster.PALETTE[rng.integers(0, 10, (view, view))]
        else:
            idx = np.ones((view, view), dtype=np.int64)
            for _ in range(40):
                x, y, w, h = (rng.integers(0, view - 8), rng.integers(0, view - 8), rng.integers(1, view // 3),
                              rng.integers(1, view // 3))
                idx[y:y + h, x:x + w] = rng.integers(0, 10)
            img = raster.PALETTE[idx]
        want = cv2.resize(img, (size[1], size[0]), interpolation=cv2.INTER_AREA)
        assert np.array_equal(raster.resize_area(img, size), want)


def test_mask_coincidence_is_kept():
    # 1/2 white + 1/4 gray150 + 1/4 sidewalk220 = 220 = SIDEWALK: the blend path must report it
    img = np.zeros((128, 128, 3), np.uint8)
    img[:] = raster.PALETTE[raster.PAL_DRIVABLE]
    img[1, 1] = raster.PALETTE[raster.PAL_NON_DRIVABLE]
    img[2, 1] = raster.PALETTE[raster.PAL_SIDEWALK]
    small = raster.resize_area(img, (96, 96))
    m = raster.semantic_masks(small, "6-class")
    assert m.sum(axis=0).max() <= 2  # drivable+route may coincide, nothing else
    assert m.shape == (6, 96, 96) and m.dtype == np.float32


@pytest.mark.parametrize("anchor", [(0.5, 0.5), (0.5, 0.2), (0.5, 0.75)])
@pytest.mark.parametrize("yaw", [0.0, math.pi / 6, -math.pi / 4, -math.pi / 2, 2.5])
def test_anchor_pixel_is_hero_colour_and_crop_alignment(anchor, yaw):
    """validate_simulator_semantics.py:366-414 + tests/test_seeded_scene_consistency.py:128-138."""
    cls = load_map()
    geom = raster.FovGeometry(128, *anchor)
    x, y = 851.3, 949.6
    img = raster.render_fov(cls, geom, x, y, yaw, [])
    ax, ay = geom.anchor
    assert img[ay, ax] == raster.PAL_BLACK                       # hero colour (0, 0, 0) at the anchor
    H, W = cls.shape
    xmin, ymin = raster.crop_origin(x, y, geom, W, H)
    assert 0 <= xmin and xmin + geom.crop <= W + 2 * geom.pad    # crop strictly inside the render surface
    assert abs((xmin + geom.crop / 2) - (x + geom.pad)) <= 1.5   # ego within 1.5 px of the crop centre
    assert abs((ymin + geom.crop / 2) - (y + geom.pad)) <= 1.5


def test_bicycle_yaw_update_closed_form():
    """validate_simulator_semantics.py:60-89."""
    b = sim.Body()
    b.x = b.y = b.yaw = 0.0
    b.v = 5.0
    b.target = 20.0
    b.update(0.0, 0.2)
    assert math.isclose(b.yaw, (5.0 / 2.9) * math.tan(0.2) * 0.1, rel_tol=1e-12)


def test_carl_speed_penalty_monotone():
    """validate_simulator_semantics.py:122-183 (the reference's hand-built info fixture)."""
    cls = load_map()
    scene = {
        "ego_state0": np.array([1.0, 0.0, 0.0, 0.0]), "ego_target_speed": np.float64(100.0), "ego_tidx0": np.int32(0),
        "ego_cx": np.array([0.0, 2.0, 4.0, 6.0, 8.0]), "ego_cy": np.zeros(5), "ego_cyaw": np.zeros(5),
        "rew_rx": np.array([0, 10, 20], np.int32), "rew_ry": np.zeros(3, np.int32),
        "act_kind": np.zeros(0, np.uint8), "act_state0": np.zeros((0, 4)), "act_tidx0": np.zeros(0, np.int32),
        "act_cruise_px": np.zeros(0), "act_cruise_mps": np.zeros(0), "act_beh": np.zeros(0, np.uint8),
        "act_beh_p": np.zeros((0, 4)), "act_route_off": np.zeros(1, np.int32), "act_cx": np.zeros(0),
        "act_cy": np.zeros(0), "act_cyaw": np.zeros(0), "act_raw_off": np.zeros(1, np.int32),
        "act_raw_x": np.zeros(0), "act_raw_y": np.zeros(0),
    }
    s = sim.SceneSim(scene, cls, 182)
    ps = []
    for v in (10.0, 36.0, 80.0):
        s.s_prev = 0.0
        hero = dict(state=[1.0, 0.0, 0.0, v], dist2wp=0.0, next_wps=(scene["ego_cx"], scene["ego_cy"]))
        r, term, cause = s.carl_reward(sim.HIT_NONE, -1, [], sim.CLS_DRIVABLE, hero)
        over = max(v * sim.MPP - 35 / 3.6, 0.0)
        ps.append(1.0 if over <= 0 else max(0.1, math.exp(-over / 6.0)))
        assert not term and cause == sim.CAUSE_NONE
    assert ps[0] >= ps[1] >= ps[2] and len(set(round(p, 6) for p in ps)) > 1


def test_shim_surface_pipeline_equals_direct_raster():
    """Literal pygame-style pipeline (blit map, draw rects, subsurface, transform.rotate, blit, draw ego)
    vs the direct per-pixel formulation the CUDA kernel follows."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    try:
        import pygame
    finally:
        sys.path.pop(0)
    assert pygame.__doc__ and "pygame-lite" in pygame.__doc__
    cls = load_map()
    rng = np.random.default_rng(3)
    for anchor in ((0.5, 0.5), (0.5, 0.75)):
        geom = raster.FovGeometry(128, *anchor)
        pad, crop = geom.pad, geom.crop
        H, W = cls.shape
        scene = pygame.Surface((W + 2 * pad, H + 2 * pad))
        base = pygame.Surface(None, _array=raster.PALETTE[cls].copy())
        for _ in range(6):
            x, y = rng.uniform(60, W - 60), rng.uniform(60, H - 60)
            yaw = rng.uniform(-math.pi, math.pi) if rng.random() < 0.8 else rng.choice([0.0, -math.pi / 2, math.pi / 2])
            rects = []
            for _ in range(12):
                size = int(rng.choice([2, 4]))
                rects.append((int(x + pad + rng.integers(-70, 70)), int(y + pad + rng.integers(-70, 70)), size, size,
                              int(rng.choice([raster.PAL_VEHICLE, raster.PAL_PEDESTRIAN, raster.PAL_ROUTE]))))
            scene.fill(tuple(raster.PALETTE[raster.PAL_NON_DRIVABLE]))
            scene.blit(base, (pad, pad))
            for rx, ry, rw, rh, pal in rects:
                pygame.draw.rect(scene, tuple(int(v) for v in raster.PALETTE[pal]), pygame.Rect(rx, ry, rw, rh))
            xmin, ymin = raster.crop_origin(x, y, geom, W, H)
            sub = scene.subsurface(pygame.Rect(xmin, ymin, crop, crop))
            rot = pygame.transform.rotate(sub, math.degrees(yaw) + 90)
            out = pygame.Surface((128, 128))
            out.fill((0, 0, 0))
            out.blit(rot, rot.get_rect(center=geom.anchor))
            ego = pygame.Rect(0, 0, 4, 4)
            ego.center = geom.anchor
            pygame.draw.rect(out, (0, 0, 0), ego)
            direct = raster.PALETTE[raster.render_fov(cls, geom, x, y, yaw, rects)]
            assert np.array_equal(out._a, direct)
