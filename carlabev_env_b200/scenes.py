"""Host-side scene generation for the scripted scenarios -> scene-pool entries.

The device never generates scenes: resets draw from a device-resident pool.
This module rebuilds, on the host, exactly the state the reference reaches at the end of
`CarlaBEV.reset` for the closed-form scenarios:

  * seeding        -- src/randomness.py:13-65 (sha256-derived sub-seeds, RNGBundle)
  * lead_brake     -- src/scenes/scenarios/lead_brake.py:18-129
  * jaywalk        -- src/scenes/scenarios/jaywalk.py:29-117
  * reset pipeline -- envs/carlabev.py:96-148 (retry loop, spawn validation),
                      scenes/scene.py:61-88,142-196 (int32 ego route, hero, targets),
                      control/stanley_controller.py:34-49 + control/utils.py:200-269 (smoothing, jitter),
                      actors/actor.py:86-108 (actor controllers start at cruise speed)

`rdm` and `red_light_runner` scenes need the reference's lane graphs (networkx pickles); they are
exported from the reference by oracle/gen_golden.py / tools and shipped as packed pool files.
"""
from __future__ import annotations

import hashlib
import random

import numpy as np

from .pool import empty_scene

MPP = 40.0 / 128.0  # envs/geometry.py:6-10
WHEEL_BASE = 2.9
BEH_NONE, BEH_LEAD_BRAKE, BEH_CROSS, BEH_STOP_MID, BEH_STOP_RETURN = 0, 1, 2, 3, 4
KIND_IDS = {"rdm": 0, "lead_brake": 1, "jaywalk": 2, "red_light_runner": 3}
_SEED_MODULUS = 2**31 - 1


def derive_seed(base_seed: int, *parts) -> int:
    """randomness.py:13-16."""
    token = ":".join([str(int(base_seed)), *(str(p) for p in parts)])
    return int(hashlib.sha256(token.encode("utf-8")).hexdigest()[:16], 16) % _SEED_MODULUS


class RNGBundle:
    """randomness.py:35-65 (the streams the scripted scenarios consume)."""

    def __init__(self, scene_seed, route_seed=None, traffic_seed=None, scenario_seed=None):
        self.scene_seed = int(scene_seed)
        self.route_seed = derive_seed(scene_seed, "route") if route_seed is None else int(route_seed)
        self.traffic_seed = derive_seed(scene_seed, "traffic") if traffic_seed is None else int(traffic_seed)
        self.scenario_seed = derive_seed(scene_seed, "scenario") if scenario_seed is None else int(scenario_seed)
        self.scenario_rng = random.Random(self.scenario_seed)
        self.route_np_rng = np.random.default_rng(self.route_seed)
        self.scenario_np_rng = np.random.default_rng(self.scenario_seed)


def m2s(d: float) -> float:
    """distance_meters_to_surface, envs/geometry.py:49-50."""
    return float(d) / MPP


def smooth_and_compute(ax, ay, window=11, poly=3):
    """control/utils.py:200-269 -> (cx, cy, cyaw)."""
    from scipy.signal import savgol_filter

    ax = np.asarray(ax, dtype=float)
    ay = np.asarray(ay, dtype=float)
    d = np.hypot(np.diff(ax), np.diff(ay))
    keep = np.concatenate(([True], d > 1e-9))
    ax, ay = ax[keep], ay[keep]
    if len(ax) < 2:
        ax = np.array([ax[0], ax[0] + 1e-3])
        ay = np.array([ay[0], ay[0]])
    if window % 2 == 0:
        window += 1
    if window > len(ax):
        window = len(ax) if len(ax) % 2 == 1 else len(ax) - 1
    if window < 3:
        window = 3
    poly = min(poly, window - 1)
    if len(ax) >= window:
        cx = savgol_filter(ax, window_length=window, polyorder=poly)
        cy = savgol_filter(ay, window_length=window, polyorder=poly)
    else:
        cx, cy = ax.copy(), ay.copy()
    s = np.concatenate(([0.0], np.cumsum(np.hypot(np.diff(cx), np.diff(cy)))))
    if s[-1] <= 1e-9:
        return cx, cy, np.zeros_like(cx)
    cyaw = np.unwrap(np.arctan2(np.gradient(cy, s), np.gradient(cx, s)))
    return cx, cy, cyaw


def _nearest(x, y, yaw, cx, cy):
    fx = x + WHEEL_BASE * np.cos(yaw)
    fy = y + WHEEL_BASE * np.sin(yaw)
    return int(np.argmin(np.hypot(fx - cx, fy - cy)))


def controller_init(rx, ry, v0, np_rng):
    """Controller.set_route with jitter (stanley_controller.py:34-49): returns (state0, tidx0, cx, cy, cyaw)."""
    cx, cy, cyaw = smooth_and_compute(rx, ry, window=11, poly=3)
    x = cx[0] + int(np_rng.integers(-1, 2))
    y = cy[0] + int(np_rng.integers(-1, 2))
    tidx = _nearest(x, y, 0.0, cx, cy)  # State() starts with yaw = 0.0
    yaw = cyaw[tidx]
    return np.array([x, y, yaw, v0], dtype=np.float64), tidx, cx, cy, cyaw


def _round_half_even(v: float) -> int:
    return int(round(float(v)))


def _rect_left(c, pad, size):
    return _round_half_even(float(pad) + float(c)) - (size >> 1)


class _ActorSpec:
    def __init__(self, kind, rx, ry, speed_mps, beh=BEH_NONE, beh_p=(0.0, 0.0, 0.0, 0.0)):
        self.kind, self.rx, self.ry = kind, list(rx), list(ry)
        self.speed_mps = max(0.0, float(speed_mps))  # set_cruise_speed_mps, actor.py:135-138
        self.beh, self.beh_p = beh, tuple(float(v) for v in beh_p)


def sample_lead_brake(level, np_rng):
    """LeadBrakeScenario.sample, lead_brake.py:18-129 (draw order preserved)."""
    ego_start_y = int(np_rng.integers(900, 1000))
    lead_gap_m = float(np_rng.uniform(4.5, 12.5))
    ego_speed = float(np_rng.uniform(8.0, 16.0))
    lead_speed = ego_speed + float(np_rng.uniform(-2.0, 2.0))
    brake_delay = float(np_rng.uniform(1.5, 4.0))
    brake_strength = float(np_rng.uniform(2.0, 6.0))
    x_center = 850
    lane_width = m2s(2.2)
    ego_step, lead_step, rear_step = m2s(6.25), m2s(1.56), m2s(3.12)
    ego_rx = [x_center] * 6
    ego_ry = [ego_start_y - i * ego_step for i in range(6)]
    lead_ry_start = ego_ry[0] - m2s(lead_gap_m)
    actors = [_ActorSpec(0, [x_center - 1] * 6, [lead_ry_start - i * lead_step for i in range(6)], lead_speed,
                         BEH_LEAD_BRAKE, (brake_delay, brake_strength, 0, 0))]
    if level >= 2:
        left_rx = [x_center - lane_width] * 7
        left_ry = [ego_start_y - i * 20 for i in range(7)]
        left_rx.reverse()
        left_ry.reverse()
        left_speed = float(np_rng.uniform(10.0, 18.0))
        actors.append(_ActorSpec(0, left_rx, left_ry, left_speed))
    if level >= 3:
        rear_gap_m = float(np_rng.uniform(3.0, 6.0))
        rear_ry_start = ego_ry[0] + m2s(rear_gap_m)
        rear_ry = [rear_ry_start - i * rear_step for i in range(6)]
        rear_speed = max(ego_speed - float(np_rng.uniform(1.0, 3.0)), 4.0)
        rear_delay = float(np_rng.uniform(2.0, 5.0))
        actors.append(_ActorSpec(0, [x_center] * 6, rear_ry, rear_speed, BEH_LEAD_BRAKE,
                                 (rear_delay, brake_strength, 0, 0)))
    return (ego_rx, ego_ry, ego_speed, ego_speed), actors


def sample_jaywalk(level, np_rng):
    """JaywalkScenario.sample, jaywalk.py:29-117 (draw order preserved)."""
    ego_start_y = int(np_rng.integers(900, 1000))
    ego_speed = float(np_rng.uniform(8.0, 14.0))
    ped_x_base = 850
    lane_width = m2s(1.6)
    cross_offset_m = float(np_rng.uniform(-3.0, 3.0))
    cross_delay = float(np_rng.uniform(1.0, 2.5))
    pedestrian_speed = float(np_rng.uniform(1.2, 2.2))
    ego_step, rear_step = m2s(6.25), m2s(3.12)
    yield_duration = float(np_rng.uniform(0.8, 1.6))
    ego_rx = [ped_x_base] * 6
    ego_ry = [ego_start_y - i * ego_step for i in range(6)]
    cross_offset = m2s(cross_offset_m)
    ped_start_x = ped_x_base + lane_width + cross_offset
    ped_end_x = ped_x_base - lane_width + cross_offset
    ped_y = ego_ry[2] + m2s(float(np_rng.uniform(-1.0, 1.6)))
    ped_rx = np.linspace(ped_start_x, ped_end_x, 8)
    ped_ry = np.ones_like(ped_rx) * ped_y
    if level == 1:
        beh, p = BEH_CROSS, (cross_delay, 2.0, 0.0, 0.0)
    elif level == 2:
        beh, p = BEH_STOP_MID, (cross_delay, 0.5, -1.0, 0.0)
    else:
        beh, p = BEH_STOP_RETURN, (cross_delay, 1.0 / 3.0, yield_duration, 1.0)
    peds = [_ActorSpec(1, ped_rx, ped_ry, pedestrian_speed, beh, p)]
    vehicles = []
    if level >= 4:
        rear_gap_m = float(np_rng.uniform(3.0, 6.0))
        rear_ry_start = ego_ry[0] + m2s(rear_gap_m)
        rear_ry = [rear_ry_start - i * rear_step for i in range(6)]
        rear_speed = max(ego_speed - float(np_rng.uniform(1.0, 3.0)), 4.0)
        vehicles.append(_ActorSpec(0, [ped_x_base] * 6, rear_ry, rear_speed))
    return (ego_rx, ego_ry, ego_speed, ego_speed), vehicles + peds


_SAMPLERS = {"lead_brake": sample_lead_brake, "jaywalk": sample_jaywalk}


def _route_length_m(rx, ry) -> float:
    """route_length_meters, envs/geometry.py:61-69."""
    total = 0.0
    for i in range(1, len(rx)):
        total += np.hypot(float(rx[i]) - float(rx[i - 1]), float(ry[i]) - float(ry[i - 1]))
    return float(total) * MPP


def build_scripted_scene(kind: str, scene_seed: int, level: int | None = None, cls_map=None, pad: int = 182,
                         max_reset_attempts: int = 10) -> dict:
    """CarlaBEV.reset for scene in {"lead_brake", "jaywalk"} -> pool entry.

    `cls_map` (H, W) uint8 classes enables the reference's spawn validation / retry loop
    (carlabev.py:108-131, scene.py:142-170); without it the first sample is accepted."""
    if kind not in _SAMPLERS:
        raise KeyError(f"Unknown scenario '{kind}'")
    bundle = RNGBundle(scene_seed)
    last = None
    for _ in range(max_reset_attempts):
        lvl = level
        if lvl is None:
            lvl = bundle.scenario_rng.choice([1, 2, 3, 4])  # scene_generator.py:171-176
        agent, specs = _SAMPLERS[kind](int(lvl), bundle.scenario_np_rng)
        ego_rx, ego_ry, init_mps, target_mps = agent
        len_route = _route_length_m(ego_rx, ego_ry)              # compute_total_dist_m
        rx_i = np.array(ego_rx, dtype=np.int32)                  # scene.py:192-193: truncation to int32
        ry_i = np.array(ego_ry, dtype=np.int32)
        s = empty_scene()
        v0 = float(init_mps) / MPP
        st0, _, cx, cy, cyaw = controller_init(rx_i, ry_i, v0, bundle.route_np_rng)
        # BaseAgent.__init__: second stanley_control() with the route yaw (hero.py:84-86)
        t0 = _nearest(st0[0], st0[1], 0.0, cx, cy)
        t1 = _nearest(st0[0], st0[1], st0[2], cx, cy)
        tidx = t0 if t0 >= t1 else t1
        s.update(ego_state0=st0, ego_target_speed=np.float64(float(target_mps) / MPP), ego_tidx0=np.int32(tidx),
                 ego_cx=cx, ego_cy=cy, ego_cyaw=cyaw, rew_rx=rx_i, rew_ry=ry_i,
                 route_length_m=np.float64(_route_length_m(rx_i, ry_i)), len_ego_route=np.float64(len_route),
                 num_vehicles=np.int32(sum(1 for a in specs if a.kind == 0)), kind=np.int32(KIND_IDS[kind]),
                 level=np.int32(lvl), seed=np.int64(scene_seed))
        # ActorManager.reset_all: vehicles then pedestrians, each Controller.set_route(v0=cruise) draws 2 jitter ints
        st, ti, rcx, rcy, rcyaw, roff, wx, wy, woff = [], [], [], [], [], [0], [], [], [0]
        for a in specs:
            cruise_px = a.speed_mps / MPP
            a_st, a_t, acx, acy, acyaw = controller_init(a.rx, a.ry, cruise_px, bundle.scenario_np_rng)
            st.append(a_st)
            ti.append(a_t)
            rcx.append(acx), rcy.append(acy), rcyaw.append(acyaw)
            roff.append(roff[-1] + len(acx))
            wx.append(np.asarray(a.rx, dtype=np.float64)), wy.append(np.asarray(a.ry, dtype=np.float64))
            woff.append(woff[-1] + len(a.rx))
        n = len(specs)
        s.update(act_kind=np.array([a.kind for a in specs], dtype=np.uint8),
                 act_state0=np.array(st, dtype=np.float64).reshape(n, 4), act_tidx0=np.array(ti, dtype=np.int32),
                 act_cruise_px=np.array([a.speed_mps / MPP for a in specs], dtype=np.float64),
                 act_cruise_mps=np.array([a.speed_mps for a in specs], dtype=np.float64),
                 act_beh=np.array([a.beh for a in specs], dtype=np.uint8),
                 act_beh_p=np.array([a.beh_p for a in specs], dtype=np.float64).reshape(n, 4),
                 act_cx=np.concatenate(rcx), act_cy=np.concatenate(rcy), act_cyaw=np.concatenate(rcyaw),
                 act_route_off=np.array(roff, dtype=np.int32), act_raw_x=np.concatenate(wx),
                 act_raw_y=np.concatenate(wy), act_raw_off=np.array(woff, dtype=np.int32))
        last = s
        if cls_map is None or _spawn_valid(s, cls_map, pad):
            return s
    raise RuntimeError(f"Failed to reset into a valid initial state after {max_reset_attempts} attempts "
                       f"(kind={kind}, seed={scene_seed}, last level={int(last['level'])})")


def _spawn_valid(s, cls_map, pad) -> bool:
    """Scene.spawn_validation_info, scene.py:142-170."""
    h, w = cls_map.shape
    x, y = float(s["ego_state0"][0]), float(s["ego_state0"][1])
    tx = int(np.clip(_round_half_even(x), 0, w - 1))
    ty = int(np.clip(_round_half_even(y), 0, h - 1))
    if cls_map[ty, tx] == 0:
        return False
    hx, hy = _rect_left(x, pad, 4), _rect_left(y, pad, 4)
    for i, kind in enumerate(s["act_kind"]):
        size = 4 if kind == 0 else 2
        ax, ay = _rect_left(s["act_state0"][i, 0], pad, size), _rect_left(s["act_state0"][i, 1], pad, size)
        if hx < ax + size and hy < ay + size and hx + 4 > ax and hy + 4 > ay:
            return False
    return True


def build_scripted_pool(kind_level_seed, cls_map=None, pad: int = 182) -> list[dict]:
    """[(kind, level | None, scene_seed), ...] -> list of pool entries."""
    return [build_scripted_scene(k, seed, level=lv, cls_map=cls_map, pad=pad) for k, lv, seed in kind_level_seed]
