"""Host-side scene generation -> scene-pool entries.

The device never generates scenes: resets draw from a device-resident pool.
This module rebuilds, on the host, exactly the state the reference reaches at the end of
`CarlaBEV.reset` for every generated scene kind:

  * seeding        -- src/randomness.py:13-65 (sha256-derived sub-seeds, RNGBundle)
  * lead_brake     -- src/scenes/scenarios/lead_brake.py:18-129
  * jaywalk        -- src/scenes/scenarios/jaywalk.py:29-117
  * reset pipeline -- envs/carlabev.py:96-148 (retry loop, spawn validation),
                      scenes/scene.py:61-88,142-196 (int32 ego route, hero, targets),
                      control/stanley_controller.py:34-49 + control/utils.py:200-269 (smoothing, jitter),
                      actors/actor.py:86-108 (actor controllers start at cruise speed)

  * rdm            -- managers/scene_generator.py:196-344, scenes/utils.py:75-214 (random routes on the lane
                      graphs; lanegraph.py holds the graphs and the shortest-path search)
  * red_light_runner -- src/scenes/scenarios/red_light_running.py:13-245
  * actor start jitter draws from a deep copy of the bundle's generator (managers/actor_manager.py:88-95)
"""
from __future__ import annotations

import copy
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
    """randomness.py:35-65 (the streams scene generation consumes)."""

    def __init__(self, scene_seed, route_seed=None, traffic_seed=None, scenario_seed=None):
        self.scene_seed = int(scene_seed)
        self.route_seed = derive_seed(scene_seed, "route") if route_seed is None else int(route_seed)
        self.traffic_seed = derive_seed(scene_seed, "traffic") if traffic_seed is None else int(traffic_seed)
        self.scenario_seed = derive_seed(scene_seed, "scenario") if scenario_seed is None else int(scenario_seed)
        self.route_rng = random.Random(self.route_seed)
        self.traffic_rng = random.Random(self.traffic_seed)
        self.scenario_rng = random.Random(self.scenario_seed)
        self.traffic_np_rng = np.random.default_rng(self.traffic_seed)
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


def sample_lead_brake(level, np_rng, **kw):
    """LeadBrakeScenario.sample, lead_brake.py:18-129 (draw order preserved).  Explicit parameters (`ego_speed`,
    `lead_gap`, `anchor_y`, ... -- the scenario-preset / scenario-config fields, scenarios/specs.py:60-75) replace
    the drawn value AFTER the draw: the reference evaluates `kwargs.get(key, <draw>)`, so the stream advances
    either way."""
    def pick(key, drawn):
        return kw[key] if key in kw else drawn

    ego_start_y = pick("anchor_y", int(np_rng.integers(900, 1000)))
    lead_gap_m = pick("lead_gap", float(np_rng.uniform(4.5, 12.5)))
    ego_speed = pick("ego_speed", float(np_rng.uniform(8.0, 16.0)))
    lead_speed = pick("lead_speed", ego_speed + float(np_rng.uniform(-2.0, 2.0)))
    brake_delay = pick("brake_delay", float(np_rng.uniform(1.5, 4.0)))
    brake_strength = pick("brake_strength", float(np_rng.uniform(2.0, 6.0)))
    x_center = pick("anchor_x", 850)
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
        left_speed = pick("left_speed", float(np_rng.uniform(10.0, 18.0)))
        actors.append(_ActorSpec(0, left_rx, left_ry, left_speed))
    if level >= 3:
        rear_gap_m = pick("rear_gap", float(np_rng.uniform(3.0, 6.0)))
        rear_ry_start = ego_ry[0] + m2s(rear_gap_m)
        rear_ry = [rear_ry_start - i * rear_step for i in range(6)]
        rear_speed = pick("rear_speed", max(ego_speed - float(np_rng.uniform(1.0, 3.0)), 4.0))
        rear_delay = pick("rear_brake_delay", float(np_rng.uniform(2.0, 5.0)))
        actors.append(_ActorSpec(0, [x_center] * 6, rear_ry, rear_speed, BEH_LEAD_BRAKE,
                                 (rear_delay, brake_strength, 0, 0)))
    return (ego_rx, ego_ry, ego_speed, ego_speed), actors


def sample_jaywalk(level, np_rng, **kw):
    """JaywalkScenario.sample, jaywalk.py:29-117 (draw order preserved; overrides as in sample_lead_brake)."""
    def pick(key, drawn):
        return kw[key] if key in kw else drawn

    ego_start_y = pick("anchor_y", int(np_rng.integers(900, 1000)))
    ego_speed = pick("ego_speed", float(np_rng.uniform(8.0, 14.0)))
    ped_x_base = pick("anchor_x", 850)
    lane_width = m2s(1.6)
    cross_offset_m = pick("cross_offset", float(np_rng.uniform(-3.0, 3.0)))
    cross_delay = pick("cross_delay", float(np_rng.uniform(1.0, 2.5)))
    pedestrian_speed = pick("pedestrian_speed", float(np_rng.uniform(1.2, 2.2)))
    ego_step, rear_step = m2s(6.25), m2s(3.12)
    yield_duration = pick("yield_duration", float(np_rng.uniform(0.8, 1.6)))
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
        rear_gap_m = pick("rear_gap", float(np_rng.uniform(3.0, 6.0)))
        rear_ry_start = ego_ry[0] + m2s(rear_gap_m)
        rear_ry = [rear_ry_start - i * rear_step for i in range(6)]
        rear_speed = pick("rear_speed", max(ego_speed - float(np_rng.uniform(1.0, 3.0)), 4.0))
        vehicles.append(_ActorSpec(0, [ped_x_base] * 6, rear_ry, rear_speed))
    return (ego_rx, ego_ry, ego_speed, ego_speed), vehicles + peds


_SAMPLERS = {"lead_brake": sample_lead_brake, "jaywalk": sample_jaywalk}


def _route_length_m(rx, ry) -> float:
    """route_length_meters, envs/geometry.py:61-69."""
    total = 0.0
    for i in range(1, len(rx)):
        total += np.hypot(float(rx[i]) - float(rx[i - 1]), float(ry[i]) - float(ry[i - 1]))
    return float(total) * MPP


def route_profile_metrics(ax, ay, turn_rate_thresh=0.12, min_turn_segment_m=4.0) -> dict:
    """compute_route_profile_metrics (src/control/route_profile.py:57-159): turning rate of the smoothed route
    [rad/m] labelled left (+) / right (-) / straight, turn segments of at least 4 m, profile name."""
    straight = dict(straight_fraction=1.0, left_turn_fraction=0.0, right_turn_fraction=0.0, turn_count=0,
                    has_left_turn=False, has_right_turn=False, intersection_like=False, route_profile="mostly_straight")
    cx, cy, cyaw = smooth_and_compute(ax, ay, window=11, poly=3)
    cx, cy = np.asarray(cx, dtype=float), np.asarray(cy, dtype=float)
    cyaw = np.unwrap(np.asarray(cyaw, dtype=float))
    if cx.size < 2 or cy.size < 2 or cyaw.size < 2:
        return straight
    ds_m = np.hypot(np.diff(cx), np.diff(cy)) * MPP
    valid = ds_m > 1e-6
    if not np.any(valid):
        return straight
    dtheta = np.diff(cyaw)
    dtheta = (dtheta + np.pi) % (2.0 * np.pi) - np.pi
    ds_valid = ds_m[valid]
    turn_rate = dtheta[valid] / ds_valid
    labels = np.where(turn_rate > turn_rate_thresh, 1, np.where(turn_rate < -turn_rate_thresh, -1, 0))
    total = float(ds_valid.sum())
    if total <= 1e-9:
        return straight
    segments, cur_sign, cur_len = [], 0, 0.0       # _normalize_turn_segments (:21-54)
    for sign, seg in zip(labels, ds_valid):
        sign = int(sign)
        if sign == 0:
            if cur_sign != 0 and cur_len >= min_turn_segment_m:
                segments.append(cur_sign)
            cur_sign, cur_len = 0, 0.0
        elif sign == cur_sign:
            cur_len += float(seg)
        else:
            if cur_sign != 0 and cur_len >= min_turn_segment_m:
                segments.append(cur_sign)
            cur_sign, cur_len = sign, float(seg)
    if cur_sign != 0 and cur_len >= min_turn_segment_m:
        segments.append(cur_sign)
    turn_count = len(segments)
    has_left, has_right = any(sg > 0 for sg in segments), any(sg < 0 for sg in segments)
    sf = float(ds_valid[labels == 0].sum()) / total
    lf = float(ds_valid[labels == 1].sum()) / total
    rf = float(ds_valid[labels == -1].sum()) / total
    if turn_count == 0 or sf >= 0.9:
        name = "mostly_straight"
    elif turn_count == 1 and lf >= rf:
        name = "single_left"
    elif turn_count == 1 and rf > lf:
        name = "single_right"
    elif turn_count >= 2:
        name = "multi_turn"
    else:
        name = "mixed"
    return dict(straight_fraction=sf, left_turn_fraction=lf, right_turn_fraction=rf, turn_count=turn_count,
                has_left_turn=has_left, has_right_turn=has_right,
                intersection_like=turn_count >= 2 or (has_left and has_right), route_profile=name)


def matches_route_profile(m, route_profile=None, min_turns=None, max_turns=None, intersection_required=None) -> bool:
    """src/control/route_profile.py:162-182."""
    if route_profile is not None and route_profile != "any" and m.get("route_profile") != route_profile:
        return False
    turns = int(m.get("turn_count", 0))
    if min_turns is not None and turns < min_turns:
        return False
    if max_turns is not None and turns > max_turns:
        return False
    if intersection_required is True and not bool(m.get("intersection_like", False)):
        return False
    if intersection_required is False and bool(m.get("intersection_like", False)):
        return False
    return True


def _assemble(agent, specs, len_route, hero_np_rng, actor_np_rngs, kind, level, scene_seed, lights=()) -> dict:
    """Scene.load_scene (scenes/scene.py:61-88) on a sampled actor dict -> pool entry: int32 ego route, hero spawn
    with its start jitter, then ActorManager.reset_all (vehicles, then pedestrians; each Controller.set_route draws
    two jitter integers from that actor's generator)."""
    ego_rx, ego_ry, init_mps, target_mps = agent
    rx_i = np.array(ego_rx, dtype=np.int32)                  # scene.py:192-193: truncation to int32
    ry_i = np.array(ego_ry, dtype=np.int32)
    s = empty_scene()
    v0 = float(init_mps) / MPP
    st0, _, cx, cy, cyaw = controller_init(rx_i, ry_i, v0, hero_np_rng)
    # BaseAgent.__init__: second stanley_control() with the route yaw (hero.py:84-86)
    t0 = _nearest(st0[0], st0[1], 0.0, cx, cy)
    t1 = _nearest(st0[0], st0[1], st0[2], cx, cy)
    tidx = t0 if t0 >= t1 else t1
    s.update(ego_state0=st0, ego_target_speed=np.float64(float(target_mps) / MPP), ego_tidx0=np.int32(tidx),
             ego_cx=cx, ego_cy=cy, ego_cyaw=cyaw, rew_rx=rx_i, rew_ry=ry_i,
             route_length_m=np.float64(_route_length_m(rx_i, ry_i)), len_ego_route=np.float64(len_route),
             num_vehicles=np.int32(sum(1 for a in specs if a.kind == 0)), kind=np.int32(KIND_IDS[kind]),
             level=np.int32(level), seed=np.int64(scene_seed))
    st, ti, rcx, rcy, rcyaw, roff, wx, wy, woff = [], [], [], [], [], [0], [], [], [0]
    # ActorManager.load deep-copies the actor dict (actor_manager.py:88-95): every actor of one load shares ONE copy
    # of its generator, so start jitter never advances the bundle's own stream (a retry re-draws the same jitter).
    copies = {id(g): copy.deepcopy(g) for g in actor_np_rngs}
    for a, np_rng in zip(specs, (copies[id(g)] for g in actor_np_rngs)):
        cruise_px = a.speed_mps / MPP
        a_st, a_t, acx, acy, acyaw = controller_init(a.rx, a.ry, cruise_px, np_rng)
        st.append(a_st)
        ti.append(a_t)
        rcx.append(acx), rcy.append(acy), rcyaw.append(acyaw)
        roff.append(roff[-1] + len(acx))
        wx.append(np.asarray(a.rx, dtype=np.float64)), wy.append(np.asarray(a.ry, dtype=np.float64))
        woff.append(woff[-1] + len(a.rx))
    n = len(specs)
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)  # noqa: E731
    s.update(act_kind=np.array([a.kind for a in specs], dtype=np.uint8),
             act_state0=np.array(st, dtype=np.float64).reshape(n, 4), act_tidx0=np.array(ti, dtype=np.int32),
             act_cruise_px=np.array([a.speed_mps / MPP for a in specs], dtype=np.float64),
             act_cruise_mps=np.array([a.speed_mps for a in specs], dtype=np.float64),
             act_beh=np.array([a.beh for a in specs], dtype=np.uint8),
             act_beh_p=np.array([a.beh_p for a in specs], dtype=np.float64).reshape(n, 4),
             act_cx=cat(rcx), act_cy=cat(rcy), act_cyaw=cat(rcyaw),
             act_route_off=np.array(roff, dtype=np.int32), act_raw_x=cat(wx),
             act_raw_y=cat(wy), act_raw_off=np.array(woff, dtype=np.int32))
    if lights:
        s["tl_rect"] = np.array([r for r, _ in lights], dtype=np.int32).reshape(-1, 4)
        s["tl_color"] = np.array([c for _, c in lights], dtype=np.uint8)
    return s


def _reset_loop(sample, bundle, cls_map, pad, max_reset_attempts, what):
    """CarlaBEV.reset's retry loop (carlabev.py:108-131): resample on an invalid spawn, generators keep running."""
    for _ in range(max_reset_attempts):
        s = sample()
        if cls_map is None or _spawn_valid(s, cls_map, pad):
            return s
    raise RuntimeError(f"Failed to reset into a valid initial state after {max_reset_attempts} attempts ({what})")


_SCENARIO_FIELDS = {  # scenarios/specs.py:47-90 (+ the anchor the samplers read)
    "lead_brake": ("ego_speed", "lead_gap", "lead_speed", "brake_delay", "brake_strength", "left_speed", "rear_gap",
                   "rear_speed", "rear_brake_delay", "anchor_x", "anchor_y"),
    "jaywalk": ("ego_speed", "cross_delay", "pedestrian_speed", "cross_offset", "yield_duration", "rear_gap",
                "rear_speed", "anchor_x", "anchor_y"),
}


def build_scripted_scene(kind: str, scene_seed: int, level: int | None = None, cls_map=None, pad: int = 182,
                         max_reset_attempts: int = 10, **params) -> dict:
    """CarlaBEV.reset for scene in {"lead_brake", "jaywalk"} -> pool entry.

    `cls_map` (H, W) uint8 classes enables the reference's spawn validation / retry loop
    (carlabev.py:108-131, scene.py:142-170); without it the first sample is accepted."""
    if kind not in _SAMPLERS:
        raise KeyError(f"Unknown scenario '{kind}'")
    bundle = RNGBundle(scene_seed, params.get("route_seed"), params.get("traffic_seed"), params.get("scenario_seed"))
    fields = {k: params[k] for k in _SCENARIO_FIELDS[kind] if params.get(k) is not None}

    def sample():
        lvl = level
        if lvl is None:
            lvl = bundle.scenario_rng.choice([1, 2, 3, 4])  # scene_generator.py:171-176
        agent, specs = _SAMPLERS[kind](int(lvl), bundle.scenario_np_rng, **fields)
        len_route = _route_length_m(agent[0], agent[1])             # compute_total_dist_m
        return _assemble(agent, specs, len_route, bundle.route_np_rng, [bundle.scenario_np_rng] * len(specs),
                         kind, lvl, scene_seed)

    return _reset_loop(sample, bundle, cls_map, pad, max_reset_attempts, f"kind={kind}, seed={scene_seed}")


# ---- scenes that need the lane graphs ----------------------------------------------------------------------------------
_EGO_GRAPHS = {"full_vehicle": ("vehicle-full", "vehicle"), "right_lane": ("vehicle-R", "R"),
               "left_lane": ("vehicle-L", "L")}  # scene_generator.py:252-268


def _ego_route_in_range(graph, node_cls, lo_m, hi_m, rng, max_attempts=100, filters=None):
    """find_route_in_range (scenes/utils.py:121-214): two random nodes, shortest path thinned at 10 raw px,
    waypoints of path[1:] in surface pixels, accepted when lo <= length [m] <= hi and the route-profile filters
    (route_profile / min_turns / max_turns / intersection_required) hold."""
    for _ in range(max_attempts):
        a = graph.random_node(node_cls, rng)
        b = graph.random_node(node_cls, rng)
        if a == b:
            continue
        path = graph.find_path(a, b)
        if len(path) < 2:
            continue
        pts = [graph.pos_surface(n) for n in path[1:]]
        rx, ry = [p[0] for p in pts], [p[1] for p in pts]
        total = _route_length_m(rx, ry)
        if lo_m <= total <= hi_m:
            if filters and not matches_route_profile(route_profile_metrics(rx, ry), **filters):
                continue
            return rx, ry, total
    return None


def sample_rdm(bundle, num_vehicles, dist_range, ego_target_speed=12.0, ego_route_graph="full_vehicle",
               traffic_enabled=True, max_route_attempts=20, route_profile=None, route_profile_mix=None, min_turns=None,
               max_turns=None, intersection_required=None):
    """SceneGenerator.generate_random (managers/scene_generator.py:196-330) + get_actor (:333-344)."""
    from .lanegraph import load_graph

    if ego_route_graph not in _EGO_GRAPHS:
        raise ValueError(f"Unsupported ego_route_graph={ego_route_graph!r}. "
                         "Expected one of: full_vehicle, right_lane, left_lane.")
    key, node_cls = _EGO_GRAPHS[ego_route_graph]
    graph = load_graph(key)
    if route_profile_mix:  # _sample_route_profile (scene_generator.py:78-93): one weighted draw from route_rng
        labels = list(route_profile_mix.keys())
        weights = [float(route_profile_mix[k]) for k in labels]
        if any(w < 0.0 for w in weights):
            raise ValueError(f"route_profile_mix must use non-negative weights: {route_profile_mix}")
        if sum(weights) <= 0.0:
            raise ValueError(f"route_profile_mix must contain at least one positive weight: {route_profile_mix}")
        route_profile = bundle.route_rng.choices(labels, weights=weights, k=1)[0]
    filters = dict(route_profile=route_profile, min_turns=min_turns, max_turns=max_turns,
                   intersection_required=intersection_required)
    if all(v is None for v in filters.values()):
        filters = None
    ego = None
    for _ in range(max_route_attempts):
        ego = _ego_route_in_range(graph, node_cls, dist_range[0], dist_range[1], bundle.route_rng, filters=filters)
        if ego is not None and len(ego[0]) > 1:
            break
        ego = None
    if ego is None:
        raise RuntimeError(f"Failed to generate a valid ego route in range {list(dist_range)} after "
                           f"{max_route_attempts} attempts.")
    rx, ry, len_route = ego
    specs = []
    for _ in range(int(num_vehicles) if traffic_enabled else 0):
        lane = bundle.traffic_rng.choice(["L", "R"])
        g = load_graph(f"vehicle-{lane}")
        n1 = g.random_node(lane, bundle.traffic_rng)
        n2 = g.random_node(lane, bundle.traffic_rng)
        path = g.find_path(n1, n2)
        pts = [g.pos_surface(n) for n in path[1:-1]]      # find_route, scenes/utils.py:94-108
        if len(pts) > 5:
            specs.append(_ActorSpec(0, [p[0] for p in pts], [p[1] for p in pts], 12.0))  # Vehicle(target_speed=12.0)
    return (rx, ry, 0.0, float(ego_target_speed)), specs, len_route


def build_rdm_scene(scene_seed: int, difficulty_id: str | None = None, num_vehicles: int | None = None,
                    route_dist_range=None, ego_target_speed: float | None = None, ego_route_graph="full_vehicle",
                    traffic_enabled: bool | None = None, max_route_attempts: int | None = None, cls_map=None,
                    pad: int = 182, max_reset_attempts: int = 10, max_vehicles: int = 50, **unsupported) -> dict:
    """CarlaBEV.reset(options={"scene": "rdm", ...}) -> pool entry (random navigation with background traffic).

    Option meaning and defaults follow build_scene (scene_generator.py:95-168): `num_vehicles` defaults to
    EnvConfig.max_vehicles, `route_dist_range` to [30, 100], `ego_target_speed` to 12 m/s.  Vehicles draw lanes / nodes from
    traffic_rng and start jitter from traffic_np_rng, the ego route from route_rng, the hero jitter from
    route_np_rng (src/randomness.py)."""
    prof = {k: unsupported.get(k) for k in ("route_profile", "route_profile_mix", "min_turns", "max_turns",
                                             "intersection_required")}
    # `difficulty_id` in raw reset options is context metadata only (scene_generator.py:144-150): the preset is
    # expanded into num_vehicles / route_dist_range / traffic_enabled by the typed request
    # (reset.build_random_navigation_options, config/reset.py:104-116), not here.
    num_vehicles = max_vehicles if num_vehicles is None else num_vehicles
    route_dist_range = [30, 100] if route_dist_range is None else route_dist_range
    bundle = RNGBundle(scene_seed, unsupported.get("route_seed"), unsupported.get("traffic_seed"),
                       unsupported.get("scenario_seed"))

    def sample():
        agent, specs, len_route = sample_rdm(bundle, num_vehicles, route_dist_range,
                                             12.0 if ego_target_speed is None else ego_target_speed,
                                             ego_route_graph, True if traffic_enabled is None else traffic_enabled,
                                             20 if max_route_attempts is None else int(max_route_attempts), **prof)
        return _assemble(agent, specs, len_route, bundle.route_np_rng, [bundle.traffic_np_rng] * len(specs),
                         "rdm", 0, scene_seed)

    return _reset_loop(sample, bundle, cls_map, pad, max_reset_attempts, f"kind=rdm, seed={scene_seed}")


# red_light_running.py:24-41: intersection centres as (raw_y, raw_x)
_INTERSECTIONS = ((8642, 1564), (8654, 6755), (7250, 1552), (7241, 2446), (7242, 3652), (7242, 4704), (7257, 6773),
                  (6199, 1552), (6197, 2439), (3349, 1545), (3350, 2456), (3350, 3639), (3335, 4714), (3315, 6773),
                  (2456, 1563), (2446, 6757))
PAL_ROUTE, PAL_TL_RED = 5, 6  # palette indices (include/cbev.h CBEV_PAL_*): a green strip is drawn in the route colour


def _direction_key(dx, dy):
    if abs(dx) > abs(dy):
        return "east" if dx > 0 else "west"
    return "south" if dy > 0 else "north"


def _select_intersection(graph, intersection_index=None, anchor_x=None, anchor_y=None):
    """RedLightRunningScenario._select_intersection (red_light_running.py:73-107): candidates by distance to the
    requested centre / anchor (else list order); the first with lane nodes in all four directions within 1200 raw px."""
    centres = [np.array([float(x), float(y)]) for y, x in _INTERSECTIONS]
    if intersection_index is not None:
        idx = int(intersection_index)
        if not 0 <= idx < len(centres):
            raise IndexError(f"intersection_index {idx} out of range for red_light_runner.")
        order = [i for _, i in sorted((float(np.linalg.norm(c - centres[idx])), i) for i, c in enumerate(centres))]
    elif anchor_x is not None and anchor_y is not None:
        anchor = np.array([anchor_x * 8.0, anchor_y * 8.0], dtype=float)
        order = [i for _, i in sorted((float(np.linalg.norm(c - anchor)), i) for i, c in enumerate(centres))]
    else:
        order = list(range(len(centres)))
    for i in order:
        delta = graph.pos - centres[i]
        near = np.linalg.norm(delta, axis=1) < 1200.0
        seen = {_direction_key(dx, dy) for dx, dy in delta[near]}
        if len(seen) == 4:
            return i, centres[i]
    raise RuntimeError("No valid 4-way intersection candidate found for red_light_runner.")


def _candidate_nodes(graph, centre, direction, min_dist=150.0, max_dist=1500.0):
    """_candidate_nodes (red_light_running.py:109-128): nodes of one approach ordered by |dist - 950| + 0.2 lateral."""
    cands = []
    for n in range(len(graph.names)):
        delta = graph.pos[n] - centre
        dist = np.linalg.norm(delta)
        if not (min_dist <= dist <= max_dist) or _direction_key(delta[0], delta[1]) != direction:
            continue
        lateral = abs(delta[0]) if direction in ("north", "south") else abs(delta[1])
        cands.append((abs(dist - 950.0) + 0.2 * lateral, n))
    cands.sort(key=lambda item: item[0])
    return [n for _, n in cands]


def _straight_route(graph, centre, start_dir, end_dir):
    """_sample_straight_route (red_light_running.py:139-165): first of the 25 x 25 best node pairs whose shortest
    path passes within 180 raw px of the centre with at least 6 nodes; waypoints = float raw positions / 8."""
    from .lanegraph import NoPath

    starts = _candidate_nodes(graph, centre, start_dir)[:25]
    ends = _candidate_nodes(graph, centre, end_dir)[:25]
    for a in starts:
        for b in ends:
            try:
                path = graph.shortest_path(a, b)
            except NoPath:
                continue
            coords = graph.pos[path]
            if min(float(np.linalg.norm(p - centre)) for p in coords) > 180.0 or len(coords) < 6:
                continue
            return [float(p[0]) / 8.0 for p in coords], [float(p[1]) / 8.0 for p in coords]
    raise RuntimeError(f"Unable to build a valid {start_dir}->{end_dir} route through the selected 4-way intersection.")


def _stop_line(centre_surface, direction, color):
    """_build_stop_line + TrafficLight._update_rect (red_light_running.py:167-199, traffic_light.py:60-72):
    8 m x (0.45 m + 1 px) strip, 4 m before the centre; pygame.Rect truncates its float arguments."""
    off, length, width = m2s(4.0), m2s(8.0), m2s(0.45) + 1.0
    x, y = float(centre_surface[0]), float(centre_surface[1])
    if direction == "south":
        y, horizontal = y + off, True
    elif direction == "north":
        y, horizontal = y - off, True
    elif direction == "west":
        x, horizontal = x - off, False
    else:
        x, horizontal = x + off, False
    w, h = (length, width) if horizontal else (width, length)
    return [int(x - w / 2), int(y - h / 2), int(w), int(h)], color


def build_red_light_scene(scene_seed: int, intersection_index=None, anchor_x=None, anchor_y=None, ego_speed=10.0,
                          adv_speed=16.0, level=None, cls_map=None, pad: int = 182, max_reset_attempts: int = 10,
                          **seeds) -> dict:
    """CarlaBEV.reset(options={"scene": "red_light_runner", ...}) -> pool entry (red_light_running.py:201-245).

    The geometry is fixed by the intersection; only the +-1 px start jitters vary.  The hero's comes from
    route_np_rng as in the reference.  The adversary's generator is `None` in the reference, so its jitter is drawn
    from an unseeded `np.random.default_rng()` there (quirk C-10): here it is drawn from a stream derived from the
    scene seed, which is one of the nine realisations the reference can produce for that seed."""
    from .lanegraph import load_graph

    graph = load_graph("vehicle")
    bundle = RNGBundle(scene_seed, seeds.get("route_seed"), seeds.get("traffic_seed"), seeds.get("scenario_seed"))
    adv_np_rng = np.random.default_rng(derive_seed(scene_seed, "adversary_jitter"))

    def sample():
        if level is None:
            bundle.scenario_rng.choice([1, 2, 3, 4])            # scene_generator.py:171-176 (the level is unused)
        _, centre = _select_intersection(graph, intersection_index, anchor_x, anchor_y)
        ego_rx, ego_ry = _straight_route(graph, centre, "south", "north")
        adv_rx, adv_ry = _straight_route(graph, centre, "west", "east")
        centre_s = centre / 8.0
        lights = [_stop_line(centre_s, "south", PAL_ROUTE), _stop_line(centre_s, "west", PAL_TL_RED)]
        return _assemble((ego_rx, ego_ry, ego_speed, ego_speed), [_ActorSpec(0, adv_rx, adv_ry, adv_speed)],
                         _route_length_m(ego_rx, ego_ry), bundle.route_np_rng, [adv_np_rng], "red_light_runner",
                         level or 0, scene_seed, lights)

    return _reset_loop(sample, bundle, cls_map, pad, max_reset_attempts, f"kind=red_light_runner, seed={scene_seed}")


# ---- authored scenes and scenario-config files (JSON) ----------------------------------------------------------------
SCENARIO_PRESETS = {  # scenarios/specs.py:95-146
    "jaywalk_debug": dict(scene="jaywalk", level=3, ego_speed=10.0, cross_delay=1.2, pedestrian_speed=1.6,
                          yield_duration=1.2),
    "lead_brake_debug": dict(scene="lead_brake", level=2, ego_speed=12.0, lead_gap=8.0, lead_speed=11.0,
                             brake_delay=2.0, brake_strength=4.0),
    "red_light_debug": dict(scene="red_light_runner", intersection_index=11, ego_speed=10.0, adv_speed=16.0),
    "rdm_navigation": dict(scene="rdm", num_vehicles=25, route_dist_range=[30, 130]),
}
_SPEC_DEFAULTS = {  # ScenarioSpec field defaults (scenarios/specs.py:47-92): a scenario-config file fixes ALL of them
    "jaywalk": dict(ego_speed=12.0, cross_delay=1.5, pedestrian_speed=1.6, cross_offset=0.0, yield_duration=1.2,
                    rear_gap=5.0, rear_speed=10.0),
    "lead_brake": dict(ego_speed=12.0, lead_gap=7.5, lead_speed=12.0, brake_delay=2.5, brake_strength=4.0,
                       left_speed=14.0, rear_gap=5.0, rear_speed=10.0, rear_brake_delay=3.0),
    "red_light_runner": dict(ego_speed=10.0, adv_speed=16.0, intersection_index=11),
}
_LEGACY_BEHAVIORS = {"Normal": "constant_speed", "CrossBehavior": "cross", "StopMidBehavior": "stop_mid",
                     "StopReturnBehavior": "yield_return", "LeadBrakeBehavior": "timed_brake"}
_BEHAVIOR_FIELDS = {  # actors/behavior/registry.py:33-68
    "vehicle": {"constant_speed": {}, "timed_brake": {"start_brake_t": 3.5, "decel_mps2": 1.0}},
    "pedestrian": {"cross": {"start_delay": 0.0}, "stop_mid": {"start_delay": 0.0},
                   "yield_return": {"start_delay": 0.0, "yield_duration": 1.0}},
}


def scenario_preset_options(preset_id: str, overrides: dict | None = None) -> dict:
    """build_runtime_scenario_options (scenarios/specs.py:167-183)."""
    if preset_id not in SCENARIO_PRESETS:
        raise KeyError(f"Unknown scenario preset '{preset_id}'")
    options = copy.deepcopy(SCENARIO_PRESETS[preset_id])
    options.update({k: v for k, v in (overrides or {}).items() if v is not None})
    return options


def scenario_config_options(data: dict, overrides: dict | None = None) -> dict:
    """normalize_scenario_config + build_scenario_options_from_config (scenarios/specs.py:218-272): a scenario
    config ({"scenario_id", "level", "anchor", "parameters"} or the legacy {"scenario", "kwargs"}) -> reset options
    with every ScenarioSpec field filled in (missing ones take the spec default, so nothing is left to the draw)."""
    if data.get("type") == "scenario_config" or "scenario_id" in data:
        scenario_id, level = data.get("scenario_id"), int(data.get("level", 1))
        anchor, raw = data.get("anchor", {}) or {}, data.get("parameters", {}) or {}
    elif "scenario" in data and "kwargs" in data:
        raw = dict(data.get("kwargs", {}))
        scenario_id, level = data.get("scenario"), int(raw.pop("level", 1))
        anchor = {"x": raw.pop("anchor_x", None), "y": raw.pop("anchor_y", None)}
        raw.pop("scene", None)
    else:
        raise ValueError("Unsupported scenario config format.")
    if scenario_id not in _SPEC_DEFAULTS:
        raise KeyError(f"Unknown scenario '{scenario_id}'")
    options = {k: type(d)(d if raw.get(k) in (None, "") else raw[k]) for k, d in _SPEC_DEFAULTS[scenario_id].items()}
    if anchor.get("x") is not None:
        options["anchor_x"] = int(anchor["x"])
    if anchor.get("y") is not None:
        options["anchor_y"] = int(anchor["y"])
    options["level"], options["scene"] = level, scenario_id
    for k, v in (overrides or {}).items():
        if k not in ("config_file", "scene", "reset_mask") and v is not None:
            options[k] = v
    return options


def _behavior_of(actor_type, behavior):
    """normalize_behavior_spec + build_behavior (actors/behavior/registry.py:101-145) -> (BEH_*, 4 parameters)."""
    specs = _BEHAVIOR_FIELDS[actor_type]
    if behavior in (None, "", "Normal"):
        bid, raw = next(iter(specs)), {}
    elif isinstance(behavior, str):
        bid, raw = _LEGACY_BEHAVIORS.get(behavior, behavior), {}
    else:
        bid = _LEGACY_BEHAVIORS.get(behavior.get("type", ""), behavior.get("type", ""))
        raw = behavior.get("params", {}) or behavior.get("behavior_kwargs", {}) or {}
    if bid not in specs:
        bid = next(iter(specs))
    p = {k: float(d if raw.get(k) in (None, "") else raw[k]) for k, d in specs[bid].items()}
    if isinstance(behavior, str) or behavior in (None, ""):
        p = dict(specs[bid])  # a bare name carries no parameters: build_behavior falls back to its own defaults
    if bid == "cross":
        return BEH_CROSS, (p["start_delay"], 2.0, 0.0, 0.0)
    if bid == "stop_mid":
        return BEH_STOP_MID, (p["start_delay"], 0.5, -1.0, 0.0)
    if bid == "yield_return":
        return BEH_STOP_RETURN, (p["start_delay"], 1.0 / 3.0, p["yield_duration"], 1.0)
    if bid == "timed_brake":
        return BEH_LEAD_BRAKE, (p["start_brake_t"], p["decel_mps2"], 0.0, 0.0)
    return BEH_NONE, (0.0, 0.0, 0.0, 0.0)


def _linear_route(start, end, step_px=8):
    """_build_linear_route (scenarios/__init__.py:12-20)."""
    dx, dy = end[0] - start[0], end[1] - start[1]
    n = max(2, int(max(abs(dx), abs(dy)) / max(1, step_px)) + 1)
    return (np.linspace(start[0], end[0], n).round().astype(int).tolist(),
            np.linspace(start[1], end[1], n).round().astype(int).tolist())


def _route_from_waypoints(waypoints):
    rx, ry = [], []
    for i in range(len(waypoints) - 1):
        sx, sy = _linear_route(waypoints[i], waypoints[i + 1])
        rx.extend(sx[1:] if i else sx)
        ry.extend(sy[1:] if i else sy)
    return rx, ry


def _variation_value(spec, rng, fallback=None):
    """_sample_variation_value (scenarios/__init__.py:43-66)."""
    if spec is None:
        return fallback
    if not isinstance(spec, dict):
        return spec
    mode = spec.get("mode", "fixed")
    if mode == "fixed":
        return spec.get("value", fallback)
    if mode == "uniform":
        return rng.uniform(float(spec["low"]), float(spec["high"]))
    if mode == "normal":
        value = rng.normalvariate(float(spec["mean"]), float(spec["std"]))
        clip = spec.get("clip")
        if clip is not None and len(clip) == 2:
            value = max(float(clip[0]), min(float(clip[1]), value))
        return value
    if mode == "choice":
        values = spec.get("values", [])
        return rng.choice(list(values)) if values else fallback
    return fallback


def _vary_actor(actor_data, enabled, seed, global_spec, index):
    """_apply_actor_variation (scenarios/__init__.py:113-199): one random.Random per actor, seeded
    variation_seed + seed_offset; waypoint jitter, then speed, then behaviour parameters, then signal state."""
    actor = copy.deepcopy(actor_data)
    var = actor.get("variation") or {}
    if not enabled or not var.get("enabled", False):
        return actor
    rng = random.Random(seed + int(var.get("seed_offset", index)))
    if actor.get("waypoints"):
        wps = [[int(round(p[0])), int(round(p[1]))] for p in actor["waypoints"]]
    else:
        rx, ry = actor.get("rx", []), actor.get("ry", [])
        start = actor.get("start") or ({"x": rx[0], "y": ry[0]} if rx and ry else None)
        goal = actor.get("goal") or ({"x": rx[-1], "y": ry[-1]} if rx and ry else None)
        wps = [] if start is None or goal is None else [[int(round(start["x"])), int(round(start["y"]))],
                                                        [int(round(goal["x"])), int(round(goal["y"]))]]
    lock = (var.get("constraints", {}) or {}).get("lock_endpoints", True)
    jitter = var.get("waypoint_jitter_px", global_spec.get("waypoint_jitter_px"))
    if jitter and wps:
        r = float(jitter)
        varied = []
        for i, p in enumerate(wps):
            if lock and i in (0, len(wps) - 1):
                varied.append(list(p))
            else:
                varied.append([int(round(p[0] + rng.uniform(-r, r))), int(round(p[1] + rng.uniform(-r, r)))])
        actor["waypoints"] = varied
        actor["start"] = {"x": varied[0][0], "y": varied[0][1]}
        actor["goal"] = {"x": varied[-1][0], "y": varied[-1][1]}
    speed = float(actor.get("cruise_speed", actor.get("initial_speed", actor.get("speed", 0.0))))
    scale = _variation_value(global_spec.get("speed_scale"), rng, fallback=1.0)
    if var.get("speed") is not None:
        speed = float(_variation_value(var.get("speed"), rng, fallback=speed))
    else:
        speed = speed * float(scale)
    speed = max(0.0, speed)
    actor["speed"] = actor["initial_speed"] = actor["cruise_speed"] = speed
    behavior = copy.deepcopy(actor.get("behavior") or {})
    if isinstance(behavior, dict):
        params = copy.deepcopy(behavior.get("params") or {})
        changed = False
        for key, spec in (var.get("behavior_params", {}) or {}).items():
            if key in params:
                params[key] = _variation_value(spec, rng, fallback=params[key])
                changed = True
        if changed:
            behavior["params"] = params
            actor["behavior"] = behavior
    if actor.get("type") == "traffic_light" and var.get("signal_state"):
        actor["signal_state"] = _variation_value(var.get("signal_state"), rng, fallback=actor.get("signal_state", "red"))
    return actor


def bundled_authored_files() -> dict:
    """The reference's authored scene files (assets/scenes/*.json), parsed: {file name: scene dict}."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "authored_scene_files.json")
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def build_authored_scene(config, scene_seed: int = 0, variation_enabled=None, variation_seed=None, cls_map=None,
                         pad: int = 182, max_reset_attempts: int = 10, **overrides) -> dict:
    """CarlaBEV.reset(options={"config_file": path, ...}) -> pool entry.  `config` is a path or the parsed JSON.

    Authored scenes ({"actors": [...]}, scenarios/__init__.py:210-338): routes from `rx`/`ry` or piecewise-linear
    waypoints, optional seeded variation of waypoints / speeds / behaviour parameters / signal states.  Their
    scripted actors are built without a generator, so the reference draws their +-1 px start jitter from an
    unseeded stream; here it comes from a stream derived from (scene_seed, actor index).  Files without "actors"
    are scenario configs and resolve to the scenario sampler with every parameter pinned."""
    if isinstance(config, (str, bytes)) or hasattr(config, "__fspath__"):
        import json

        with open(config, "r", encoding="utf-8") as f:
            data = json.load(f)
    else:
        data = config
    if "actors" not in data:
        opts = scenario_config_options(data, overrides)
        opts.setdefault("scene_seed", scene_seed)
        return build_scene(opts, cls_map=cls_map, pad=pad)
    scenario_id = data.get("scenario_id") or data.get("scenario")
    if scenario_id not in _SPEC_DEFAULTS:
        raise KeyError(f"Unknown scenario '{scenario_id}' in authored config")
    variation = data.get("variation") or {}
    enabled = bool(variation.get("enabled", False)) if variation_enabled is None else bool(variation_enabled)
    vseed = variation_seed if variation_seed is not None else variation.get("default_seed")
    vseed = int(vseed) if vseed is not None else 0
    global_spec = variation.get("global", {}) or {}
    bundle = RNGBundle(scene_seed, overrides.get("route_seed"), overrides.get("traffic_seed"),
                       overrides.get("scenario_seed"))

    def sample():
        agent, vehicles, peds, lights = None, [], [], []
        for idx, actor_data in enumerate(data["actors"]):
            a = _vary_actor(actor_data, enabled, vseed, global_spec, idx)
            atype = actor_data["type"]
            rx, ry = a.get("rx"), a.get("ry")
            if (not rx or not ry) and a.get("waypoints"):
                rx, ry = _route_from_waypoints(a["waypoints"])
            rx, ry = rx or [], ry or []
            speed = a.get("cruise_speed", a.get("initial_speed", a.get("speed", 2.0)))
            if atype == "agent":
                agent = (rx, ry, speed, speed)
            elif atype in ("vehicle", "pedestrian"):
                beh, p = _behavior_of(atype, a.get("behavior", "constant_speed" if atype == "vehicle" else "cross"))
                if len(rx) > 0 and len(rx) == len(ry):  # ActorManager._has_valid_route
                    (vehicles if atype == "vehicle" else peds).append(_ActorSpec(0 if atype == "vehicle" else 1,
                                                                                 rx, ry, speed, beh, p))
            elif atype == "traffic_light":
                start = a.get("start") or ({"x": rx[0], "y": ry[0]} if rx and ry else None)
                goal = a.get("goal") or ({"x": rx[-1], "y": ry[-1]} if rx and ry else None)
                if start is None or goal is None:
                    continue
                dx, dy = float(goal["x"]) - float(start["x"]), float(goal["y"]) - float(start["y"])
                cx, cy = 0.5 * (float(start["x"]) + float(goal["x"])), 0.5 * (float(start["y"]) + float(goal["y"]))
                horizontal = a.get("orientation", "horizontal" if abs(dx) >= abs(dy) else "vertical") == "horizontal"
                # TrafficLight.__init__ defaults (traffic_light.py:33-42)
                width = float(a["width"]) if a.get("width") is not None else max(1.0, m2s(0.45)) + 1.0
                length = float(a["length"]) if a.get("length") is not None else max(4.0, m2s(8.5))
                w, h = (length, width) if horizontal else (width, length)
                color = {"red": PAL_TL_RED, "yellow": 7, "green": PAL_ROUTE}.get(a.get("signal_state", "red"), PAL_TL_RED)
                lights.append(([int(cx - w / 2), int(cy - h / 2), int(w), int(h)], color))
        if agent is None:
            raise ValueError("authored scene has no agent")
        # compute_total_dist_px([rx, ry]) (scenes/utils.py:217-224) pairs the two coordinate LISTS as if they were
        # two points: the reference's `len_ego_route` of an authored scene is hypot(ry[0]-rx[0], ry[1]-rx[1])
        rx, ry = agent[0], agent[1]
        len_route = float(np.hypot(ry[0] - rx[0], ry[1] - rx[1]))
        specs = vehicles + peds
        rngs = [np.random.default_rng(derive_seed(scene_seed, "authored_jitter", i)) for i in range(len(specs))]
        return _assemble(agent, specs, len_route, bundle.route_np_rng, rngs, scenario_id,
                         int(data.get("level", 0) or 0), scene_seed, lights)

    return _reset_loop(sample, bundle, cls_map, pad, max_reset_attempts, f"authored scene, seed={scene_seed}")


def build_scene(options: dict, cls_map=None, pad: int = 182, max_vehicles: int = 50) -> dict:
    """Host mirror of CarlaBEV.reset(options=...) -> SceneGenerator.build_scene (scene_generator.py:95-194) for
    the generated scene kinds: one reset-options dict -> one pool entry."""
    o = dict(options)
    o.pop("reset_mask", None)
    scene = o.pop("scene", "rdm")
    seed = int(o.pop("scene_seed", 0))
    common = dict(cls_map=cls_map, pad=pad, max_reset_attempts=o.pop("max_reset_attempts", 10))
    if o.get("config_file") or str(scene).endswith(".json"):
        return build_authored_scene(o.pop("config_file", None) or scene, scene_seed=seed, **common, **o)
    if scene == "rdm":
        return build_rdm_scene(seed, max_vehicles=max_vehicles, **common, **o)
    if scene in _SAMPLERS:
        return build_scripted_scene(scene, seed, level=o.pop("level", None), **common, **o)
    if scene == "red_light_runner":
        return build_red_light_scene(seed, **common, **o)
    raise KeyError(f"Unknown scenario '{scene}' (authored JSON scenes ship as a pool: pool.load_shipped_pool)")


def _build_chunk(args):
    requests, pad, max_vehicles = args[:3]
    size = args[3] if len(args) > 3 else 128
    from .vector_env import load_town01_map

    skip_invalid = args[4] if len(args) > 4 else False
    cls = load_town01_map(size)
    out = []
    for r in requests:
        try:
            out.append(build_scene(r, cls_map=cls, pad=pad, max_vehicles=max_vehicles))
        except RuntimeError as ex:
            # "Failed to reset into a valid initial state after N attempts": what CarlaBEV.reset raises for this seed
            if not skip_invalid or "valid initial state" not in str(ex):
                raise
            out.append(None)
    return out


def build_pool(requests: list[dict], pad: int = 182, max_vehicles: int = 50, workers: int | None = None,
               size: int = 128, skip_invalid: bool = False) -> list[dict]:
    """Reset-option dicts -> pool entries, on `workers` host processes (default: one per core, serial for small
    pools).  Scene generation is host work by design (the device steps scenes, it does not build them).
    `size` = EnvConfig.size: the generators keep their 128-scale coordinates at every size (the reference hard-codes
    map_size=128, scene_generator.py:65-76), only the spawn validation reads the map of that scale.
    `skip_invalid`: requests whose reset the reference itself gives up on (RuntimeError after max_reset_attempts) come
    back as None instead of raising (pools for the other map scales, where many seeds spawn the ego off the road).

    The workers are plain child interpreters running this module (`python -m carlabev_env_b200.scenes --worker`),
    fed and read through files: nothing of the caller's `__main__` is re-imported, so this is safe to call from any
    script (no `if __name__ == "__main__"` guard needed) and from a process that already holds a CUDA context."""
    import os
    import pickle
    import subprocess
    import sys
    import tempfile

    n = len(requests)
    workers = min(os.cpu_count() or 1, 32) if workers is None else workers
    workers = min(workers, max(1, n // 16))
    if workers <= 1:
        return _build_chunk((requests, pad, max_vehicles, size, skip_invalid))
    chunks = [list(range(w, n, workers)) for w in range(workers)]
    pkg_parent = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = pkg_parent + os.pathsep + env.get("PYTHONPATH", "")
    env.setdefault("OMP_NUM_THREADS", "1")  # one core per worker: the routes are tiny, BLAS threads only collide
    out = [None] * n
    with tempfile.TemporaryDirectory(prefix="cbev_pool_") as tmp:
        procs = []
        for w, c in enumerate(chunks):
            req, res = os.path.join(tmp, f"req{w}.pkl"), os.path.join(tmp, f"res{w}.pkl")
            with open(req, "wb") as f:
                pickle.dump(([requests[i] for i in c], pad, max_vehicles, size, skip_invalid), f)
            procs.append((c, res, subprocess.Popen([sys.executable, "-m", "carlabev_env_b200.scenes", "--worker", req, res],
                                                   env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)))
        errors = []
        for c, res, p in procs:
            _, err = p.communicate()
            if p.returncode != 0:
                errors.append(err.decode(errors="replace")[-800:])
                continue
            with open(res, "rb") as f:
                for i, sc in zip(c, pickle.load(f)):
                    out[i] = sc
        if errors:
            raise RuntimeError("scene generation worker failed:\n" + errors[0])
    return out


def _spawn_valid(s, cls_map, pad) -> bool:
    """Scene.spawn_validation_info, scene.py:142-170."""
    h, w = cls_map.shape
    x, y = float(s["ego_state0"][0]), float(s["ego_state0"][1])
    tx = int(np.clip(_round_half_even(x), 0, w - 1))
    ty = int(np.clip(_round_half_even(y), 0, h - 1))
    if cls_map[ty, tx] == 0:
        return False
    # the ego square follows EnvConfig.size (hero.py:14-17: int(32 / int(1024 / size)) = 2 / 4 / 8 px; the map is 8 x size
    # wide), the scripted actors do not: the scene generator builds them with map_size = 128 at every scale
    # (scene_generator.py:65, vehicle.py:19-24, pedestrian.py:19-24)
    hw = max(1, 32 // max(1, 1024 // max(1, w // 8)))
    hx, hy = _rect_left(x, pad, hw), _rect_left(y, pad, hw)
    for i, kind in enumerate(s["act_kind"]):
        size = 4 if kind == 0 else 2
        ax, ay = _rect_left(s["act_state0"][i, 0], pad, size), _rect_left(s["act_state0"][i, 1], pad, size)
        if hx < ax + size and hy < ay + size and hx + hw > ax and hy + hw > ay:
            return False
    return True


def build_scripted_pool(kind_level_seed, cls_map=None, pad: int = 182) -> list[dict]:
    """[(kind, level | None, scene_seed), ...] -> list of pool entries."""
    return [build_scripted_scene(k, seed, level=lv, cls_map=cls_map, pad=pad) for k, lv, seed in kind_level_seed]


def _main(argv=None):
    """python -m carlabev_env_b200.scenes --scene rdm --difficulty-id rt_hard_v1 --count 4096 --out pool.npz"""
    import argparse
    import json
    import sys
    import time

    from .pool import save_pool

    if argv is None and len(sys.argv) == 4 and sys.argv[1] == "--worker":  # child of build_pool: chunk in, scenes out
        import pickle

        with open(sys.argv[2], "rb") as f:
            args = pickle.load(f)
        with open(sys.argv[3], "wb") as f:
            pickle.dump(_build_chunk(args), f)
        return
    ap = argparse.ArgumentParser(description="Generate a scene pool on the host cores (entry i has scene_seed = seed0 + i).")
    ap.add_argument("--scene", default="rdm", help="rdm | lead_brake | jaywalk | red_light_runner | path to an authored / scenario-config JSON")
    ap.add_argument("--count", type=int, default=256)
    ap.add_argument("--seed0", type=int, default=0)
    ap.add_argument("--difficulty-id", default=None)
    ap.add_argument("--level", type=int, default=None)
    ap.add_argument("--options", default="{}", help="JSON dict of further reset options (num_vehicles, route_dist_range, ...)")
    ap.add_argument("--pad", type=int, default=182, help="crop size of the camera (182 centred, 230 lookahead_75)")
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--out", required=True)
    a = ap.parse_args(argv)
    base = dict(json.loads(a.options))
    if a.scene.endswith(".json"):
        base["config_file"] = a.scene
    else:
        base["scene"] = a.scene
    if a.difficulty_id is not None:  # expanded like the typed request does (a bare difficulty_id is metadata only)
        from .reset import RandomNavigationReset, build_random_navigation_options

        preset = build_random_navigation_options(RandomNavigationReset(difficulty_id=a.difficulty_id))
        base = {**preset, **base}
    if a.level is not None:
        base["level"] = a.level
    t0 = time.time()
    scenes = build_pool([{**base, "scene_seed": a.seed0 + i} for i in range(a.count)], pad=a.pad, workers=a.workers)
    save_pool(a.out, scenes)
    print(f"{a.out}: {len(scenes)} scenes, actors {min(len(s['act_kind']) for s in scenes)}.."
          f"{max(len(s['act_kind']) for s in scenes)}, {time.time() - t0:.1f} s")


if __name__ == "__main__":
    _main()
