"""Generate tests/golden/* by stepping the UNMODIFIED reference (under oracle/shims).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py lead_brake # cases whose name contains the substring

Each case writes tests/golden/<case>.npz holding: the packed scene pool (post-reset
snapshots taken from the reference's own objects), the action sequence, and per-step
state / flags / rewards / palette-index frames / observation checksums recorded from
the reference.  The map fixture carlabev_env_b200/assets/town01_128_cls.npz is derived
from the reference's Town01-128-sem.png by `dump_map`.
"""
from __future__ import annotations

import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

load_reference()

from CarlaBEV.config import EnvConfig, RunConfig  # noqa: E402
from CarlaBEV.envs import make_env  # noqa: E402

from carlabev_env_b200.pool import pack_pool  # noqa: E402
from oracle import raster, sim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
KIND_IDS = {"rdm": 0, "lead_brake": 1, "jaywalk": 2, "red_light_runner": 3}
CAUSE_IDS = {name: i for i, name in enumerate(sim.CAUSE_NAMES)}
FSM_IDS = {"idle": 0, "waiting": 1, "entering": 2, "yielding": 3, "crossing": 4, "stalled": 5, "retreating": 6,
           "cleared": 7, "retreated": 8}


def dump_map(size=128):
    from PIL import Image

    sem = np.array(Image.open(os.path.join(REFERENCE_ROOT, f"CarlaBEV/assets/Town01/Town01-{size}-sem.png")))
    rgb = np.array(Image.open(os.path.join(REFERENCE_ROOT, f"CarlaBEV/assets/Town01/Town01-{size}-rgb.png")).convert("RGB"))
    cls = np.zeros(sem.shape, dtype=np.uint8)
    cls[sem == 127] = 1   # DRIVABLE   (semantics.py:34-38)
    cls[sem == 255] = 2   # SIDEWALK
    assert set(np.unique(sem)) <= {0, 127, 255}
    assert np.array_equal(raster.PALETTE[cls], rgb), "rgb map is not LUT(sem)"
    out = os.path.join(ROOT, "carlabev_env_b200", "assets")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, f"town01_{size}_cls.npz"), cls=cls)
    print("map", size, cls.shape, np.bincount(cls.ravel()))


def behaviour_of(actor):
    b = actor.behavior
    if b is None:
        return sim.BEH_NONE, (0, 0, 0, 0)
    name = type(b).__name__
    if name == "LeadBrakeBehavior":
        return sim.BEH_LEAD_BRAKE, (b.start_brake_t, b.dec_rate, 0, 0)
    beh = {"CrossBehavior": sim.BEH_CROSS, "StopMidBehavior": sim.BEH_STOP_MID,
           "StopReturnBehavior": sim.BEH_STOP_RETURN}[name]
    return beh, (b.start_delay, b.trigger_fraction, -1.0 if b.stop_duration is None else b.stop_duration,
                 1.0 if b.retreat else 0.0)


def tl_palette(color):
    color = tuple(int(c) for c in color)
    for i, p in enumerate(raster.PALETTE):
        if tuple(int(v) for v in p) == color:
            return i
    raise ValueError(color)


def extract_scene(base_env, options):
    """Post-reset snapshot of the reference's objects."""
    import pygame

    m = base_env.map
    hero = m.hero
    am = m.actor_manager
    s = {
        "ego_state0": np.array([hero.x, hero.y, hero.yaw, hero.v], dtype=np.float64),
        "ego_target_speed": np.float64(hero._target_speed),
        "ego_tidx0": np.int32(hero.target_idx),
        "ego_cx": np.asarray(hero.cx, dtype=np.float64),
        "ego_cy": np.asarray(hero.cy, dtype=np.float64),
        "ego_cyaw": np.asarray(hero.cyaw, dtype=np.float64),
        "rew_rx": np.asarray(m.route[0], dtype=np.int32),
        "rew_ry": np.asarray(m.route[1], dtype=np.int32),
        "route_length_m": np.float64(am.route_length),
        "len_ego_route": np.float64(base_env.len_ego_route),
        "num_vehicles": np.int32(base_env.num_vehicles),
        "kind": np.int32(KIND_IDS.get(options.get("scene", "rdm"), 4)),
        "level": np.int32(options.get("level", 0) or 0),
        "seed": np.int64(options.get("scene_seed", 0)),
    }
    kinds, st, tidx, cpx, cmps, beh, behp = [], [], [], [], [], [], []
    rcx, rcy, rcyaw, roff = [], [], [], [0]
    wx, wy, woff = [], [], [0]
    for key, kind in (("vehicle", sim.KIND_VEHICLE), ("pedestrian", sim.KIND_PEDESTRIAN)):
        for a in am.actors[key]:
            c = a._controller
            kinds.append(kind)
            st.append([c.x, c.y, c.yaw, c.v])
            tidx.append(c.target_idx)
            cpx.append(a.cruise_speed)
            cmps.append(a.cruise_speed_mps)
            b, p = behaviour_of(a)
            beh.append(b)
            behp.append(p)
            rcx.append(np.asarray(c.cx, dtype=np.float64))
            rcy.append(np.asarray(c.cy, dtype=np.float64))
            rcyaw.append(np.asarray(c.cyaw, dtype=np.float64))
            roff.append(roff[-1] + len(c.cx))
            wx.append(np.asarray(a._initial_rx, dtype=np.float64))
            wy.append(np.asarray(a._initial_ry, dtype=np.float64))
            woff.append(woff[-1] + len(a._initial_rx))
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)  # noqa: E731
    s.update(
        act_kind=np.array(kinds, dtype=np.uint8), act_state0=np.array(st, dtype=np.float64).reshape(-1, 4),
        act_tidx0=np.array(tidx, dtype=np.int32), act_cruise_px=np.array(cpx, dtype=np.float64),
        act_cruise_mps=np.array(cmps, dtype=np.float64), act_beh=np.array(beh, dtype=np.uint8),
        act_beh_p=np.array(behp, dtype=np.float64).reshape(-1, 4),
        act_cx=cat(rcx), act_cy=cat(rcy), act_cyaw=cat(rcyaw), act_route_off=np.array(roff, dtype=np.int32),
        act_raw_x=cat(wx), act_raw_y=cat(wy), act_raw_off=np.array(woff, dtype=np.int32),
    )
    rects, cols = [], []
    for tl in am.actors["traffic_light"]:
        if tl.orientation == "horizontal":
            w, h = tl.length, tl.width
        else:
            w, h = tl.width, tl.length
        r = pygame.Rect(tl.x - w / 2, tl.y - h / 2, w, h)  # traffic_light.py:81-90
        rects.append([r.x, r.y, r.w, r.h])
        cols.append(tl_palette(tl._color))
    s["tl_rect"] = np.array(rects, dtype=np.int32).reshape(-1, 4)
    s["tl_color"] = np.array(cols, dtype=np.uint8)
    return s


def rgb_to_index(rgb):
    idx = np.full(rgb.shape[:2], 255, dtype=np.uint8)
    for i, p in enumerate(raster.PALETTE):
        idx[np.all(rgb == p, axis=-1)] = i
    assert idx.max() < len(raster.PALETTE), "frame colour outside the palette"
    return idx


def run_case(name, env_kwargs, reset_options, actions, frame_every=1, obs_every=25):
    """Step one reference env through `actions`, resetting (masked reset) after terminal steps."""
    cfg = RunConfig(env=EnvConfig(render_mode="rgb_array", **env_kwargs), num_envs=1)
    envs = make_env(cfg)
    wrapped = envs.envs[0]
    base = wrapped.unwrapped
    T = len(actions)
    scenes, reset_steps, reset_scene = [], [], []
    episode = 0

    def do_reset(step):
        nonlocal episode
        opts = reset_options(episode) if callable(reset_options) else dict(reset_options)
        obs, info = envs.reset(options={**opts, "reset_mask": np.array([True])})
        scenes.append(extract_scene(base, opts))
        reset_steps.append(step)
        reset_scene.append(len(scenes) - 1)
        episode += 1
        return obs

    rec = {k: [] for k in ("ego_state", "acc", "tidx", "reward", "term", "trunc", "cause", "hit", "hit_id", "tile",
                           "dist2wp", "comfort", "n_nearby", "obs_crc", "rgb_crc")}
    amax = 0
    act_states, act_tidx, act_fsm, tgt_vis = [], [], [], []
    frames, frame_steps, obs_full, obs_steps, reset_frames, reset_obs = [], [], [], [], [], []
    ep_infos = {}

    obs = do_reset(0)
    reset_frames.append(rgb_to_index(base.observation))
    reset_obs.append(np.asarray(obs[0]))
    for t in range(T):
        obs, rew, term, trunc, infos = envs.step([actions[t]])
        info = base.current_info
        hero = base.map.hero
        rec["ego_state"].append([hero.x, hero.y, hero.yaw, hero.v])
        rec["acc"].append(float(hero.acc))
        rec["tidx"].append(int(hero.target_idx))
        rec["reward"].append(float(rew[0]))
        rec["term"].append(bool(term[0]))
        rec["trunc"].append(bool(trunc[0]))
        rec["cause"].append(CAUSE_IDS[info["reward"]["cause"]])
        col = info["collision"]
        hit = {None: 0, "vehicle": 1, "pedestrian": 2, "target": 3}[col["collided"]]
        hid = col["actor_id"]
        n_t = len(base.map.actor_manager.actors["target"])
        rec["hit"].append(hit)
        rec["hit_id"].append(-1 if hid is None else (n_t - 1 if hid == "goal" else int(hid)))
        rec["tile"].append(int(col["tile_class"]))
        rec["dist2wp"].append(float(info["hero"]["dist2wp"]))
        rec["comfort"].append([float(info["hero"][k]) for k in ("speed_mps",) + sim.COMFORT_KEYS])
        rec["n_nearby"].append(len(col["actors_state"]))
        rec["obs_crc"].append(zlib.crc32(np.ascontiguousarray(obs[0]).tobytes()))
        rec["rgb_crc"].append(zlib.crc32(np.ascontiguousarray(base.observation).tobytes()))
        am = base.map.actor_manager
        acts = list(am.actors["vehicle"]) + list(am.actors["pedestrian"])
        amax = max(amax, len(acts))
        act_states.append([[a._controller.x, a._controller.y, a._controller.yaw, a._controller.v] for a in acts])
        act_tidx.append([a._controller.target_idx for a in acts])
        act_fsm.append([FSM_IDS[a.behavior_state] for a in acts])
        tgt_vis.append([bool(tg.visible) for tg in am.actors["target"]])
        if t % frame_every == 0 or term[0]:
            frames.append(rgb_to_index(base.observation))
            frame_steps.append(t)
        if t % obs_every == 0 or term[0]:
            obs_full.append(np.asarray(obs[0]))
            obs_steps.append(t)
        if term[0] or trunc[0]:
            ei = dict(infos["episode_info"])
            ei = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in ei.items() if not k.startswith("_")}
            ei = {k: (v.item() if isinstance(v, np.generic) else v) for k, v in ei.items()}
            epi = infos["episode"]
            ei["_episode_r"] = float(epi["r"][0])
            ei["_episode_l"] = int(epi["l"][0])
            ep_infos[t] = ei
            if t + 1 < T:
                obs = do_reset(t + 1)
                reset_frames.append(rgb_to_index(base.observation))
                reset_obs.append(np.asarray(obs[0]))

    tmax = max(len(v) for v in tgt_vis)
    A = np.full((T, max(amax, 1), 4), np.nan)
    AT = np.full((T, max(amax, 1)), -1, dtype=np.int32)
    AF = np.full((T, max(amax, 1)), -1, dtype=np.int8)
    TV = np.zeros((T, tmax), dtype=bool)
    for t in range(T):
        n = len(act_states[t])
        if n:
            A[t, :n] = act_states[t]
            AT[t, :n] = act_tidx[t]
            AF[t, :n] = act_fsm[t]
        TV[t, :len(tgt_vis[t])] = tgt_vis[t]
    out = {f"pool_{k}": v for k, v in pack_pool(scenes).items()}
    out.update(
        actions=np.asarray(actions), reset_steps=np.array(reset_steps, dtype=np.int32),
        reset_scene=np.array(reset_scene, dtype=np.int32),
        ego_state=np.array(rec["ego_state"], dtype=np.float64), acc=np.array(rec["acc"]),
        tidx=np.array(rec["tidx"], dtype=np.int32), reward=np.array(rec["reward"]),
        term=np.array(rec["term"]), trunc=np.array(rec["trunc"]), cause=np.array(rec["cause"], dtype=np.int8),
        hit=np.array(rec["hit"], dtype=np.int8), hit_id=np.array(rec["hit_id"], dtype=np.int32),
        tile=np.array(rec["tile"], dtype=np.int8), dist2wp=np.array(rec["dist2wp"]),
        comfort=np.array(rec["comfort"]), n_nearby=np.array(rec["n_nearby"], dtype=np.int32),
        obs_crc=np.array(rec["obs_crc"], dtype=np.uint32), rgb_crc=np.array(rec["rgb_crc"], dtype=np.uint32),
        actor_state=A, actor_tidx=AT, actor_fsm=AF, tgt_visible=TV,
        frames=np.array(frames, dtype=np.uint8), frame_steps=np.array(frame_steps, dtype=np.int32),
        obs_full=np.array(obs_full), obs_steps=np.array(obs_steps, dtype=np.int32),
        reset_frames=np.array(reset_frames, dtype=np.uint8), reset_obs=np.array(reset_obs),
        episode_infos=np.array(json.dumps({str(k): v for k, v in ep_infos.items()}, default=str)),
        env_kwargs=np.array(json.dumps(env_kwargs)),
    )
    if out["obs_full"].dtype == np.float32 and not np.isin(out["obs_full"], (0.0, 1.0)).all():
        out["obs_full"] = out["obs_full"].astype(np.float16)   # weighted history: multiples of 0.25, exact in f16
        out["reset_obs"] = out["reset_obs"].astype(np.float16)
    elif out["obs_full"].dtype == np.float32:  # 0/1 masks: store bit-packed
        out["obs_full"] = np.packbits(out["obs_full"].astype(bool), axis=-1)
        out["reset_obs"] = np.packbits(out["reset_obs"].astype(bool), axis=-1)
        out["obs_packed"] = np.array(True)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    n_term = int(np.sum(out["term"]))
    print(f"{name}: T={T} resets={len(scenes)} terminals={n_term} actors<= {amax} -> {os.path.getsize(path) / 1024:.0f} KiB")
    envs.close()


def cont_actions(seed, T, brake_first=0, gas_bias=0.0):
    rng = np.random.default_rng(seed)
    a = np.stack([rng.uniform(0, 1, T), rng.uniform(-1, 1, T), rng.uniform(0, 1, T)], axis=1).astype(np.float32)
    a[:, 0] = np.clip(a[:, 0] + gas_bias, 0, 1)
    if brake_first:
        a[:brake_first] = np.array([0.0, 0.0, 1.0], dtype=np.float32)
    return a


CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


@case
def rdm_medium_discrete():
    # BASELINE config 1 (shortened): rdm rt_medium_v1, scene_seed 0, 6-class F=4, discrete9, random policy
    acts = np.random.default_rng(0).integers(0, 9, 400)
    opts = lambda ep: dict(scene="rdm", difficulty_id="rt_medium_v1", num_vehicles=16, route_dist_range=(40, 100),  # noqa: E731
                           scene_seed=ep)
    run_case("rdm_medium_discrete", dict(obs_mode="bev_semantic"), opts, acts, frame_every=4)


@case
def lead_brake_continuous():
    # BASELINE config 2 shape: lead_brake levels 1..3 round-robin, continuous actions
    acts = cont_actions(0, 240, gas_bias=0.2)
    opts = lambda ep: dict(scene="lead_brake", level=1 + ep % 3, scene_seed=ep)  # noqa: E731
    run_case("lead_brake_continuous", dict(obs_mode="bev_semantic", action_mode="continuous"), opts, acts,
             frame_every=2)


@case
def jaywalk_levels():
    # ego brakes to a stop so the pedestrian FSM (incl. StopReturn retreat) plays out
    T = 4 * 110
    acts = cont_actions(1, T)
    acts[:, 0] *= 0.15
    for k in range(4):
        acts[k * 110:k * 110 + 70] = np.array([0.0, 0.0, 1.0], dtype=np.float32)
    opts = lambda ep: dict(scene="jaywalk", level=1 + ep % 4, scene_seed=100 + ep)  # noqa: E731
    run_case("jaywalk_levels", dict(obs_mode="bev_semantic", action_mode="continuous"), opts, acts, frame_every=3)


@case
def jaywalk_drive():
    acts = cont_actions(2, 160, gas_bias=0.1)
    opts = lambda ep: dict(scene="jaywalk", level=1 + ep % 4, scene_seed=200 + ep)  # noqa: E731
    run_case("jaywalk_drive", dict(obs_mode="bev_semantic", action_mode="continuous", semantic_mask_ch="7-class"),
             opts, acts, frame_every=2)


@case
def red_light_runner():
    acts = cont_actions(3, 160, gas_bias=0.3)
    acts[:, 1] *= 0.3
    opts = lambda ep: dict(scene="red_light_runner", scene_seed=ep)  # noqa: E731
    run_case("red_light_runner", dict(obs_mode="bev_semantic", action_mode="continuous", semantic_mask_ch="7-class"),
             opts, acts, frame_every=2)


@case
def rdm_shaping_discrete13():
    acts = np.random.default_rng(5).integers(0, 13, 300)
    opts = lambda ep: dict(scene="rdm", difficulty_id="rt_easy_v1", num_vehicles=8, route_dist_range=(30, 80),  # noqa: E731
                           scene_seed=50 + ep)
    run_case("rdm_shaping_discrete13",
             dict(obs_mode="bev_semantic", reward_mode="shaping", action_profile_id="discrete13_v1",
                  semantic_mask_ch="4-class", frame_stack=2),
             opts, acts, frame_every=4)


@case
def rdm_rgb_lookahead():
    # BASELINE config 5 shape: bev_rgb, lookahead_75 camera, max actor density, continuous
    acts = cont_actions(7, 200, gas_bias=0.3)
    acts[:, 2] *= 0.2
    opts = lambda ep: dict(scene="rdm", num_vehicles=50, route_dist_range=(30, 130), scene_seed=70 + ep)  # noqa: E731
    run_case("rdm_rgb_lookahead",
             dict(obs_mode="bev_rgb", action_mode="continuous", ego_anchor_x_frac=0.5, ego_anchor_y_frac=0.75),
             opts, acts, frame_every=3)


@case
def fusion_temporal_masked():
    # vehicle_temporal fusion + fov_masked corner mask, lead_brake level 3 (vehicles in view)
    acts = cont_actions(11, 150, gas_bias=0.1)
    opts = lambda ep: dict(scene="lead_brake", level=3, scene_seed=300 + ep)  # noqa: E731
    run_case("fusion_temporal_masked",
             dict(obs_mode="bev_semantic", action_mode="continuous", temporal_fusion_mode="vehicle_temporal",
                  fov_masked=True), opts, acts, frame_every=3, obs_every=10)


@case
def fusion_weighted():
    acts = cont_actions(12, 150, gas_bias=0.1)
    opts = lambda ep: dict(scene="lead_brake", level=2 + ep % 2, scene_seed=400 + ep)  # noqa: E731
    run_case("fusion_weighted",
             dict(obs_mode="bev_semantic", action_mode="continuous", temporal_fusion_mode="vehicle_weighted",
                  semantic_mask_ch="5-class"), opts, acts, frame_every=3, obs_every=10)


# ---- other map scales.  The unmodified reference resets and steps `rdm` scenes at size 64 and 256 for
# the seeds below (its generators hard-code map_size=128, quirk C-11, so routes keep their 128-scale coordinates; the
# other seeds, the scripted scenarios and sizes 512 / 1024 end in "hero_on_obstacle").
VALID_SEEDS = {64: [0, 2, 3, 5, 6, 10, 11, 12, 16, 18, 19, 20, 21, 22, 23, 26, 29, 32, 34, 36, 37, 38],
               256: [0, 1, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21, 22, 23, 24, 25, 26, 27, 28]}


def _scale_case(name, size, env_kwargs, acts, frame_every):
    seeds = VALID_SEEDS[size]
    opts = lambda ep: dict(scene="rdm", num_vehicles=12, route_dist_range=(30, 100), scene_seed=seeds[ep % len(seeds)])  # noqa: E731
    run_case(name, dict(size=size, **env_kwargs), opts, acts, frame_every=frame_every)


@case
def size256_rdm_discrete():
    acts = np.random.default_rng(21).integers(0, 9, 300)
    acts[40:120] = 1  # a stretch of plain throttle so that the ego gets somewhere
    _scale_case("size256_rdm_discrete", 256, dict(obs_mode="bev_semantic"), acts, 6)


@case
def size256_rdm_gray_lookahead():
    acts = cont_actions(22, 160, gas_bias=0.3)
    acts[:, 2] *= 0.2
    _scale_case("size256_rdm_gray_lookahead", 256,
                dict(obs_mode="bev_rgb", action_mode="continuous", ego_anchor_x_frac=0.5, ego_anchor_y_frac=0.75), acts, 6)


@case
def size64_rdm_discrete():
    acts = np.random.default_rng(23).integers(0, 9, 300)
    acts[40:120] = 1
    _scale_case("size64_rdm_discrete", 64, dict(obs_mode="bev_semantic"), acts, 3)


@case
def size64_rdm_gray_lookahead():
    acts = cont_actions(24, 160, gas_bias=0.3)
    acts[:, 2] *= 0.2
    _scale_case("size64_rdm_gray_lookahead", 64,
                dict(obs_mode="bev_rgb", action_mode="continuous", ego_anchor_x_frac=0.5, ego_anchor_y_frac=0.75), acts, 3)


if __name__ == "__main__":
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    if pat in ("", "map"):
        for size in (128, 64, 256):
            dump_map(size)
    for name, fn in CASES.items():
        if pat in name:
            fn()
