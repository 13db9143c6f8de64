"""Fuzz the oracle's step against the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE:  python oracle/fuzz_steps.py [n_episodes] [seed] [max_steps] [pursuit|random] [scales]
(`scales`: EnvConfig.size is drawn from {64, 128, 256}; off the 128 scale only `rdm` scenes exist, quirk C-11)
Random env configuration (action profile, reward mode, mask channels, camera anchor, fov mask), random reset
options (oracle/fuzz_scenes.random_options) and random actions; the reference env and OracleEnv are stepped in
lock-step and must agree exactly on the observation, reward, flags, ego state and every actor state of every step.
This widens the pin of the oracle beyond the nine recorded goldens (tests/golden/)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.fuzz_scenes import random_options  # noqa: E402
from oracle.gen_golden import extract_scene  # noqa: E402  (loads the reference)
from CarlaBEV.config import EnvConfig, RunConfig  # noqa: E402
from CarlaBEV.envs import make_env  # noqa: E402

from oracle.env import OracleEnv  # noqa: E402
from carlabev_env_b200.vector_env import load_town01_map  # noqa: E402


def random_env(rng):
    kw = {}
    mode = rng.choice(["continuous", "discrete9_v1", "discrete13_v1"])
    if mode == "continuous":
        kw["action_mode"] = "continuous"
    else:
        kw["action_mode"] = "discrete"
        kw["action_profile_id"] = str(mode)
    if rng.random() < 0.3:
        kw["reward_mode"] = "shaping"
    kw["semantic_mask_ch"] = str(rng.choice(["6-class", "7-class", "5-class", "4-class", "binary"]))
    if rng.random() < 0.3:
        kw["ego_anchor_x_frac"], kw["ego_anchor_y_frac"] = 0.5, 0.75
    if rng.random() < 0.2:
        kw["fov_masked"] = True
    if rng.random() < 0.2:
        kw["obs_mode"] = "bev_rgb"
    elif rng.random() < 0.3 and kw["semantic_mask_ch"] != "binary":
        kw["temporal_fusion_mode"] = str(rng.choice(["vehicle_temporal", "vehicle_weighted"]))
    if rng.random() < 0.3:
        kw["frame_stack"] = int(rng.choice([3, 5, 6]))
    if rng.random() < 0.3:
        kw["obs_size"] = tuple(int(v) for v in rng.choice([(84, 84), (64, 64), (128, 128), (112, 100), (48, 64)]))
    if SCALES and rng.random() < 0.7:
        kw["size"] = int(rng.choice([64, 256]))
        if rng.random() < 0.5:
            kw["obs_size"] = tuple(int(v) for v in rng.choice([(96, 96), (24, 24), (128, 128), (200, 200), (100, 60), (64, 64)]))
    return kw


PURSUIT = len(sys.argv) > 4 and sys.argv[4] == "pursuit"
SCALES = len(sys.argv) > 5 and sys.argv[5] == "scales"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 160
    maps = {}
    bad = steps = 0
    causes, by_size = {}, {}
    for ep in range(n):
        kw = random_env(rng)
        o = random_options(rng)
        o.pop("route_profile", None)
        size = kw.get("size", 128)
        if size != 128:
            o = {"scene": "rdm", "scene_seed": int(rng.integers(0, 1_000_000)), "num_vehicles": int(rng.integers(0, 16)),
                 "route_dist_range": [30, 100]}
        cls = maps.setdefault(size, load_town01_map(size))
        cfg = RunConfig(env=EnvConfig(render_mode="rgb_array", **kw), num_envs=1)
        envs = make_env(cfg)
        base = envs.envs[0].unwrapped
        try:
            obs, _ = envs.reset(options={**o, "reset_mask": np.array([True])})
        except Exception:  # noqa: BLE001  (route search gave up: covered by fuzz_scenes)
            envs.close()
            continue
        scene = extract_scene(base, o)
        by_size[size] = by_size.get(size, 0) + 1
        am = kw.get("action_mode", "discrete")
        ora = OracleEnv(cls, obs_mode="bev_semantic" if kw.get("obs_mode", "bev_semantic") == "bev_semantic" else "bev_gray",
                        semantic_mask_ch=kw["semantic_mask_ch"], action_mode=am,
                        action_profile=kw.get("action_profile_id") or ("continuous_gsb_v1" if am == "continuous" else "discrete9_v1"),
                        reward_mode=kw.get("reward_mode", "carl"),
                        anchor=(kw.get("ego_anchor_x_frac", 0.5), kw.get("ego_anchor_y_frac", 0.5)),
                        fov_masked=kw.get("fov_masked", False), frame_stack=kw.get("frame_stack", 4),
                        obs_size=kw.get("obs_size", (96, 96)), temporal_fusion_mode=kw.get("temporal_fusion_mode", "stack"),
                        size=size)
        o0 = ora.reset(scene)
        what = None
        if not np.array_equal(np.asarray(obs[0]), o0):
            what = "reset observation"
        n_act = envs.single_action_space.n if am == "discrete" else 0
        style = rng.random()
        pursuit = PURSUIT and rng.random() < 0.7   # steer towards the route so that episodes get long / succeed
        table = np.asarray(ora.table, dtype=np.float64) if am == "discrete" else None
        for t in range(max_steps):
            if what:
                break
            if pursuit:
                h = base.map.hero
                k = min(int(h.target_idx) + 2, len(h.cx) - 1)
                err = np.arctan2(h.cy[k] - h.y, h.cx[k] - h.x) - h.yaw
                err = (err + np.pi) % (2 * np.pi) - np.pi
                want = np.array([0.7 if h.v < 25.0 else 0.0, np.clip(2.0 * err, -1, 1) + rng.normal(0, 0.05),
                                 1.0 if h.v > 32.0 else 0.0])
                if am == "continuous":
                    a = np.clip(want, [0, -1, 0], [1, 1, 1]).astype(np.float32)
                else:
                    a = int(np.argmin(((table - want) ** 2).sum(1)))
            elif am == "continuous":
                a = np.array([rng.uniform(0.2 if style < 0.6 else 0, 1), rng.uniform(-1, 1) * (0.3 if style < 0.6 else 1.0),
                              rng.uniform(0, 1) * (rng.random() < 0.2)], dtype=np.float32)
            else:
                a = int(rng.integers(0, n_act))
            obs, rew, term, trunc, _ = envs.step([a])
            oo, orew, oterm, otrunc, _ = ora.step(a)
            steps += 1
            hero = base.map.hero
            e = ora.sim.ego
            acts = list(base.map.actor_manager.actors["vehicle"]) + list(base.map.actor_manager.actors["pedestrian"])
            ref_a = np.array([[c._controller.x, c._controller.y, c._controller.yaw, c._controller.v] for c in acts]).reshape(-1, 4)
            got_a = np.array([[c.x, c.y, c.yaw, c.v] for c in ora.sim.actors]).reshape(-1, 4)
            if not np.array_equal(np.asarray(obs[0]), oo):
                what = f"observation at step {t}"
            elif float(rew[0]) != float(orew):
                what = f"reward at step {t}: {float(rew[0])!r} vs {float(orew)!r}"
            elif bool(term[0]) != bool(oterm) or bool(trunc[0]) != bool(otrunc):
                what = f"flags at step {t}"
            elif [hero.x, hero.y, hero.yaw, hero.v] != [e.x, e.y, e.yaw, e.v]:
                what = f"ego state at step {t}"
            elif ref_a.shape != got_a.shape or not np.array_equal(ref_a, got_a):
                what = f"actor states at step {t}"
            if term[0] or trunc[0]:
                c = base.current_info["reward"]["cause"]
                causes[c] = causes.get(c, 0) + 1
                break
        if what:
            bad += 1
            print("MISMATCH", what, kw, o)
        envs.close()
    print(f"endings: {causes}; episodes per EnvConfig.size: {by_size}")
    print(f"{n} episodes, {steps} steps compared, {bad} mismatching episodes")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
