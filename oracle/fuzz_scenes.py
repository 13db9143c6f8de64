"""Fuzz the host scene generators against the UNMODIFIED reference (build container only).

TEST / DATA INFRASTRUCTURE:  python oracle/fuzz_scenes.py [n_cases] [seed]
Resets the reference with random reset options / seeds, snapshots its objects (gen_golden.extract_scene) and
requires carlabev_env_b200.scenes.build_scene to return the same pool entry bit for bit (the unseeded start jitter
of red_light_runner's adversary excepted)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.gen_golden import extract_scene  # noqa: E402  (loads the reference)
from CarlaBEV.config import EnvConfig, RunConfig  # noqa: E402
from CarlaBEV.envs import make_env  # noqa: E402

from carlabev_env_b200 import scenes as S  # noqa: E402
from carlabev_env_b200.vector_env import load_town01_map  # noqa: E402


AUTHORED = ("jaywalk-01.01.json", "jaywalk-01.02.json", "jaywalk-01.03.json", "leadbrake-01.01.json",
            "leadbrake-01.02.json", "leadbrake-01.03.json", "redlightrunner-01.01.json")


def random_options(rng, authored=False):
    if authored and rng.random() < 0.35:   # authored scene file of the reference, with / without seeded variation
        from oracle.ref_loader import REFERENCE_ROOT

        o = {"config_file": os.path.join(REFERENCE_ROOT, "CarlaBEV", "assets", "scenes", str(rng.choice(AUTHORED))),
             "scene_seed": int(rng.integers(0, 1000))}
        if rng.random() < 0.7:
            o["variation_enabled"] = True
            if rng.random() < 0.8:
                o["variation_seed"] = int(rng.integers(0, 10000))
        return o
    kind = rng.choice(["rdm", "rdm", "rdm", "lead_brake", "jaywalk", "red_light_runner"])
    o = {"scene": str(kind), "scene_seed": int(rng.integers(0, 1_000_000))}
    if kind == "rdm":
        if rng.random() < 0.4:
            o["difficulty_id"] = str(rng.choice(["rt_no_traffic_v1", "rt_easy_v1", "rt_medium_v1", "rt_hard_v1"]))
        else:
            o["num_vehicles"] = int(rng.integers(0, 30))
            lo = int(rng.integers(25, 70))
            o["route_dist_range"] = [lo, lo + int(rng.integers(20, 70))]
        if rng.random() < 0.3:
            o["ego_route_graph"] = str(rng.choice(["full_vehicle", "right_lane", "left_lane"]))
        if rng.random() < 0.2:
            o["ego_target_speed"] = float(rng.uniform(6, 14))
        if rng.random() < 0.2:
            o["route_profile"] = str(rng.choice(["mostly_straight", "single_left", "single_right", "any"]))
        if rng.random() < 0.1:
            o["traffic_seed"] = int(rng.integers(0, 1000))
    elif kind in ("lead_brake", "jaywalk"):
        if rng.random() < 0.7:
            o["level"] = int(rng.integers(1, 4 if kind == "lead_brake" else 5))
        for key, lo, hi in (("ego_speed", 6.0, 16.0), ("rear_gap", 3.0, 7.0), ("rear_speed", 4.0, 12.0)):
            if rng.random() < 0.25:
                o[key] = float(rng.uniform(lo, hi))
        if kind == "lead_brake" and rng.random() < 0.25:
            o["lead_gap"] = float(rng.uniform(4.0, 14.0))
        if kind == "jaywalk" and rng.random() < 0.25:
            o["cross_delay"] = float(rng.uniform(0.5, 3.0))
    else:
        if rng.random() < 0.5:
            o["intersection_index"] = int(rng.integers(0, 16))
        if rng.random() < 0.3:
            o["adv_speed"] = float(rng.uniform(8, 18))
    return o


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 128  # EnvConfig.size: 64 / 128 / 256
    cfg = RunConfig(env=EnvConfig(render_mode="rgb_array", size=size), num_envs=1)
    envs = make_env(cfg)
    base = envs.envs[0].unwrapped
    cls = load_town01_map(size)
    pad = {64: 91, 128: 182, 256: 363}[size]
    bad = errors = 0
    for case in range(n):
        o = random_options(rng, authored=size == 128)
        ref_err = got_err = None
        try:
            envs.reset(options={**o, "reset_mask": np.array([True])})
            ref = extract_scene(base, o)
        except Exception as ex:  # noqa: BLE001
            ref_err = type(ex).__name__
        try:
            got = S.build_scene(o, cls_map=cls, pad=pad)
        except Exception as ex:  # noqa: BLE001
            got_err = type(ex).__name__
        if ref_err or got_err:
            errors += 1
            if ref_err != got_err:
                bad += 1
                print("ERROR MISMATCH", o, ref_err, got_err)
            continue
        diff = [k for k in ref if k not in ("kind", "level")
                and (np.asarray(ref[k]).shape != np.asarray(got[k]).shape or not np.array_equal(ref[k], got[k]))]
        if (o.get("scene") == "red_light_runner" or "config_file" in o) and diff == ["act_state0"]:
            d = np.abs(ref["act_state0"] - got["act_state0"])
            if d[:, :2].max() <= 2.0 and d[:, 3].max() == 0.0:
                diff = []
        if diff:
            bad += 1
            print("MISMATCH", o, diff)
    print(f"size {size}: {n} cases, {errors} raised on both sides, {bad} mismatches")
    envs.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
