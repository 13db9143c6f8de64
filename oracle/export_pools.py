"""Export scene pools that need the reference's lane graphs (rdm, red_light_runner) by resetting the
UNMODIFIED reference (under oracle/shims) and snapshotting its objects (oracle/gen_golden.py:extract_scene).

TEST / DATA INFRASTRUCTURE, run in the build container only:  python oracle/export_pools.py
Writes carlabev_env_b200/assets/pools/*.npz (packed pools, carlabev_env_b200/pool.py)."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.gen_golden import extract_scene  # noqa: E402  (loads the reference)
from CarlaBEV.config import EnvConfig, RunConfig  # noqa: E402
from CarlaBEV.envs import make_env  # noqa: E402

from carlabev_env_b200.pool import save_pool  # noqa: E402

OUT = os.path.join(ROOT, "carlabev_env_b200", "assets", "pools")


def export(name, n, options_of, env_kwargs=None):
    cfg = RunConfig(env=EnvConfig(render_mode="rgb_array", **(env_kwargs or {})), num_envs=1)
    envs = make_env(cfg)
    base = envs.envs[0].unwrapped
    scenes = []
    t0 = time.time()
    for i in range(n):
        opts = options_of(i)
        envs.reset(options={**opts, "reset_mask": np.array([True])})
        scenes.append(extract_scene(base, opts))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    save_pool(path, scenes)
    acts = [len(s["act_kind"]) for s in scenes]
    print(f"{name}: {n} scenes, actors {min(acts)}..{max(acts)}, ego route {min(len(s['ego_cx']) for s in scenes)}.."
          f"{max(len(s['ego_cx']) for s in scenes)} pts, {os.path.getsize(path) / 1e6:.2f} MB, {time.time() - t0:.0f}s")
    envs.close()


def export_authored(variations=4):
    """The reference's 7 authored scenes (assets/scenes/*.json) x seeded variations
    (scenarios/__init__.py:210-338).  Scripted actors of authored scenes carry no np_rng, so their +-1 px start
    jitter is unseeded in the reference: the pool holds snapshots of one realisation each."""
    import glob
    import json

    from oracle.ref_loader import REFERENCE_ROOT

    files = sorted(glob.glob(os.path.join(REFERENCE_ROOT, "CarlaBEV", "assets", "scenes", "*.json")))
    cfg = RunConfig(env=EnvConfig(render_mode="rgb_array"), num_envs=1)
    envs = make_env(cfg)
    base = envs.envs[0].unwrapped
    scenes, manifest = [], []
    for f in files:
        scenario_id = json.load(open(f))["scenario_id"]
        for v in range(variations):
            opts = dict(scene=scenario_id, config_file=f, variation_enabled=True, variation_seed=v, scene_seed=v)
            envs.reset(options={**opts, "reset_mask": np.array([True])})
            scenes.append(extract_scene(base, opts))
            manifest.append({"config_file": os.path.basename(f), "scenario_id": scenario_id, "variation_seed": v})
    os.makedirs(OUT, exist_ok=True)
    save_pool(os.path.join(OUT, "authored_scenes.npz"), scenes)
    json.dump(manifest, open(os.path.join(OUT, "authored_scenes.json"), "w"), indent=1)
    # the scene files themselves (DATA assets of the reference), bundled for the host loader and its tests
    bundle = {os.path.basename(f): json.load(open(f)) for f in files}
    json.dump(bundle, open(os.path.join(os.path.dirname(OUT), "authored_scene_files.json"), "w"), separators=(",", ":"))
    print(f"authored_scenes: {len(scenes)} scenes from {len(files)} files, actors "
          f"{min(len(s['act_kind']) for s in scenes)}..{max(len(s['act_kind']) for s in scenes)}, "
          f"traffic lights <= {max(len(s['tl_color']) for s in scenes)}")
    envs.close()


OPTION_CASES = [  # reset options beyond the bench configs: presets, parameter overrides, scenario-config files
    dict(scene="jaywalk", level=3, ego_speed=10.0, cross_delay=1.2, pedestrian_speed=1.6, yield_duration=1.2,
         scene_seed=3),                                                            # preset jaywalk_debug
    dict(scene="lead_brake", level=2, ego_speed=12.0, lead_gap=8.0, lead_speed=11.0, brake_delay=2.0,
         brake_strength=4.0, scene_seed=4),                                        # preset lead_brake_debug
    dict(scene="red_light_runner", intersection_index=11, ego_speed=10.0, adv_speed=16.0, scene_seed=5),
    dict(scene="rdm", num_vehicles=25, route_dist_range=[30, 130], scene_seed=6),  # preset rdm_navigation
    dict(scene="rdm", num_vehicles=6, route_dist_range=[30, 90], ego_route_graph="right_lane", scene_seed=7),
    dict(scene="rdm", num_vehicles=6, route_dist_range=[30, 90], ego_route_graph="left_lane", scene_seed=8),
    dict(scene="rdm", num_vehicles=9, route_dist_range=[40, 70], route_seed=5, traffic_seed=9, scene_seed=9,
         ego_target_speed=9.0),
    dict(scene="rdm", difficulty_id="rt_easy_v1", traffic_enabled=True, num_vehicles=8, route_dist_range=[30, 80],
         scene_seed=10),
    dict(scene="lead_brake", scene_seed=11),                                       # level drawn from scenario_rng
    dict(scene="jaywalk", scene_seed=12),
    dict(scene="jaywalk", level=4, anchor_x=848, anchor_y=930, rear_gap=4.0, rear_speed=7.0, cross_offset=1.0,
         scene_seed=13),
    dict(scene="lead_brake", level=3, rear_brake_delay=2.2, left_speed=15.0, rear_gap=5.5, scene_seed=14),
    dict(scene="red_light_runner", anchor_x=300, anchor_y=420, adv_speed=12.0, scene_seed=15),
    dict(config={"type": "scenario_config", "scenario_id": "lead_brake", "level": 3,
                 "anchor": {"x": 850, "y": 950}, "parameters": {"ego_speed": 9.0, "lead_gap": 6.0}}, scene_seed=16),
    dict(config={"scenario": "jaywalk", "kwargs": {"level": 2, "anchor_y": 940, "cross_delay": 2.0}}, scene_seed=17),
    # route-profile filters of the random-navigation reset (src/control/route_profile.py)
    dict(scene="rdm", num_vehicles=3, route_dist_range=[30, 130], route_profile="single_left", scene_seed=20),
    dict(scene="rdm", num_vehicles=3, route_dist_range=[30, 130], route_profile="single_right", scene_seed=21),
    dict(scene="rdm", num_vehicles=3, route_dist_range=[30, 130], route_profile="mostly_straight", scene_seed=22),
    dict(scene="rdm", num_vehicles=2, route_dist_range=[40, 130], min_turns=1, max_turns=2, scene_seed=23),
    dict(scene="rdm", num_vehicles=2, route_dist_range=[40, 130], intersection_required=True, scene_seed=24),
    dict(scene="rdm", num_vehicles=2, route_dist_range=[30, 130], intersection_required=False, scene_seed=25),
    dict(scene="rdm", num_vehicles=2, route_dist_range=[30, 130],
         route_profile_mix={"mostly_straight": 0.5, "single_left": 0.25, "single_right": 0.25}, scene_seed=26),
    dict(scene="rdm", num_vehicles=2, route_dist_range=[30, 130],
         route_profile_mix={"mostly_straight": 0.5, "single_left": 0.25, "single_right": 0.25}, scene_seed=27),
]


def export_options():
    """Snapshots for tests/test_host_logic.py::test_reset_option_variants_match_reference."""
    import json
    import tempfile

    cfg = RunConfig(env=EnvConfig(render_mode="rgb_array"), num_envs=1)
    envs = make_env(cfg)
    base = envs.envs[0].unwrapped
    scenes = []
    for case in OPTION_CASES:
        opts = dict(case)
        if "config" in opts:
            tmp = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
            json.dump(opts.pop("config"), tmp)
            tmp.close()
            opts["config_file"] = tmp.name
        envs.reset(options={**opts, "reset_mask": np.array([True])})
        scenes.append(extract_scene(base, opts))
    out = os.path.join(ROOT, "tests", "golden", "option_scenes.npz")
    save_pool(out, scenes)
    json.dump(OPTION_CASES, open(os.path.join(ROOT, "tests", "golden", "option_scenes.json"), "w"), indent=1)
    print(f"option_scenes: {len(scenes)} snapshots, {os.path.getsize(out) / 1e3:.0f} kB")
    envs.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else ""
    if which in ("", "rdm_hard"):
        # BASELINE configs[2]: rdm rt_hard_v1 (25 vehicles, routes 50-130 m), scene_seed = i
        export("rdm_rt_hard_v1", 96, lambda i: dict(scene="rdm", difficulty_id="rt_hard_v1", num_vehicles=25,
                                                    route_dist_range=(50, 130), scene_seed=i))
    if which in ("", "rdm_medium"):
        # BASELINE configs[0]: rdm rt_medium_v1 (16 vehicles, routes 40-100 m)
        export("rdm_rt_medium_v1", 32, lambda i: dict(scene="rdm", difficulty_id="rt_medium_v1", num_vehicles=16,
                                                      route_dist_range=(40, 100), scene_seed=i))
    if which in ("", "rdm_dense"):
        # BASELINE configs[4]: max actor density (num_vehicles = max_vehicles = 50), lookahead_75 camera
        export("rdm_dense_50", 48, lambda i: dict(scene="rdm", num_vehicles=50, route_dist_range=(30, 130), scene_seed=i),
               env_kwargs=dict(ego_anchor_x_frac=0.5, ego_anchor_y_frac=0.75))
    if which in ("", "authored"):
        export_authored()
    if which in ("", "options"):
        export_options()
    if which in ("", "red_light"):
        # BASELINE configs[3]: red_light_runner (first valid 4-way intersection); the adversary's start jitter is
        # unseeded in the reference (quirk C-10), so these are snapshots, not re-derivable from the seed
        export("red_light_runner", 32, lambda i: dict(scene="red_light_runner", scene_seed=i))
