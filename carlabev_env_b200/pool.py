"""Scene pool: host-side container for the device-resident pool of pre-generated scenes.

A scene is the post-reset snapshot of one reference episode start:
ego pose + smoothed route + raw reward route, every scripted actor's post-reset state,
smoothed route, authored route and behaviour parameters, and the traffic-light strips.
The reference produces exactly this state in CarlaBEV.reset (envs/carlabev.py:96-148 ->
scenes/scene.py:61-88 -> managers/actor_manager.py:36-110 -> actors/actor.py:86-108).

`pack_pool` concatenates a list of scene dicts into flat arrays + offset tables -- the
layout `cbev_upload_scene_pool` copies to HBM unchanged (include/cbev.h).
"""
from __future__ import annotations

import numpy as np

SCENE_KINDS = ("rdm", "lead_brake", "jaywalk", "red_light_runner", "authored")

_PER_SCENE = (
    ("ego_state0", np.float64, (4,)),
    ("ego_target_speed", np.float64, ()),
    ("ego_tidx0", np.int32, ()),
    ("route_length_m", np.float64, ()),
    ("len_ego_route", np.float64, ()),
    ("num_vehicles", np.int32, ()),
    ("kind", np.int32, ()),
    ("level", np.int32, ()),
    ("seed", np.int64, ()),
)
_PER_ACTOR = (
    ("act_kind", np.uint8, ()),
    ("act_state0", np.float64, (4,)),
    ("act_tidx0", np.int32, ()),
    ("act_cruise_px", np.float64, ()),
    ("act_cruise_mps", np.float64, ()),
    ("act_beh", np.uint8, ()),
    ("act_beh_p", np.float64, (4,)),
)


def empty_scene() -> dict:
    s = {k: np.zeros(shape, dtype=dt) for k, dt, shape in _PER_SCENE}
    for k in ("ego_cx", "ego_cy", "ego_cyaw", "act_cx", "act_cy", "act_cyaw", "act_raw_x", "act_raw_y"):
        s[k] = np.zeros(0, dtype=np.float64)
    for k in ("rew_rx", "rew_ry"):
        s[k] = np.zeros(0, dtype=np.int32)
    for k, dt, shape in _PER_ACTOR:
        s[k] = np.zeros((0,) + shape, dtype=dt)
    s["act_route_off"] = np.zeros(1, dtype=np.int32)
    s["act_raw_off"] = np.zeros(1, dtype=np.int32)
    s["tl_rect"] = np.zeros((0, 4), dtype=np.int32)
    s["tl_color"] = np.zeros(0, dtype=np.uint8)
    return s


def pack_pool(scenes: list[dict]) -> dict:
    """List of scene dicts -> dict of flat arrays (the HBM layout)."""
    n = len(scenes)
    out = {"n_scenes": np.int32(n)}
    for k, dt, shape in _PER_SCENE:
        out[k] = np.array([np.asarray(s.get(k, 0), dtype=dt) for s in scenes], dtype=dt).reshape((n,) + shape)

    def offsets(lengths):
        return np.concatenate(([0], np.cumsum(lengths))).astype(np.int32)

    def cat(key, dt, tail=()):
        parts = [np.asarray(s[key], dtype=dt).reshape((-1,) + tail) for s in scenes]
        return np.concatenate(parts) if parts else np.zeros((0,) + tail, dtype=dt)

    out["ego_off"] = offsets([len(s["ego_cx"]) for s in scenes])
    for k in ("ego_cx", "ego_cy", "ego_cyaw"):
        out[k] = cat(k, np.float64)
    out["rew_off"] = offsets([len(s["rew_rx"]) for s in scenes])
    for k in ("rew_rx", "rew_ry"):
        out[k] = cat(k, np.int32)
    out["actor_off"] = offsets([len(s["act_kind"]) for s in scenes])
    for k, dt, shape in _PER_ACTOR:
        out[k] = cat(k, dt, shape)
    # per-actor route offsets are global into the concatenated route arrays; one extra
    # closing entry per scene keeps [actor_off[i] + i .. actor_off[i+1] + i] self-contained.
    route_off, raw_off = [], []
    rbase = wbase = 0
    for s in scenes:
        ro = np.asarray(s["act_route_off"], dtype=np.int64)
        wo = np.asarray(s["act_raw_off"], dtype=np.int64)
        route_off.append(ro[:-1] + rbase)
        raw_off.append(wo[:-1] + wbase)
        rbase += int(ro[-1])
        wbase += int(wo[-1])
    out["act_route_off"] = np.concatenate(route_off + [np.array([rbase])]).astype(np.int32)
    out["act_raw_off"] = np.concatenate(raw_off + [np.array([wbase])]).astype(np.int32)
    for k in ("act_cx", "act_cy", "act_cyaw", "act_raw_x", "act_raw_y"):
        out[k] = cat(k, np.float64)
    out["tl_off"] = offsets([len(s["tl_color"]) for s in scenes])
    out["tl_rect"] = cat("tl_rect", np.int32, (4,))
    out["tl_color"] = cat("tl_color", np.uint8)
    return out


def unpack_pool(pool) -> list[dict]:
    """Inverse of pack_pool (accepts a dict or an open np.load handle)."""
    # an open NpzFile decompresses the WHOLE array on every pool[key]: materialise each array exactly once
    # (a 4096-scene pool took tens of minutes to unpack otherwise -- the round-1 "host-side pool bottleneck")
    pool = {k: np.asarray(pool[k]) for k in (pool.files if hasattr(pool, "files") else pool.keys())}
    n = int(pool["n_scenes"])
    scenes = []
    aro, awo = pool["act_route_off"], pool["act_raw_off"]
    for i in range(n):
        s = {k: np.array(pool[k][i]) for k, _, _ in _PER_SCENE}
        lo, hi = int(pool["ego_off"][i]), int(pool["ego_off"][i + 1])
        for k in ("ego_cx", "ego_cy", "ego_cyaw"):
            s[k] = np.array(pool[k][lo:hi])
        lo, hi = int(pool["rew_off"][i]), int(pool["rew_off"][i + 1])
        for k in ("rew_rx", "rew_ry"):
            s[k] = np.array(pool[k][lo:hi])
        a0, a1 = int(pool["actor_off"][i]), int(pool["actor_off"][i + 1])
        for k, _, _ in _PER_ACTOR:
            s[k] = np.array(pool[k][a0:a1])
        ro = np.array(aro[a0:a1 + 1], dtype=np.int64)
        wo = np.array(awo[a0:a1 + 1], dtype=np.int64)
        for k in ("act_cx", "act_cy", "act_cyaw"):
            s[k] = np.array(pool[k][ro[0]:ro[-1]])
        for k in ("act_raw_x", "act_raw_y"):
            s[k] = np.array(pool[k][wo[0]:wo[-1]])
        s["act_route_off"] = (ro - ro[0]).astype(np.int32)
        s["act_raw_off"] = (wo - wo[0]).astype(np.int32)
        t0, t1 = int(pool["tl_off"][i]), int(pool["tl_off"][i + 1])
        s["tl_rect"] = np.array(pool["tl_rect"][t0:t1]).reshape(-1, 4)
        s["tl_color"] = np.array(pool["tl_color"][t0:t1])
        scenes.append(s)
    return scenes


def save_pool(path, scenes: list[dict], compress: bool = True) -> None:
    (np.savez_compressed if compress else np.savez)(path, **pack_pool(scenes))


def load_pool(path) -> list[dict]:
    with np.load(path) as z:
        return unpack_pool(z)


# ---- pools shipped with the package (exported from the reference by oracle/export_pools.py) -------------
SHIPPED_POOLS = {
    # name: (file, reset options the entries were generated with; entry i has scene_seed = i)
    "rdm_rt_hard_v1": dict(scene="rdm", difficulty_id="rt_hard_v1", num_vehicles=25, route_dist_range=(50, 130)),
    "rdm_rt_medium_v1": dict(scene="rdm", difficulty_id="rt_medium_v1", num_vehicles=16, route_dist_range=(40, 100)),
    "rdm_dense_50": dict(scene="rdm", num_vehicles=50, route_dist_range=(30, 130)),
    "red_light_runner": dict(scene="red_light_runner"),
    # the reference's 7 authored scenes (assets/scenes/*.json) x 4 seeded variations; see authored_manifest()
    "authored_scenes": dict(config_file="*.json", variation_enabled=True),
}


def authored_manifest() -> list[dict]:
    """[{config_file, scenario_id, variation_seed}] for the entries of the `authored_scenes` pool."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "pools", "authored_scenes.json")) as f:
        return json.load(f)


def load_shipped_pool(name: str) -> list[dict]:
    import os

    if name not in SHIPPED_POOLS:
        raise KeyError(f"Unknown shipped pool {name!r}. Available: {', '.join(sorted(SHIPPED_POOLS))}")
    return load_pool(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "pools", f"{name}.npz"))


_GENERATION_KEYS = ("ego_route_graph", "ego_target_speed", "route_profile", "route_profile_mix", "min_turns", "max_turns",
                    "intersection_required", "max_route_attempts", "route_seed", "traffic_seed", "scenario_seed",
                    "intersection_index", "anchor_x", "anchor_y", "ego_speed", "adv_speed", "max_reset_attempts")


def shipped_pool_for(options: dict, max_vehicles: int = 50) -> str | None:
    """Name of the shipped pool whose entries the reference generates for exactly these reset options (entry i <->
    scene_seed i), if any.  The match is on the EFFECTIVE generation parameters (scene_generator.py:95-107:
    `num_vehicles` defaults to EnvConfig.max_vehicles, `route_dist_range` to [30, 100]; a bare `difficulty_id` is
    context metadata), and any option that changes the generated scene rules the snapshots out."""
    scene = options.get("scene", "rdm")
    if options.get("config_file") or str(scene).endswith(".json"):
        return "authored_scenes"
    if any(options.get(k) is not None for k in _GENERATION_KEYS):
        if not (options.get("ego_route_graph") in (None, "full_vehicle")
                and all(options.get(k) is None for k in _GENERATION_KEYS if k != "ego_route_graph")):
            return None
    if scene == "red_light_runner":
        return "red_light_runner"
    if scene == "rdm" and options.get("traffic_enabled", True):
        nv = int(options.get("num_vehicles", max_vehicles))
        rng = tuple(int(v) for v in options.get("route_dist_range", (30, 100)))
        for name in ("rdm_rt_hard_v1", "rdm_rt_medium_v1", "rdm_dense_50"):
            ref = SHIPPED_POOLS[name]
            if nv == ref["num_vehicles"] and rng == tuple(ref["route_dist_range"]):
                return name
    return None
