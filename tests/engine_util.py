"""Shared driver: replay a golden case through the CUDA engine (C ABI via ctypes) and compare."""
from __future__ import annotations

import numpy as np

from golden_util import Golden, crc, load_map

CAUSES = (None, "ckpt", "collision", "success", "out_of_bounds", "off_road", "max_actions", "unknown")


def make_engine_for(g: Golden, num_envs=1, autoreset=0, ring_slots=None, raw_rgb=False, seed=0, trajectory_steps=1024):
    import torch  # noqa: F401

    from carlabev_env_b200 import engine as E
    from carlabev_env_b200.config import ACTION_PROFILES

    kw = g.env_kwargs
    obs_mode = kw.get("obs_mode", "bev_semantic")
    action_mode = kw.get("action_mode", "discrete")
    profile = kw.get("action_profile_id") or ("continuous_gsb_v1" if action_mode == "continuous" else "discrete9_v1")
    eng = E.Engine(
        num_envs,
        obs_mode=E.OBS_SEMANTIC if obs_mode == "bev_semantic" else (E.OBS_RGB if raw_rgb else E.OBS_GRAY),
        mask_mode=kw.get("semantic_mask_ch", "6-class"),
        frame_stack=kw.get("frame_stack", 4),
        ring_slots=ring_slots,
        ring_budget_bytes=64 << 20,
        action_mode=E.ACTION_DISCRETE if action_mode == "discrete" else E.ACTION_CONTINUOUS,
        discrete_table=ACTION_PROFILES[profile].get("discrete_actions"),
        reward_mode=E.REWARD_CARL if kw.get("reward_mode", "carl") == "carl" else E.REWARD_SHAPING,
        autoreset=autoreset,
        anchor=(kw.get("ego_anchor_x_frac", 0.5), kw.get("ego_anchor_y_frac", 0.5)),
        max_actors=int(max(1, np.diff(g.pool["actor_off"]).max())),
        seed=seed,
        trajectory_steps=trajectory_steps,
        size=g.size,
    )
    eng.upload_map(load_map(g.size))
    eng.upload_pool(g.pool)
    eng.keep_fov(True)
    if kw.get("fov_masked"):
        from carlabev_env_b200.fovmask import corner_mask

        eng.upload_fov_mask(corner_mask(g.size, 0.5))
    fusion = kw.get("temporal_fusion_mode", "stack")
    eng.current_obs = (lambda: eng.obs()) if fusion == "stack" else (lambda: eng.fuse(fusion))
    return eng


def replay_golden(name, pose_rtol=1e-9, check_frames=True, max_report=12, trajectory_steps=1024, debug_flags=0):
    """Free-running replay of one golden case on the GPU; returns a list of mismatch strings.

    trajectory_steps = 0 steps the scripted actors live; > 0 reads their poses from the per-scene tables rolled
    out at pool upload for that many steps and continues live afterwards (target indices / FSM states of the
    actors are only observable through cbev_get_state while they are stepped live)."""
    import torch

    from carlabev_env_b200 import engine as E

    g = Golden(name)
    eng = make_engine_for(g, trajectory_steps=trajectory_steps)
    eng.set_debug_flags(debug_flags)
    dev = eng.device
    since_reset = 0
    resets = dict(zip(g["reset_steps"].tolist(), g["reset_scene"].tolist()))
    fsteps = {int(s): i for i, s in enumerate(g["frame_steps"])}
    osteps = {int(s): i for i, s in enumerate(g["obs_steps"])}
    obs_full = g.full_obs()
    reset_obs = g.full_obs("reset_obs")
    amax = g["actor_state"].shape[1]
    bad = []
    H = {k: i for i, k in enumerate(E.HERO_FIELDS)}
    discrete = eng.action_mode == E.ACTION_DISCRETE
    ri = 0
    first = True
    for t in range(g.T):
        if t in resets:
            ids = torch.tensor([resets[t]], dtype=torch.int32)
            eng.reset(ids, None if first else np.array([True]))
            obs = eng.current_obs()
            first = False
            since_reset = 0
            fr = eng.fov()[0].cpu().numpy()
            if not np.array_equal(fr, g["reset_frames"][ri]):
                bad.append(f"t={t} reset frame differs in {(fr != g['reset_frames'][ri]).sum()} px")
            if not np.array_equal(obs[0].cpu().numpy(), reset_obs[ri]):
                bad.append(f"t={t} reset obs differs")
            ri += 1
        a = g["actions"][t]
        at = torch.tensor([int(a)], dtype=torch.int64, device=dev) if discrete else \
            torch.tensor(np.asarray(a, dtype=np.float32)[None], device=dev)
        eng.step(at)
        since_reset += 1
        live = since_reset > trajectory_steps
        hero = eng.hero[0].cpu().numpy()
        st = hero[[H["x"], H["y"], H["yaw"], H["v"]]]
        ref = g["ego_state"][t]
        if not np.allclose(st, ref, rtol=pose_rtol, atol=1e-9):
            bad.append(f"t={t} ego state {st} vs {ref}")
        if abs(hero[H["acc"]] - g["acc"][t]) > 1e-9:
            bad.append(f"t={t} acc {hero[H['acc']]} vs {g['acc'][t]}")
        for k, gk in (("target_idx", "tidx"), ("hit", "hit"), ("hit_id", "hit_id"), ("tile_class", "tile"),
                      ("n_nearby", "n_nearby")):
            if int(hero[H[k]]) != int(g[gk][t]):
                bad.append(f"t={t} {k} {int(hero[H[k]])} vs {int(g[gk][t])}")
        if abs(hero[H["dist2wp"]] - g["dist2wp"][t]) > 1e-9 * max(1.0, abs(g["dist2wp"][t])):
            bad.append(f"t={t} dist2wp")
        cf = hero[[H[k] for k in ("speed_mps", "accel_long", "accel_lat", "jerk_long", "jerk_lat", "yaw_rate", "yaw_acc")]]
        if not np.allclose(cf, g["comfort"][t], rtol=1e-7, atol=1e-7):
            bad.append(f"t={t} comfort {cf} vs {g['comfort'][t]}")
        r = float(eng.reward[0])
        if abs(r - g["reward"][t]) > 1e-9:
            bad.append(f"t={t} reward {r} vs {g['reward'][t]}")
        term, trunc, cause = bool(eng.terminated[0]), bool(eng.truncated[0]), int(eng.cause[0])
        if term != bool(g["term"][t]) or trunc != bool(g["trunc"][t]) or cause != int(g["cause"][t]):
            bad.append(f"t={t} flags {term, trunc, cause} vs {bool(g['term'][t]), bool(g['trunc'][t]), int(g['cause'][t])}")
        ego, act = eng.get_state(amax)
        ga = g["actor_state"][t]
        n = int(np.sum(~np.isnan(ga[:, 0])))
        if n and not np.allclose(act[0, :n, :4], ga[:n], rtol=pose_rtol, atol=1e-9):
            bad.append(f"t={t} actor state max abs diff {np.abs(act[0, :n, :4] - ga[:n]).max()}")
        if live and n and not np.array_equal(act[0, :n, 4].astype(int), g["actor_tidx"][t][:n]):
            bad.append(f"t={t} actor tidx")
        fsm = g["actor_fsm"][t][:n]
        if live and n and not np.array_equal(act[0, :n, 5].astype(int)[fsm > 0], fsm[fsm > 0]):
            bad.append(f"t={t} actor fsm {act[0, :n, 5].astype(int)} vs {fsm}")
        if check_frames and t in fsteps:
            fr = eng.fov()[0].cpu().numpy()
            if not np.array_equal(fr, g["frames"][fsteps[t]]):
                d = np.argwhere(fr != g["frames"][fsteps[t]])
                bad.append(f"t={t} frame differs in {len(d)} px, first {d[:3].tolist()}")
        obs = eng.current_obs()[0].cpu().numpy()
        if crc(obs) != int(g["obs_crc"][t]):
            msg = f"t={t} obs crc differs"
            if t in osteps:
                msg += f" ({int((obs != obs_full[osteps[t]]).sum())} values)"
            bad.append(msg)
        if term:
            gi = g.episode_infos[t]
            ep = dict(zip(E.EPISODE_FIELDS, eng.episode[0].cpu().numpy()))
            if CAUSES[int(ep["cause"])] != gi["termination"] or int(ep["length"]) != gi["length"]:
                bad.append(f"t={t} episode termination/length {ep['cause'], ep['length']} vs {gi['termination'], gi['length']}")
            for k in ("return", "mean_speed", "mean_abs_accel_long", "mean_abs_accel_lat", "mean_abs_jerk_long",
                      "mean_abs_jerk_lat", "mean_abs_yaw_rate", "mean_abs_yaw_acc", "comfort_violation_rate",
                      "harsh_brake_rate", "num_vehicles", "len_ego_route"):
                if abs(ep[k] - float(gi[k])) > 1e-7 * max(1.0, abs(float(gi[k]))):
                    bad.append(f"t={t} episode_info[{k}] {ep[k]} vs {gi[k]}")
        if len(bad) >= max_report:
            break
    eng.close()
    return bad
