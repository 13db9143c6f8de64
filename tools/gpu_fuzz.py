"""Engine vs oracle under random configurations with route-pursuit driving (long episodes, success / checkpoint
endings).  Run on the B200 box:  python tools/gpu_fuzz.py [n_rounds] [seed] [steps] [scales]
`scales` = 1 also draws EnvConfig.size from {64, 128, 256} and the observation size from a list (k_render_any).
Same comparison as tests/test_gpu_engine.py::test_random_configurations_against_the_oracle, more of it."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")


def main():
    import torch

    from golden_util import load_map
    from carlabev_env_b200 import engine as E
    from carlabev_env_b200 import scenes as S
    from carlabev_env_b200.config import ACTION_PROFILES
    from carlabev_env_b200.fovmask import corner_mask
    from carlabev_env_b200.pool import pack_pool
    from oracle import raster
    from oracle.env import OracleEnv

    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    scales = len(sys.argv) > 4 and sys.argv[4] == "1"
    bad = compared = 0
    endings = {}
    for rd in range(rounds):
        profile = str(rng.choice(["continuous_gsb_v1", "discrete9_v1", "discrete13_v1"]))
        continuous = profile == "continuous_gsb_v1"
        reward = "shaping" if rng.random() < 0.4 else "carl"
        mask = str(rng.choice(["6-class", "7-class", "5-class", "4-class", "binary"]))
        anchor = (0.5, 0.75) if rng.random() < 0.4 else (0.5, 0.5)
        fov_masked, gray = bool(rng.random() < 0.3), bool(rng.random() < 0.25)
        frame_stack = int(rng.choice([3, 4, 4, 5]))
        size, obs_size, raw = 128, (96, 96), False
        traj, nveh_hi, fuse = 1024, 12, None
        if scales:
            # actor stepping: trajectory tables, a table -> live hand-over inside the episode, or live stepping (with
            # more than 8 actor slots: the 32-lane groups of k_move); temporal fusion of the vehicle channel
            traj = int(rng.choice([1024, 1024, 40, 0]))
            nveh_hi = int(rng.choice([12, 28]))
            if rng.random() < 0.25 and mask != "binary":
                fuse = str(rng.choice(["vehicle_temporal", "vehicle_weighted"]))
            raw = bool(rng.random() < 0.2)  # raw render() frames (size, size, 3) uint8, no resize / frame stack
            if rng.random() < 0.35:  # any camera anchor, also on a border
                anchor = (float(rng.choice([0.0, 0.25, 0.5, 0.6, 1.0])), float(rng.choice([0.0, 0.4, 0.9, 1.0])))
            size = int(rng.choice([64, 128, 128, 256]))
            sizes = {64: [(96, 96), (96, 96), (24, 24), (64, 64), (32, 32), (48, 40)],
                     128: [(96, 96), (84, 84), (64, 64), (48, 48), (128, 128), (112, 100), (160, 160), (36, 36)],
                     256: [(96, 96), (96, 96), (84, 84), (128, 128), (64, 64), (256, 256), (100, 60)]}[size]
            obs_size = sizes[int(rng.integers(0, len(sizes)))]
        cls = load_map(size)
        pad = raster.FovGeometry(size, anchor[0], anchor[1]).pad  # the crop side (= vector_env.py:_crop_size)
        reqs = [dict(scene="rdm", num_vehicles=int(rng.integers(0, nveh_hi)), route_dist_range=[30, 90],
                     scene_seed=int(rng.integers(0, 10**6))) for _ in range(8 if size == 128 else 20)]
        if size == 128:  # the scripted scenarios exist at the 128 scale only (quirk C-11)
            reqs += [dict(scene="lead_brake", level=int(rng.integers(1, 4)), scene_seed=int(rng.integers(0, 10**6))) for _ in range(3)]
            reqs += [dict(scene="jaywalk", level=int(rng.integers(1, 5)), scene_seed=int(rng.integers(0, 10**6))) for _ in range(3)]
            reqs += [dict(scene="red_light_runner", scene_seed=int(rng.integers(0, 10**6))) for _ in range(2)]
            scenes = [S.build_scene(r, cls_map=cls, pad=pad) for r in reqs]
        else:
            scenes = [sc for sc in S.build_pool(reqs, pad=pad, size=size, skip_invalid=True, workers=1) if sc is not None][:12]
            reqs = [dict(size=size, n=len(scenes))]
        n = len(scenes)
        table = ACTION_PROFILES[profile].get("discrete_actions")
        if raw:
            gray, frame_stack, fuse = False, 1, None
        if gray:
            fuse = None
        eng = E.Engine(n, obs_mode=E.OBS_RGB if raw else E.OBS_GRAY if gray else E.OBS_SEMANTIC, mask_mode=mask,
                       frame_stack=frame_stack,
                       action_mode=E.ACTION_CONTINUOUS if continuous else E.ACTION_DISCRETE, discrete_table=table,
                       reward_mode=E.REWARD_SHAPING if reward == "shaping" else E.REWARD_CARL, anchor=anchor,
                       max_actors=16 if nveh_hi <= 12 else 32, ring_budget_bytes=64 << 20, size=size, obs_size=obs_size,
                       trajectory_steps=traj)
        eng.upload_map(cls)
        eng.upload_pool(pack_pool(scenes))
        if fov_masked:
            eng.upload_fov_mask(corner_mask(size, 0.5))
        oracles = [OracleEnv(cls, obs_mode="bev_raw" if raw else "bev_gray" if gray else "bev_semantic", semantic_mask_ch=mask,
                             action_mode="continuous" if continuous else "discrete", action_profile=profile,
                             reward_mode=reward, anchor=anchor, fov_masked=fov_masked, frame_stack=frame_stack,
                             size=size, obs_size=obs_size, temporal_fusion_mode=fuse or "stack")
                   for _ in range(n)]
        obs = eng.reset(torch.arange(n, dtype=torch.int32))
        obs = (eng.fuse(fuse) if fuse else obs).cpu().numpy()
        what = None
        frame = (lambda o: o[0]) if raw else (lambda o: o)  # the raw mode has no frame stack
        for i in range(n):
            if not np.array_equal(obs[i], frame(oracles[i].reset(scenes[i]))):
                what = f"reset observation of env {i}"
        alive = np.ones(n, bool)
        tab = None if continuous else np.asarray(table, dtype=np.float64)
        for t in range(steps):
            if what or not alive.any():
                break
            acts = []
            for i in range(n):
                e = oracles[i].sim.ego
                k = min(int(e.tidx) + 2, len(e.cx) - 1)
                err = np.arctan2(e.cy[k] - e.y, e.cx[k] - e.x) - e.yaw
                err = (err + np.pi) % (2 * np.pi) - np.pi
                want = np.array([0.7 if e.v < 25.0 else 0.0, np.clip(2.0 * err, -1, 1) + rng.normal(0, 0.05),
                                 1.0 if e.v > 32.0 else 0.0])
                acts.append(np.clip(want, [0, -1, 0], [1, 1, 1]) if continuous else int(np.argmin(((tab - want) ** 2).sum(1))))
            a = np.asarray(acts, dtype=np.float32 if continuous else np.int64)
            eng.step(torch.from_numpy(a).cuda())
            obs, rew = (eng.fuse(fuse) if fuse else eng.obs()).cpu().numpy(), eng.reward.cpu().numpy()
            term, trunc = eng.terminated.cpu().numpy().astype(bool), eng.truncated.cpu().numpy().astype(bool)
            hero = eng.hero.cpu().numpy()
            for i in range(n):
                if not alive[i]:
                    continue
                o, r, te, tr, _ = oracles[i].step(a[i] if continuous else int(a[i]))
                e = oracles[i].sim.ego
                compared += 1
                if not np.array_equal(obs[i], frame(o)):
                    what = f"observation env {i} step {t}"
                elif abs(r - rew[i]) > 1e-9 or te != term[i] or tr != trunc[i]:
                    what = f"reward / flags env {i} step {t}: {r} {rew[i]} {te} {term[i]}"
                elif not np.allclose(hero[i, :4], [e.x, e.y, e.yaw, e.v], rtol=1e-9, atol=1e-9):
                    what = f"pose env {i} step {t}"
                if te or tr:
                    alive[i] = False
                    c = int(eng.cause.cpu()[i])
                    endings[c] = endings.get(c, 0) + 1
        eng.close()
        if what:
            bad += 1
            print("MISMATCH", what, dict(profile=profile, reward=reward, mask=mask, anchor=anchor, fov_masked=fov_masked,
                                         gray=gray, frame_stack=frame_stack, size=size, obs_size=obs_size, raw=raw, traj=traj, fuse=fuse,
                                         nveh_hi=nveh_hi), reqs)
        elif scales:
            print(f"round {rd}: size {size} obs {'raw' if raw else obs_size} mask {mask} gray {gray} fov_masked {fov_masked} anchor {anchor} tables {traj} fusion {fuse} vehicles < {nveh_hi}: ok", flush=True)
    print(f"endings (cause id -> count): {endings}")
    print(f"{rounds} rounds, {compared} env-steps compared, {bad} mismatching rounds")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
