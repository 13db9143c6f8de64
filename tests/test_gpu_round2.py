"""Round-2 GPU tests: limits lifted since round 1 (ego routes with more than 64 targets, camera anchors whose crop
exceeds one TMA box), pool growth while environments are mid-episode, the actor half of cbev_set_state, the set_point
columns of the `vector` observation, and a second engine on another device."""
import numpy as np
import pytest

from golden_util import load_map

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _check(t, i, eng_hero, rew, term, obs, oracle, action):
    o, r, te, tr, _ = oracle.step(action)
    e = oracle.sim.ego
    assert np.allclose(eng_hero[:4], [e.x, e.y, e.yaw, e.v], rtol=1e-9, atol=1e-9), (t, i, "pose")
    assert abs(r - rew) < 1e-9 and te == term, (t, i, "reward / done", r, rew)
    assert np.array_equal(obs, o), (t, i, "observation", int((obs != o).sum()))
    return te


def _pursuit(oracle, rng):
    e = oracle.sim.ego
    k = min(int(e.tidx) + 2, len(e.cx) - 1)
    err = np.arctan2(e.cy[k] - e.y, e.cx[k] - e.x) - e.yaw
    err = (err + np.pi) % (2 * np.pi) - np.pi
    return np.clip([0.7 if e.v < 25.0 else 0.0, 2.0 * err + rng.normal(0, 0.05), 1.0 if e.v > 32.0 else 0.0],
                   [0, -1, 0], [1, 1, 1]).astype(np.float32)


def test_ego_routes_with_more_than_64_targets():
    """Round 1 rejected ego routes of more than 64 points (one 64-bit visibility word); now two words = 128 targets.
    The ego drives the route (pursuit), so targets beyond index 63 are drawn, consumed and rewarded."""
    import torch

    from carlabev_env_b200 import engine as E
    from carlabev_env_b200 import scenes as S
    from carlabev_env_b200.pool import pack_pool
    from oracle.env import OracleEnv

    cls = load_map()
    scenes = []
    seed = 0
    while len(scenes) < 4 and seed < 200:
        try:
            s = S.build_scene(dict(scene="rdm", num_vehicles=3, route_dist_range=[230, 330], scene_seed=seed), cls_map=cls)
            if len(s["ego_cx"]) > 70:
                scenes.append(s)
        except RuntimeError:
            pass
        seed += 1
    assert len(scenes) == 4, "no long routes found"
    assert max(len(s["ego_cx"]) for s in scenes) <= 128
    n = len(scenes)
    eng = E.Engine(n, action_mode=E.ACTION_CONTINUOUS, max_actors=4, ring_budget_bytes=64 << 20, trajectory_steps=0)
    eng.upload_map(cls)
    eng.upload_pool(pack_pool(scenes))
    oracles = [OracleEnv(cls, action_mode="continuous") for _ in range(n)]
    obs = eng.reset(torch.arange(n, dtype=torch.int32)).cpu().numpy()
    for i in range(n):
        assert np.array_equal(obs[i], oracles[i].reset(scenes[i]))
    rng = np.random.default_rng(0)
    alive = np.ones(n, bool)
    best = 0
    for t in range(700):
        a = np.stack([_pursuit(o, rng) for o in oracles])
        eng.step(torch.from_numpy(a).cuda())
        obs, rew = eng.obs().cpu().numpy(), eng.reward.cpu().numpy()
        term, hero = eng.terminated.cpu().numpy().astype(bool), eng.hero.cpu().numpy()
        for i in range(n):
            if alive[i]:
                alive[i] = not _check(t, i, hero[i], rew[i], term[i], obs[i], oracles[i], a[i])
                best = max(best, int(oracles[i].sim.ego.tidx))
        if not alive.any():
            break
    assert best > 64, f"no env drove past target 64 (best {best})"
    eng.close()


@pytest.mark.parametrize("anchor", [(0.1, 0.9), (0.0, 1.0), (0.95, 0.5)])
def test_camera_anchor_near_a_border(anchor):
    """Anchors close to a border need crops of up to 360 px (round 1 rejected anything above 241 px): the raster kernel
    fetches only the 184 x 208 window a frame can sample, whatever the crop size."""
    import torch

    from carlabev_env_b200 import engine as E
    from carlabev_env_b200.pool import pack_pool
    from carlabev_env_b200.scenes import build_scene
    from oracle.env import OracleEnv
    from oracle.raster import FovGeometry

    cls = load_map()
    pad = FovGeometry(128, *anchor).pad
    assert pad > 241
    reqs = [dict(scene="rdm", num_vehicles=12, route_dist_range=[30, 90], scene_seed=300 + i) for i in range(6)]
    reqs += [dict(scene="lead_brake", level=3, scene_seed=7), dict(scene="jaywalk", level=3, scene_seed=8)]
    scenes = [build_scene(r, cls_map=cls, pad=pad) for r in reqs]
    n = len(scenes)
    eng = E.Engine(n, action_mode=E.ACTION_CONTINUOUS, max_actors=14, ring_budget_bytes=64 << 20, anchor=anchor,
                   mask_mode="7-class")
    eng.upload_map(cls)
    eng.upload_pool(pack_pool(scenes))
    oracles = [OracleEnv(cls, action_mode="continuous", anchor=anchor, semantic_mask_ch="7-class") for _ in range(n)]
    obs = eng.reset(torch.arange(n, dtype=torch.int32)).cpu().numpy()
    for i in range(n):
        assert np.array_equal(obs[i], oracles[i].reset(scenes[i])), (i, "reset observation")
    rng = np.random.default_rng(1)
    alive = np.ones(n, bool)
    for t in range(120):
        a = np.stack([_pursuit(o, rng) if i % 2 == 0 else
                      np.array([rng.uniform(0, 1), rng.uniform(-1, 1), rng.uniform(0, 0.3)], np.float32)
                      for i, o in enumerate(oracles)])
        eng.step(torch.from_numpy(a).cuda())
        obs, rew = eng.obs().cpu().numpy(), eng.reward.cpu().numpy()
        term, hero = eng.terminated.cpu().numpy().astype(bool), eng.hero.cpu().numpy()
        for i in range(n):
            if alive[i]:
                alive[i] = not _check(t, i, hero[i], rew[i], term[i], obs[i], oracles[i], a[i])
        if not alive.any():
            break
    assert t > 20
    eng.close()


def test_pool_grows_while_envs_are_mid_episode():
    """re-uploading the pool while environments run must not disturb them -- in particular not the
    StopReturn retreat route a pedestrian is following, and not the table -> live hand-over step."""
    import torch

    from carlabev_env_b200 import engine as E
    from carlabev_env_b200.pool import pack_pool
    from carlabev_env_b200.scenes import build_scripted_scene
    from oracle.env import OracleEnv

    cls = load_map()
    first = [build_scripted_scene("jaywalk", 40 + i, level=3, cls_map=cls) for i in range(4)]
    first += [build_scripted_scene("lead_brake", 60 + i, level=2, cls_map=cls) for i in range(2)]
    more = [build_scripted_scene("jaywalk", 80 + i, level=3 + i % 2, cls_map=cls) for i in range(4)]
    more += [build_scripted_scene("lead_brake", 90, level=3, cls_map=cls)]
    n = len(first)
    eng = E.Engine(n, action_mode=E.ACTION_CONTINUOUS, max_actors=4, ring_budget_bytes=64 << 20, trajectory_steps=24)
    eng.upload_map(cls)
    eng.upload_pool(pack_pool(first))
    oracles = [OracleEnv(cls, action_mode="continuous") for _ in range(n)]
    scene_of = list(range(n))
    obs = eng.reset(torch.arange(n, dtype=torch.int32)).cpu().numpy()
    for i in range(n):
        assert np.array_equal(obs[i], oracles[i].reset(first[i]))
    pool = list(first)
    crawl = np.tile(np.array([[0.0, 0.0, 1.0]], np.float32), (n, 1))  # brake: the ego waits, the pedestrians play out
    retreating = 0
    for t in range(110):
        if t in (30, 45, 70):  # past the hand-over (24 steps); some pedestrians are on their retreat route by now
            pool = pool + more[: {30: 2, 45: 4, 70: 5}[t] - (len(pool) - n)]
            eng.upload_pool(pack_pool(pool))
            # a finished or arbitrary env moves to one of the NEW scenes, everyone else keeps running
            j = t % n
            mask = np.zeros(n, bool)
            mask[j] = True
            ids = np.array(scene_of, dtype=np.int32)
            ids[j] = len(pool) - 1
            scene_of[j] = len(pool) - 1
            o = eng.reset(torch.from_numpy(ids), mask).cpu().numpy()
            assert np.array_equal(o[j], oracles[j].reset(pool[scene_of[j]])), (t, j, "masked reset onto a new scene")
        eng.step(torch.from_numpy(crawl).cuda())
        obs, rew = eng.obs().cpu().numpy(), eng.reward.cpu().numpy()
        term, hero = eng.terminated.cpu().numpy().astype(bool), eng.hero.cpu().numpy()
        _, act = eng.get_state(4)
        finished = np.zeros(n, bool)
        for i in range(n):
            finished[i] = _check(t, i, hero[i], rew[i], term[i], obs[i], oracles[i], crawl[i])
            ref = np.array([[b.x, b.y, b.yaw, b.v] for b in oracles[i].sim.actors])
            assert np.allclose(act[i, :len(ref), :4], ref, rtol=1e-9, atol=1e-9), (t, i, "actor poses")
            retreating += int(any(int(f) & 32 for f in act[i, :len(ref), 7]))
        if finished.any():  # e.g. a pedestrian walked into the standing ego: SyncVectorEnv-style masked reset
            o = eng.reset(torch.from_numpy(np.array(scene_of, dtype=np.int32)), finished).cpu().numpy()
            for i in np.flatnonzero(finished):
                assert np.array_equal(o[i], oracles[i].reset(pool[scene_of[i]])), (t, i, "masked reset")
    assert retreating > 0, "no pedestrian was on its retreat route during the test"
    with pytest.raises(E.CbevError, match="extend"):
        eng.upload_pool(pack_pool(first[:2]))
    eng.close()


def test_set_state_round_trip_and_vector_observation_set_point():
    import torch

    from carlabev_env_b200 import EnvConfig, RunConfig, make_env
    from carlabev_env_b200.pool import load_shipped_pool
    from oracle.env import OracleEnv

    cls = load_map()
    scenes = load_shipped_pool("rdm_rt_medium_v1")[:5]
    n = len(scenes)
    envs = make_env(RunConfig(env=EnvConfig(action_mode="continuous"), num_envs=n), scenes=scenes,
                    ring_budget_bytes=64 << 20)
    eng = envs.engine
    envs.reset(options={"scene_ids": np.arange(n)})
    oracles = [OracleEnv(cls, action_mode="continuous") for _ in range(n)]
    for i in range(n):
        oracles[i].reset(scenes[i])
    rng = np.random.default_rng(5)
    for t in range(25):
        a = np.stack([_pursuit(o, rng) for o in oracles])
        obs, rew, term, trunc, infos = envs.step(a)
        for i in range(n):
            oracles[i].step(a[i])
        vec = envs.vector_observation().cpu().numpy()   # state ++ set_point (carlabev.py:237-244), float32
        for i in range(n):
            e = oracles[i].sim.ego
            want = np.array([e.x, e.y, e.yaw, e.v, e.cx[e.tidx], e.cy[e.tidx], e.cyaw[e.tidx]]).astype(np.float32)
            assert np.array_equal(vec[i], want), (t, i, vec[i], want)
        if t == 10:  # cbev_get_state -> cbev_set_state is the identity, for the ego and for every actor column
            ego, act = eng.get_state()
            eng.set_state(ego, act)
            ego2, act2 = eng.get_state()
            assert np.array_equal(ego, ego2) and np.array_equal(act, act2)
    # actors can be moved (live stepping): park actor 0 of env 0 somewhere else and read it back
    ego, act = eng.get_state()
    act[0, 0, 0] += 5.0
    eng.set_state(None, act)
    assert eng.get_state()[1][0, 0, 0] == act[0, 0, 0]
    envs.close()


def test_two_engines_on_two_devices():
    """constant tables / shared-memory opt-in are per device, entry points select the engine's device."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from carlabev_env_b200 import engine as E
    from carlabev_env_b200.pool import pack_pool
    from carlabev_env_b200.scenes import build_scripted_scene

    cls = load_map()
    scenes = [build_scripted_scene("lead_brake", 3 + i, level=1 + i % 3, cls_map=cls) for i in range(4)]
    outs = []
    for dev in (0, 1):
        eng = E.Engine(4, action_mode=E.ACTION_CONTINUOUS, max_actors=4, ring_budget_bytes=64 << 20, device=dev)
        eng.upload_map(cls)
        eng.upload_pool(pack_pool(scenes))
        torch.cuda.set_device(0)  # the caller's current device is not the engine's
        eng.reset(torch.arange(4, dtype=torch.int32, device=f"cuda:{dev}"))
        eng.step(torch.full((4, 3), 0.5, device=f"cuda:{dev}"))
        outs.append((eng.obs().cpu().numpy().copy(), eng.reward.cpu().numpy().copy()))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_vector_env_step_paths_agree():
    """VectorEnv.step with host NumPy actions (cbev_step_host_ex), with a CUDA tensor (cbev_step_ex) and with
    host_infos=False (cbev_step) gives the same observations / rewards / flags; to_numpy returns host copies of the
    pinned mirrors; the terminal infos come from the pinned episode block."""
    import torch

    from carlabev_env_b200 import EnvConfig, RunConfig, make_env
    from carlabev_env_b200.scenes import build_scripted_scene

    cls = load_map()
    scenes = [build_scripted_scene("lead_brake", 50 + i, level=1 + i % 3, cls_map=cls) for i in range(12)]
    cfg = RunConfig(env=EnvConfig(action_mode="continuous"), num_envs=12)
    kinds = [dict(), dict(), dict(host_infos=False), dict(to_numpy=True)]
    envs = [make_env(cfg, scenes=scenes, autoreset="next_step", ring_budget_bytes=64 << 20, seed=3, **kw) for kw in kinds]
    for e in envs:
        e.reset(options={"scene_ids": np.arange(12)})
    rng = np.random.default_rng(9)
    finished = 0
    for t in range(60):
        a = np.stack([rng.uniform(0.3, 1, 12), rng.uniform(-1, 1, 12), rng.uniform(0, 0.3, 12)], axis=1).astype(np.float32)
        outs = [envs[0].step(a), envs[1].step(torch.from_numpy(a).cuda()), envs[2].step(torch.from_numpy(a).cuda()),
                envs[3].step(a)]
        torch.cuda.synchronize()
        obs0, rew0, term0, trunc0, info0 = outs[0]
        for k in (1, 2):
            o, r, te, tr, _ = outs[k]
            assert torch.equal(o, obs0) and torch.equal(r, rew0) and torch.equal(te, term0) and torch.equal(tr, trunc0), (t, k)
        o, r, te, tr, info3 = outs[3]
        assert isinstance(o, np.ndarray) and r.dtype == np.float64 and te.dtype == np.bool_
        assert np.array_equal(o, obs0.cpu().numpy()) and np.array_equal(r, rew0.cpu().numpy())
        assert np.array_equal(te, term0.cpu().numpy())
        assert "episode_block" in outs[2][4] and "episode_info" not in outs[2][4]
        if bool(term0.any()):
            done = term0.cpu().numpy()
            finished += int(done.sum())
            for info in (info0, outs[1][4], info3):
                assert np.array_equal(info["_episode"], done)
                i = int(np.flatnonzero(done)[0])
                assert info["episode"]["l"][i] >= 1 and info["episode_info"]["termination"][i] in ("collision", "success", "out_of_bounds")
                assert abs(info["episode"]["r"][i] - info["episode_info"]["return"][i]) < 1e-12
            blk = outs[2][4]["episode_block"].cpu().numpy()
            assert np.allclose(blk[done, 0], info0["episode_info"]["return"][done])
        else:
            assert "episode" not in info0
    assert finished > 0
    for e in envs:
        e.close()


@pytest.mark.parametrize("size,seeds", [(256, (0, 1, 3, 4, 5, 6)), (64, (0, 2, 3, 5, 6, 10))])
def test_make_env_at_other_map_scales(size, seeds):
    """SURVEY.md section 8 row f4 through the reference-facing boundary: make_env(EnvConfig(size=...)).reset(options=
    {"scene": "rdm", ...}).step(actions) against the oracle (pinned on goldens recorded from the reference at that size):
    stacked masks, rewards, flags; then the raw (size, size, 3) frames of the raw_rgb engine mode."""
    import numpy as np

    from carlabev_env_b200 import EnvConfig, RunConfig, make_env
    from carlabev_env_b200 import scenes as S
    from golden_util import load_map
    from oracle.env import OracleEnv

    n = len(seeds)
    cls = load_map(size)
    for raw in (False, True):
        env_cfg = EnvConfig(size=size, action_mode="continuous", obs_mode="bev_rgb" if raw else "bev_semantic")
        envs = make_env(RunConfig(env=env_cfg, num_envs=n), ring_budget_bytes=64 << 20, to_numpy=True, raw_rgb=raw)
        assert envs.single_observation_space.shape == ((size, size, 3) if raw else (24, 96, 96))
        opts = dict(scene="rdm", num_vehicles=12, route_dist_range=(30, 100))
        oracles = [OracleEnv(cls, action_mode="continuous", size=size, obs_mode="bev_raw" if raw else "bev_semantic",
                             frame_stack=1 if raw else 4) for _ in range(n)]
        frame = (lambda o: o[0]) if raw else (lambda o: o)  # the raw mode has no frame stack
        want, scenes = [], []
        for i, sd in enumerate(seeds):  # SyncVectorEnv hands every env the same options: one reset per seed, masked
            mask = np.ones(n, bool) if i == 0 else np.arange(n) == i   # the first reset covers every env
            obs, _ = envs.reset(options={**opts, "scene_seed": sd, "reset_mask": mask})
            scenes.append(S.build_scene({**opts, "scene_seed": sd}, cls_map=cls, pad=envs.pad))
            want.append(frame(oracles[i].reset(scenes[i])))
        assert np.array_equal(np.asarray(obs), np.stack(want)), "reset observations"
        rng = np.random.default_rng(size)
        episodes = 0
        for t in range(60):
            a = np.stack([rng.uniform(0.3, 1, n), rng.uniform(-0.4, 0.4, n), rng.uniform(0, 0.2, n)], axis=1).astype(np.float32)
            obs, rew, term, trunc, _ = envs.step(a)
            obs, rew, term, trunc = np.asarray(obs), np.asarray(rew), np.asarray(term), np.asarray(trunc)
            for i in range(n):
                o, r, te, tr, _ = oracles[i].step(a[i])
                assert np.array_equal(obs[i], frame(o)), (size, raw, t, i)
                assert abs(r - rew[i]) <= 1e-9 and te == term[i] and tr == trunc[i], (size, raw, t, i)
                if te or tr:  # masked reset of the finished env, as a SyncVectorEnv caller does
                    episodes += 1
                    obs_r, _ = envs.reset(options={**opts, "scene_seed": seeds[i], "reset_mask": np.arange(n) == i})
                    assert np.array_equal(np.asarray(obs_r)[i], frame(oracles[i].reset(scenes[i]))), (size, raw, t, i, "reset")
        assert episodes >= 1
        envs.close()


@pytest.mark.parametrize("size,obs_size,rotate", [(256, (96, 96), 0), (64, (24, 24), 0), (256, (96, 96), 1),
                                                  (128, (84, 84), 0), (256, (84, 84), 0), (128, (100, 72), 1),
                                                  (128, (16, 24), 0)])
def test_block_shortcut_equals_table_resize(size, obs_size, rotate):
    """k_render_any's 8 : 3 block shortcut (single-colour 8 x 8 source blocks resolved in the rotate's registers) and, at
    the other ratios, its separable single-colour test (vertical uniformity per 4-texel word, then one window) against
    the same kernel with the shortcut off (debug flag 512: every output through the table resize, which the goldens and
    the oracle tests pin): 96 envs, 40 steps, device auto-reset -- identical stacked masks, whatever the heading
    (`rotate` = 1: through the range-tested rotate, debug flag 1)."""
    import numpy as np
    import torch

    from carlabev_env_b200 import engine as E
    from carlabev_env_b200.config import ACTION_PROFILES
    from carlabev_env_b200.pool import pack_pool
    from carlabev_env_b200.scenes import build_pool
    from golden_util import load_map

    pad = {64: 91, 128: 182, 256: 363}[size]
    scenes = [s for s in build_pool([dict(scene="rdm", num_vehicles=12, route_dist_range=(30, 100), scene_seed=i)
                                     for i in range(24)], pad=pad, size=size, skip_invalid=True) if s is not None]
    assert len(scenes) >= 8
    n = 96
    out = []
    for flags in (rotate, rotate | 512):
        eng = E.Engine(n, action_mode=E.ACTION_DISCRETE, discrete_table=ACTION_PROFILES["discrete9_v1"]["discrete_actions"],
                       max_actors=16, autoreset=E.AUTORESET_NEXT_STEP, ring_slots=8, size=size, obs_size=obs_size)
        eng.upload_map(load_map(size))
        eng.upload_pool(pack_pool(scenes))
        eng.set_debug_flags(flags)
        frames = [eng.reset(torch.arange(n, dtype=torch.int32) % len(scenes)).clone()]
        g = torch.Generator(device="cuda")
        g.manual_seed(7)
        for _ in range(40):
            eng.step(torch.randint(0, 9, (n,), device="cuda", generator=g))
            frames.append(eng.obs().clone())
        out.append(torch.stack(frames).cpu().numpy())
        eng.close()
    assert out[0].shape == (41, n, 24, *obs_size)
    assert out[0].any() and np.array_equal(out[0], out[1])


@pytest.mark.parametrize("size,obs_size,mode", [(256, (256, 256), "semantic"), (256, (200, 200), "gray"),
                                                (256, (128, 128), "semantic"), (256, (64, 64), "gray"),
                                                (64, (24, 24), "semantic")])
def test_large_and_small_observations_at_other_map_scales(size, obs_size, mode):
    """Observation sizes whose work list of mixed outputs cannot hold every output next to the output bytes in the dead
    tile (size 256 with 200 x 200 or the 256 x 256 copy: the resize runs in bands), the halving and quartering of the
    256-px view, and the 8 : 3 block shortcut at size 64 -- against the oracle (resize branches pinned on cv2 in
    tests/test_oracle_contracts.py)."""
    import numpy as np
    import torch

    from carlabev_env_b200 import engine as E
    from carlabev_env_b200.pool import pack_pool
    from carlabev_env_b200.scenes import build_pool
    from golden_util import load_map
    from oracle.env import OracleEnv

    pad = {64: 91, 256: 363}[size]
    cls = load_map(size)
    scenes = [s for s in build_pool([dict(scene="rdm", num_vehicles=10, route_dist_range=(30, 100), scene_seed=i)
                                     for i in range(16)], pad=pad, size=size, skip_invalid=True) if s is not None][:5]
    n = len(scenes)
    assert n >= 3
    gray = mode == "gray"
    eng = E.Engine(n, action_mode=E.ACTION_CONTINUOUS, max_actors=16, ring_budget_bytes=256 << 20, size=size,
                   obs_size=obs_size, obs_mode=E.OBS_GRAY if gray else E.OBS_SEMANTIC, frame_stack=2)
    eng.upload_map(cls)
    eng.upload_pool(pack_pool(scenes))
    oracles = [OracleEnv(cls, action_mode="continuous", size=size, obs_size=obs_size, frame_stack=2,
                         obs_mode="bev_rgb" if gray else "bev_semantic") for _ in range(n)]
    obs = eng.reset(torch.arange(n, dtype=torch.int32)).cpu().numpy()
    for i in range(n):
        assert np.array_equal(obs[i], oracles[i].reset(scenes[i])), (i, "reset")
    rng = np.random.default_rng(size + obs_size[0])
    alive = np.ones(n, bool)
    for t in range(16):
        a = np.stack([rng.uniform(0.3, 1, n), rng.uniform(-0.5, 0.5, n), rng.uniform(0, 0.2, n)], axis=1).astype(np.float32)
        eng.step(torch.from_numpy(a).cuda())
        obs = eng.obs().cpu().numpy()
        for i in range(n):
            if alive[i]:
                o, _, te, tr, _ = oracles[i].step(a[i])
                assert np.array_equal(obs[i], o), (t, i)
                alive[i] = not (te or tr)
    eng.close()
