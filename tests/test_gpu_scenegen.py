"""Row f2 on the GPU: lead_brake / jaywalk scenes generated on the device (cbev_generate_scenes, csrc/scenegen.cu)
against the host generator (carlabev_env_b200/scenes.py, bit-identical to the reference's post-reset state) for 4096
seeds over all levels, and an engine-vs-oracle stepping check on the generated pool."""
import numpy as np
import pytest

from golden_util import load_map

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _ang_close(a, b, tol=1e-9):
    d = (np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64) + np.pi) % (2 * np.pi) - np.pi
    return bool(np.all(np.abs(d) < tol))


def _engine(n, **kw):
    from carlabev_env_b200 import engine as E

    eng = E.Engine(n, obs_mode=E.OBS_SEMANTIC, mask_mode="6-class", frame_stack=4, action_mode=E.ACTION_CONTINUOUS,
                   max_actors=4, ring_budget_bytes=64 << 20, **kw)
    eng.upload_map(load_map())
    return eng


@pytest.mark.parametrize("kind,nlev", [("lead_brake", 3), ("jaywalk", 4)])
def test_device_generated_pool_matches_the_host_generator(kind, nlev):
    from carlabev_env_b200.pool import pack_pool
    from carlabev_env_b200.scenes import build_pool

    n = 4096
    seeds = np.arange(n, dtype=np.int64)
    levels = 1 + (seeds % nlev).astype(np.int32)
    eng = _engine(8)
    att = eng.generate_scripted_pool([kind] * n, levels, seeds)
    assert att.min() >= 1 and att.max() <= 10
    dev = eng.read_pool()
    ref = pack_pool(build_pool([dict(scene=kind, level=int(lv), scene_seed=int(s)) for lv, s in zip(levels, seeds)]))
    # exact: layout, every drawn parameter, raw routes, behaviour parameters, integer spawn data
    for key in ("ego_off", "rew_off", "actor_off", "tl_off", "act_route_off", "act_raw_off", "num_vehicles", "rew_rx",
                "rew_ry", "ego_target_speed", "len_ego_route", "ego_tidx0", "act_kind", "act_beh", "act_beh_p",
                "act_cruise_px", "act_cruise_mps", "act_raw_x", "act_raw_y", "act_tidx0"):
        assert np.array_equal(dev[key], np.asarray(ref[key]).reshape(dev[key].shape)), key
    assert np.array_equal(dev["ego_state0"][:, 3], ref["ego_state0"][:, 3])
    assert np.array_equal(dev["act_state0"][:, 3], ref["act_state0"][:, 3])
    # smoothed routes / start poses: SciPy's LAPACK edge fit vs the linear Savitzky-Golay operator (scenegen.h)
    for key in ("ego_cx", "ego_cy", "act_cx", "act_cy"):
        assert np.allclose(dev[key], ref[key], rtol=0, atol=1e-9), key
    for key in ("ego_cyaw", "act_cyaw"):
        assert _ang_close(dev[key], ref[key]), key
    assert np.allclose(dev["ego_state0"][:, :2], ref["ego_state0"][:, :2], rtol=0, atol=1e-9)
    assert np.allclose(dev["act_state0"][:, :2], ref["act_state0"][:, :2], rtol=0, atol=1e-9)
    assert _ang_close(dev["ego_state0"][:, 2], ref["ego_state0"][:, 2]) and _ang_close(dev["act_state0"][:, 2], ref["act_state0"][:, 2])
    # the spawn jitter is an exact integer in {-1, 0, 1} on top of the smoothed start point
    jit_dev = np.rint(dev["ego_state0"][:, 0] - dev["ego_cx"][dev["ego_off"][:-1]])
    jit_ref = np.rint(ref["ego_state0"][:, 0] - ref["ego_cx"][ref["ego_off"][:-1]])
    assert np.array_equal(jit_dev, jit_ref) and set(np.unique(jit_dev)) <= {-1.0, 0.0, 1.0}
    eng.close()


def test_stepping_on_a_device_generated_pool_matches_the_oracle():
    import torch

    from oracle.env import OracleEnv, unpack_pool

    n = 24
    kinds = ["lead_brake" if i % 2 == 0 else "jaywalk" for i in range(n)]
    levels = [1 + (i // 2) % (3 if i % 2 == 0 else 4) for i in range(n)]
    seeds = [9000 + i for i in range(n)]
    eng = _engine(n)
    eng.generate_scripted_pool(kinds, levels, seeds)
    pool = eng.read_pool()
    scenes = unpack_pool(pool)
    cls = load_map()
    oracles = [OracleEnv(cls, action_mode="continuous") for _ in range(n)]
    obs = eng.reset(torch.arange(n, dtype=torch.int32)).cpu().numpy()
    for i in range(n):
        assert np.array_equal(obs[i], oracles[i].reset(scenes[i])), (i, "reset observation")
    rng = np.random.default_rng(3)
    alive = np.ones(n, bool)
    for t in range(70):
        a = np.stack([rng.uniform(0, 0.6, n), rng.uniform(-0.3, 0.3, n), rng.uniform(0, 0.5, n)], axis=1).astype(np.float32)
        eng.step(torch.from_numpy(a).cuda())
        obs, rew = eng.obs().cpu().numpy(), eng.reward.cpu().numpy()
        term, hero = eng.terminated.cpu().numpy().astype(bool), eng.hero.cpu().numpy()
        for i in range(n):
            if not alive[i]:
                continue
            o, r, te, tr, _ = oracles[i].step(a[i])
            e = oracles[i].sim.ego
            assert np.allclose(hero[i, :4], [e.x, e.y, e.yaw, e.v], rtol=1e-9, atol=1e-9), (t, i, "pose")
            assert abs(r - rew[i]) < 1e-9 and te == term[i], (t, i, "reward / done")
            assert np.array_equal(obs[i], o), (t, i, "observation")
            alive[i] = not te
    eng.close()
