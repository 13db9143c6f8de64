"""Oracle pin: the NumPy restatement (oracle/) against golden trajectories recorded from the UNMODIFIED
reference (tests/golden/*.npz; generator oracle/gen_golden.py).  Float64 state, rewards and comfort signals
must be bit-identical on the CPU; flags, palette frames and observations exact."""
import numpy as np
import pytest

from golden_util import ALL_CASES, SCALE_CASES, Golden, crc, load_map
from oracle import raster
from oracle.env import OracleEnv, unpack_pool

STEP_LIMIT = {"rdm_medium_discrete": 160, "jaywalk_levels": 330, "rdm_rgb_lookahead": 120}  # keep the CPU suite short


@pytest.mark.parametrize("case", ALL_CASES + SCALE_CASES)
def test_oracle_replays_reference_golden(case):
    g = Golden(case)
    scenes = unpack_pool(g.pool)
    env = OracleEnv(load_map(g.size), **g.oracle_kwargs())
    resets = dict(zip(g["reset_steps"].tolist(), g["reset_scene"].tolist()))
    fsteps = {int(s): i for i, s in enumerate(g["frame_steps"])}
    osteps = {int(s): i for i, s in enumerate(g["obs_steps"])}
    obs_full = g.full_obs()
    reset_obs = g.full_obs("reset_obs")
    ri = 0
    T = min(g.T, STEP_LIMIT.get(case, g.T))
    for t in range(T):
        if t in resets:
            obs = env.reset(scenes[resets[t]])
            assert np.array_equal(env.render_index(reset_frame=True), g["reset_frames"][ri]), f"reset frame t={t}"
            assert np.array_equal(obs, reset_obs[ri]), f"reset obs t={t}"
            ri += 1
        obs, r, term, trunc, info = env.step(g["actions"][t])
        s = env.sim
        e = s.ego
        assert np.array_equal(np.array([e.x, e.y, e.yaw, e.v], dtype=np.float64), g["ego_state"][t]), f"ego t={t}"
        assert float(s.acc) == g["acc"][t] and e.tidx == g["tidx"][t], f"acc/tidx t={t}"
        A = np.array([[a.x, a.y, a.yaw, a.v] for a in s.actors], dtype=np.float64).reshape(-1, 4)
        assert np.array_equal(A, g["actor_state"][t][: len(A)]), f"actors t={t}"
        assert r == g["reward"][t], f"reward t={t}: {r} vs {g['reward'][t]}"
        assert term == g["term"][t] and trunc == g["trunc"][t], f"flags t={t}"
        for k in ("hit", "hit_id", "tile", "n_nearby"):
            assert s.last[k] == g[k][t], f"{k} t={t}"
        assert s.last["dist2wp"] == g["dist2wp"][t]
        cf = np.array([s.comfort[k] for k in ("speed_mps", "accel_long", "accel_lat", "jerk_long", "jerk_lat",
                                              "yaw_rate", "yaw_acc")])
        assert np.array_equal(cf, g["comfort"][t]), f"comfort t={t}"
        assert crc(env.last_rgb) == g["rgb_crc"][t], f"rgb frame t={t}"
        if t in fsteps:
            assert np.array_equal(raster.PALETTE[g["frames"][fsteps[t]]], env.last_rgb)
        assert crc(obs) == g["obs_crc"][t], f"observation t={t}"
        if t in osteps:
            assert np.array_equal(obs, obs_full[osteps[t]])
        if term:
            gi = g.episode_infos[t]
            oi = info["episode_info"]
            for k, v in oi.items():
                if k in gi:
                    assert gi[k] == v or str(gi[k]) == str(v), f"episode_info[{k}] t={t}: {v} vs {gi[k]}"
            assert info["episode"]["r"] == gi["_episode_r"] and info["episode"]["l"] == gi["_episode_l"]
