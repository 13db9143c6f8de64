"""Small driver for the ncu pass over the kernel variants the bench line does not launch: grayscale and generic-size
(GEN) raster, temporal fusion, live actor stepping with wide groups (k_move<32>), raw RGB.  Each variant runs a few
steps at 2048 envs; run under `ncu --set full -k regex:... ` (tools/r2_ncu.sh)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from carlabev_env_b200 import engine as E  # noqa: E402
from carlabev_env_b200.config import ACTION_PROFILES  # noqa: E402
from carlabev_env_b200.pool import load_shipped_pool, pack_pool  # noqa: E402
from carlabev_env_b200.vector_env import load_town01_map  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
N, STEPS = 2048, 3
cls = load_town01_map()
scenes = load_shipped_pool("rdm_rt_hard_v1")
packed = pack_pool(scenes)
table = ACTION_PROFILES["discrete9_v1"]["discrete_actions"]
variants = {
    "gray": dict(obs_mode=E.OBS_GRAY),
    "gen84": dict(obs_mode=E.OBS_SEMANTIC, obs_size=(84, 84)),
    "fuse": dict(obs_mode=E.OBS_SEMANTIC, mask_mode="5-class"),
    "live32": dict(obs_mode=E.OBS_SEMANTIC, trajectory_steps=0),
    "rgb": dict(obs_mode=E.OBS_RGB, anchor=(0.5, 0.75)),
}
for name, kw in variants.items():
    if which not in ("all", name):
        continue
    eng = E.Engine(N, action_mode=E.ACTION_DISCRETE, discrete_table=table, max_actors=25, autoreset=E.AUTORESET_NEXT_STEP,
                   ring_budget_bytes=8 << 30, **kw)
    eng.upload_map(cls)
    eng.upload_pool(packed)
    eng.reset(torch.arange(N, dtype=torch.int32) % len(scenes))
    g = torch.Generator().manual_seed(0)
    for t in range(STEPS):
        eng.step(torch.randint(0, 9, (N,), generator=g).cuda())
        if name == "fuse":
            eng.fuse("vehicle_temporal")
    torch.cuda.synchronize()
    print(name, "ok", float(eng.reward.sum()))
    eng.close()
