"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-golden (GPU) tests."""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

ALL_CASES = ("lead_brake_continuous", "rdm_medium_discrete", "jaywalk_levels", "jaywalk_drive", "red_light_runner",
             "rdm_shaping_discrete13", "rdm_rgb_lookahead", "fusion_temporal_masked", "fusion_weighted")
# EnvConfig.size 64 / 256 (the reference runs `rdm` scenes there; oracle/gen_golden.py)
SCALE_CASES = ("size256_rdm_discrete", "size256_rdm_gray_lookahead", "size64_rdm_discrete", "size64_rdm_gray_lookahead")


def load_map(size=128):
    with np.load(os.path.join(ROOT, "carlabev_env_b200", "assets", f"town01_{size}_cls.npz")) as z:
        return np.ascontiguousarray(z["cls"])


class Golden:
    def __init__(self, name):
        self.name = name
        with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")) as z:
            self.d = {k: z[k] for k in z.files}
        self.env_kwargs = json.loads(str(self.d["env_kwargs"]))
        self.episode_infos = {int(k): v for k, v in json.loads(str(self.d["episode_infos"])).items()}
        self.pool = {k[len("pool_"):]: v for k, v in self.d.items() if k.startswith("pool_")}
        self.T = len(self.d["reward"])
        self.size = int(self.env_kwargs.get("size", 128))

    def __getitem__(self, k):
        return self.d[k]

    def oracle_kwargs(self):
        kw = self.env_kwargs
        obs_mode = kw.get("obs_mode", "bev_semantic")
        action_mode = kw.get("action_mode", "discrete")
        return dict(
            obs_mode="bev_semantic" if obs_mode == "bev_semantic" else "bev_gray",
            semantic_mask_ch=kw.get("semantic_mask_ch", "6-class"),
            frame_stack=kw.get("frame_stack", 4),
            action_mode=action_mode,
            action_profile=kw.get("action_profile_id") or ("continuous_gsb_v1" if action_mode == "continuous" else "discrete9_v1"),
            reward_mode=kw.get("reward_mode", "carl"),
            anchor=(kw.get("ego_anchor_x_frac", 0.5), kw.get("ego_anchor_y_frac", 0.5)),
            fov_masked=kw.get("fov_masked", False),
            temporal_fusion_mode=kw.get("temporal_fusion_mode", "stack"),
            size=self.size,
        )

    def full_obs(self, key="obs_full"):
        o = self.d[key]
        if "obs_packed" in self.d:
            o = np.unpackbits(o, axis=-1)[..., :96].astype(np.float32)
        elif o.dtype == np.float16:
            o = o.astype(np.float32)
        return o


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())
