"""ctypes binding of libcbev.so (include/cbev.h) + a thin torch-tensor front end.

PyTorch is plumbing here: it owns device buffers and streams; every computation of the step
happens inside the hand-written CUDA kernels behind the C ABI.  There is NO CPU fallback: if the
library is missing or no CUDA device is present, constructing an Engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .config import CARL_DEFAULTS, SHAPING_DEFAULTS

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcbev.so")

# enums of include/cbev.h
OBS_SEMANTIC, OBS_GRAY, OBS_RGB = 0, 1, 2
MASK_MODES = {"binary": 0, "2-class": 1, "4-class": 2, "5-class": 3, "6-class": 4, "7-class": 5}
MASK_CHANNELS = {"binary": 1, "2-class": 2, "4-class": 4, "5-class": 5, "6-class": 6, "7-class": 7}
ACTION_DISCRETE, ACTION_CONTINUOUS = 0, 1
REWARD_CARL, REWARD_SHAPING = 0, 1
AUTORESET_DISABLED, AUTORESET_NEXT_STEP = 0, 1
CAUSE_NAMES = (None, "ckpt", "collision", "success", "out_of_bounds", "off_road", "max_actions", "unknown")
HERO_FIELDS = ("x", "y", "yaw", "v", "x_1", "y_1", "yaw_1", "v_1", "dist2wp", "set_point_x", "set_point_y",
               "set_point_yaw", "cmd_gas", "cmd_steer", "cmd_brake", "applied_delta", "speed_mps", "accel_long",
               "accel_lat", "jerk_long", "jerk_lat", "yaw_rate", "yaw_acc", "acc", "target_idx", "hit", "hit_id",
               "tile_class", "n_nearby", "dist2goal", "t", "scene")
EPISODE_FIELDS = ("return", "length", "cause", "mean_speed", "mean_abs_accel_long", "mean_abs_accel_lat",
                  "mean_abs_jerk_long", "mean_abs_jerk_lat", "mean_abs_yaw_rate", "mean_abs_yaw_acc",
                  "comfort_violation_rate", "harsh_brake_rate", "scene", "num_vehicles", "len_ego_route", "episode")
STATS_FIELDS = 4 + 8 + 6 + 3
SG_MAX = 12
EXPORTS = ("cbev_version", "cbev_last_error", "cbev_create", "cbev_destroy", "cbev_upload_map",
           "cbev_upload_scene_pool", "cbev_frame_bytes", "cbev_bind_obs_ring", "cbev_reset", "cbev_step",
           "cbev_step_host", "cbev_obs_head", "cbev_get_state", "cbev_set_ego_state", "cbev_copy_fov",
           "cbev_read_stats", "cbev_launch_count", "cbev_profile_enable", "cbev_profile_read", "cbev_abi_sizes",
           "cbev_keep_fov", "cbev_upload_fov_mask", "cbev_fuse", "cbev_debug_rerender", "cbev_set_debug_flags",
           "cbev_debug_read_trace", "cbev_step_host_ex", "cbev_wait_host_outputs", "cbev_set_state",
           "cbev_profile_read_ex", "cbev_step_ex", "cbev_invalidate", "cbev_generate_scenes", "cbev_pool_counts",
           "cbev_read_scene_pool")


class CbevConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("fov_size", C.c_int32),
        ("anchor_x_frac", C.c_double), ("anchor_y_frac", C.c_double),
        ("obs_h", C.c_int32), ("obs_w", C.c_int32), ("obs_mode", C.c_int32), ("mask_mode", C.c_int32),
        ("frame_stack", C.c_int32), ("ring_slots", C.c_int32), ("action_mode", C.c_int32), ("n_discrete", C.c_int32),
        ("discrete_table", C.c_float * 48),
        ("reward_mode", C.c_int32), ("autoreset", C.c_int32), ("max_actors", C.c_int32), ("trajectory_steps", C.c_int32),
        ("lane_center_exponent", C.c_double), ("lane_center_floor", C.c_double), ("off_lane_penalty", C.c_double),
        ("speed_penalty_scale", C.c_double), ("speed_penalty_floor", C.c_double), ("ttc_threshold", C.c_double),
        ("ttc_penalty_floor", C.c_double),
        ("max_actions", C.c_int32), ("offroad_terminate_after", C.c_int32),
        ("sidewalk_step_penalty", C.c_double), ("sidewalk_penalty_scale", C.c_double),
        ("k_lat_quadratic", C.c_double), ("k_progress", C.c_double), ("k_flow", C.c_double),
        ("k_align_bonus", C.c_double), ("k_reverse", C.c_double), ("k_ttc", C.c_double), ("alive_bias", C.c_double),
        ("k_smooth", C.c_double), ("k_steer_smooth", C.c_double), ("k_steer_jerk", C.c_double),
        ("k_route_dev", C.c_double), ("route_dev_start", C.c_double),
        ("max_speed_for_flow", C.c_double), ("lat_clip", C.c_double), ("yaw_small", C.c_double),
        ("lat_small", C.c_double),
        ("seed", C.c_uint64),
    ]


_P = C.c_void_p


class CbevPoolDesc(C.Structure):
    _fields_ = [("n_scenes", C.c_int32), ("n_actors_total", C.c_int32), ("n_tl_total", C.c_int32)] + [
        (name, _P) for name in (
            "ego_state0", "ego_target_speed", "ego_tidx0", "len_ego_route", "num_vehicles", "ego_off", "rew_off",
            "actor_off", "tl_off", "ego_cx", "ego_cy", "ego_cyaw", "rew_rx", "rew_ry", "rew_cum", "act_kind",
            "act_state0", "act_tidx0", "act_cruise_px", "act_cruise_mps", "act_beh", "act_beh_p", "act_route_off",
            "act_raw_off", "act_cx", "act_cy", "act_cyaw", "act_raw_x", "act_raw_y", "tl_rect", "tl_color", "sg_mat")
    ]


class CbevStepOut(C.Structure):
    _fields_ = [("reward", _P), ("terminated", _P), ("truncated", _P), ("cause", _P), ("hero", _P), ("episode", _P)]


class CbevHostOut(C.Structure):
    _fields_ = [("reward", _P), ("terminated", _P), ("truncated", _P), ("cause", _P), ("episode", _P)]


class CbevError(RuntimeError):
    pass


_lib = None


def load_library(build_if_missing: bool = True):
    """dlopen libcbev.so; raises if it cannot be built/loaded (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or _build.needs_build():
        # a library older than csrc/ or include/cbev.h would run stale kernels behind an unchanged ABI
        if not build_if_missing:
            raise CbevError(f"{LIB_PATH} is missing or older than its sources; run `python -m carlabev_env_b200.build`")
        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.cbev_last_error.restype = C.c_char_p
    lib.cbev_frame_bytes.restype = C.c_int64
    lib.cbev_launch_count.restype = C.c_int64
    lib.cbev_frame_bytes.argtypes = [_P]
    lib.cbev_launch_count.argtypes = [_P]
    lib.cbev_create.argtypes = [C.POINTER(CbevConfig), C.POINTER(_P)]
    lib.cbev_destroy.argtypes = [_P]
    lib.cbev_upload_map.argtypes = [_P, _P, C.c_int32, C.c_int32]
    lib.cbev_upload_scene_pool.argtypes = [_P, C.POINTER(CbevPoolDesc)]
    lib.cbev_bind_obs_ring.argtypes = [_P, _P, C.c_int64]
    lib.cbev_reset.argtypes = [_P, _P, _P, _P]
    lib.cbev_step.argtypes = [_P, _P, C.POINTER(CbevStepOut), _P]
    lib.cbev_step_host.argtypes = [_P, _P, _P, _P, _P, _P]
    lib.cbev_step_host_ex.argtypes = [_P, _P, C.POINTER(CbevStepOut), C.POINTER(CbevHostOut), _P]
    lib.cbev_step_ex.argtypes = [_P, _P, C.POINTER(CbevStepOut), C.POINTER(CbevHostOut), _P]
    lib.cbev_invalidate.argtypes = [_P]
    lib.cbev_generate_scenes.argtypes = [_P, C.c_int32, _P, _P, _P, _P, _P]
    lib.cbev_pool_counts.argtypes = [_P, _P]
    lib.cbev_read_scene_pool.argtypes = [_P, C.POINTER(CbevPoolDesc)]
    lib.cbev_wait_host_outputs.argtypes = [_P]
    lib.cbev_set_state.argtypes = [_P, _P, _P]
    lib.cbev_profile_read_ex.argtypes = [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int64)]
    lib.cbev_obs_head.argtypes = [_P, C.POINTER(C.c_int32)]
    lib.cbev_get_state.argtypes = [_P, _P, _P]
    lib.cbev_set_ego_state.argtypes = [_P, _P]
    lib.cbev_copy_fov.argtypes = [_P, _P, _P]
    lib.cbev_keep_fov.argtypes = [_P, C.c_int32]
    lib.cbev_debug_read_trace.argtypes = [_P, _P]
    lib.cbev_upload_fov_mask.argtypes = [_P, _P]
    lib.cbev_fuse.argtypes = [_P, C.c_int32, _P, _P]
    lib.cbev_debug_rerender.argtypes = [_P, C.c_int32, _P]
    lib.cbev_set_debug_flags.argtypes = [_P, C.c_int32]
    lib.cbev_read_stats.argtypes = [_P, _P, C.c_int32, _P]
    lib.cbev_profile_enable.argtypes = [_P, C.c_int32]
    lib.cbev_profile_read.argtypes = [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.cbev_abi_sizes.argtypes = [C.POINTER(C.c_int32)] * 3
    sizes = [C.c_int32(), C.c_int32(), C.c_int32()]
    lib.cbev_abi_sizes(*[C.byref(v) for v in sizes])
    expect = (C.sizeof(CbevConfig), C.sizeof(CbevPoolDesc), C.sizeof(CbevStepOut))
    if tuple(v.value for v in sizes) != expect:
        raise CbevError(f"ABI mismatch between include/cbev.h and engine.py: {[v.value for v in sizes]} vs {expect}")
    _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise CbevError(f"cbev error {rc}: {lib.cbev_last_error().decode()}")


def savgol_operators() -> np.ndarray:
    """Linear operators of control/utils.py:smooth_and_compute's savgol stage for n-point routes,
    n <= SG_MAX (window 11 -> adapted, polyorder 3 -> adapted): sg[n] @ ax == smoothed ax."""
    from scipy.signal import savgol_filter

    sg = np.zeros((SG_MAX + 1, SG_MAX, SG_MAX), dtype=np.float64)
    for n in range(2, SG_MAX + 1):
        window = 11
        if window > n:
            window = n if n % 2 == 1 else n - 1
        if window < 3:
            window = 3
        poly = min(3, window - 1)
        if n >= window:
            m = savgol_filter(np.eye(n), window_length=window, polyorder=poly, axis=0)
        else:
            m = np.eye(n)
        sg[n, :n, :n] = m
    return sg


class Engine:
    """One engine = one GPU = one shard of environments."""

    def __init__(self, num_envs, *, obs_mode=OBS_SEMANTIC, mask_mode="6-class", frame_stack=4, ring_slots=None,
                 action_mode=ACTION_DISCRETE, discrete_table=None, reward_mode=REWARD_CARL, reward_params=None,
                 autoreset=AUTORESET_DISABLED, anchor=(0.5, 0.5), max_actors=0, seed=0, device=None,
                 ring_budget_bytes=None, size=128, obs_size=(96, 96), trajectory_steps=1024):
        import torch

        if not torch.cuda.is_available():
            raise CbevError("carlabev_env_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        self.N = int(num_envs)
        self.obs_mode = obs_mode
        self.mask_mode = mask_mode
        self.channels = MASK_CHANNELS[mask_mode] if obs_mode == OBS_SEMANTIC else 1
        self.F = 1 if obs_mode == OBS_RGB else int(frame_stack)
        self.size = int(size)
        self.obs_hw = tuple(obs_size)
        if obs_mode == OBS_SEMANTIC:
            frame = self.channels * obs_size[0] * obs_size[1] * 4
        elif obs_mode == OBS_GRAY:
            frame = obs_size[0] * obs_size[1]
        else:
            frame = size * size * 3
        if ring_slots is None:
            if self.F == 1:
                ring_slots = 1
            else:
                if ring_budget_bytes is None:
                    free, _ = torch.cuda.mem_get_info(self.device)
                    ring_budget_bytes = int(free * 0.35)
                # L >= 2F: the previous observation window stays intact for one more step (a caller holding both
                # `obs` and `next_obs` -- replay buffers, GAE bootstrap -- reads two valid windows); with the minimum
                # L = 2F - 1 the next head or mirror write lands inside the previous window
                ring_slots = max(2 * self.F, min(64, ring_budget_bytes // max(1, self.N * frame)))
        self.L = int(ring_slots)
        cfg = CbevConfig()
        cfg.num_envs = self.N
        cfg.fov_size = size
        cfg.anchor_x_frac, cfg.anchor_y_frac = float(anchor[0]), float(anchor[1])
        cfg.obs_h, cfg.obs_w = int(obs_size[0]), int(obs_size[1])
        cfg.obs_mode = obs_mode
        cfg.mask_mode = MASK_MODES[mask_mode]
        cfg.frame_stack = self.F
        cfg.ring_slots = self.L
        cfg.action_mode = action_mode
        table = np.zeros((16, 3), dtype=np.float32)
        if action_mode == ACTION_DISCRETE:
            dt = np.asarray(discrete_table, dtype=np.float32).reshape(-1, 3)
            table[: len(dt)] = dt
            cfg.n_discrete = len(dt)
        for i, v in enumerate(table.ravel()):
            cfg.discrete_table[i] = float(v)
        cfg.reward_mode = reward_mode
        cfg.autoreset = autoreset
        cfg.max_actors = int(max_actors)
        cfg.trajectory_steps = int(trajectory_steps)
        params = dict(CARL_DEFAULTS)
        params.update(SHAPING_DEFAULTS)
        fixed = {"zero_speed_reward_offroad": True, "zero_progress_reward_offroad": True}  # what the kernels implement
        for k, v in (reward_params or {}).items():
            if k in params:
                params[k] = v
            elif k in fixed:
                if bool(v) != fixed[k]:
                    raise CbevError(f"reward parameter {k}={v!r} is not implemented (the engine computes {k}={fixed[k]})")
            else:
                raise CbevError(f"unknown reward parameter {k!r}; known: {sorted(params) + sorted(fixed)}")
        for k, v in params.items():
            setattr(cfg, k, v)
        cfg.seed = int(seed) & (2**64 - 1)
        self.cfg = cfg
        self.action_mode = action_mode
        self.handle = _P()
        _check(self.lib, self.lib.cbev_create(C.byref(cfg), C.byref(self.handle)))
        self.frame_bytes = int(self.lib.cbev_frame_bytes(self.handle))
        assert self.frame_bytes == frame
        # caller-owned buffers
        dev = self.device
        self.ring = torch.empty(self.N * self.L * frame, dtype=torch.uint8, device=dev)
        _check(self.lib, self.lib.cbev_bind_obs_ring(self.handle, self.ring.data_ptr(), self.ring.numel()))
        # reward f64[N] | terminated u8[N] | truncated u8[N] back to back: one D2H copy brings all three to the host
        self._rtt = torch.zeros(self.N * 10, dtype=torch.uint8, device=dev)
        self.reward = self._rtt[: self.N * 8].view(torch.float64)
        self.terminated = self._rtt[self.N * 8: self.N * 9]
        self.truncated = self._rtt[self.N * 9:]
        self._host = None  # pinned mirrors, allocated by the first step_host_full
        self.cause = torch.zeros(self.N, dtype=torch.uint8, device=dev)
        self.hero = torch.zeros(self.N, len(HERO_FIELDS), dtype=torch.float64, device=dev)
        self.episode = torch.zeros(self.N, len(EPISODE_FIELDS), dtype=torch.float64, device=dev)
        self.stats_buf = torch.zeros(STATS_FIELDS, dtype=torch.float64, device=dev)
        self._out = CbevStepOut(self.reward.data_ptr(), self.terminated.data_ptr(), self.truncated.data_ptr(),
                                self.cause.data_ptr(), self.hero.data_ptr(), self.episode.data_ptr())
        self._keep = []
        self.n_scenes = 0

    # ------------------------------------------------------------------ uploads
    def upload_map(self, cls_map: np.ndarray):
        m = np.ascontiguousarray(cls_map, dtype=np.uint8)
        _check(self.lib, self.lib.cbev_upload_map(self.handle, m.ctypes.data, m.shape[1], m.shape[0]))
        self.map_hw = m.shape

    def upload_pool(self, packed: dict):
        """packed = carlabev_env_b200.pool.pack_pool(scenes)."""
        d = CbevPoolDesc()
        keep = []

        def arr(key, dtype):
            a = np.ascontiguousarray(packed[key], dtype=dtype)
            keep.append(a)
            return a.ctypes.data if a.size else None

        d.n_scenes = int(packed["n_scenes"])
        d.n_actors_total = int(len(packed["act_kind"]))
        d.n_tl_total = int(len(packed["tl_color"]))
        for key, dt in (("ego_state0", np.float64), ("ego_target_speed", np.float64), ("ego_tidx0", np.int32),
                        ("len_ego_route", np.float64), ("num_vehicles", np.int32), ("ego_off", np.int32),
                        ("rew_off", np.int32), ("actor_off", np.int32), ("tl_off", np.int32), ("ego_cx", np.float64),
                        ("ego_cy", np.float64), ("ego_cyaw", np.float64), ("rew_rx", np.int32), ("rew_ry", np.int32),
                        ("act_kind", np.uint8), ("act_state0", np.float64), ("act_tidx0", np.int32),
                        ("act_cruise_px", np.float64), ("act_cruise_mps", np.float64), ("act_beh", np.uint8),
                        ("act_beh_p", np.float64), ("act_route_off", np.int32), ("act_raw_off", np.int32),
                        ("act_cx", np.float64), ("act_cy", np.float64), ("act_cyaw", np.float64),
                        ("act_raw_x", np.float64), ("act_raw_y", np.float64), ("tl_rect", np.int32),
                        ("tl_color", np.uint8)):
            setattr(d, key, arr(key, dt))
        # cumulative reward-route lengths exactly as carl_reward_fn.py:20-26 computes them (np.hypot of int deltas)
        rx, ry, off = packed["rew_rx"], packed["rew_ry"], packed["rew_off"]
        cum = np.zeros(len(rx), dtype=np.float64)
        for s in range(d.n_scenes):
            lo, hi = int(off[s]), int(off[s + 1])
            acc = 0.0
            for i in range(lo + 1, hi):
                acc = acc + np.hypot(rx[i] - rx[i - 1], ry[i] - ry[i - 1])
                cum[i] = acc
        keep.append(cum)
        d.rew_cum = cum.ctypes.data if cum.size else None
        sg = savgol_operators()
        keep.append(sg)
        d.sg_mat = sg.ctypes.data
        _check(self.lib, self.lib.cbev_upload_scene_pool(self.handle, C.byref(d)))
        self.n_scenes = d.n_scenes

    def generate_scripted_pool(self, kinds, levels, seeds):
        """Row f2: generate lead_brake / jaywalk scenes ON THE DEVICE from (kind, level, scene_seed) and make them the
        resident pool (cbev_generate_scenes).  kinds: "lead_brake" | "jaywalk" (or 1 | 2).  Returns the number of
        samples every scene needed (the reference's reset retry loop)."""
        ids = {"lead_brake": 1, "jaywalk": 2}
        k = np.ascontiguousarray([ids.get(v, v) for v in kinds], dtype=np.uint8)
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        sd = np.ascontiguousarray(seeds, dtype=np.int64)
        assert len(k) == len(lv) == len(sd)
        sg = np.ascontiguousarray(savgol_operators())
        att = np.zeros(len(k), dtype=np.int32)
        _check(self.lib, self.lib.cbev_generate_scenes(self.handle, len(k), k.ctypes.data, lv.ctypes.data, sd.ctypes.data,
                                                       sg.ctypes.data, att.ctypes.data))
        self.n_scenes = len(k)
        return att

    def read_pool(self) -> dict:
        """The resident pool read back into host arrays (keys of pool.pack_pool that live on the device)."""
        c = np.zeros(8, dtype=np.int32)
        _check(self.lib, self.lib.cbev_pool_counts(self.handle, c.ctypes.data))
        n, na, nr, nq, npts, nw, ntl = (int(v) for v in c[:7])
        spec = (("ego_state0", np.float64, n * 4), ("ego_target_speed", np.float64, n), ("ego_tidx0", np.int32, n),
                ("len_ego_route", np.float64, n), ("num_vehicles", np.int32, n), ("ego_off", np.int32, n + 1),
                ("rew_off", np.int32, n + 1), ("actor_off", np.int32, n + 1), ("tl_off", np.int32, n + 1),
                ("ego_cx", np.float64, nr), ("ego_cy", np.float64, nr), ("ego_cyaw", np.float64, nr),
                ("rew_rx", np.int32, nq), ("rew_ry", np.int32, nq), ("rew_cum", np.float64, nq),
                ("act_kind", np.uint8, na), ("act_state0", np.float64, na * 4), ("act_tidx0", np.int32, na),
                ("act_cruise_px", np.float64, na), ("act_cruise_mps", np.float64, na), ("act_beh", np.uint8, na),
                ("act_beh_p", np.float64, na * 4), ("act_route_off", np.int32, na + 1), ("act_raw_off", np.int32, na + 1),
                ("act_cx", np.float64, npts), ("act_cy", np.float64, npts), ("act_cyaw", np.float64, npts),
                ("act_raw_x", np.float64, nw), ("act_raw_y", np.float64, nw), ("tl_rect", np.int32, ntl * 4),
                ("tl_color", np.uint8, ntl))
        d = CbevPoolDesc()
        out = {"n_scenes": n}
        for key, dt, cnt in spec:
            a = np.zeros(max(cnt, 1), dtype=dt)
            out[key] = a[:cnt]
            setattr(d, key, a.ctypes.data)
        _check(self.lib, self.lib.cbev_read_scene_pool(self.handle, C.byref(d)))
        out["ego_state0"] = out["ego_state0"].reshape(n, 4)
        out["act_state0"] = out["act_state0"].reshape(na, 4)
        out["act_beh_p"] = out["act_beh_p"].reshape(na, 4)
        out["tl_rect"] = out["tl_rect"].reshape(ntl, 4)
        return out

    # ------------------------------------------------------------------ stepping
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, scene_ids, mask=None):
        t = self.torch
        ids = t.as_tensor(scene_ids, dtype=t.int32, device=self.device).contiguous()
        m = None
        if mask is not None:
            m = t.as_tensor(mask, device=self.device).to(t.uint8).contiguous()
        _check(self.lib, self.lib.cbev_reset(self.handle, None if m is None else m.data_ptr(), ids.data_ptr(),
                                             self._stream()))
        self._keep = [ids, m]
        return self.obs()

    def step(self, actions):
        """actions: device tensor int64[N] (discrete) or float32[N,3] (continuous)."""
        _check(self.lib, self.lib.cbev_step(self.handle, actions.data_ptr(), C.byref(self._out), self._stream()))

    def step_host(self, actions_host, reward_host, term_host, trunc_host):
        """Pinned host tensors in / out (the e2e path of bench.py)."""
        _check(self.lib, self.lib.cbev_step_host(self.handle, actions_host.data_ptr(), reward_host.data_ptr(),
                                                 term_host.data_ptr(), trunc_host.data_ptr(), self._stream()))

    def step_host_full(self, actions, episode=True):
        """Pinned host actions (or a CUDA tensor) in; reward / terminated / truncated (and the episode block) come back into pinned host
        mirrors (`host_reward`, `host_terminated`, `host_truncated`, `host_episode`) through D2H copies that overlap
        the raster kernel.  All device outputs (hero, cause, episode, ...) are written as in `step`.  Call
        `wait_host_outputs()` before reading the mirrors."""
        t = self.torch
        if self._host is None:
            blk = t.zeros(self.N * 10, dtype=t.uint8).pin_memory()
            self.host_reward = blk[: self.N * 8].view(t.float64)
            self.host_terminated = blk[self.N * 8: self.N * 9]
            self.host_truncated = blk[self.N * 9:]
            self.host_episode = t.zeros(self.N, len(EPISODE_FIELDS), dtype=t.float64).pin_memory()
            self._host = CbevHostOut(self.host_reward.data_ptr(), self.host_terminated.data_ptr(),
                                     self.host_truncated.data_ptr(), None, self.host_episode.data_ptr())
            self._host_noep = CbevHostOut(self.host_reward.data_ptr(), self.host_terminated.data_ptr(),
                                          self.host_truncated.data_ptr(), None, None)
            self._host_blk = blk
        fn = self.lib.cbev_step_ex if actions.is_cuda else self.lib.cbev_step_host_ex
        _check(self.lib, fn(self.handle, actions.data_ptr(), C.byref(self._out),
                            C.byref(self._host if episode else self._host_noep), self._stream()))

    def invalidate(self):
        """Every env must be reset before the next step (after the pool was replaced by an unrelated one)."""
        _check(self.lib, self.lib.cbev_invalidate(self.handle))

    def wait_host_outputs(self):
        """Block until the host copies of the last step_host / step_host_full have landed (the raster kernel of that
        step may still be running)."""
        _check(self.lib, self.lib.cbev_wait_host_outputs(self.handle))

    @property
    def head(self) -> int:
        h = C.c_int32(-1)
        _check(self.lib, self.lib.cbev_obs_head(self.handle, C.byref(h)))
        return h.value

    def obs(self):
        """Zero-copy view of the current observation window in the ring."""
        t = self.torch
        head, F, L, N = self.head, self.F, self.L, self.N
        if self.obs_mode == OBS_SEMANTIC:
            ring = self.ring.view(t.float32).view(N, L, self.channels, *self.obs_hw)
            return ring[:, head - F + 1: head + 1].view(N, F * self.channels, *self.obs_hw)
        if self.obs_mode == OBS_GRAY:
            ring = self.ring.view(N, L, *self.obs_hw)
            return ring[:, head - F + 1: head + 1]
        return self.ring.view(N, self.size, self.size, 3)

    def upload_fov_mask(self, mask):
        """mask: uint8 [S, S], non-zero = painted black (fov_masked); None removes it."""
        if mask is None:
            _check(self.lib, self.lib.cbev_upload_fov_mask(self.handle, None))
            return
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        assert m.shape == (self.size, self.size)
        _check(self.lib, self.lib.cbev_upload_fov_mask(self.handle, m.ctypes.data))

    def fuse(self, mode: str):
        """Temporal fusion of the current window: 'vehicle_temporal' -> [N, C+2, h, w], 'vehicle_weighted' -> [N, C, h, w]."""
        code = {"vehicle_temporal": 1, "vehicle_weighted": 2}[mode]
        cout = self.channels + 2 if code == 1 else self.channels
        key = ("fuse", code)
        buf = getattr(self, "_fuse_buf", {}).get(key)
        if buf is None:
            buf = self.torch.empty(self.N, cout, *self.obs_hw, dtype=self.torch.float32, device=self.device)
            self._fuse_buf = {**getattr(self, "_fuse_buf", {}), key: buf}
        _check(self.lib, self.lib.cbev_fuse(self.handle, code, buf.data_ptr(), self._stream()))
        return buf

    def set_debug_flags(self, flags: int):
        _check(self.lib, self.lib.cbev_set_debug_flags(self.handle, int(flags)))

    def read_trace(self):
        """Per-CTA phase timestamps of the last raster launch, uint64 [N, 8] (set_debug_flags(4) first)."""
        out = np.zeros((self.N, 8), dtype=np.uint64)
        _check(self.lib, self.lib.cbev_debug_read_trace(self.handle, out.ctypes.data))
        return out

    def keep_fov(self, on=True):
        _check(self.lib, self.lib.cbev_keep_fov(self.handle, int(on)))

    def fov(self):
        """Last rendered palette-index frames, uint8 [N, S, S] (needs keep_fov(True) before the step)."""
        out = self.torch.empty(self.N, self.size, self.size, dtype=self.torch.uint8, device=self.device)
        _check(self.lib, self.lib.cbev_copy_fov(self.handle, out.data_ptr(), self._stream()))
        return out

    def get_state(self, max_actors=None):
        """(ego [N, 16], actors [N, max_actors, 8]) of cbev_get_state; the library always writes the engine's full
        actor capacity, the returned block is cut to `max_actors` columns."""
        cap = max(1, int(self.cfg.max_actors))
        ego = np.zeros((self.N, 16), dtype=np.float64)
        act = np.zeros((self.N, cap, 8), dtype=np.float64)
        _check(self.lib, self.lib.cbev_get_state(self.handle, ego.ctypes.data, act.ctypes.data))
        return ego, (act if max_actors is None else act[:, :max(1, max_actors)])

    def set_ego_state(self, ego):
        e = np.ascontiguousarray(ego, dtype=np.float64)
        _check(self.lib, self.lib.cbev_set_ego_state(self.handle, e.ctypes.data))

    def set_state(self, ego=None, actors=None):
        """Debug / parity: write back blocks in the layout of get_state (either may be None)."""
        e = None if ego is None else np.ascontiguousarray(ego, dtype=np.float64)
        a = None if actors is None else np.ascontiguousarray(actors, dtype=np.float64)
        if a is not None:
            assert a.shape == (self.N, max(1, int(self.cfg.max_actors)), 8), a.shape
        _check(self.lib, self.lib.cbev_set_state(self.handle, None if e is None else e.ctypes.data,
                                                 None if a is None else a.ctypes.data))

    def read_stats(self, reset=False):
        _check(self.lib, self.lib.cbev_read_stats(self.handle, self.stats_buf.data_ptr(), int(reset), self._stream()))
        return self.stats_buf

    def profile(self, on: bool):
        _check(self.lib, self.lib.cbev_profile_enable(self.handle, int(on)))

    def profile_read(self):
        """(move_ms, render_ms, judge_ms, steps) summed over the profiled steps since the last read; k_judge runs on
        the side stream concurrently with k_render."""
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        _check(self.lib, self.lib.cbev_profile_read_ex(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        return a.value, b.value, c.value, n.value

    @property
    def launches(self) -> int:
        return int(self.lib.cbev_launch_count(self.handle))

    def close(self):
        if self.handle:
            self.torch.cuda.synchronize(self.device)
            self.lib.cbev_destroy(self.handle)
            self.handle = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
