"""Configuration surface, kept field-for-field with the reference so callers can switch over.

Mirrors CarlaBEV/config/env.py:43-207 (EnvConfig, RunConfig, legacy-name normalisation and
validation), config/action_profiles.py:35-76, config/reward_profiles.py:19-45 and
config/difficulty.py:22-47.  Plain dataclasses (no pydantic dependency); invalid values raise
ValueError / KeyError like the reference's validators do.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field, fields, is_dataclass
from typing import Any

OBS_MODES = ("bev_rgb", "bev_semantic", "vector")
SEMANTIC_MASK_CH = ("binary", "2-class", "4-class", "5-class", "6-class", "7-class")
TEMPORAL_FUSION_MODES = ("stack", "vehicle_temporal", "vehicle_weighted")
ACTION_MODES = ("discrete", "continuous")
REWARD_MODES = ("shaping", "carl")

# config/action_profiles.py:35-76
ACTION_PROFILES: dict[str, dict[str, Any]] = {
    "discrete9_v1": dict(action_mode="discrete", discrete_actions=[
        (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0), (1.0, -1.0, 0.0),
        (0.0, 1.0, 0.0), (0.0, -1.0, 0.0), (0.0, 1.0, 1.0), (0.0, -1.0, 1.0)]),
    "discrete13_v1": dict(action_mode="discrete", discrete_actions=[
        (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0), (1.0, 0.5, 0.0), (1.0, -0.5, 0.0),
        (1.0, -1.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.5, 0.0), (0.0, -0.5, 0.0), (0.0, -1.0, 0.0), (0.0, 1.0, 1.0),
        (0.0, -1.0, 1.0)]),
    "continuous_gsb_v1": dict(action_mode="continuous", low=(0.0, -1.0, 0.0), high=(1.0, 1.0, 1.0)),
}
# config/reward_profiles.py:19-45
REWARD_PROFILES: dict[str, dict[str, Any]] = {
    "carl_base_v1": dict(family="carl", parameters={}),
    "carl_safety_v1": dict(family="carl", parameters=dict(
        lane_center_exponent=1.5, lane_center_floor=0.15, off_lane_penalty=0.05, speed_penalty_scale=4.0,
        speed_penalty_floor=0.05, ttc_threshold=5.0, ttc_penalty_floor=0.05, reward_scale=0.85,
        comfort_penalty_floor=0.25)),
    "shaping_base_v1": dict(family="shaping", parameters={}),
}
# config/difficulty.py:22-47
DIFFICULTY_PRESETS: dict[str, dict[str, Any]] = {
    "rt_no_traffic_v1": dict(traffic_enabled=False, num_vehicles=0, route_dist_range=(30, 80)),
    "rt_easy_v1": dict(traffic_enabled=True, num_vehicles=8, route_dist_range=(30, 80)),
    "rt_medium_v1": dict(traffic_enabled=True, num_vehicles=16, route_dist_range=(40, 100)),
    "rt_hard_v1": dict(traffic_enabled=True, num_vehicles=25, route_dist_range=(50, 130)),
}
LEGACY_ACTION_PROFILE_IDS = {"discrete": "discrete9_v1", "continuous": "continuous_gsb_v1"}
LEGACY_REWARD_PROFILE_IDS = {"carl": "carl_base_v1", "shaping": "shaping_base_v1"}

# defaults of CaRLRewardFn (carl_reward_fn.py:73-88) and RewardFn (reward.py:14-47)
CARL_DEFAULTS = dict(lane_center_exponent=1.0, lane_center_floor=0.2, off_lane_penalty=0.0, speed_penalty_scale=6.0,
                     speed_penalty_floor=0.1, ttc_threshold=4.0, ttc_penalty_floor=0.1)
SHAPING_DEFAULTS = dict(max_actions=5000, offroad_terminate_after=40, sidewalk_step_penalty=-0.12,
                        sidewalk_penalty_scale=-0.006, k_lat_quadratic=0.004, k_progress=0.06, k_flow=0.010,
                        k_align_bonus=0.02, k_reverse=0.03, k_ttc=0.03, alive_bias=0.0025, k_smooth=0.0006,
                        k_steer_smooth=0.003, k_steer_jerk=0.01, k_route_dev=0.006, route_dev_start=8.0,
                        max_speed_for_flow=6.0, lat_clip=4.0, yaw_small=0.12, lat_small=0.8)


def get_action_profile_spec(action_profile_id: str) -> dict:
    try:
        return dict(ACTION_PROFILES[action_profile_id], action_profile_id=action_profile_id)
    except KeyError as exc:
        raise KeyError(f"Unknown action_profile_id={action_profile_id!r}. Available action profiles: "
                       f"{', '.join(sorted(ACTION_PROFILES))}") from exc


def get_reward_profile_spec(reward_profile_id: str) -> dict:
    try:
        return dict(REWARD_PROFILES[reward_profile_id], reward_profile_id=reward_profile_id)
    except KeyError as exc:
        raise KeyError(f"Unknown reward_profile_id={reward_profile_id!r}. Available reward profiles: "
                       f"{', '.join(sorted(REWARD_PROFILES))}") from exc


def get_difficulty_spec(difficulty_id: str) -> dict:
    try:
        return dict(DIFFICULTY_PRESETS[difficulty_id], difficulty_id=difficulty_id)
    except KeyError as exc:
        raise KeyError(f"Unknown difficulty_id={difficulty_id!r}. Available difficulty presets: "
                       f"{', '.join(sorted(DIFFICULTY_PRESETS))}") from exc


def list_action_profile_ids() -> list:
    return sorted(ACTION_PROFILES)


def list_reward_profile_ids() -> list:
    return sorted(REWARD_PROFILES)


def list_difficulty_ids() -> list:
    return sorted(DIFFICULTY_PRESETS)


def get_env_capabilities() -> dict:
    """config/env.py:328-355: what this implementation of the environment accepts (Town01 at the map scales
    64 / 128 / 256; `vector` observations are reachable through `CarlaBEVVectorEnv.vector_observation()`, not make_env)."""
    from .scenes import SCENARIO_PRESETS, _SPEC_DEFAULTS

    masks = ["binary", "2-class", "4-class", "5-class", "6-class", "7-class"]
    fusion = ["stack", "vehicle_temporal", "vehicle_weighted"]
    return {
        "maps": ["Town01"], "obs_modes": ["bev_rgb", "bev_semantic", "vector"],
        "semantic_mask_channels": masks, "semantic_mask_ch": masks,
        "temporal_fusion_modes": fusion, "temporal_fusion_mode": fusion,
        "action_modes": ["discrete", "continuous"], "action_profile_ids": list_action_profile_ids(),
        "reward_modes": ["shaping", "carl"], "reward_profile_ids": list_reward_profile_ids(),
        "difficulty_ids": list_difficulty_ids(), "render_modes": ["rgb_array"],
        "supports_vector_make_env": False, "scenario_ids": list(_SPEC_DEFAULTS),
        "scenario_preset_ids": list(SCENARIO_PRESETS),
    }


def resolve_env_profiles(env_cfg) -> dict:
    """config/env.py:318-325."""
    cfg = validate_env_config(env_cfg)
    return {"action": get_action_profile_spec(cfg.action_profile_id),
            "reward": get_reward_profile_spec(cfg.reward_profile_id)}


@dataclass
class EnvConfig:
    """config/env.py:43-74 (same names and defaults)."""

    seed: int = 0
    fps: int = 15
    size: int = 128
    env_id: str = "CarlaBEV-v0"
    map_name: str = "Town01"
    obs_size: tuple = (96, 96)
    obs_mode: str = "bev_semantic"
    semantic_mask_ch: str = "6-class"
    temporal_fusion_mode: str = "stack"
    fov_masked: bool = False
    ego_anchor_x_frac: float = 0.5
    ego_anchor_y_frac: float = 0.5
    frame_stack: int = 4
    action_mode: str = "discrete"
    action_profile_id: str | None = None
    render_mode: str = "rgb_array"
    max_actions: int = 5000
    scenes_path: str = "assets/scenes"
    reward_mode: str = "carl"
    reward_profile_id: str | None = None
    traffic_enabled: bool = True
    max_vehicles: int = 50
    route_direction_metrics_enabled: bool = False

    def __post_init__(self):
        self.obs_size = tuple(int(v) for v in self.obs_size)
        if self.action_profile_id is None:
            self.action_profile_id = LEGACY_ACTION_PROFILE_IDS.get(self.action_mode, "discrete9_v1")
        if self.reward_profile_id is None:
            self.reward_profile_id = LEGACY_REWARD_PROFILE_IDS.get(self.reward_mode, "carl_base_v1")
        self.validate()

    # config/env.py:105-160
    def validate(self):
        if self.obs_mode not in OBS_MODES:
            raise ValueError(f"obs_mode must be one of {OBS_MODES}, got {self.obs_mode!r}")
        if self.semantic_mask_ch not in SEMANTIC_MASK_CH:
            raise ValueError(f"Unsupported semantic_mask_ch={self.semantic_mask_ch!r}. Expected one of: "
                             f"{', '.join(sorted(SEMANTIC_MASK_CH))}")
        if self.temporal_fusion_mode not in TEMPORAL_FUSION_MODES:
            raise ValueError(f"temporal_fusion_mode must be one of {TEMPORAL_FUSION_MODES}")
        if self.action_mode not in ACTION_MODES:
            raise ValueError(f"action_mode must be one of {ACTION_MODES}")
        if self.reward_mode not in REWARD_MODES:
            raise ValueError(f"reward_mode must be one of {REWARD_MODES}")
        if self.frame_stack < 1:
            raise ValueError("frame_stack must be >= 1")
        if self.temporal_fusion_mode != "stack":
            if self.obs_mode != "bev_semantic":
                raise ValueError("temporal_fusion_mode requires obs_mode='bev_semantic'")
            if self.frame_stack < 3:
                raise ValueError("temporal_fusion_mode requires frame_stack >= 3")
            if self.semantic_mask_ch not in {"4-class", "5-class", "6-class", "7-class"}:
                raise ValueError("temporal_fusion_mode requires a semantic_mask_ch with a vehicle channel "
                                 "(one of: '4-class', '5-class', '6-class', '7-class')")
        if self.obs_size[0] < 1 or self.obs_size[1] < 1:
            raise ValueError("obs_size dimensions must be >= 1")
        if not 0.0 <= self.ego_anchor_x_frac <= 1.0:
            raise ValueError("ego_anchor_x_frac must be within [0.0, 1.0]")
        if not 0.0 <= self.ego_anchor_y_frac <= 1.0:
            raise ValueError("ego_anchor_y_frac must be within [0.0, 1.0]")
        action_spec = get_action_profile_spec(self.action_profile_id)
        reward_spec = get_reward_profile_spec(self.reward_profile_id)
        if action_spec["action_mode"] != self.action_mode:
            raise ValueError(f"action_profile_id={self.action_profile_id!r} resolves to action_mode="
                             f"{action_spec['action_mode']!r}, but EnvConfig.action_mode={self.action_mode!r}")
        if reward_spec["family"] != self.reward_mode:
            raise ValueError(f"reward_profile_id={self.reward_profile_id!r} resolves to reward_mode="
                             f"{reward_spec['family']!r}, but EnvConfig.reward_mode={self.reward_mode!r}")
        if self.map_name != "Town01":
            raise ValueError(f"map_name='{self.map_name}' is missing required assets (only Town01 ships)")
        if self.size not in (64, 128, 256):
            # config/env.py:146-160 checks that Town01-<size>-{sem,rgb}.png exist (64 ... 1024 do); the unmodified
            # reference cannot reset at 512 / 1024 ("hero_on_obstacle" for every seed: quirk C-11), so those are not served
            raise ValueError(f"size={self.size}: Town01 ships at the map scales 64, 128 and 256")

    # legacy computed fields, config/env.py:162-181
    @property
    def obs_space(self) -> str:
        return "vector" if self.obs_mode == "vector" else "bev"

    @property
    def masked(self) -> bool:
        return self.obs_mode == "bev_semantic"

    @property
    def action_space(self) -> str:
        return self.action_mode

    @property
    def reward_type(self) -> str:
        return "carl" if self.reward_mode == "carl" else "shaping"


@dataclass
class RunConfig:
    """config/env.py:184-207, plus engine-only knobs (prefixed `engine_`)."""

    env: EnvConfig = field(default_factory=EnvConfig)
    exp_name: str = "carlabev-run"
    num_envs: int = 1
    seed: int = 1
    capture_video: bool = False
    capture_every: int = 50
    video_output_dir: str | None = None
    video_episode_indices: list | None = None
    video_name_prefix: str = "rl-video"
    cuda: bool = True
    torch_deterministic: bool = True

    def __post_init__(self):
        if isinstance(self.env, dict):
            self.env = validate_env_config(self.env)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")


_ENV_FIELDS = {f.name for f in fields(EnvConfig)}
_RUN_FIELDS = {f.name for f in fields(RunConfig)} - {"env"}


def _to_mapping(value: Any):
    if isinstance(value, (EnvConfig, RunConfig, dict)):
        return value
    if is_dataclass(value):
        return asdict(value)
    if hasattr(value, "model_dump"):
        return value.model_dump(mode="python")
    if hasattr(value, "__dict__"):
        return {k: v for k, v in vars(value).items() if not k.startswith("_")}
    return value


def validate_env_config(cfg) -> EnvConfig:
    """config/env.py:298-301 incl. legacy normalisation (:76-103, :226-264)."""
    raw = _to_mapping(cfg)
    if isinstance(raw, EnvConfig):
        return raw
    if not isinstance(raw, dict):
        raise ValueError(f"cannot build EnvConfig from {type(cfg).__name__}")
    env = {k: v for k, v in raw.items() if k in _ENV_FIELDS}
    if "obs_mode" not in raw:
        if raw.get("obs_space") == "vector":
            env["obs_mode"] = "vector"
        elif raw.get("masked") is False:
            env["obs_mode"] = "bev_rgb"
        else:
            env["obs_mode"] = "bev_semantic"
    if "action_mode" not in raw and "action_space" in raw:
        env["action_mode"] = raw["action_space"]
    if "reward_mode" not in raw and "reward_type" in raw:
        env["reward_mode"] = "carl" if raw["reward_type"] == "carl" else "shaping"
    unknown = set(raw) - _ENV_FIELDS - {"obs_space", "masked", "action_space", "reward_type"}
    if unknown and not hasattr(cfg, "__dict__"):
        raise ValueError(f"Extra inputs are not permitted: {sorted(unknown)}")
    return EnvConfig(**env)


def validate_run_config(cfg) -> RunConfig:
    """config/env.py:304-315: vector observations are not available through make_env."""
    raw = _to_mapping(cfg)
    if isinstance(raw, RunConfig):
        run = raw
    else:
        if not isinstance(raw, dict):
            raise ValueError(f"cannot build RunConfig from {type(cfg).__name__}")
        mapping = {k: v for k, v in raw.items() if k in _RUN_FIELDS}
        if "env" in raw:
            mapping["env"] = validate_env_config(raw["env"])
        run = RunConfig(**mapping)
    if run.env.obs_mode == "vector":
        raise ValueError("obs_mode='vector' is not supported through make_env()/wrap_env() yet. "
                         "Use CarlaBEV() directly if you need vector observations.")
    return run
