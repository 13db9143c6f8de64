"""carlabev_env_b200 -- B200-native batched stepping engine behind the CarlaBEV VectorEnv surface.

Only what the hot path needs lives here: `csrc/` (sm_100a CUDA kernels + the C ABI of include/cbev.h),
the ctypes binding (`engine.py`), the host-side mirror of the reference interface (`config.py`,
`spaces.py`, `vector_env.py`), the scene pool container / scripted-scenario generators
(`pool.py`, `scenes.py`) and the episode-statistics all-reduce (`distributed.py`).
Importing the package does not import torch or load the CUDA library; constructing an env does,
and fails loudly when the library or a CUDA device is missing (there is no CPU fallback).
"""
from .config import (EnvConfig, RunConfig, get_action_profile_spec, get_difficulty_spec, get_env_capabilities,  # noqa: F401
                     get_reward_profile_spec, resolve_env_profiles, validate_env_config, validate_run_config)
from .reset import (AuthoredSceneReset, RandomNavigationReset, ScenarioConfigReset, ScenarioPresetReset,  # noqa: F401
                    build_authored_scene_options, build_random_navigation_options, build_reset_options,
                    build_scenario_config_options, build_scenario_options_from_config,
                    build_scenario_preset_options)

__version__ = "0.1.0"


def make_env(cfg=None, eval=False, **kw):  # noqa: A002
    from .vector_env import make_env as _make_env

    return _make_env(cfg, eval, **kw)


__all__ = ["EnvConfig", "RunConfig", "make_env", "validate_env_config", "validate_run_config", "__version__",
           "get_action_profile_spec", "get_difficulty_spec", "get_env_capabilities", "get_reward_profile_spec",
           "resolve_env_profiles", "AuthoredSceneReset", "RandomNavigationReset", "ScenarioConfigReset",
           "ScenarioPresetReset", "build_authored_scene_options", "build_random_navigation_options",
           "build_reset_options", "build_scenario_config_options", "build_scenario_options_from_config",
           "build_scenario_preset_options"]
