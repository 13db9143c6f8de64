"""Typed reset requests -> `reset(options=...)` dicts.

Host-side mirror of the reference's config/reset.py:15-196 (same class and field names, same option keys),
so callers that build their reset options through it keep working:

    envs.reset(options=build_reset_options(RandomNavigationReset(difficulty_id="rt_hard_v1", scene_seed=7)))
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .config import get_difficulty_spec


@dataclass
class RandomNavigationReset:
    difficulty_id: str | None = None
    num_vehicles: int = 25
    route_dist_range: tuple = (30, 130)
    ego_route_graph: str = "full_vehicle"
    route_profile: str | None = None
    route_profile_mix: dict | None = None
    min_turns: int | None = None
    max_turns: int | None = None
    intersection_required: bool | None = None
    max_route_attempts: int | None = None
    scene_seed: int | None = None
    route_seed: int | None = None
    traffic_seed: int | None = None
    scenario_seed: int | None = None


@dataclass
class ScenarioPresetReset:
    preset_id: str
    overrides: dict = field(default_factory=dict)


@dataclass
class AuthoredSceneReset:
    config_file: str
    variation_enabled: bool = False
    variation_seed: int | None = None


@dataclass
class ScenarioConfigReset:
    scenario_id: str
    level: int = 1
    anchor_x: int | None = None
    anchor_y: int | None = None
    parameters: dict = field(default_factory=dict)


def _with_mask(options: dict, reset_mask) -> dict:
    if reset_mask is not None:
        options["reset_mask"] = np.asarray(reset_mask, dtype=bool)
    return options


def build_random_navigation_options(request: RandomNavigationReset, *, reset_mask=None) -> dict[str, Any]:
    """config/reset.py:72-117: a difficulty preset overrides traffic_enabled / num_vehicles / route_dist_range."""
    options = {"scene": "rdm", "num_vehicles": int(request.num_vehicles),
               "route_dist_range": list(request.route_dist_range), "ego_route_graph": request.ego_route_graph}
    for key, cast in (("route_profile", str), ("route_profile_mix", dict), ("min_turns", int), ("max_turns", int),
                      ("intersection_required", bool), ("max_route_attempts", int), ("scene_seed", int),
                      ("route_seed", int), ("traffic_seed", int), ("scenario_seed", int)):
        value = getattr(request, key)
        if value is not None:
            options[key] = cast(value)
    if request.difficulty_id is not None:
        spec = get_difficulty_spec(request.difficulty_id)
        options.update(difficulty_id=spec["difficulty_id"], traffic_enabled=spec["traffic_enabled"],
                       num_vehicles=int(spec["num_vehicles"]), route_dist_range=list(spec["route_dist_range"]))
        if spec.get("ego_target_speed") is not None:
            options["ego_target_speed"] = float(spec["ego_target_speed"])
    return _with_mask(options, reset_mask)


def build_scenario_preset_options(request: ScenarioPresetReset, *, reset_mask=None) -> dict[str, Any]:
    from .scenes import SCENARIO_PRESETS, scenario_preset_options

    options = scenario_preset_options(request.preset_id, request.overrides)
    options["scenario_preset_id"] = request.preset_id                    # scenarios/specs.py:174-176
    options["scenario_preset_scene"] = SCENARIO_PRESETS[request.preset_id]["scene"]
    return _with_mask(options, reset_mask)


def build_authored_scene_options(request: AuthoredSceneReset, *, reset_mask=None) -> dict[str, Any]:
    options = {"config_file": request.config_file, "variation_enabled": request.variation_enabled}
    if request.variation_seed is not None:
        options["variation_seed"] = int(request.variation_seed)
    return _with_mask(options, reset_mask)


def build_scenario_config_options(request: ScenarioConfigReset, *, reset_mask=None) -> dict[str, Any]:
    options = dict(request.parameters)
    options["scene"] = request.scenario_id
    options["level"] = int(request.level)
    if request.anchor_x is not None:
        options["anchor_x"] = int(request.anchor_x)
    if request.anchor_y is not None:
        options["anchor_y"] = int(request.anchor_y)
    return _with_mask(options, reset_mask)


def build_scenario_options_from_config(config: dict, *, overrides: dict | None = None, reset_mask=None) -> dict[str, Any]:
    """config/reset.py:161-172 -> scenarios/specs.py:248-272: a loaded scenario config + overrides -> options
    (only the keys the config carries; scenes.scenario_config_options is the file loader that also fills defaults)."""
    overrides = dict(overrides or {})
    if reset_mask is None and "reset_mask" in overrides:
        reset_mask = overrides.pop("reset_mask")
    options = dict(config.get("parameters", {}))
    anchor = config.get("anchor", {}) or {}
    if anchor.get("x") is not None:
        options["anchor_x"] = anchor["x"]
    if anchor.get("y") is not None:
        options["anchor_y"] = anchor["y"]
    options["level"] = int(config.get("level", 1))
    options["scene"] = config["scenario_id"]
    for key, value in overrides.items():
        if key in ("config_file", "scene", "reset_mask") or value is None:
            continue
        options[key] = value
    return _with_mask(options, reset_mask)


def build_reset_options(request, *, reset_mask=None) -> dict[str, Any]:
    if isinstance(request, RandomNavigationReset):
        return build_random_navigation_options(request, reset_mask=reset_mask)
    if isinstance(request, ScenarioPresetReset):
        return build_scenario_preset_options(request, reset_mask=reset_mask)
    if isinstance(request, AuthoredSceneReset):
        return build_authored_scene_options(request, reset_mask=reset_mask)
    if isinstance(request, ScenarioConfigReset):
        return build_scenario_config_options(request, reset_mask=reset_mask)
    raise TypeError(f"Unsupported reset request type: {type(request)!r}")
