"""gymnasium-lite SyncVectorEnv restating gymnasium 1.2.2 (AutoresetMode.DISABLED path)."""
from copy import deepcopy
from enum import Enum

import numpy as np


class AutoresetMode(Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


class VectorEnv:
    pass


class SyncVectorEnv(VectorEnv):
    def __init__(self, env_fns, copy=True, observation_mode="same", autoreset_mode=AutoresetMode.NEXT_STEP):
        assert autoreset_mode == AutoresetMode.DISABLED, "only the reference's mode is restated"
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.autoreset_mode = autoreset_mode
        self.copy = copy
        self.single_observation_space = self.envs[0].observation_space
        self.single_action_space = self.envs[0].action_space
        self.metadata = getattr(self.envs[0], "metadata", {})
        self._env_obs = [None] * self.num_envs
        self._rewards = np.zeros((self.num_envs,), dtype=np.float64)
        self._terminations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._truncations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._autoreset_envs = np.zeros((self.num_envs,), dtype=np.bool_)

    def _add_info(self, infos, env_info, i):
        for key, value in env_info.items():
            if isinstance(value, dict):
                array = self._add_info(infos.get(key, {}), value, i)
            else:
                if key not in infos:
                    if isinstance(value, (int, float, bool, np.number, np.bool_)):
                        array = np.zeros(self.num_envs, dtype=type(value) if not isinstance(value, np.generic) else value.dtype)
                    elif isinstance(value, np.ndarray):
                        array = np.zeros((self.num_envs, *value.shape), dtype=value.dtype)
                    else:
                        array = np.full(self.num_envs, fill_value=None, dtype=object)
                else:
                    array = infos[key]
                array[i] = value
            mask = infos.get(f"_{key}", np.zeros(self.num_envs, dtype=np.bool_))
            mask[i] = True
            infos[key], infos[f"_{key}"] = array, mask
        return infos

    def reset(self, *, seed=None, options=None):
        if seed is None:
            seed = [None] * self.num_envs
        elif isinstance(seed, int):
            seed = [seed + i for i in range(self.num_envs)]
        infos = {}
        if options is not None and "reset_mask" in options:
            options = dict(options)
            reset_mask = options.pop("reset_mask")
            assert isinstance(reset_mask, np.ndarray) and reset_mask.dtype == np.bool_
            assert reset_mask.shape == (self.num_envs,) and np.any(reset_mask)
            self._terminations[reset_mask] = False
            self._truncations[reset_mask] = False
            self._autoreset_envs[reset_mask] = False
            for i, (env, s, m) in enumerate(zip(self.envs, seed, reset_mask)):
                if m:
                    self._env_obs[i], env_info = env.reset(seed=s, options=options)
                    infos = self._add_info(infos, env_info, i)
        else:
            self._terminations[:] = False
            self._truncations[:] = False
            self._autoreset_envs[:] = False
            for i, (env, s) in enumerate(zip(self.envs, seed)):
                self._env_obs[i], env_info = env.reset(seed=s, options=options)
                infos = self._add_info(infos, env_info, i)
        obs = np.stack(self._env_obs)
        return (deepcopy(obs) if self.copy else obs), infos

    def step(self, actions):
        infos = {}
        for i, action in enumerate(actions):
            assert not self._autoreset_envs[i], f"{self._autoreset_envs=}"
            (self._env_obs[i], self._rewards[i], self._terminations[i], self._truncations[i], env_info) = self.envs[i].step(action)
            infos = self._add_info(infos, env_info, i)
        obs = np.stack(self._env_obs)
        self._autoreset_envs = np.logical_or(self._terminations, self._truncations)
        return (
            deepcopy(obs) if self.copy else obs,
            np.copy(self._rewards),
            np.copy(self._terminations),
            np.copy(self._truncations),
            infos,
        )

    def close(self):
        for env in self.envs:
            env.close()
