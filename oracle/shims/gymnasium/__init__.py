"""gymnasium-lite: TEST INFRASTRUCTURE ONLY (never imported by the product path).

Restates the slice of gymnasium 1.2.2 (reference uv.lock:251-252) that
CarlaBEV.envs.make_env / wrap_env touch, so the unmodified reference can be
stepped in this container where gymnasium is not installed.  Wrapper values are
PARITY UNPINNED by the reference's tests (they pin only shapes).
"""
from __future__ import annotations

import numpy as np

from . import spaces  # noqa: F401


class Env:
    metadata: dict = {}
    render_mode = None
    spec = None
    observation_space = None
    action_space = None
    _np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)
        return None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._observation_space = None
        self._action_space = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def observation_space(self):
        return self._observation_space if self._observation_space is not None else self.env.observation_space

    @observation_space.setter
    def observation_space(self, v):
        self._observation_space = v

    @property
    def action_space(self):
        return self._action_space if self._action_space is not None else self.env.action_space

    @action_space.setter
    def action_space(self, v):
        self._action_space = v

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        return self.observation(obs), r, term, trunc, info

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        return obs, self.reward(r), term, trunc, info

    def reward(self, reward):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


from . import wrappers  # noqa: E402,F401
from . import vector  # noqa: E402,F401
from . import envs  # noqa: E402,F401
