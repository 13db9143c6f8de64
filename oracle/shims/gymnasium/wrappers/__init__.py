"""gymnasium-lite wrappers restating gymnasium 1.2.2 semantics."""
import time
from collections import deque

import numpy as np

import gymnasium as gym
from gymnasium import spaces


class ResizeObservation(gym.ObservationWrapper):
    def __init__(self, env, shape):
        super().__init__(env)
        import cv2  # noqa: F401

        self.shape = tuple(shape)
        old = env.observation_space
        self.observation_space = spaces.Box(0, 255, self.shape + tuple(old.shape[2:]), dtype=np.uint8)

    def observation(self, obs):
        import cv2

        return cv2.resize(obs, self.shape[::-1], interpolation=cv2.INTER_AREA)


class GrayscaleObservation(gym.ObservationWrapper):
    def __init__(self, env, keep_dim=False):
        super().__init__(env)
        assert not keep_dim
        self.observation_space = spaces.Box(0, 255, tuple(env.observation_space.shape[:2]), dtype=np.uint8)

    def observation(self, obs):
        return np.sum(np.multiply(obs, np.array([0.2125, 0.7154, 0.0721])), axis=-1).astype(np.uint8)


class FrameStackObservation(gym.Wrapper):
    def __init__(self, env, stack_size, *, padding_type="reset"):
        super().__init__(env)
        assert padding_type == "reset"
        self.stack_size = int(stack_size)
        old = env.observation_space
        low = np.stack([old.low] * self.stack_size)
        high = np.stack([old.high] * self.stack_size)
        self.observation_space = spaces.Box(low, high, dtype=old.dtype)
        self.obs_queue = deque(maxlen=self.stack_size)

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        self.obs_queue.append(obs)
        return np.stack(list(self.obs_queue)), r, term, trunc, info

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        for _ in range(self.stack_size - 1):
            self.obs_queue.append(obs)
        self.obs_queue.append(obs)
        return np.stack(list(self.obs_queue)), info


class RecordEpisodeStatistics(gym.Wrapper):
    def __init__(self, env, buffer_length=100, stats_key="episode"):
        super().__init__(env)
        self._stats_key = stats_key
        self.episode_count = 0
        self.episode_start_time = -1.0
        self.episode_returns = 0.0
        self.episode_lengths = 0

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        self.episode_returns += r
        self.episode_lengths += 1
        if term or trunc:
            assert self._stats_key not in info
            info[self._stats_key] = {
                "r": self.episode_returns,
                "l": self.episode_lengths,
                "t": round(time.perf_counter() - self.episode_start_time, 6),
            }
            self.episode_count += 1
            self.episode_start_time = time.perf_counter()
        return obs, r, term, trunc, info

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        self.episode_start_time = time.perf_counter()
        self.episode_returns = 0.0
        self.episode_lengths = 0
        return obs, info


class RecordVideo(gym.Wrapper):
    def __init__(self, env, *a, **k):
        super().__init__(env)
