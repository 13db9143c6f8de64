"""Frame dump of env 0 -- the role `gym.wrappers.RecordVideo` plays in the reference's wrap_env
(envs/__init__.py:40-60, make_carlabev_env :93-100: only env 0 captures).

Same configuration (RunConfig.capture_video / capture_every / video_output_dir / video_episode_indices /
video_name_prefix, `exp_name` for the default directory), same episode trigger, same frames: the 128x128 RGB field of
view of `render()` after the reset and after every step of a selected episode.  No video encoder is assumed on the
box: an episode is written as `<prefix>-episode-<id>.npy`, uint8 [T, S, S, 3] (one `np.save`), which any encoder can
consume.  The engine keeps the palette-index field of view of every env while a recorder is attached
(cbev_keep_fov) -- 16 KB per env-step of extra stores, so leave capture off for throughput runs.
"""
from __future__ import annotations

import os

import numpy as np

# RGB of the CBEV_PAL_* palette indices (include/cbev.h; CarlaBEV/semantics.py:20-40, traffic_light.py:44-54)
PALETTE = np.array([(150, 150, 150), (255, 255, 255), (220, 220, 220), (0, 7, 175), (255, 0, 0), (0, 255, 0),
                    (255, 64, 64), (255, 255, 0), (0, 0, 0), (100, 100, 100)], dtype=np.uint8)


def build_episode_trigger(episode_indices=None, every=50):
    """envs/__init__.py:25-37."""
    if episode_indices:
        selected = {int(v) for v in episode_indices}
        return lambda episode_id: episode_id in selected
    return lambda episode_id: episode_id % every == 0


class FrameRecorder:
    def __init__(self, cfg, eval=False):  # noqa: A002
        base = getattr(cfg, "video_output_dir", None)
        if base is None:
            base = f"videos/{getattr(cfg, 'exp_name', 'carlabev-run')}"
            base = f"{base}/eval" if eval else base
        self.dir = base
        self.prefix = getattr(cfg, "video_name_prefix", "rl-video")
        self.trigger = build_episode_trigger(getattr(cfg, "video_episode_indices", None),
                                             50 if eval else getattr(cfg, "capture_every", 50))
        self.episode_id = -1
        self.frames = None
        self.written = []

    def on_reset(self, index_frame):
        """Env 0 was reset: close the running recording, start the next episode if the trigger selects it."""
        self.flush()
        self.episode_id += 1
        self.frames = [PALETTE[index_frame]] if self.trigger(self.episode_id) else None

    def on_step(self, index_frame, done):
        if self.frames is not None:
            self.frames.append(PALETTE[index_frame])
            if done:
                self.flush()

    def flush(self):
        if self.frames:
            os.makedirs(self.dir, exist_ok=True)
            path = os.path.join(self.dir, f"{self.prefix}-episode-{self.episode_id}.npy")
            np.save(path, np.stack(self.frames))
            self.written.append(path)
        self.frames = None
