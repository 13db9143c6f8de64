"""A pool of worker processes, each owning a slice of `OracleEnv`s, so that the scale-parity GPU tests step their
64-env oracle sample on all host cores instead of one (the oracle costs ~10 ms per env-step).  TEST INFRASTRUCTURE.

Protocol per call: the parent sends every worker a list of commands, one per env it owns --
("reset", scene_index) or ("step", action) -- and gets back, per env, a dict with the stacked observation, reward,
flags, ego pose and actor poses after the command."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(conn, scenes, oracle_kw, n_envs):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from golden_util import load_map
    from oracle.env import OracleEnv

    cls = load_map()
    envs = [OracleEnv(cls, **oracle_kw) for _ in range(n_envs)]
    while True:
        cmds = conn.recv()
        if cmds is None:
            break
        out = []
        for env, (op, arg) in zip(envs, cmds):
            if op == "reset":
                res = dict(obs=env.reset(scenes[int(arg)]), reward=0.0, term=False, trunc=False)
            else:
                o, r, te, tr, _ = env.step(arg)
                res = dict(obs=o, reward=float(r), term=bool(te), trunc=bool(tr))
            e = env.sim.ego
            res["ego"] = np.array([e.x, e.y, e.yaw, e.v])
            res["actors"] = np.array([[b.x, b.y, b.yaw, b.v] for b in env.sim.actors]).reshape(-1, 4)
            out.append(res)
        conn.send(out)
    conn.close()


class OraclePool:
    def __init__(self, n_envs, scenes, oracle_kw, workers=None, timeout=300.0):
        self.timeout = timeout
        workers = max(1, min(workers or (os.cpu_count() or 2) - 1, 16, n_envs))
        ctx = mp.get_context("spawn")  # the parent holds a CUDA context: never fork it
        self.slices = np.array_split(np.arange(n_envs), workers)
        self.conns, self.procs = [], []
        for sl in self.slices:
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, scenes, oracle_kw, len(sl)), daemon=True)
            p.start()
            child.close()
            self.conns.append(parent)
            self.procs.append(p)

    def run(self, cmds):
        """cmds[j] = ("reset", scene_index) | ("step", action) for oracle env j; returns the per-env result dicts."""
        for conn, sl in zip(self.conns, self.slices):
            conn.send([cmds[j] for j in sl])
        out = [None] * len(cmds)
        for conn, sl, proc in zip(self.conns, self.slices, self.procs):
            waited = 0.0
            while not conn.poll(1.0):  # never block for ever on a dead worker (a hung test would hang the GPU box)
                waited += 1.0
                if not proc.is_alive() or waited > self.timeout:
                    raise RuntimeError(f"oracle worker {proc.pid} died or timed out (exit code {proc.exitcode})")
            for j, res in zip(sl, conn.recv()):
                out[j] = res
        return out

    def close(self):
        for conn in self.conns:
            try:
                conn.send(None)
                conn.close()
            except Exception:  # noqa: BLE001
                pass
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
