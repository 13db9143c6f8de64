"""Throughput of the public VectorEnv surface (make_env / reset / step) at 4096 envs."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from carlabev_env_b200 import EnvConfig, RunConfig, make_env
if __name__ == "__main__":
    N = 4096
    scenes = bench.build_pool(1024)
    host_infos = "--device-infos" not in sys.argv
    envs = make_env(RunConfig(env=EnvConfig(action_mode="continuous"), num_envs=N), scenes=scenes, autoreset="next_step",
                    host_infos=host_infos)
    obs, _ = envs.reset(options={"scene": "pool"})
    g = torch.Generator().manual_seed(0)
    acts = [torch.rand(N, 3, generator=g).cuda() for _ in range(16)]
    for a in acts: a[:, 1] = a[:, 1] * 2 - 1
    for i in range(20): envs.step(acts[i % 16])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = 300
    for i in range(K):
        obs, rew, term, trunc, infos = envs.step(acts[i % 16])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("VectorEnv.step: %.1f us/step -> %.3e env-steps/s; episodes finished in last step: %d" % (dt / K * 1e6, N * K / dt, int(term.sum())))
