"""Timing probe: how much of k_sim's in-step duration is a cold start (instruction fetch, descriptors)?
debug flag 2 makes cbev_step launch k_sim twice back to back; the difference of the profiled sim time is the
duration of an instruction-warm launch.  Not a parity configuration (state advances twice per step)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from carlabev_env_b200 import engine as E
from carlabev_env_b200.pool import pack_pool
from carlabev_env_b200.vector_env import load_town01_map
if __name__ == "__main__":
    N = 4096
    scenes = bench.build_pool(1024)
    eng = E.Engine(N, action_mode=E.ACTION_CONTINUOUS, max_actors=4, autoreset=E.AUTORESET_NEXT_STEP)
    eng.upload_map(load_town01_map()); eng.upload_pool(pack_pool(scenes))
    eng.reset(torch.arange(N, dtype=torch.int32) % len(scenes))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.rand(N, 3, device="cuda", generator=g) * torch.tensor([1, 2, 1], device="cuda") - torch.tensor([0, 1, 0], device="cuda") for _ in range(64)]
    for flags in (0, 2, 0, 2):
        eng.set_debug_flags(flags)
        for i in range(30): eng.step(acts[i % 64])
        eng.profile(True)
        for i in range(200): eng.step(acts[i % 64])
        torch.cuda.synchronize()
        sim_ms, render_ms, n = eng.profile_read()
        eng.profile(False)
        print("flags=%d: sim %.2f us  render %.2f us  (%d steps)" % (flags, sim_ms / n * 1e3, render_ms / n * 1e3, n))
