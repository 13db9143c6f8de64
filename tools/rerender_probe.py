import sys, torch, numpy as np
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
    a = torch.zeros(N, 3, device="cuda"); a[:, 2] = 1.0
    for _ in range(10): eng.step(a)
    s = torch.cuda.current_stream().cuda_stream
    E._check(eng.lib, eng.lib.cbev_debug_rerender(eng.handle, 5, s)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E._check(eng.lib, eng.lib.cbev_debug_rerender(eng.handle, 50, s)); e1.record(); torch.cuda.synchronize()
    print("re-render only: %.1f us per launch" % (e0.elapsed_time(e1) / 50 * 1e3))
    for mode, name in ((1, "advance head"), (2, "spin kernel between"), (3, "advance head + spin kernel")):
        E._check(eng.lib, eng.lib.cbev_debug_rerender(eng.handle, 5 | (mode << 16), s)); torch.cuda.synchronize()
        e0.record(); E._check(eng.lib, eng.lib.cbev_debug_rerender(eng.handle, 50 | (mode << 16), s)); e1.record(); torch.cuda.synchronize()
        print("re-render, %s: %.1f us per iteration" % (name, e0.elapsed_time(e1) / 50 * 1e3))
    e0.record()
    for _ in range(50): eng.step(a)
    e1.record(); torch.cuda.synchronize()
    print("full step (brake): %.1f us" % (e0.elapsed_time(e1) / 50 * 1e3))
