"""Per-CTA phase timeline of k_render_any (debug flag 4): python tools/render_trace_any.py <size> [envs] [obs_h obs_w]"""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from carlabev_env_b200 import engine as E
from carlabev_env_b200.pool import pack_pool
from carlabev_env_b200.scenes import build_pool
from carlabev_env_b200.vector_env import load_town01_map
from carlabev_env_b200.config import ACTION_PROFILES
if __name__ == "__main__":
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    obs = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (96, 96)
    pad = {64: 91, 128: 182, 256: 363}[size]
    scenes = [s for s in build_pool([dict(scene="rdm", num_vehicles=12, route_dist_range=(30, 100), scene_seed=i) for i in range(96)],
                                    pad=pad, size=size, skip_invalid=True) if s is not None]
    eng = E.Engine(N, action_mode=E.ACTION_DISCRETE, discrete_table=ACTION_PROFILES["discrete9_v1"]["discrete_actions"], max_actors=16, autoreset=E.AUTORESET_NEXT_STEP, ring_slots=16, size=size, obs_size=obs)
    eng.upload_map(load_town01_map(size)); eng.upload_pool(pack_pool(scenes))
    eng.reset(torch.arange(N, dtype=torch.int32) % len(scenes))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.randint(0, 9, (N,), device="cuda", generator=g) for _ in range(16)]
    eng.set_debug_flags(4 | (256 if size == 128 else 0))
    for i in range(12): eng.step(acts[i % 16])
    torch.cuda.synchronize()
    eng.step(acts[3]); torch.cuda.synchronize()
    tr = eng.read_trace().astype(np.int64)
    t = tr[:, :6]; t0 = t[:, 0].min()
    ph = np.diff(t, axis=1) / 1e3
    life = (t[:, 5] - t[:, 0]) / 1e3
    names = ["tma wait", "draw list", "rotate", "resize", "stores"]
    print(f"size {size}, obs {obs}, {N} envs: span {(t[:, 5].max() - t0) / 1e3:.1f} us; CTA life mean {life.mean():.1f} p10 {np.percentile(life, 10):.1f} p90 {np.percentile(life, 90):.1f} us")
    print("   phase means (us): " + ", ".join(f"{n} {ph[:, i].mean():.2f} (p90 {np.percentile(ph[:, i], 90):.2f})" for i, n in enumerate(names)))
    wl = tr[:, 7]
    print(f"   work lists (table resize): flagged 8x8 blocks {(wl & 0xffffffff).mean():.0f} of {(size // 8) ** 2}, mixed outputs {(wl >> 32).mean():.0f} of 9216")
