"""Per-CTA phase timeline of the raw-RGB k_render (debug flag 4) on the configs[4] shape: python tools/render_trace_rgb.py [envs] [extra flags]"""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from carlabev_env_b200 import engine as E
from carlabev_env_b200.pool import load_shipped_pool, pack_pool
from carlabev_env_b200.vector_env import load_town01_map
if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    extra = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
    scenes = load_shipped_pool("rdm_dense_50")
    eng = E.Engine(N, obs_mode=E.OBS_RGB, action_mode=E.ACTION_CONTINUOUS, max_actors=50, anchor=(0.5, 0.75), autoreset=E.AUTORESET_NEXT_STEP)
    eng.upload_map(load_town01_map()); eng.upload_pool(pack_pool(scenes))
    eng.reset(torch.arange(N, dtype=torch.int32) % len(scenes))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.rand(N, 3, device="cuda", generator=g) * torch.tensor([1, 2, 1], device="cuda") - torch.tensor([0, 1, 0], device="cuda") for _ in range(16)]
    eng.set_debug_flags(4 | extra)
    for i in range(20): eng.step(acts[i % 16])
    torch.cuda.synchronize()
    eng.step(acts[3]); torch.cuda.synchronize()
    tr = eng.read_trace().astype(np.int64)
    t = tr[:, :4]; t0 = t[:, 0].min()
    ph = np.diff(t, axis=1) / 1e3
    life = (t[:, 3] - t[:, 0]) / 1e3
    names = ["tma wait", "draw list", "rotate + stores"]
    print(f"raw RGB, {N} envs, {len(scenes)} scenes: span {(t[:, 3].max() - t0) / 1e3:.1f} us; CTA life mean {life.mean():.1f} p10 {np.percentile(life, 10):.1f} p90 {np.percentile(life, 90):.1f} us; concurrent CTAs {N * life.mean() / ((t[:, 3].max() - t0) / 1e3):.0f}")
    print("   phase means (us): " + ", ".join(f"{n} {ph[:, i].mean():.2f} (p90 {np.percentile(ph[:, i], 90):.2f})" for i, n in enumerate(names)))
    print(f"   draw list, warp 0 only (before the barrier): {((tr[:, 7] - tr[:, 1]) / 1e3).mean():.2f} us")
