import os, sys
os.environ.setdefault("CUDA_LAUNCH_BLOCKING", "1")
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from golden_util import Golden
from engine_util import make_engine_for
g = Golden("lead_brake_continuous")
eng = make_engine_for(g)
print("engine ok; crop", eng.cfg.ring_slots, flush=True)
ids = torch.tensor([0], dtype=torch.int32)
try:
    eng.reset(ids)
    torch.cuda.synchronize()
    print("reset ok", flush=True)
    fr = eng.fov()[0].cpu().numpy()
    print("fov", np.unique(fr, return_counts=True), (fr == g["reset_frames"][0]).mean())
    a = torch.tensor(np.asarray(g["actions"][0], dtype=np.float32)[None], device=eng.device)
    eng.step(a); torch.cuda.synchronize(); print("step ok", eng.reward, eng.hero[0, :4])
except Exception as ex:
    print("FAILED:", ex)
