"""Sensitivity of a retreating StopReturn pedestrian IN THE ORACLE ALONE (CPU): the smoothed retreat route of one of
two identical oracle envs is perturbed by 1e-13 px / 1e-12 rad when the retreat starts; prints the heading difference
over the following steps (grows ~2.4x per step to O(1) rad in the scenes whose retreat route starts with a reversal).
Evidence for the RETREAT NOTE of tests/test_gpu_scale.py and DESIGN.md section 2.   python tools/retreat_chaos.py"""
import sys, copy
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from golden_util import load_map
from carlabev_env_b200.scenes import build_scripted_scene
from oracle.env import OracleEnv
from oracle import sim as OS
cls=load_map()
worst=[]
for seed in range(7000,7064,2):
    lvl=1+seed%4
    if lvl<3: continue
    sc=build_scripted_scene("jaywalk", seed, level=lvl, cls_map=cls)
    a=OracleEnv(cls, action_mode="continuous"); b=OracleEnv(cls, action_mode="continuous")
    a.reset(sc); b.reset(sc)
    pert=False; maxd=0; hist=[]
    for t in range(160):
        act=np.array([0.0,0.0,1.0],np.float32)
        a.step(act); b.step(act)
        pa=[x for x in a.sim.actors if x.kind==1][0]; pb=[x for x in b.sim.actors if x.kind==1][0]
        if not pert and pa.fsm==OS.ST_RETREATING:
            rng=np.random.default_rng(seed); pb.cy = np.asarray(pb.cy, dtype=float) + rng.normal(0,1e-13,len(pb.cy)); pb.cx = np.asarray(pb.cx, dtype=float) + rng.normal(0,1e-13,len(pb.cx)); pb.cyaw = np.asarray(pb.cyaw, dtype=float) + rng.normal(0,1e-12,len(pb.cyaw))
            pert=True; t0=t
        if pert:
            d=abs(pa.yaw-pb.yaw); hist.append(d)
    if pert:
        worst.append((max(hist), seed, lvl, t0, [f"{h:.1e}" for h in hist[:60:3]], [round(float(v),3) for v in np.asarray(pa.cyaw)]))
worst.sort(reverse=True)
for w in worst[:8]: print(w)
