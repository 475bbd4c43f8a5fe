"""Runs a few eager (un-graphed) device-resident steps of a bench workload; the command ncu wraps.
usage: profile_step.py <workload> <steps>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, StepConfig

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = bench.make_workload(name)
spec = wl["spec"]
cfg = StepConfig(grid=wl["grid"], spat_res=spec["h"] or 1.0, diffuse_dt=spec["diffuse_dt"] or 1.0,
                 diffuse_steps=spec["sd"], bond_stride=16 if spec["dim"] == 2 else 32,
                 search_cols_per_cell=12.0 if spec["dim"] == 2 else 24.0,
                 overlap_morphogen=len(sys.argv) > 3 and sys.argv[3] == "overlap")
eng = ColonyEngine(wl["size"])
eng.upload(wl["loc"], wl["radii"], wl["mot"])
pipe = HotPathStep(eng, cfg)
for _ in range(steps):
    pipe.step()
torch.cuda.synchronize()
print("ok", name, steps, pipe.check(), "launches/step", pipe.launches_per_step)
