"""Event-timed 3-D diffusion substeps (k_diffuse3d) on a cubic grid, TMA and plain loads, L2 flushed before each run.
usage: time_diffuse3d.py [g=226] [substeps=10]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hipsc_abm_b200.engine import ColonyEngine

g = int(sys.argv[1]) if len(sys.argv) > 1 else 226
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
h, D = 7.07106, 50.0
dt = 0.15 * h * h / D
eng = ColonyEngine([1.0, 1.0, 1.0])
field = torch.rand((g, g, g), dtype=torch.float64, device="cuda") * 2
scratch = torch.empty_like(field)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"grid": [g, g, g], "substeps": steps, "algorithmic_bytes_per_substep": 16 * g ** 3}
for tma in (True, False):
    ts = []
    for _ in range(12):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.diffuse3d(field, steps, dt, dt, D, h * h, 2.0, 0.1, use_tma=tma, scratch=scratch)
        z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z) * 1e3 / steps)
    us = float(np.median(ts[2:]))
    out["tma" if tma else "plain"] = {"us_per_substep": round(us, 2), "GBps_algorithmic": round(16 * g ** 3 / us / 1e3, 1)}
print(json.dumps(out))
