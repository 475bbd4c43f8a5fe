"""Times the DEVICE-RESIDENT model step (rules.DeviceColony: search r=15 -> division -> death -> pathway ->
differentiate -> motility -> 11 mechanics substeps; reference cell_simulation.py:77-126 without outputs and without the
morphogen stage) on a bench workload.  The reference runs the rule methods as CPython loops (about 33 us per cell and
step); here they are kernels and nothing visits the host inside a step.
usage: bench_rules.py [workload=c3] [steps=20]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from hipsc_abm_b200.rules import DeviceColony, INT_FIELDS, default_rule_params

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wl = bench.make_workload(name)
n, size = wl["n"], wl["size"]
P = default_rule_params(size, seed=3, dox_step=2)
rng = np.random.default_rng(3)
draw = lambda hi: rng.integers(0, hi, n).astype(np.int32)
zero = np.zeros(n, dtype=np.int32)
# a colony in the shape of cell_simulation.py:128-157, counters spread over their ranges so that every step sees
# divisions (about 1/36 of the pluripotent cells per step) as in a running simulation
state = dict(FGF4=draw(P.field), FGFR=draw(P.field), ERK=draw(P.field), GATA6=zero.copy(), NANOG=draw(P.field),
             states=zero.copy(), death_counters=draw(P.death_thresh), diff_counters=draw(P.pluri_to_diff),
             div_counters=draw(P.pluri_div_thresh), fds_counters=draw(P.fds_thresh), gid=np.arange(n, dtype=np.int32))
assert set(state) == set(INT_FIELDS)
three_d = size[2] != 0
dev = DeviceColony(size, capacity=int(2.2 * n), params=P, bond_stride=32 if three_d else 16,
                   search_cols_per_cell=30.0 if three_d else 14.0)
dev.load(wl["loc"], 5.0, state)
for k in range(1, 4):                                   # warm-up (also lets the random start relax)
    dev.step(k)
torch.cuda.synchronize()
dev.check()
n_start = int(dev.sets[dev.cur]["n"].item())
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for k, (a, b) in enumerate(evs):
    a.record()
    dev.step(4 + k)
    b.record()
torch.cuda.synchronize()
dev.check()
ms = [a.elapsed_time(b) for a, b in evs]
n_end = int(dev.sets[dev.cur]["n"].item())
h = dev.host_state()
print(json.dumps({"workload": name, "cells_start": n_start, "cells_end": n_end, "steps": steps,
                  "ms_per_step_mean": float(np.mean(ms)), "ms_per_step_first": ms[0], "ms_per_step_last": ms[-1],
                  "cell_updates_per_s": float(0.5 * (n_start + n_end) * steps / (sum(ms) / 1e3)),
                  "launches_per_step": dev.launches_per_step, "differentiated": int(h["states"].sum()),
                  "bonds": int(len(h["bond_ids"])), "host_bytes_inside_steps": 0,
                  "reference_rule_methods_estimate_ms_per_step": 33e-3 * 0.5 * (n_start + n_end),
                  "note": "eager launches; whole model step on the device, no morphogen stage"}))
