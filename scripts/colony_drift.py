"""Why bench.py alternates the sign of the motility forces: bonds and neighbours per cell of the synthetic colony over
many steps, with the CPU oracle (no GPU needed; ~1 min for 300 steps of 20 000 cells on 8 cores).
   python scripts/colony_drift.py            # constant force: crowding, > 16 bonds per cell after ~190 steps
   python scripts/colony_drift.py flip       # alternating sign: stationary (3.9 bonds, 5.7 neighbours per cell)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from oracle import oracle as orc

n = 20000
wl = bench.make_workload("c3", seed=0, n_override=n)
orc.set_threads(orc.max_threads())
sim = orc.OracleSim(wl["loc"], wl["radii"], wl["size"])
sim.multithread = True
flip = len(sys.argv) > 1 and sys.argv[1] == "flip"
t0 = time.time()
for s in range(1, 301):
    sim.motility_forces[:] = wl["mot"] * (-1 if (flip and s % 2 == 0) else 1)
    sim.apply_forces()
    if s % 20 == 0 or s < 4:
        deg = sim.jkr_graph.degrees()
        sim.get_neighbors("neighbor_graph", 15)
        d15 = sim.neighbor_graph.degrees()
        at_wall = int(((sim.locations[:, :2] <= 0) | (sim.locations[:, :2] >= wl["size"][:2])).any(axis=1).sum())
        print(f"step {s}: bonds per cell max {deg.max()} mean {deg.mean():.2f} | r=15 neighbours max {d15.max()} "
              f"mean {d15.mean():.2f} | cells on a wall {at_wall} | {time.time() - t0:.0f} s", flush=True)
