"""Prints how far the fused loop is from the oracle on the mixed-radii colony of tests/test_gpu_edges.py (debug aid for
kernel variants: ABM_B200_LIB=<so> python scripts/dbg_trim.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hipsc_abm_b200  # noqa
from test_gpu_edges import Cells
from oracle import oracle as orc
rng = np.random.default_rng(2)
n = 4000
size = np.array([500.0, 500.0, 0.0])
loc = rng.random((n, 3)) * size
radii = rng.uniform(5 / 2 ** 0.5, 5.0, n)
mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
sim = Cells(loc, radii, size)
ref = orc.OracleSim(loc, radii, size)
for step in range(2):
    sim.motility_forces[:] = mot; ref.motility_forces[:] = mot
    sim.apply_forces(); ref.apply_forces()
    d = np.abs(sim.locations - ref.locations)
    eq = np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array())
    print("step", step, "max|dx|", d.max(), "cells off by >1e-9:", int((d.max(axis=1) > 1e-9).sum()), "nan:", int(np.isnan(sim.locations).sum()), "bonds equal", eq)
