"""Parity at the scale bench.py measures (BASELINE.json configs C3, C4, C5): the CUDA path against the CPU oracle on
the bench's own workload recipe, not only on colonies of a few ten thousand cells.

  C3  500,000 cells 2-D   r=15 edges identical, JKR bonds identical, positions 1e-12 -- through the reference-facing
                          API (host arrays) AND through the device-resident pipeline bench.py times (HotPathStep,
                          eager and as the captured CUDA graph), field + sampled concentrations included
  C4  3-D                 200,000 cells: edges / bonds / positions vs the oracle; 4,000,000 cells (the stated size):
                          per-cell neighbour counts of the r=15 search against an independent KD-tree
  C5  2048 x 2048 grid    10 diffusion substeps bit-identical to the oracle (TMA and plain staging)
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch
import hipsc_abm_b200  # noqa: F401
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
from oracle import oracle as orc

from conftest import ROOT
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from test_gpu_parity import Cells, close  # noqa: E402


def oracle_for(wl, threads=None):
    orc.set_threads(threads or bench.host_cores())
    ref = orc.OracleSim(wl["loc"], wl["radii"], wl["size"])
    ref.multithread = False              # serial force pass (the parallel one is racy by construction)
    return ref


def test_c3_apply_forces_500k_vs_oracle():
    wl = bench.make_workload("c3", seed=3)
    sim = Cells(wl["loc"], wl["radii"], wl["size"])
    ref = oracle_for(wl)
    # The random start holds cell pairs almost on top of each other whose first separation is ill-conditioned (see
    # bench.parity_check): the first two steps are compared at 1e-10, the relaxed colony at the 1e-12 gate.
    for step in range(4):                                  # bonds carried over, alternating motility
        mot = wl["mot"] if step % 2 == 0 else -wl["mot"]
        sim.get_neighbors("neighbor_graph", 15)
        ref.get_neighbors("neighbor_graph", 15)
        assert np.array_equal(sim.neighbor_graph.edge_array(), orc.canonical_edges(ref.neighbor_graph.edge_array()))
        sim.motility_forces[:] = mot
        ref.motility_forces[:] = mot
        sim.apply_forces()
        ref.apply_forces()
        assert np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array())
        close(sim.locations, ref.locations, rel=1e-10 if step < 2 else 1e-12)     # fp64 mode gate, per step
        sim.locations = ref.locations.copy()              # teacher forcing: the gate is per step (see bench.parity_check)
    assert sim.jkr_graph.ecount() > 400_000


@pytest.mark.parametrize("captured", [False, True], ids=["eager", "graph"])
def test_c3_device_resident_step_is_what_the_oracle_computes(captured):
    """The pipeline bench.py's `value` times -- HotPathStep on device-resident state, morphogen stage on a second
    stream, replayed as a CUDA graph -- against the oracle running the same recipe (reference order: neighbour
    search, deposit/diffuse/sample on the step's starting positions, then the mechanics)."""
    wl = bench.make_workload("c3", seed=5)
    spec, n = wl["spec"], wl["n"]
    h, dt, sd = spec["h"], spec["diffuse_dt"], spec["sd"]
    cfg = StepConfig(grid=wl["grid"], spat_res=h, diffuse_dt=dt, diffuse_steps=sd, bond_stride=16,
                     search_cols_per_cell=12.0, overlap_morphogen=True)
    eng = ColonyEngine(wl["size"])
    eng.upload(wl["loc"], wl["radii"], wl["mot"])
    pipe = HotPathStep(eng, cfg, motility_by_parity=(eng.mot, -eng.mot))
    ref = oracle_for(wl)
    field = np.zeros(wl["grid"] + (1,))
    steps = 2
    if captured:
        pipe.capture()
    for step in range(steps):
        pipe.replay() if captured else pipe.step()
        torch.cuda.synchronize()
        # oracle: same order
        ref.get_neighbors("neighbor_graph", 15)
        orc.adjust_morphogens(ref.locations, 1.0, field, h)
        field[:, :, 0] = orc.diffuse_field(field[:, :, 0], sd, dt, dt, 50.0, h * h, 2.0, 0.1)
        sampled = orc.get_concentrations(ref.locations, field, h)
        ref.motility_forces[:] = wl["mot"] if step % 2 == 0 else -wl["mot"]
        ref.apply_forces()
        # compare
        got_edges = eng.edge_list(pipe.neighbor_graph)
        assert np.array_equal(got_edges, orc.canonical_edges(ref.neighbor_graph.edge_array())), f"step {step}"
        bonds = pipe.current_bonds()
        assert np.array_equal(eng.edge_list(bonds), ref.jkr_graph.edge_array()), f"step {step}"
        # free-running from the random start (the captured chain cannot be reset to the oracle): 1e-10, see above
        close(eng.loc.cpu().numpy(), ref.locations, rel=1e-10)
        assert np.abs(pipe.field.cpu().numpy() - field[:, :, 0]).max() <= 1e-6     # gate 1e-6 (deposit uses atomics)
        assert np.abs(pipe.sampled.cpu().numpy() - sampled).max() <= 1e-6
    pipe.check()


def test_c4_apply_forces_200k_3d_vs_oracle():
    wl = bench.make_workload("c4", seed=11, n_override=200_000)
    sim = Cells(wl["loc"], wl["radii"], wl["size"])
    ref = oracle_for(wl)
    sim.get_neighbors("neighbor_graph", 15)
    ref.get_neighbors("neighbor_graph", 15)
    assert np.array_equal(sim.neighbor_graph.edge_array(), orc.canonical_edges(ref.neighbor_graph.edge_array()))
    for step in range(3):
        mot = wl["mot"] if step % 2 == 0 else -wl["mot"]
        sim.motility_forces[:] = mot
        ref.motility_forces[:] = mot
        sim.apply_forces()
        ref.apply_forces()
        assert np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array())
        close(sim.locations, ref.locations, rel=1e-10 if step < 2 else 1e-12)
        sim.locations = ref.locations.copy()


def test_c4_search_4m_cells_3d_degrees_match_a_kdtree():
    """The stated C4 size.  The oracle's dense buckets and edge slots would need several GB here, so the independent
    check is a KD-tree: the number of cells within 15 um of EVERY cell must equal that cell's row length, and the
    rows must be symmetric."""
    from scipy.spatial import cKDTree
    wl = bench.make_workload("c4", seed=0)
    n, loc = wl["n"], wl["loc"]
    eng = ColonyEngine(wl["size"])
    eng.upload(loc, wl["radii"])
    g = eng.search(eng.sort(15.0), 15.0, capacity=int(16 * n))
    row_len = g.row_len.cpu().numpy().astype(np.int64)
    assert row_len.sum() == g.entries and g.entries % 2 == 0
    tree = cKDTree(loc)
    want = tree.query_ball_point(loc, 15.0, return_length=True, workers=-1).astype(np.int64) - 1
    assert np.array_equal(row_len, want)
    # symmetry on a sample of rows (the full edge list is 2 x 28 M entries)
    row_start = g.row_start.cpu().numpy().astype(np.int64)
    cols = g.cols
    rng = np.random.default_rng(0)
    for i in rng.integers(0, n, 2000):
        for j in cols[row_start[i]:row_start[i] + row_len[i]].cpu().numpy():
            row_j = cols[row_start[j]:row_start[j] + row_len[j]].cpu().numpy()
            assert i in row_j
            assert np.linalg.norm(loc[i] - loc[j]) <= 15.0


@pytest.mark.parametrize("tma", [False, True], ids=["plain", "tma"])
def test_c5_diffusion_2048_bit_exact_vs_oracle(tma):
    gx = gy = 2048
    rng = np.random.default_rng(5)
    field = rng.random((gx, gy)) * 2.4 - 0.2          # values below 0 and above max_concentration get clamped
    h, D = 1.5441, 50.0
    dt = 0.24 * h * h / D
    ref = orc.diffuse_field(field, 10, dt, dt, D, h * h, 2.0, 0.1)
    eng = ColonyEngine([1, 1, 0])
    out = eng.diffuse(torch.from_numpy(field.copy()).cuda(), 10, dt, dt, D, h * h, 2.0, 0.1, use_tma=tma)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_c5_step_100k_cells_on_the_fine_grid_vs_oracle():
    wl = bench.make_workload("c5", seed=2)
    spec = wl["spec"]
    h, dt, sd = spec["h"], spec["diffuse_dt"], spec["sd"]
    sim = Cells(wl["loc"], wl["radii"], wl["size"])
    ref = oracle_for(wl)
    for o in (sim, ref):
        o.spat_res, o.spat_res2, o.diffuse_const, o.diffuse_dt = h, h * h, 50.0, dt
        o.max_concentration, o.degradation = 2.0, 0.1
        o.fgf4_values = np.zeros(wl["grid"] + (1,))
        o.step_dt = dt * (sd - 1) + 0.5 * dt          # divmod -> sd substeps, the last with dt/2
    for _ in range(2):
        sim.update_concentrations("fgf4_values", 1.0, sample=False)
        orc.adjust_morphogens(ref.locations, 1.0, ref.fgf4_values, h)
        sim.update_diffusion("fgf4_values")
        ref.update_diffusion("fgf4_values")
        got = sim.update_concentrations("fgf4_values", None)
        want = orc.get_concentrations(ref.locations, ref.fgf4_values, h)
        assert np.abs(sim.fgf4_values - ref.fgf4_values).max() <= 1e-12
        assert np.abs(got - want).max() <= 1e-12
