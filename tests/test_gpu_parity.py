"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes), against the CPU oracle and the golden
vectors generated from the reference.  Gates (BASELINE.json north star):
  edge sets           bit-exact after (min,max) canonical sort
  forces / positions  |delta| <= 1e-12 * max|ref|  (fp64 mode)
  concentrations      bit-exact here (the gate is 1e-6)
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

import torch
import hipsc_abm_b200  # noqa: F401
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.simulation import ColonySimulation
from hipsc_abm_b200.cell_methods import B200CellMethods
from oracle import oracle as orc

KERNEL_CASES = ["2d_uniform", "2d_ragged", "3d_uniform", "3d_ragged", "tiny"]


def close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(initial=0.0), 1e-300)
    err = np.abs(a - b).max(initial=0.0)
    assert err <= rel * scale, f"max abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


class Cells(B200CellMethods, ColonySimulation):
    """What a reference subclass looks like to the hot path: numpy agent arrays + two graphs."""

    def __init__(self, locations, radii, size):
        ColonySimulation.__init__(self, size)
        self.add_agents(len(locations))
        self.agent_array("locations", override=np.array(locations, dtype=float))
        self.agent_array("radii", override=np.array(np.broadcast_to(radii, (len(locations),)), dtype=float))
        self.agent_array("motility_forces", vector=3)
        self.agent_array("jkr_forces", vector=3)
        self.agent_graph("neighbor_graph")
        self.agent_graph("jkr_graph")
        self.max_radius = 5.0
        self.step_dt, self.move_dt = 1800, 180


def colony(seed, n, dim, density=None):
    rng = np.random.default_rng(seed)
    if dim == 2:
        side = (n / (density or 0.01)) ** 0.5
        size = np.array([side, side, 0.0])
    else:
        side = (n / (density or 0.001)) ** (1 / 3)
        size = np.array([side, side, side])
    return rng.random((n, 3)) * size, np.full(n, 5.0), size


# ---- a1/a2 ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KERNEL_CASES)
def test_assign_bins_matches_reference(case):
    g = load_golden("kernels_" + case)
    sim = Cells(g["locations"], g["radii"], g["size"])
    bins, bins_help, bin_locations, max_agents = sim.assign_bins(1, 15)
    assert max_agents == int(g["max_agents15"])
    assert np.array_equal(bin_locations, g["bin_locations15"])
    assert np.array_equal(bins_help, g["bins_help15"])
    assert np.array_equal(bins, g["bins15"])


def test_sort_is_a_stable_permutation():
    loc, radii, size = colony(3, 20000, 2)
    eng = ColonyEngine(size)
    eng.upload(loc, radii)
    from hipsc_abm_b200 import _lib
    s = eng.sort(10.0, _lib.SORT_STABLE)
    n = len(loc)
    sid = s.sorted_id[:n].cpu().numpy()
    assert np.array_equal(np.sort(sid), np.arange(n))
    start = s.bin_start.cpu().numpy()
    assert start[0] == 0 and start[-1] == len(loc) and np.all(np.diff(start) >= 0)
    sbin = s.sorted_bin[:n].cpu().numpy()
    assert np.all(np.diff(sbin) >= 0)
    same = sbin[1:] == sbin[:-1]
    assert np.all(sid[1:][same] > sid[:-1][same])          # ascending id inside a bin
    pos = s.sorted_pos[:n].cpu().numpy()
    assert np.array_equal(pos[:, :3], loc[sid]) and np.array_equal(pos[:, 3], radii[sid])
    assert eng.read_i32(s.max_count) == np.diff(start).max()
    # the fast placement is the same permutation up to the order inside a bin
    f = eng.sort(10.0)
    fid = f.sorted_id[:n].cpu().numpy()
    assert np.array_equal(np.sort(fid), np.arange(n))
    assert np.array_equal(f.bin_start.cpu().numpy(), start)
    assert np.array_equal(f.sorted_bin[:n].cpu().numpy(), sbin)
    assert np.array_equal(f.sorted_pos[:n].cpu().numpy()[:, :3], loc[fid])


# ---- a3/a4/a5 ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("radius", [15, 10])
def test_neighbor_edges_bit_exact_vs_reference(case, radius):
    g = load_golden("kernels_" + case)
    sim = Cells(g["locations"], g["radii"], g["size"])
    sim.get_neighbors("neighbor_graph", radius)
    assert np.array_equal(sim.neighbor_graph.edge_array(), g[f"edges{radius}"])
    assert "get_neighbors" in sim.method_times
    # the device-side edge list kernel gives the same canonical list
    eng = sim.b200_engine()
    assert np.array_equal(eng.edge_list(sim.neighbor_graph._dev), g[f"edges{radius}"])


@pytest.mark.parametrize("n,dim,radius", [(30000, 2, 15), (30000, 2, 10), (20000, 3, 15), (20000, 3, 10)])
def test_neighbor_edges_bit_exact_vs_oracle(n, dim, radius):
    loc, radii, size = colony(11 + dim, n, dim)
    sim = Cells(loc, radii, size)
    sim.get_neighbors("neighbor_graph", radius)
    ref, _, _ = orc.neighbor_edges(loc, size, radius)
    assert np.array_equal(sim.neighbor_graph.edge_array(), orc.canonical_edges(ref))


def test_graph_facade_on_device_result():
    loc, radii, size = colony(5, 4000, 2)
    sim = Cells(loc, radii, size)
    sim.get_neighbors("neighbor_graph", 15)
    ref = orc.HostGraph(len(loc))
    ref.add_edges(orc.neighbor_edges(loc, size, 15)[0])
    ref.simplify()
    g = sim.neighbor_graph
    for i in (0, 1, 17, 3999):
        assert g.neighbors(i) == ref.neighbors(i) and g.num_neighbors(i) == ref.num_neighbors(i)
    assert g.get_edgelist() == ref.get_edgelist()
    dead = np.array([3, 17, 200, 3998])
    g.delete_vertices(dead)
    ref.delete_vertices(dead)
    g.add_vertex()
    ref.add_vertex()
    assert g.vcount() == ref.vcount() and g.get_edgelist() == ref.get_edgelist()


def test_get_neighbors_union_keeps_old_edges():
    """clear=False: new contacts are merged into the persisting graph (simulation.py:284-288)."""
    loc, radii, size = colony(6, 5000, 2)
    sim = Cells(loc, radii, size)
    o = orc.OracleSim(loc, radii, size)
    for step in range(3):
        sim.get_neighbors("jkr_graph", 10, clear=False)
        o.get_neighbors("jkr_graph", 10, clear=False)
        assert np.array_equal(sim.jkr_graph.edge_array(), o.jkr_graph.edge_array())
        rng = np.random.default_rng(step)
        move = rng.normal(0, 1.0, loc.shape) * np.array([1, 1, 0])
        sim.locations = np.clip(sim.locations + move, 0, size)
        o.locations = sim.locations.copy()


# ---- a6/a7/a9 ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KERNEL_CASES)
def test_calculate_jkr_matches_reference(case):
    g = load_golden("kernels_" + case)
    sim = Cells(g["locations"], g["radii"], g["size"])
    sim.jkr_graph.add_edges(g["jkr_edges_in"])
    sim.calculate_jkr()
    assert np.array_equal(sim.jkr_graph.edge_array(), g["jkr_edges_out"])
    close(sim.jkr_forces, g["jkr_forces"])


@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("dt", [180, 0])
def test_integrate_bit_exact(case, dt):
    g = load_golden("kernels_" + case)
    eng = ColonyEngine(g["size"])
    eng.upload(g["locations"], g["radii"], g["motility"], g["jkr_forces"])
    eng.integrate(dt)
    assert np.array_equal(eng.locations_host(), g[f"integrated_dt{dt}"])
    assert not eng.jkr.cpu().numpy().any()


# ---- a8 ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("fused", [True, False])
def test_apply_forces_matches_reference(case, fused):
    g = load_golden("kernels_" + case)
    cls = Cells
    if not fused:
        class Unfused(Cells):                       # overriding a stage forces the stage-by-stage loop
            def calculate_jkr(self):
                return B200CellMethods.calculate_jkr(self)
        cls = Unfused
    sim = cls(g["locations"], g["radii"], g["size"])
    sim.motility_forces = g["motility"].copy()
    sim.apply_forces()
    assert np.array_equal(sim.jkr_graph.edge_array(), g["apply_jkr_edges"])
    close(sim.locations, g["apply_locations"])
    assert not sim.motility_forces.any() and not sim.jkr_forces.any()
    sim.motility_forces = -g["motility"]
    sim.apply_forces()
    assert np.array_equal(sim.jkr_graph.edge_array(), g["apply2_jkr_edges"])
    close(sim.locations, g["apply2_locations"])


@pytest.mark.parametrize("n,dim", [(40000, 2), (25000, 3)])
def test_apply_forces_vs_oracle_larger(n, dim):
    loc, radii, size = colony(21 + dim, n, dim, density=0.012 if dim == 2 else 0.0014)
    rng = np.random.default_rng(9)
    mot = rng.normal(0, 2e-9, loc.shape) * (np.array([1, 1, 0]) if dim == 2 else 1)
    sim = Cells(loc, radii, size)
    o = orc.OracleSim(loc, radii, size)
    sim.motility_forces = mot.copy()
    o.motility_forces = mot.copy()
    sim.apply_forces()
    o.apply_forces()
    assert np.array_equal(sim.jkr_graph.edge_array(), o.jkr_graph.edge_array())
    close(sim.locations, o.locations)


def test_fused_row_overflow_grows_and_retries():
    rng = np.random.default_rng(2)
    n = 600
    size = np.array([60.0, 60.0, 0.0])
    loc = np.zeros((n, 3))
    loc[:, :2] = 30 + rng.normal(0, 3.0, (n, 2))        # a clump: far more than 16 bonds per cell
    sim = Cells(loc, 5.0, size)
    o = orc.OracleSim(loc, 5.0, size)
    sim.jkr_graph.row_stride_hint = 4
    sim.apply_forces()
    o.apply_forces()
    assert sim.jkr_graph.row_stride_hint > 4
    assert np.array_equal(sim.jkr_graph.edge_array(), o.jkr_graph.edge_array())
    close(sim.locations, o.locations, rel=1e-11)


# ---- a10-a13 ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["small", "odd"])
@pytest.mark.parametrize("tma", [False, True], ids=["plain", "tma"])
def test_diffusion_bit_exact_vs_reference(case, tma):
    g = load_golden("diffusion_" + case)
    step_dt, diffuse_dt, D, h, cmax, deg = g["method_args"]
    gx, gy = g["field_in"].shape[:2]
    if tma and gy % 2:
        pytest.skip("TMA path needs an even row length")
    eng = ColonyEngine(g["size"])
    steps, last_dt = divmod(step_dt, diffuse_dt)
    plane = torch.from_numpy(g["field_in"][:, :, 0].copy()).cuda()
    out = eng.diffuse(plane, int(steps) + 1, diffuse_dt, last_dt, D, h * h, cmax, deg, use_tma=tma)
    assert np.array_equal(out.cpu().numpy(), g["field_out"][:, :, 0])


@pytest.mark.parametrize("gx,gy,steps", [(318, 318, 10), (500, 258, 3), (97, 1030, 5)])
@pytest.mark.parametrize("tma", [False, True], ids=["plain", "tma"])
def test_diffusion_bit_exact_vs_oracle(gx, gy, steps, tma):
    rng = np.random.default_rng(gx)
    field = rng.random((gx, gy)) * 3 - 0.5
    h, D = 7.07106, 50.0
    dt = 0.24 * h * h / D
    ref = orc.diffuse_field(field, steps, dt, dt * 0.3, D, h * h, 2.0, 0.1)
    eng = ColonyEngine([1, 1, 0])
    out = eng.diffuse(torch.from_numpy(field.copy()).cuda(), steps, dt, dt * 0.3, D, h * h, 2.0, 0.1, use_tma=tma)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_update_diffusion_method_and_concentrations():
    g = load_golden("diffusion_small")
    step_dt, diffuse_dt, D, h, cmax, deg = g["method_args"]
    sim = Cells(g["locations"], 5.0, g["size"])
    sim.step_dt, sim.diffuse_dt, sim.diffuse_const, sim.spat_res, sim.spat_res2 = step_dt, diffuse_dt, D, h, h * h
    sim.max_concentration, sim.degradation = cmax, deg
    sim.fgf4_values = g["field_in"].copy()
    sim.update_diffusion("fgf4_values")
    assert np.array_equal(sim.fgf4_values, g["field_out"])
    sampled = sim.update_concentrations("fgf4_values", g["amounts"])
    assert np.array_equal(sampled, g["sampled"])
    close(sim.fgf4_values, g["field_deposited"], rel=1e-14)
    assert sim.get_concentration("fgf4_values", 5) == sim.fgf4_values[tuple(
        np.ceil(np.floor(2 * g["locations"][5] / h) / 2).astype(int))]
    before = sim.fgf4_values.sum()
    sim.adjust_morphogens("fgf4_values", 5, 2.5)
    assert abs(sim.fgf4_values.sum() - before - 2.5) < 1e-9


def test_concentrations_odd_grid_with_out_of_bounds_corners():
    g = load_golden("diffusion_odd")
    step_dt, diffuse_dt, D, h, cmax, deg = g["method_args"]
    sim = Cells(g["locations"], 5.0, g["size"])
    sim.spat_res = h
    sim.fgf4_values = g["field_out"].copy()
    sampled = sim.update_concentrations("fgf4_values", g["amounts"])
    assert np.array_equal(sampled, g["sampled"])
    close(sim.fgf4_values, g["field_deposited"], rel=1e-14)


# ---- full-size, size-independent properties (BASELINE configs C3 / C5) ---------------------------------------
def test_c3_scale_search_properties():
    n = 500_000
    loc, radii, size = colony(0, n, 2)
    eng = ColonyEngine(size)
    eng.upload(loc, radii)
    g = eng.search(eng.sort(15.0), 15.0)
    row_len = g.row_len.cpu().numpy().astype(np.int64)
    row_start = g.row_start.cpu().numpy().astype(np.int64)
    cols = g.cols.cpu().numpy()
    assert row_len.sum() == g.entries and g.entries % 2 == 0
    # rows tile [0, entries) exactly once (packed, no overlap)
    order = np.argsort(row_start, kind="stable")
    nz = order[row_len[order] > 0]
    assert row_start[nz][0] == 0 and np.array_equal(row_start[nz][1:], (row_start[nz] + row_len[nz])[:-1])
    # symmetric, ascending, within radius
    rows = np.repeat(np.arange(n), row_len)
    src = np.repeat(row_start - (np.cumsum(row_len) - row_len), row_len) + np.arange(row_len.sum())
    c = cols[src].astype(np.int64)
    d = np.linalg.norm(loc[rows] - loc[c], axis=1)
    assert d.max() <= 15.0
    key = rows * n + c
    assert np.all(np.diff(key) > 0)                       # ascending rows, no duplicates
    assert np.array_equal(np.sort(c * n + rows), key)     # j in row i <=> i in row j
    # count agrees with an independent KD-tree
    from scipy.spatial import cKDTree
    assert cKDTree(loc).count_neighbors(cKDTree(loc), 15.0) - n == g.entries


def test_c5_scale_diffusion_mass_conservation():
    gx = gy = 2048
    rng = np.random.default_rng(1)
    field = torch.from_numpy(rng.random((gx, gy)) * 2).cuda()
    total = field.sum().item()
    eng = ColonyEngine([1, 1, 0])
    h, D = 1.5441, 50.0
    dt = 0.24 * h * h / D
    a = eng.diffuse(field.clone(), 10, dt, dt, D, h * h, 2.0, 0.1, use_tma=True)
    b = eng.diffuse(field.clone(), 10, dt, dt, D, h * h, 2.0, 0.1, use_tma=False)
    assert torch.equal(a, b)
    assert abs(a.sum().item() - 0.9 * total) <= 1e-9 * total
    assert a.min().item() >= 0 and a.max().item() <= 2.0
