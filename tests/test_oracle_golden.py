"""Pins the CPU oracle (oracle/) to vectors produced by the reference's own Numba functions
(oracle/make_golden.py).  Integer/index results must be identical; fp64 results within 1e-12
relative to the array's max magnitude (the C restatement uses libm sqrt/pow where Numba used
BLAS dnrm2 / LLVM pow)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as orc

KERNEL_CASES = ["2d_uniform", "2d_ragged", "3d_uniform", "3d_ragged", "tiny"]


def close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return np.abs(a - b).max(initial=0.0) <= rel * scale


@pytest.mark.parametrize("case", KERNEL_CASES)
def test_assign_bins_matches_reference(case):
    g = load_golden("kernels_" + case)
    bins, bins_help, bin_locations, max_agents = orc.assign_bins(g["locations"], g["size"], 15, 1)
    assert max_agents == int(g["max_agents15"])
    assert np.array_equal(bin_locations, g["bin_locations15"])
    assert np.array_equal(bins_help, g["bins_help15"])
    assert np.array_equal(bins, g["bins15"])


@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("radius", [15, 10])
def test_neighbor_edges_bit_exact(case, radius):
    g = load_golden("kernels_" + case)
    edges, max_agents, max_neighbors = orc.neighbor_edges(g["locations"], g["size"], radius)
    assert np.array_equal(orc.canonical_edges(edges), g[f"edges{radius}"])
    assert max_agents == int(g[f"max_agents{radius}"])
    assert max_neighbors == int(g[f"max_neighbors{radius}"])


@pytest.mark.parametrize("case", KERNEL_CASES)
def test_jkr_forces_and_bond_breaks(case):
    g = load_golden("kernels_" + case)
    forces = np.zeros_like(g["locations"])
    delete = orc.jkr_forces(g["jkr_edges_in"], g["locations"], g["radii"], forces)
    assert np.array_equal(g["jkr_edges_in"][~delete], g["jkr_edges_out"])
    assert close(forces, g["jkr_forces"])


@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("dt", [180, 0])
def test_integrate(case, dt):
    g = load_golden("kernels_" + case)
    loc = g["locations"].copy()
    orc.integrate(g["jkr_forces"], g["motility"], loc, g["radii"], g["size"], dt)
    assert np.array_equal(loc, g[f"integrated_dt{dt}"])


@pytest.mark.parametrize("case", KERNEL_CASES)
def test_apply_forces_eleven_substeps(case):
    g = load_golden("kernels_" + case)
    sim = orc.OracleSim(g["locations"], g["radii"], g["size"])
    sim.motility_forces = g["motility"].copy()
    sim.apply_forces()
    assert np.array_equal(sim.jkr_graph.edge_array(), g["apply_jkr_edges"])
    assert close(sim.locations, g["apply_locations"])
    assert not sim.motility_forces.any() and not sim.jkr_forces.any()
    sim.motility_forces = -g["motility"]
    sim.apply_forces()
    assert np.array_equal(sim.jkr_graph.edge_array(), g["apply2_jkr_edges"])
    assert close(sim.locations, g["apply2_locations"])


@pytest.mark.parametrize("case", ["small", "odd"])
def test_diffusion_matches_reference(case):
    g = load_golden("diffusion_" + case)
    steps, dt, last_dt, D, h2 = g["jit_args"]
    out = orc.diffuse_padded(g["base_in"].copy(), int(steps), dt, last_dt, D, h2)
    assert np.array_equal(out, g["base_out"])          # same operation order, no FMA -> bit exact
    step_dt, diffuse_dt, D, h, cmax, deg = g["method_args"]
    sim = orc.OracleSim(g["locations"], 5.0, g["size"], step_dt=step_dt)
    sim.diffuse_dt, sim.diffuse_const, sim.spat_res, sim.spat_res2 = diffuse_dt, D, h, h * h
    sim.max_concentration, sim.degradation = cmax, deg
    sim.fgf4_values = g["field_in"].copy()
    sim.update_diffusion("fgf4_values")
    assert np.array_equal(sim.fgf4_values, g["field_out"])
    # mass: diffusion with reflecting walls conserves the clamped total, degradation scales it
    clamped = np.clip(g["field_in"], 0, cmax).sum()
    assert abs(sim.fgf4_values.sum() - clamped * (1 - deg)) <= 1e-9 * clamped
    sampled = sim.update_concentrations("fgf4_values", g["amounts"])
    assert np.array_equal(sampled, g["sampled"])
    assert close(sim.fgf4_values, g["field_deposited"], rel=1e-15)
    # per-cell entry points agree with the batched one
    assert sim.get_concentration("fgf4_values", 7) == orc.get_concentrations(
        g["locations"], sim.fgf4_values, h)[7]


def test_numpy_floor_divide_restatement():
    rng = np.random.default_rng(0)
    for d in (10.0, 15.0, 0.1, 7.07106):
        k = rng.integers(-3, 500, 20000).astype(float)
        x = np.concatenate([k * d, np.nextafter(k * d, np.inf), np.nextafter(k * d, -np.inf),
                            rng.random(20000) * 7000 - 5])
        loc = np.zeros((len(x), 3))
        loc[:, 0] = x
        out = np.zeros((len(x), 3), dtype=np.int64)
        orc.lib().orc_bin_locations(len(x), loc, d, out)
        assert np.array_equal(out[:, 0], np.floor_divide(x, d).astype(int) + 1)


def test_host_graph_facade():
    g = orc.HostGraph(6)
    g.add_edges(np.array([[3, 1], [1, 3], [0, 5], [2, 2], [4, 5]]))
    g.simplify()
    assert g.get_edgelist() == [(0, 5), (1, 3), (4, 5)]
    assert g.neighbors(5) == [0, 4] and g.num_neighbors(1) == 1 and g.neighbors(2) == []
    g.add_vertex()
    assert g.vcount() == 7 and g.neighbors(6) == []
    g.delete_vertices([0, 2])                      # renumbers: 1->0, 3->1, 4->2, 5->3, 6->4
    assert g.get_edgelist() == [(0, 1), (2, 3)] and g.vcount() == 5
    g.delete_edges(np.array([0]))
    assert g.get_edgelist() == [(2, 3)]
    g.delete_edges(None)
    assert g.ecount() == 0
