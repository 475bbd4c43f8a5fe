"""Edge cases of the CUDA path: empty and tiny colonies, cells outside the box, coincident cells, one
overfull bin, capacity growth, tiny / ragged morphogen grids."""
import numpy as np
import pytest
import torch

from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.simulation import ColonySimulation
from hipsc_abm_b200.cell_methods import B200CellMethods
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


class Cells(B200CellMethods, ColonySimulation):
    def __init__(self, locations, radii, size):
        ColonySimulation.__init__(self, size)
        self.add_agents(len(locations))
        self.agent_array("locations", override=np.array(locations, dtype=float).reshape(-1, 3))
        self.agent_array("radii", override=np.array(np.broadcast_to(radii, (len(locations),)), dtype=float))
        self.agent_array("motility_forces", vector=3)
        self.agent_array("jkr_forces", vector=3)
        self.agent_graph("neighbor_graph")
        self.agent_graph("jkr_graph")
        self.max_radius, self.step_dt, self.move_dt = 5.0, 1800, 180


def same_as_oracle(loc, radii, size, mot=None, steps=1):
    sim = Cells(loc, radii, size)
    ref = orc.OracleSim(np.array(loc, dtype=float).reshape(-1, 3), radii, size)
    for _ in range(steps):
        if mot is not None:
            sim.motility_forces[:] = mot
            ref.motility_forces[:] = mot
        sim.get_neighbors("neighbor_graph", 15)
        ref.get_neighbors("neighbor_graph", 15)
        assert np.array_equal(sim.neighbor_graph.edge_array(), orc.canonical_edges(ref.neighbor_graph.edge_array()))
        sim.apply_forces()
        ref.apply_forces()
        assert np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array())
        scale = max(np.abs(ref.locations).max(initial=0.0), 1.0)
        assert np.abs(sim.locations - ref.locations).max(initial=0.0) <= 1e-12 * scale
    return sim


def test_empty_colony():
    sim = Cells(np.zeros((0, 3)), 5.0, [100.0, 100.0, 0.0])
    sim.get_neighbors("neighbor_graph", 15)
    sim.apply_forces()
    sim.calculate_jkr()
    assert sim.neighbor_graph.ecount() == 0 and sim.jkr_graph.ecount() == 0 and sim.locations.shape == (0, 3)
    bins, bins_help, bin_locations, max_agents = sim.assign_bins(1, 15)
    assert bins_help.sum() == 0 and bin_locations.shape == (0, 3) and max_agents == 1


def test_one_and_two_cells():
    same_as_oracle([[50.0, 50.0, 0.0]], 5.0, [100.0, 100.0, 0.0])
    same_as_oracle([[50.0, 50.0, 0.0], [58.0, 50.0, 0.0]], 5.0, [100.0, 100.0, 0.0], steps=3)        # overlapping pair
    same_as_oracle([[50.0, 50.0, 0.0], [60.2, 50.0, 0.0]], 5.0, [100.0, 100.0, 0.0], steps=2)        # just out of contact


def test_coincident_cells_and_box_walls():
    # identical positions: mag == 0 -> zero normal but the bond exists (cell_backend.py:103-105);
    # cells pushed against / beyond the walls are clamped by the integration
    loc = [[10.0, 10.0, 0.0], [10.0, 10.0, 0.0], [0.0, 0.0, 0.0], [-1.2, 5.0, 0.0], [101.3, 99.0, 0.0], [100.0, 100.0, 0.0]]
    mot = np.array([[1e-9, 0, 0], [0, -1e-9, 0], [-3e-9, -3e-9, 0], [-1e-9, 0, 0], [2e-9, 2e-9, 0], [5e-9, 5e-9, 0]])
    sim = same_as_oracle(loc, 5.0, [100.0, 100.0, 0.0], mot, steps=2)
    assert sim.locations.min() >= 0.0 and sim.locations[:, :2].max() <= 100.0


def test_one_overfull_bin_rows_longer_than_every_buffer():
    # 150 cells inside a 3 um ball in 3-D: one bin holds them all, every row has 149 entries (> the 32-entry
    # row buffer of the search, > 16 parked partners and > the default bond stride of the fused substep)
    rng = np.random.default_rng(0)
    loc = 40.0 + rng.normal(0, 1.0, (150, 3))
    sim = Cells(loc, 5.0, [80.0, 80.0, 80.0])
    sim.get_neighbors("neighbor_graph", 15)
    assert sim.neighbor_graph.ecount() == 150 * 149 // 2 and sim.neighbor_graph.num_neighbors(7) == 149
    same_as_oracle(loc, 5.0, [80.0, 80.0, 80.0])


def test_search_capacity_grows_and_retries():
    rng = np.random.default_rng(1)
    loc = rng.random((3000, 3)) * [300.0, 300.0, 0.0]
    eng = ColonyEngine([300.0, 300.0, 0.0])
    eng.upload(loc, np.full(3000, 5.0))
    g = eng.search(eng.sort(15.0), 15.0, capacity=16)              # far too small: must grow, not truncate
    ref = orc.canonical_edges(orc.neighbor_edges(loc, [300.0, 300.0, 0.0], 15.0)[0])
    assert g.entries == 2 * len(ref)
    assert np.array_equal(eng.edge_list(g), ref)


def test_mixed_radii_in_fused_loop():
    rng = np.random.default_rng(2)
    n = 4000
    size = np.array([500.0, 500.0, 0.0])
    loc = rng.random((n, 3)) * size
    radii = rng.uniform(5 / 2 ** 0.5, 5.0, n)                      # the uniform-radius fast path must not trigger
    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    same_as_oracle(loc, radii, size, mot, steps=2)


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (5, 1), (2, 2), (33, 129), (32, 128), (31, 260)])
@pytest.mark.parametrize("steps", [1, 4])
def test_diffusion_small_and_ragged_grids(shape, steps):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    field = rng.random(shape) * 3 - 0.5
    h, D = 7.0, 50.0
    dt = 0.2 * h * h / D
    ref = orc.diffuse_field(field, steps, dt, dt * 0.4, D, h * h, 2.0, 0.1)
    eng = ColonyEngine([1, 1, 0])
    for tma in ([False, True] if shape[1] % 2 == 0 else [False]):
        out = eng.diffuse(torch.from_numpy(field.copy()).cuda(), steps, dt, dt * 0.4, D, h * h, 2.0, 0.1, use_tma=tma)
        assert np.array_equal(out.cpu().numpy(), ref), f"tma={tma}"
