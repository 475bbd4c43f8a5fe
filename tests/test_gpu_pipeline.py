"""The device-resident step (pipeline.HotPathStep) against the CPU oracle: eager and as a replayed CUDA graph, with
the morphogen stage after the mechanics (one stream) and before them on a second stream (the reference's order)."""
import numpy as np
import pytest
import torch

from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

H, D, SD = 7.07106, 50.0, 4
DT = 0.2 * H * H / D


def oracle_steps(loc, radii, mot, size, grid, steps, first):
    sim = orc.OracleSim(loc, radii, size)
    field = np.zeros(grid + (1,))
    sampled = edges15 = None

    def morphogen():
        orc.adjust_morphogens(sim.locations, 1.0, field, H)
        field[:, :, 0] = orc.diffuse_field(field[:, :, 0], SD, DT, DT, D, H * H, 2.0, 0.1)
        return orc.get_concentrations(sim.locations, field, H)
    for _ in range(steps):
        sim.get_neighbors("neighbor_graph", 15)
        edges15 = orc.canonical_edges(sim.neighbor_graph.edge_array())
        if first:
            sampled = morphogen()
        sim.motility_forces = mot.copy()
        sim.apply_forces()
        if not first:
            sampled = morphogen()
    return sim, field[:, :, 0], sampled, edges15


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
@pytest.mark.parametrize("first", [False, True], ids=["morphogen_after", "morphogen_first_second_stream"])
def test_hot_path_step_matches_oracle(first, graph):
    rng = np.random.default_rng(21)
    n, size = 5000, np.array([700.0, 700.0, 0.0])
    loc = rng.random((n, 3)) * size
    mot = rng.normal(0, 1.5e-9, (n, 3)) * [1, 1, 0]
    radii = np.full(n, 5.0)
    grid = (int(np.ceil(size[0] / H)) + 1, int(np.ceil(size[1] / H)) + 1)
    grid = (grid[0], grid[1] + grid[1] % 2)
    steps = 4 if graph else 2
    ref, ref_field, ref_sampled, ref_edges15 = oracle_steps(loc, radii, mot, size, grid, steps, first)

    eng = ColonyEngine(size)
    eng.upload(loc, radii, mot)
    cfg = StepConfig(grid=grid, spat_res=H, diffuse_dt=DT, diffuse_steps=SD, bond_stride=32, search_cols_per_cell=40.0,
                     overlap_morphogen=first)
    pipe = HotPathStep(eng, cfg)
    if graph:
        pipe.step()
        pipe.step()
        pipe.check()
        pipe.capture()
        pipe.replay()
        pipe.replay()
    else:
        for _ in range(steps):
            pipe.step()
    torch.cuda.synchronize()
    pipe.check()

    got = eng.locations_host()
    assert np.abs(got - ref.locations).max() <= 1e-10 * np.abs(ref.locations).max()
    assert np.array_equal(eng.edge_list(pipe.current_bonds()), ref.jkr_graph.edge_array())
    assert np.array_equal(eng.edge_list(pipe.neighbor_graph), ref_edges15)
    assert np.abs(pipe.field.cpu().numpy() - ref_field).max() <= 1e-12        # deposits: fp64 atomics, other order
    assert np.abs(pipe.sampled.cpu().numpy() - ref_sampled).max() <= 1e-12


def _colony(n=4000, seed=5):
    rng = np.random.default_rng(seed)
    size = np.array([600.0, 600.0, 0.0])
    loc = rng.random((n, 3)) * size
    mot = rng.normal(0, 1.5e-9, (n, 3)) * [1, 1, 0]
    return size, loc, np.full(n, 5.0), mot


def test_handed_over_cells_equal_a_fresh_sort():
    """Back-to-back device-resident steps hand the moved cells to the next step's first sort (ABM_LOOP_CARRY_*): the
    positions and bonds are bit-identical to steps that re-bin `locations` from scratch."""
    size, loc, radii, mot = _colony()
    results = []
    for carry in (True, False):
        eng = ColonyEngine(size)
        eng.upload(loc, radii, mot)
        pipe = HotPathStep(eng, StepConfig(bond_stride=32, search_cols_per_cell=40.0))
        for k in range(5):
            if not carry:
                eng._carry = None                       # forget the hand-over: the next loop starts with a histogram
            pipe.step()
            assert (eng._carry is not None)
        torch.cuda.synchronize()
        pipe.check()
        results.append((eng.locations_host(), eng.edge_list(pipe.current_bonds()), eng.launches))
    assert results[0][2] == results[1][2] - 2 * 4          # steps 2..5 skip the histogram and one scan pass
    assert np.array_equal(results[0][0], results[1][0])
    assert np.array_equal(results[0][1], results[1][1])


def test_replay_refuses_a_broken_hand_over():
    size, loc, radii, mot = _colony(n=1500)
    eng = ColonyEngine(size)
    eng.upload(loc, radii, mot)
    pipe = HotPathStep(eng, StepConfig(bond_stride=32, search_cols_per_cell=40.0))
    pipe.step()
    pipe.step()
    pipe.capture()
    pipe.replay()
    eng.integrate(0.0)                                 # anything that may move the cells drops the hand-over
    from hipsc_abm_b200._lib import AbmError
    with pytest.raises(AbmError):
        pipe.replay()
