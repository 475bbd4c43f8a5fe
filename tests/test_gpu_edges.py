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
        # the gate is per step (north star: "per-step forces and positions"): the next step starts from the oracle's
        # positions, otherwise a dense random colony amplifies the first step's last-bit differences past it
        sim.locations = ref.locations.copy()
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
@pytest.mark.parametrize("steps", [1, 2, 4, 5, 7, 8, 9])
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


@pytest.mark.parametrize("steps", [2, 3, 6, 10, 11])
def test_fused_diffusion_substeps_equal_one_launch_per_substep(steps):
    """abm_diffusion advances up to five substeps per launch in shared memory (k_diffuse_fused): the field is
    bit-identical to the substep-by-substep chain (max_fused_substeps = 1), lands in the same buffer, and
    takes the announced number of launches."""
    rng = np.random.default_rng(steps)
    field = rng.random((203, 390)) * 3 - 0.5               # ragged in both directions, negative and >max values
    h, D = 7.0, 50.0
    dt = 0.23 * h * h / D
    eng = ColonyEngine([1, 1, 0])
    out = {}
    for fuse in (1, 5, 3):
        old, eng.diffusion_fuse = eng.diffusion_fuse, fuse          # a per-engine argument of abm_diffusion
        try:
            assert eng.lib.abm_diffusion_launches(steps, fuse) % 2 == steps % 2
            if fuse == 1:
                assert eng.lib.abm_diffusion_launches(steps, fuse) == steps
            else:
                assert eng.lib.abm_diffusion_launches(steps, fuse) <= (steps + fuse - 1) // fuse + 1
            for tma in (False, True):
                src = torch.from_numpy(field.copy()).cuda()
                scratch = torch.full_like(src, float("nan"))
                res = eng.diffuse(src, steps, dt, dt * 0.4, D, h * h, 2.0, 0.1, use_tma=tma, scratch=scratch)
                assert (res.data_ptr() == scratch.data_ptr()) == bool(steps % 2)
                out[fuse, tma] = res.cpu().numpy()
        finally:
            eng.diffusion_fuse = old
    ref = orc.diffuse_field(field, steps, dt, dt * 0.4, D, h * h, 2.0, 0.1)
    for key, got in out.items():
        assert np.array_equal(got, ref), key


# ---- second-session kernels: warp tile, slot-order hand-over, multiply-high decode, slab plumbing ------------------
def test_candidate_runs_longer_than_the_warp_tile():
    # a dense strip: the stencil runs of a warp hold more than the 64 candidates the shared-memory tile takes, so
    # those cells go through the slow path (exact walk from global memory).  The strip is crowded (cells overlap
    # by several um), which makes the explicit integration amplify last-bit differences: pair sets must be
    # identical, positions are compared after one step at 1e-9
    rng = np.random.default_rng(3)
    n = 700
    size = [500.0, 200.0, 0.0]
    loc = np.stack([rng.uniform(100.0, 400.0, n), rng.uniform(100.0, 112.0, n), np.zeros(n)], axis=1)
    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    sim = Cells(loc, 5.0, size)
    ref = orc.OracleSim(loc, 5.0, size)
    sim.motility_forces[:] = mot
    ref.motility_forces[:] = mot
    sim.get_neighbors("neighbor_graph", 15)
    ref.get_neighbors("neighbor_graph", 15)
    assert np.array_equal(sim.neighbor_graph.edge_array(), orc.canonical_edges(ref.neighbor_graph.edge_array()))
    assert max(sim.neighbor_graph.num_neighbors(i) for i in range(n)) > 64
    sim.apply_forces()
    ref.apply_forces()
    assert np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array())
    assert np.abs(sim.locations - ref.locations).max() <= 1e-9 * np.abs(ref.locations).max()


def test_sparse_colony_on_a_huge_lattice():
    # 60000 um box: 5822^2 bins at the mechanics edge, bin ids beyond what the multiply-high decode covers
    # (falls back to the division), almost all bins empty, warps straddling many bin rows
    rng = np.random.default_rng(4)
    n = 600
    centres = rng.uniform(1000.0, 59000.0, (30, 2))
    loc = np.zeros((n, 3))
    loc[:, :2] = centres[rng.integers(0, 30, n)] + rng.normal(0, 14.0, (n, 2))
    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    same_as_oracle(loc, 5.0, [60000.0, 60000.0, 0.0], mot, steps=2)


def test_mixed_radii_3d_in_fused_loop():
    rng = np.random.default_rng(5)
    n = 5000
    size = np.array([170.0, 170.0, 170.0])
    loc = rng.random((n, 3)) * size
    radii = rng.uniform(5 / 2 ** 0.5, 5.0, n)
    mot = rng.normal(0, 2e-9, (n, 3))
    same_as_oracle(loc, radii, size, mot, steps=2)


def test_mechanics_loop_equals_substep_by_substep():
    # abm_mechanics_loop (cells handed from substep to sort in slot order, histogram sums accumulated by the kernel,
    # loose binning, locations written once) must give the bits of the plain sequence: fresh sort from
    # `locations` + one abm_jkr_substep, 11 times
    from hipsc_abm_b200 import _lib
    from hipsc_abm_b200.engine import GraphTensors
    rng = np.random.default_rng(6)
    for dim, n, size in ((2, 6000, [700.0, 700.0, 0.0]), (3, 6000, [180.0, 180.0, 180.0])):
        size = np.array(size)
        loc = rng.random((n, 3)) * size
        mot = rng.normal(0, 2e-9, (n, 3)) * ([1, 1, 0] if dim == 2 else [1, 1, 1])
        radii = np.full(n, 5.0)
        dts = [180.0] * 10 + [0.0]
        stride = 32
        a = ColonyEngine(size)
        a.upload(loc, radii, mot)
        bonds_a, _ = a.mechanics(None, 10.0, dts, stride)

        b = ColonyEngine(size)
        b.upload(loc, radii, mot)
        reach = b.jkr_reach(10.0)
        rows = [torch.zeros((n, stride), dtype=torch.int32, device="cuda") for _ in range(2)]
        lens = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(2)]
        worst = torch.zeros(2, dtype=torch.int32, device="cuda")
        old = None
        for k, dt in enumerate(dts):
            srt = b.sort(reach)
            _lib.check(b.lib.abm_jkr_substep(
                srt.bin_start.data_ptr(), srt.sorted_id.data_ptr(), srt.sorted_pos.data_ptr(),
                srt.sorted_bin.data_ptr(), n, _lib.i3(srt.dims), 10.0, reach, 0, stride if old is not None else 0,
                old[0].data_ptr() if old is not None else 0, old[1].data_ptr() if old is not None else 0,
                lens[k & 1].data_ptr(), rows[k & 1].data_ptr(), stride, worst.data_ptr(), int(k == len(dts) - 1),
                b.mot.data_ptr(), b.loc.data_ptr(), 0.5, 1000.0, 0.000107, 10000.0, _lib.d3(size), dt, 0,
                b.uniform_radius, None, None, 0, 0, 0, None, b.stream), "abm_jkr_substep")
            old = (lens[k & 1], rows[k & 1])
        assert int(worst[0].item()) <= stride
        assert torch.equal(a.loc, b.loc), f"{dim}-D positions differ"
        assert torch.equal(bonds_a.row_len[:n], old[0])
        m = torch.arange(stride, device="cuda")[None, :] < old[0][:, None]
        assert torch.equal(bonds_a.cols[:n][m], old[1][m])


def test_slab_message_overflow_is_reported():
    from hipsc_abm_b200 import _lib, slab
    from hipsc_abm_b200.pipeline import NativeSlabOps
    rng = np.random.default_rng(7)
    n = 4000
    size = np.array([800.0, 200.0, 0.0])
    loc = rng.random((n, 3)) * size
    dev = torch.device("cuda")
    geom = slab.SlabGeometry.uniform(size, 2, 0, 150.0)
    own = loc[:, 0] < geom.x_hi
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    state = slab.SlabState(t(loc[own]), t(np.full(own.sum(), 5.0)), t(np.zeros((own.sum(), 3))),
                           t(np.arange(n)[own]), 16)
    eng = ColonyEngine(size, dev)
    ops = NativeSlabOps(eng, send_cap=8)                       # far fewer rows than cells near the interface
    out_l, out_r, counts = ops.outgoing(state, geom)
    assert int(counts[1].item()) == int((loc[own, 0] >= geom.x_hi - geom.halo).sum()) > 8
    with pytest.raises(_lib.AbmError, match="capacity too small"):
        ops.check()
    ok = NativeSlabOps(eng, send_cap=4096)
    out_l, out_r, counts = ok.outgoing(state, geom)
    ok.check()
    sent = slab.unpack(out_r[:int(counts[1].item())], 16)["gid"].cpu().numpy()
    assert np.array_equal(np.sort(sent), np.arange(n)[own][loc[own, 0] >= geom.x_hi - geom.halo])
