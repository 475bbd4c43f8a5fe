"""The device-count slab path (slab_device.py): several ranks emulated on one GPU -- every count on the device, no
host synchronisation inside a step, phases optionally replayed as CUDA graphs -- must reproduce the single-domain run:
bonds identical by global id, positions bit-identical (integer force sums), morphogen field within fp64 round-off."""
import numpy as np
import pytest
import torch

from hipsc_abm_b200 import slab
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
from hipsc_abm_b200.slab_device import DeviceSlab, SlabStepper, loopback_step
from oracle import oracle as orc

from test_gpu_slab import bonds_as_gid_edges

pytestmark = pytest.mark.gpu


def relaxed(loc, size, steps=3):
    """A uniformly random start has a few cell pairs almost on top of each other; they fly apart by tens of um in the
    first substeps, more than any reasonable halo depth allows for (DeviceSlab.allowed_substep_move).  Let the colony
    settle for a few single-domain steps first."""
    eng = ColonyEngine(size)
    eng.upload(loc, np.full(len(loc), 5.0), np.zeros_like(loc))
    one = HotPathStep(eng, StepConfig(bond_stride=32, search_cols_per_cell=40.0))
    for _ in range(steps):
        one.step()
    torch.cuda.synchronize()
    one.check()
    return eng.locations_host().copy()


@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "graphs"])
@pytest.mark.parametrize("world,dim", [(3, 2), (2, 3)])
def test_emulated_device_slabs_match_single_domain(world, dim, graphs):
    rng = np.random.default_rng(4)
    if dim == 2:
        size, n = np.array([1500.0, 400.0, 0.0]), 6000
        h, sd = 7.07106, 4
        grid = (int(np.ceil(size[0] / h)) + 1, int(np.ceil(size[1] / h)) + 2)
    else:
        size, n = np.array([700.0, 120.0, 120.0]), 9000
        h, sd, grid = 1.0, 0, None
    loc = relaxed(rng.random((n, 3)) * size, size)
    mot = rng.normal(0, 1.5e-9, (n, 3)) * (np.array([1, 1, 0]) if dim == 2 else 1)
    radii = np.full(n, 5.0)
    steps = 6
    cfg = StepConfig(grid=grid, spat_res=h, diffuse_dt=0.2 * h * h / 50.0, diffuse_steps=sd, bond_stride=32,
                     search_cols_per_cell=40.0, overlap_morphogen=True)
    dev = torch.device("cuda")

    # single domain (morphogen stage on the step's starting positions, like the slab path)
    eng = ColonyEngine(size)
    eng.upload(loc, radii, mot)
    one = HotPathStep(eng, cfg)
    for _ in range(steps):
        one.step()
    torch.cuda.synchronize()
    one.check()
    ref_loc = eng.locations_host()
    b = one.current_bonds()
    ref_edges = bonds_as_gid_edges(torch.arange(n), b.row_len[:n], b.cols[:n])
    ref_nbr = eng.edge_list(one.neighbor_graph)

    # slabs
    halo = slab.required_halo(len(cfg.substep_dts()), eng.jkr_reach(10.0), slack=5.0, spat_res=h if grid else 0.0)
    geoms = [slab.SlabGeometry.uniform(size, world, r, halo) for r in range(world)]
    steppers = []
    send_cap = int(1.2 * n / world) + 512              # the same on every rank (the receiver posts that size)
    for g in geoms:
        m = (loc[:, 0] >= g.x_lo) & ((loc[:, 0] < g.x_hi) | (g.rank == world - 1))
        field = slab.FieldSlab(g, grid, h, sd + 1, dev) if grid else None
        s = DeviceSlab(g, cfg, dev, rows=int(2.2 * m.sum()) + 512, send_cap=send_cap, uniform_radius=5.0, field=field)
        s.load(loc[m], radii[m], mot[m], np.arange(n)[m])
        steppers.append(SlabStepper(s, side_stream=False))
    done = 0
    if graphs:
        for _ in range(2):
            loopback_step(steppers)
        done = 2
        for st in steppers:
            st.capture()
    for _ in range(steps - done):
        loopback_step(steppers)
    torch.cuda.synchronize()
    for st in steppers:
        st.slab.check()
    owned = [st.slab.owned() for st in steppers]
    gid = torch.cat([o["gid"] for o in owned]).cpu().numpy()
    got_loc = torch.cat([o["loc"] for o in owned]).cpu().numpy()[np.argsort(gid)]
    assert np.array_equal(np.sort(gid), np.arange(n))
    edges = orc.canonical_edges(np.concatenate(
        [bonds_as_gid_edges(o["gid"], o["bond_len"], o["bond_gid"]) for o in owned]))
    assert np.array_equal(edges, ref_edges)
    assert np.array_equal(got_loc, ref_loc)              # fixed-point force sums: bit-identical trajectories
    # cells really migrated between ranks during the run (otherwise the test would not exercise the merge)
    start_rank = np.searchsorted(geoms[0].bounds[1:-1], loc[:, 0], side="right")
    end_rank = np.concatenate([np.full(o["n"], k) for k, o in enumerate(owned)])[np.argsort(gid)]
    assert (start_rank != end_rank).sum() > 0
    if grid:
        got = torch.cat([st.slab.field.owned() for st in steppers]).cpu().numpy()
        assert got.shape == grid
        assert np.abs(got - one.field.cpu().numpy()).max() <= 1e-11
        total = sum(st.slab.sampled[:o["n"]].sum().item() for st, o in zip(steppers, owned))
        # the sampled values belong to the step's STARTING owners; compare the colony-wide sum
        assert abs(total - one.sampled.sum().item()) <= 1e-6 * max(1.0, abs(one.sampled.sum().item()))


def test_capacity_overflows_are_reported_not_silent():
    rng = np.random.default_rng(1)
    size, n = np.array([800.0, 300.0, 0.0]), 3000
    loc = rng.random((n, 3)) * size
    cfg = StepConfig(bond_stride=32, search_cols_per_cell=40.0)
    dev = torch.device("cuda")
    halo = slab.required_halo(11, 10.3100107, slack=5.0)
    geoms = [slab.SlabGeometry.uniform(size, 2, r, halo) for r in range(2)]
    steppers = []
    for g in geoms:
        m = (loc[:, 0] >= g.x_lo) & ((loc[:, 0] < g.x_hi) | (g.rank == 1))
        s = DeviceSlab(g, cfg, dev, rows=int(m.sum()) + 2000, send_cap=16, uniform_radius=5.0)   # messages far too small
        s.load(loc[m], np.full(m.sum(), 5.0), np.zeros((m.sum(), 3)), np.arange(n)[m])
        steppers.append(SlabStepper(s, side_stream=False))
    loopback_step(steppers)
    torch.cuda.synchronize()
    with pytest.raises(Exception, match="message capacity too small"):
        steppers[0].slab.check()


def test_halo_depth_assumption_is_verified():
    """slab.required_halo assumes no ghost-zone cell moves more than `slack` per substep; the mechanics kernel records
    the largest such displacement and check() compares."""
    rng = np.random.default_rng(2)
    size, n = np.array([900.0, 300.0, 0.0]), 3500
    loc = relaxed(rng.random((n, 3)) * size, size)
    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    cfg = StepConfig(bond_stride=32, search_cols_per_cell=40.0)
    dev = torch.device("cuda")
    # two cells almost on top of each other deep inside rank 0's OWNED zone: they fly apart by far more than the
    # halo allows for, but only ghost-zone cells matter (the move window), so the check must not trip over them
    loc = np.concatenate([loc, [[200.0, 150.0, 0.0], [200.02, 150.0, 0.0]]])
    mot = np.concatenate([mot, np.zeros((2, 3))])
    n += 2
    eng = ColonyEngine(size)
    eng.upload(loc, np.full(n, 5.0), mot)
    one = HotPathStep(eng, cfg)
    one.step()
    torch.cuda.synchronize()
    assert eng.max_substep_move() > 8.0                     # no window: every cell counts
    for slack, ok in ((6.0, True), (0.05, False)):          # 2 nN moves a cell ~0.4 um per substep, contacts more
        halo = slab.required_halo(11, 10.3100107, slack=slack)
        steppers = []
        for r in range(2):
            g = slab.SlabGeometry.uniform(size, 2, r, halo)
            m = (loc[:, 0] >= g.x_lo) & ((loc[:, 0] < g.x_hi) | (r == 1))
            s = DeviceSlab(g, cfg, dev, rows=2 * n, send_cap=n, uniform_radius=5.0)
            s.load(loc[m], np.full(m.sum(), 5.0), mot[m], np.arange(n)[m])
            steppers.append(SlabStepper(s, side_stream=False))
        for _ in range(3):
            loopback_step(steppers)
        torch.cuda.synchronize()
        assert steppers[0].slab.eng.max_substep_move() > 0.1
        if ok:
            steppers[0].slab.check()
        else:
            with pytest.raises(Exception, match="halo too shallow"):
                steppers[0].slab.check()
