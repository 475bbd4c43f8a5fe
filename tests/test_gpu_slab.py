"""Slab decomposition with the real kernels: several slabs emulated on one GPU (loopback exchange) must
reproduce the single-domain device run -- bonds identical by global id, positions bit-identical (integer force
sums), morphogen field within fp64 round-off (its deposits are fp64 atomics)."""
import numpy as np
import pytest
import torch

from hipsc_abm_b200 import slab
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, SlabPipeline, StepConfig
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def bonds_as_gid_edges(gid, row_len, cols):
    gid, row_len, cols = gid.cpu().numpy(), row_len.cpu().numpy().astype(np.int64), cols.cpu().numpy()
    k = np.arange(cols.shape[1])[None, :] < row_len[:, None]
    return orc.canonical_edges(np.stack([np.repeat(gid, row_len), cols[k]], axis=1))


@pytest.mark.parametrize("first", [False, True], ids=["morphogen_after", "morphogen_first"])
@pytest.mark.parametrize("native", [False, True], ids=["torch_ops", "native_ops"])
@pytest.mark.parametrize("world,dim", [(3, 2), (2, 3)])
def test_emulated_slabs_match_single_domain(world, dim, native, first):
    if first and dim == 3:
        pytest.skip("no morphogen field in the 3-D case")
    rng = np.random.default_rng(4)
    if dim == 2:
        size, n = np.array([1500.0, 400.0, 0.0]), 6000
        h, sd = 7.07106, 4
        grid = (int(np.ceil(size[0] / h)) + 1, int(np.ceil(size[1] / h)) + 2)
    else:
        size, n = np.array([700.0, 120.0, 120.0]), 9000
        h, sd, grid = 1.0, 0, None
    loc = rng.random((n, 3)) * size
    mot = rng.normal(0, 1.5e-9, (n, 3)) * (np.array([1, 1, 0]) if dim == 2 else 1)
    radii = np.full(n, 5.0)
    steps = 2
    cfg = StepConfig(grid=grid, spat_res=h, diffuse_dt=0.2 * h * h / 50.0, diffuse_steps=sd, bond_stride=32,
                     search_cols_per_cell=40.0, overlap_morphogen=first)
    dev = torch.device("cuda")

    # single domain
    eng = ColonyEngine(size)
    eng.upload(loc, radii, mot)
    one = HotPathStep(eng, cfg)
    for _ in range(steps):
        one.step()
    one.check()
    ref_loc = eng.locations_host()
    b = one.current_bonds()
    ref_edges = bonds_as_gid_edges(torch.arange(n), b.row_len[:n], b.cols[:n])

    # slabs
    halo = slab.required_halo(len(cfg.substep_dts()), eng.jkr_reach(10.0), slack=5.0) + (h if grid else 0.0)
    geoms = [slab.SlabGeometry.uniform(size, world, r, halo) for r in range(world)]
    assert halo < size[0] / world
    states, pipes, fields = [], [], []
    for g in geoms:
        m = (loc[:, 0] >= g.x_lo) & ((loc[:, 0] < g.x_hi) | (g.rank == world - 1))
        t = lambda a: torch.from_numpy(a[m].copy()).to(dev)
        states.append(slab.SlabState(t(loc), t(radii), t(mot), torch.from_numpy(np.arange(n)[m]).to(dev), 32))
        pipes.append(SlabPipeline(g, cfg, dev, uniform_radius=5.0 if native else 0.0, side_stream=False))
        fields.append(slab.FieldSlab(g, grid, h, sd + 1, dev) if grid else None)
    from hipsc_abm_b200.pipeline import NativeSlabOps
    ops = [NativeSlabOps(p.eng) for p in pipes] if native else None
    for _ in range(steps):
        slab.loopback_step(states, geoms, [p.compute for p in pipes], fields if grid else None,
                           pipes if grid else None, ops, morphogen_first=first)
    for p in pipes:
        p.check()
    for o in ops or []:
        o.check()
    gid = torch.cat([s.gid for s in states]).cpu().numpy()
    got_loc = torch.cat([s.loc for s in states]).cpu().numpy()[np.argsort(gid)]
    assert np.array_equal(np.sort(gid), np.arange(n))
    edges = orc.canonical_edges(np.concatenate(
        [bonds_as_gid_edges(s.gid, s.bond_len, s.bond_gid) for s in states]))
    assert np.array_equal(edges, ref_edges)
    # fixed-point force sums: the trajectory does not depend on the local order of the cells, so the slabs give the
    # single-domain positions bit for bit
    assert np.array_equal(got_loc, ref_loc)
    if grid:
        got = torch.cat([f.owned() for f in fields]).cpu().numpy()
        assert got.shape == grid
        assert np.abs(got - one.field.cpu().numpy()).max() <= 1e-11
        assert abs(sum(p.sampled.sum().item() for p in pipes) - one.sampled.sum().item()) <= 1e-8
