"""3-D morphogen extension on the GPU (csrc/diffusion3d.cu) against its CPU restatement (oracle/abm_oracle.c
orc_update_diffusion3d / orc_adjust_morphogens3d): bit-identical fields for the TMA and the plain staging, ragged
grids, chunk / pencil boundaries, the last-substep dt, clamp and degradation; the method surface with
b200_diffusion_3d; and the reference's plane-0 semantics for volumes when the extension is off."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch
import hipsc_abm_b200  # noqa: F401
from hipsc_abm_b200.engine import ColonyEngine
from oracle import oracle as orc

from test_gpu_parity import Cells

H, D = 7.07106, 50.0


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 2), (5, 9, 64), (33, 17, 66), (70, 20, 130), (40, 8, 4),
                                   (64, 64, 64), (31, 9, 63), (97, 33, 6)])
@pytest.mark.parametrize("steps", [1, 2, 5])
def test_diffusion3d_bit_exact_vs_oracle(shape, steps):
    rng = np.random.default_rng(shape[0] * 1000 + shape[2] + steps)
    field = rng.random(shape) * 3 - 0.5                      # values below 0 and above the cap get clamped
    dt = 0.15 * H * H / D
    ref = orc.diffuse_field3d(field, steps, dt, 0.4 * dt, D, H * H, 2.0, 0.1)
    eng = ColonyEngine([1, 1, 1])
    for tma in ([False, True] if shape[2] % 2 == 0 else [False]):
        src = torch.from_numpy(field.copy()).cuda()
        scratch = torch.full_like(src, float("nan"))
        out = eng.diffuse3d(src, steps, dt, 0.4 * dt, D, H * H, 2.0, 0.1, use_tma=tma, scratch=scratch)
        assert (out.data_ptr() == scratch.data_ptr()) == bool(steps % 2)
        assert np.array_equal(out.cpu().numpy(), ref), f"tma={tma}"


def test_diffusion3d_c4_sized_grid_mass_and_tma_equals_plain():
    g = 226                                                  # the C4 colony (1587 um cube) at h = 7.07 um
    rng = np.random.default_rng(3)
    field = torch.from_numpy(rng.random((g, g, g)) * 2).cuda()
    total = field.sum().item()
    eng = ColonyEngine([1, 1, 1])
    dt = 0.15 * H * H / D
    a = eng.diffuse3d(field.clone(), 4, dt, dt, D, H * H, 2.0, 0.1, use_tma=True)
    b = eng.diffuse3d(field.clone(), 4, dt, dt, D, H * H, 2.0, 0.1, use_tma=False)
    assert torch.equal(a, b)
    assert abs(a.sum().item() - 0.9 * total) <= 1e-9 * total
    assert a.min().item() >= 0 and a.max().item() <= 2.0
    # a slab of it against the oracle (the whole grid would take the CPU a while)
    sub = field[:40, :50, :].cpu().numpy()
    got = eng.diffuse3d(torch.from_numpy(sub.copy()).cuda(), 3, dt, 0.5 * dt, D, H * H, 2.0, 0.1).cpu().numpy()
    assert np.array_equal(got, orc.diffuse_field3d(sub, 3, dt, 0.5 * dt, D, H * H, 2.0, 0.1))


def test_method_surface_3d_extension_and_reference_plane_semantics():
    rng = np.random.default_rng(5)
    n, grid = 3000, (30, 26, 12)
    size = np.array([(grid[0] - 1) * H, (grid[1] - 1) * H, (grid[2] - 1) * H])
    loc = rng.random((n, 3)) * size
    amounts = rng.random(n)
    field0 = rng.random(grid) * 2
    dt = 0.15 * H * H / D
    for volume in (True, False):
        sim = Cells(loc, 5.0, size)
        sim.b200_diffusion_3d = volume
        sim.spat_res, sim.spat_res2, sim.diffuse_const, sim.diffuse_dt = H, H * H, D, dt
        sim.max_concentration, sim.degradation = 2.0, 0.1
        sim.step_dt = dt * 3 + 0.5 * dt                      # 4 substeps, the last with dt/2
        sim.fgf4_values = field0.copy()
        sampled = sim.update_concentrations("fgf4_values", amounts, diffuse=True)
        ref = field0.copy()
        if volume:
            orc.adjust_morphogens3d(loc, amounts, ref, H)
            ref = orc.diffuse_field3d(ref, 4, dt, 0.5 * dt, D, H * H, 2.0, 0.1)
        else:                                                # the reference: deposits and diffusion on plane 0 only
            orc.adjust_morphogens(loc, amounts, ref, H)
            ref[:, :, 0] = orc.diffuse_field(ref[:, :, 0], 4, dt, 0.5 * dt, D, H * H, 2.0, 0.1)
        assert np.abs(sim.fgf4_values - ref).max() <= 1e-12   # deposits are fp64 atomics: order-dependent last bits
        assert np.abs(sampled - orc.get_concentrations(loc, ref, H)).max() <= 1e-12


@pytest.mark.parametrize("substeps", [4])
def test_hot_path_step_carries_a_3d_field(substeps):
    """pipeline.HotPathStep with a (gx,gy,gz) grid: deposit over the cube corners, 7-point diffusion, sampling, on the
    step's starting positions, beside the mechanics -- also when replayed as a CUDA graph."""
    from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
    rng = np.random.default_rng(8)
    n, grid = 4000, (24, 20, 16)
    size = np.array([(grid[0] - 1) * H, (grid[1] - 1) * H, (grid[2] - 1) * H])
    loc = rng.random((n, 3)) * size
    dt = 0.15 * H * H / D
    cfg = StepConfig(grid=grid, spat_res=H, diffuse_dt=dt, diffuse_steps=substeps, bond_stride=32,
                     search_cols_per_cell=60.0, overlap_morphogen=True)       # the reference's order of the stages
    ref = np.zeros(grid)
    pos = loc.copy()
    want = []
    for graph in (False, True):
        eng = ColonyEngine(size)
        eng.upload(loc, np.full(n, 5.0), np.zeros((n, 3)))
        one = HotPathStep(eng, cfg)
        if graph:
            one.step(); one.step()
            one.capture()
            for _ in range(2):
                one.replay()
        else:
            for _ in range(4):
                one.step()
        torch.cuda.synchronize()
        one.check()
        want.append((one.field.cpu().numpy(), one.sampled[:n].cpu().numpy(), eng.locations_host().copy()))
    # the oracle: four steps of deposit / diffuse / sample on the positions each step starts from
    sim = orc.OracleSim(loc, np.full(n, 5.0), size)
    for _ in range(4):
        orc.adjust_morphogens3d(sim.locations, 1.0, ref, H)
        ref = orc.diffuse_field3d(ref, substeps, dt, dt, D, H * H, 2.0, 0.1)
        sampled = orc.get_concentrations(sim.locations, ref, H)
        sim.apply_forces()
    for field, got, _ in want:
        assert np.abs(field - ref).max() <= 1e-11
        assert np.abs(got - sampled).max() <= 1e-11
    assert np.array_equal(want[0][2], want[1][2])
