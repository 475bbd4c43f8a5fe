"""k_jkr_substep_plane (csrc/jkr_plane.cuh), the two-axis substep kernel of a flat world, against the three-axis
kernel it replaces: positions, bond rows and row lengths must be IDENTICAL bit for bit -- the z terms it drops are exact
no-ops of the reference's arithmetic (cell_backend.py:79-109, 155-170 with size_z == 0) -- and against the oracle."""
import numpy as np
import pytest
import torch

from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def colony(n, side, seed, radii="uniform", clump=False):
    rng = np.random.default_rng(seed)
    size = np.array([side, side, 0.0])
    loc = rng.random((n, 3)) * size
    if clump:                                   # a crowd: more than 16 partners per cell and runs longer than the tile
        k = n // 5
        loc[:k, :2] = 0.5 * side + rng.normal(0, 90.0, (k, 2))
        loc[:, :2] = np.clip(loc[:, :2], 0, side)
    mot = rng.normal(0, 1.5e-9, (n, 3)) * [1, 1, 0]
    rad = np.full(n, 5.0) if radii == "uniform" else 4.0 + 1.5 * rng.random(n)
    return size, loc, rad, mot


def run(size, loc, rad, mot, plane, steps, stride=32):
    eng = ColonyEngine(size)
    eng.plane_kernel = plane
    eng.upload(loc, rad, mot)
    pipe = HotPathStep(eng, StepConfig(bond_stride=stride, search_cols_per_cell=60.0))
    for _ in range(steps):
        pipe.step()
    torch.cuda.synchronize()
    pipe.check()
    b = pipe.current_bonds()
    n = len(loc)
    lens = b.row_len[:n].cpu().numpy()
    rows = b.cols.view(-1, stride)[:n].cpu().numpy()
    rows = np.where(np.arange(stride)[None, :] < lens[:, None], rows, -1)
    return eng.locations_host(), lens, rows, eng.edge_list(b)


@pytest.mark.parametrize("radii", ["uniform", "per_cell"])
@pytest.mark.parametrize("clump", [False, True], ids=["spread", "crowded"])
def test_plane_kernel_is_bit_identical_to_the_three_axis_kernel(radii, clump):
    size, loc, rad, mot = colony(20000, 1400.0, 3, radii, clump)
    a = run(size, loc, rad, mot, True, 3, stride=64 if clump else 32)
    b = run(size, loc, rad, mot, False, 3, stride=64 if clump else 32)
    assert np.array_equal(a[0], b[0])                   # positions: same bits
    assert np.array_equal(a[1], b[1])                   # bond row lengths
    assert np.array_equal(np.sort(a[2], axis=1), np.sort(b[2], axis=1))   # bond rows (order inside a row is free mid-loop)
    assert np.array_equal(a[3], b[3])
    assert np.all(a[0][:, 2] == 0.0)


def test_plane_kernel_matches_the_oracle():
    size, loc, rad, mot = colony(12000, 1100.0, 8)
    sim = orc.OracleSim(loc, rad, size)
    for _ in range(2):
        sim.motility_forces = mot.copy()
        sim.apply_forces()
    got = run(size, loc, rad, mot, True, 2)
    assert np.abs(got[0] - sim.locations).max() <= 1e-10 * np.abs(sim.locations).max()
    assert np.array_equal(got[3], sim.jkr_graph.edge_array())


def test_single_launch_flag_selects_the_plane_kernel():
    """abm_jkr_substep(sort_rows | ABM_SUBSTEP_FLAT_Z) on positions whose z is 0 == the same call without the flag."""
    from hipsc_abm_b200 import _lib as _l
    size, loc, rad, mot = colony(15000, 1200.0, 4)
    outs = []
    for flag in (0, _l.SUBSTEP_FLAT_Z):
        eng = ColonyEngine(size)
        eng.upload(loc, rad, mot)
        reach = eng.jkr_reach(10.0)
        srt = eng.sort(reach)
        n, stride = len(loc), 32
        b = eng._bond_buffers(n, stride)
        b["max_len"].zero_()
        _l.check(eng.lib.abm_jkr_substep(
            srt.bin_start.data_ptr(), srt.sorted_id.data_ptr(), srt.sorted_pos.data_ptr(), srt.sorted_bin.data_ptr(),
            n, _l.i3(srt.dims), 10.0, reach, 0, 0, 0, 0, b["len"][0].data_ptr(), b["cols"][0].data_ptr(), stride,
            b["max_len"].data_ptr(), _l.SUBSTEP_SORT_ROWS | flag, eng.mot.data_ptr(), eng.loc.data_ptr(), 0.5, 1000.0,
            0.000107, 10000.0, _l.d3(size), 180.0, 0, eng.uniform_radius, None, None, 0, 0, 0, None, eng.stream),
            "abm_jkr_substep")
        torch.cuda.synchronize()
        lens = b["len"][0][:n].cpu().numpy()
        rows = b["cols"][0].view(-1, stride)[:n].cpu().numpy()
        outs.append((eng.locations_host(), lens, np.where(np.arange(stride)[None, :] < lens[:, None], rows, -1)))
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


def test_flat_z_flag_and_timed_loop_change_nothing_but_the_kernel_and_the_clock():
    """ABM_LOOP_FLAT_Z (first substep on the two-axis kernel too) and abm_mechanics_loop_timed (events around every
    substep launch) give the same positions and bonds as the plain loop; the events measure something."""
    size, loc, rad, mot = colony(15000, 1200.0, 6)
    assert np.all(loc[:, 2] == 0.0)
    dts = [180.0] * 10 + [0.0]
    outs = []
    for mode in ("plain", "flat_z", "timed"):
        eng = ColonyEngine(size)
        eng.upload(loc, rad, mot)
        eng.flat_z_vouched = mode == "flat_z"
        events = None
        if mode == "timed":
            events = [torch.cuda.Event(enable_timing=True) for _ in range(2 * len(dts))]
            for e in events:
                e.record()
            eng.substep_events = events
        bonds = eng.enqueue_mechanics(None, 10.0, dts, 32)
        eng.substep_events = None
        torch.cuda.synchronize()
        if events:
            us = [events[2 * k].elapsed_time(events[2 * k + 1]) * 1e3 for k in range(len(dts))]
            assert all(0.5 < u < 5e4 for u in us), us
        outs.append((eng.locations_host(), eng.edge_list(bonds)))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0])
        assert np.array_equal(outs[0][1], other[1])
