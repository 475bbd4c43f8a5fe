"""fp32 mode of the mechanics loop (csrc/jkr32.cu, ABM_LOOP_FP32; north star: "per-step forces and positions within
rel 1e-5 in fp32").  Every step starts from the fp64 oracle's state (teacher forcing, SURVEY 4 hazard iii) and is
compared with the oracle's step:
  * positions: |x32 - x64| <= 1e-5 * max|x64| (the north star's gate), and -- much stricter -- the DISPLACEMENT of
    the step within 1e-5 of the largest displacement (positions relative to the colony size would hide a force error);
  * bonds: identical except for pairs that sit at a threshold (|dist - 10| or the bond-breaking point within 1e-4 um);
  * determinism: two runs give identical bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import hipsc_abm_b200  # noqa: F401
from oracle import oracle as orc

from test_gpu_parity import Cells, colony


def edge_set(e):
    return set(map(tuple, np.asarray(e).tolist()))


def run_case(n, dim, density, radii_mode, steps=5, seed=0):
    loc, radii, size = colony(seed + dim, n, dim, density=density)
    rng = np.random.default_rng(seed + 7)
    if radii_mode == "mixed":
        radii = rng.uniform(5 / 2 ** 0.5, 5.0, n)
    mot = rng.normal(0, 2e-9, (n, 3)) * (np.array([1, 1, 0]) if dim == 2 else 1)
    sim = Cells(loc, radii, size)
    sim.b200_dtype = "f32"
    ref = orc.OracleSim(loc, radii, size)
    report = dict(max_rel_pos=0.0, max_rel_disp=0.0, bond_diff=0, bonds=0)
    for step in range(steps):
        m = mot if step % 2 == 0 else -mot
        sim.motility_forces[:] = m
        ref.motility_forces[:] = m
        start = ref.locations.copy()
        sim.apply_forces()
        ref.apply_forces()
        dref = ref.locations - start
        d32 = sim.locations - start
        if step >= 2:         # (the random start is ill-conditioned: pairs almost on top of each other, test_gpu_scale.py)
            report["max_rel_pos"] = max(report["max_rel_pos"],
                                        np.abs(sim.locations - ref.locations).max() / np.abs(ref.locations).max())
            report["max_rel_disp"] = max(report["max_rel_disp"], np.abs(d32 - dref).max() / np.abs(dref).max())
        else:
            report["start_rel_pos"] = max(report.get("start_rel_pos", 0.0),
                                          np.abs(sim.locations - ref.locations).max() / np.abs(ref.locations).max())
        a, b = edge_set(sim.jkr_graph.edge_array()), edge_set(ref.jkr_graph.edge_array())
        diff = a ^ b
        report["bond_diff"] += len(diff)
        report["bonds"] += len(b)
        for (i, j) in diff:                              # every disagreement sits at a threshold
            dist = np.linalg.norm(ref.locations[i] - ref.locations[j])
            brk = radii[i] + radii[j] + 0.360562e6 * (np.pi * 1.07e-4 / 666.6666666666666) ** (2 / 3) * \
                (1 / (1e6 * (1 / radii[i] + 1 / radii[j]))) ** (1 / 3)
            assert min(abs(dist - 10.0), abs(dist - brk)) < 1e-3, (i, j, dist, brk)
        # teacher forcing: next step from the oracle's state (positions and bonds)
        sim.locations = ref.locations.copy()
        sim.jkr_graph.delete_edges(None)
        sim.jkr_graph.add_edges(ref.jkr_graph.edge_array())
    return report


@pytest.mark.parametrize("n,dim,density,radii_mode", [(30000, 2, 0.011, "uniform"), (20000, 3, 0.0012, "uniform"),
                                                      (8000, 2, 0.012, "mixed")])
def test_fp32_step_within_1e5_of_the_fp64_oracle(n, dim, density, radii_mode):
    r = run_case(n, dim, density, radii_mode)
    print("\n[fp32]", n, dim, radii_mode, r)
    assert r["max_rel_pos"] <= 1e-5                     # the north star's gate
    assert r["max_rel_disp"] <= 1e-5                    # the meaningful one: the step's displacement
    assert r["start_rel_pos"] <= 1e-3                   # even the violent first separation stays sane
    assert r["bond_diff"] <= max(2, 2e-5 * r["bonds"])  # bond sets differ only at thresholds, and rarely


def test_fp32_is_deterministic_and_far_colony_is_as_accurate_as_a_near_one():
    """Bin-relative coordinates: the same colony 51 mm from the origin is moved as accurately as next to it (absolute
    fp32 coordinates would lose three more digits there), and two runs give the same bits.  The colony sits 60 um
    inside its box in both cases so that no wall clamps anything."""
    n = 6000
    loc, radii, size = colony(3, n, 2, density=0.011)
    rng = np.random.default_rng(1)
    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    pad = np.array([60.0, 60.0, 0.0])

    # relax the random start first (fp64, three steps): what is compared is ONE step of a settled colony
    pre = Cells(loc + pad, radii, size + 2 * pad)
    for step in range(3):
        pre.motility_forces[:] = mot if step % 2 == 0 else -mot
        pre.apply_forces()
    start = np.array(pre.locations) - pad
    bonds = pre.jkr_graph.edge_array()

    def run(shift, dtype):
        off = pad + [shift, shift, 0.0]
        sim = Cells(start + off, radii, size + 2 * pad + [shift, shift, 0.0])
        sim.b200_dtype = dtype
        sim.jkr_graph.add_edges(bonds)
        before = sim.locations.copy()
        sim.motility_forces[:] = -mot
        sim.apply_forces()
        return sim.locations - before, sim.jkr_graph.edge_array()
    a, b = run(0.0, "f32"), run(0.0, "f32")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ref = run(0.0, "f64")[0]
    scale = np.abs(ref).max()
    near = np.abs(a[0] - ref).max() / scale
    far = np.abs(run(51200.0, "f32")[0] - run(51200.0, "f64")[0]).max() / scale
    print(f"\n[fp32] step displacement error relative to the largest displacement: at the origin {near:.2e}, "
          f"51 mm away {far:.2e}")
    # (the shifted lattice cuts the bins elsewhere, so the two fp32 runs are not bit-equal -- but equally accurate)
    assert near <= 1e-5 and far <= 1e-5


def test_fp32_walls_overfull_bins_and_empty():
    # cells pushed against / beyond the walls, coincident cells, a clump that overflows the hit column
    loc = [[10.0, 10.0, 0.0], [10.0, 10.0, 0.0], [0.0, 0.0, 0.0], [-1.2, 5.0, 0.0], [101.3, 99.0, 0.0],
           [100.0, 100.0, 0.0]]
    mot = np.array([[1e-9, 0, 0], [0, -1e-9, 0], [-3e-9, -3e-9, 0], [-1e-9, 0, 0], [2e-9, 2e-9, 0], [5e-9, 5e-9, 0]])
    sim = Cells(loc, 5.0, [100.0, 100.0, 0.0])
    sim.b200_dtype = "f32"
    ref = orc.OracleSim(np.array(loc), 5.0, [100.0, 100.0, 0.0])
    sim.motility_forces[:] = mot
    ref.motility_forces[:] = mot
    sim.apply_forces()
    ref.apply_forces()
    assert np.abs(sim.locations - ref.locations).max() < 1e-4
    assert sim.locations.min() >= 0.0 and sim.locations[:, :2].max() <= 100.0
    rng = np.random.default_rng(2)
    n = 500
    clump = np.zeros((n, 3))
    clump[:, :2] = 30 + rng.normal(0, 3.0, (n, 2))
    sim = Cells(clump, 5.0, [60.0, 60.0, 0.0])
    sim.b200_dtype = "f32"
    sim.jkr_graph.row_stride_hint = 8
    ref = orc.OracleSim(clump, 5.0, [60.0, 60.0, 0.0])
    sim.apply_forces()
    ref.apply_forces()
    assert sim.jkr_graph.row_stride_hint > 8                       # the row overflow was seen and retried
    a, b = edge_set(sim.jkr_graph.edge_array()), edge_set(ref.jkr_graph.edge_array())
    assert len(a ^ b) <= 0.01 * len(b)                             # a violently relaxing clump: thresholds are busy
    empty = Cells(np.zeros((0, 3)), 5.0, [50.0, 50.0, 0.0])
    empty.b200_dtype = "f32"
    empty.apply_forces()
