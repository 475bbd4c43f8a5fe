"""The reference-facing API with host numpy arrays: what travels between host and device per step, and that skipping an
upload never changes a result (arrays.py coherence; SURVEY 7 'host<->device coherence without per-call copies')."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import torch
import hipsc_abm_b200  # noqa: F401
from oracle import oracle as orc

from test_gpu_parity import Cells, close, colony


def field_setup(o, grid, h=7.07106, dt=0.24, sd=10):
    o.spat_res, o.spat_res2, o.diffuse_const, o.diffuse_dt = h, h * h, 50.0, dt
    o.max_concentration, o.degradation = 2.0, 0.1
    o.fgf4_values = np.zeros(grid + (1,))
    o.step_dt_mech = o.step_dt
    o.step_dt_field = dt * (sd - 1) + 0.5 * dt


def step(sim, mot, fused_field=True):
    sim.get_neighbors("neighbor_graph", 15)
    sim.motility_forces[:] = mot
    sim.apply_forces()
    sim.step_dt = sim.step_dt_field
    if fused_field:
        sampled = sim.update_concentrations("fgf4_values", 1.0, diffuse=True)
    else:
        sim.update_concentrations("fgf4_values", 1.0, sample=False)
        sim.update_diffusion("fgf4_values")
        sampled = sim.update_concentrations("fgf4_values", None)
    sim.step_dt = sim.step_dt_mech
    return sampled


def test_steady_state_traffic_and_identity_with_always_upload():
    n = 60_000
    loc, radii, size = colony(3, n, 2)
    g = int(np.ceil(size[0] / 7.07106)) + 1
    g += g % 2
    rng = np.random.default_rng(1)
    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    sims = []
    for mode, fused in (("tracked", True), ("always", False)):
        sim = Cells(loc, radii, size)
        sim.b200_coherence = mode
        field_setup(sim, (g, g))
        sims.append((sim, fused))
    outs = []
    for sim, fused in sims:
        eng = sim.b200_engine()
        per_step = []
        for k in range(4):
            h0, d0 = eng.h2d_bytes, eng.d2h_bytes
            sampled = step(sim, mot if k % 2 == 0 else -mot, fused)
            per_step.append((eng.h2d_bytes - h0, eng.d2h_bytes - d0))
        outs.append((sim.locations.copy(), sim.fgf4_values.copy(), sampled.copy(), sim.jkr_graph.edge_array(),
                     sim.neighbor_graph.edge_array(), per_step))
    a, b = outs
    for k in (0, 3, 4):
        assert np.array_equal(a[k], b[k])                         # skipping uploads changes nothing, bit for bit
    # (the field is built by fp64 atomic adds, whose order -- and with it the last bit -- is not fixed: gate 1e-6)
    assert np.abs(a[1] - b[1]).max() <= 1e-12 and np.abs(a[2] - b[2]).max() <= 1e-12
    field_bytes = g * g * 8
    # tracked, from the second step on: only the step's motility forces go up; positions, the field and the sampled
    # concentrations come down once each
    for h2d, d2h in a[5][1:]:
        assert h2d == n * 24
        assert d2h == n * 24 + field_bytes + n * 8
    # always: locations x4 + radii x2 + motility + field x3 up; field x2 down
    for h2d, d2h in b[5][1:]:
        assert h2d == 4 * n * 24 + 2 * n * 8 + n * 24 + 3 * field_bytes
        assert d2h == n * 24 + 2 * field_bytes + n * 8


def test_host_edits_between_calls_are_honoured():
    """Rule methods edit arrays in place and replace them (division / death): the path must see every such change."""
    n = 5000
    loc, radii, size = colony(8, n, 2)
    rng = np.random.default_rng(2)
    sim = Cells(loc, radii, size)
    ref = orc.OracleSim(loc, radii, size)

    def both(f):
        f(sim)
        f(ref)

    def check():
        sim.get_neighbors("neighbor_graph", 15)
        ref.get_neighbors("neighbor_graph", 15)
        assert np.array_equal(sim.neighbor_graph.edge_array(), orc.canonical_edges(ref.neighbor_graph.edge_array()))
        sim.apply_forces()
        ref.apply_forces()
        assert np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array())
        close(sim.locations, ref.locations)
        sim.locations = ref.locations.copy()          # per-step gate (a plain array: uploaded again, as it must be)

    mot = rng.normal(0, 2e-9, (n, 3)) * [1, 1, 0]
    both(lambda o: o.motility_forces.__setitem__(slice(None), mot))
    check()
    # in-place edits of single rows, of a slice through a view, and with an in-place operator
    both(lambda o: o.locations.__setitem__(17, [100.0, 100.0, 0.0]))
    both(lambda o: o.locations[40:60].__iadd__(0.25))
    both(lambda o: np.add(o.locations[:, 0], 0.125, out=o.locations[:, 0]))
    both(lambda o: o.motility_forces.__setitem__(slice(None), -mot))
    check()
    # radii edited in place (cell_growth does this in the reference when enabled)
    both(lambda o: o.radii.__setitem__(slice(0, 100), 5 / 2 ** 0.5))
    both(lambda o: o.motility_forces.__setitem__(slice(None), mot))
    check()
    # forces left in jkr_forces by the caller are consumed by the first substep (cell_methods.py:404-436)
    pre = rng.normal(0, 1e-9, (n, 3)) * [1, 1, 0]
    both(lambda o: o.jkr_forces.__setitem__(slice(None), pre))
    check()
    assert not sim.jkr_forces.any() and not sim.motility_forces.any()
    # death: arrays replaced wholesale, graphs renumbered
    dead = np.sort(rng.choice(n, 50, replace=False))
    for o in (sim, ref):
        for name in ("locations", "radii", "motility_forces", "jkr_forces"):
            setattr(o, name, np.delete(getattr(o, name), dead, axis=0))
        o.number_agents -= len(dead)
        o.neighbor_graph.delete_vertices(dead)
        o.jkr_graph.delete_vertices(dead)
    both(lambda o: o.motility_forces.__setitem__(slice(None), mot[:n - 50]))
    check()


def test_fused_morphogen_call_equals_three_calls():
    n = 20_000
    loc, radii, size = colony(5, n, 2)
    g = int(np.ceil(size[0] / 7.07106)) + 1
    g += g % 2
    res = []
    for fused in (True, False):
        sim = Cells(loc, radii, size)
        field_setup(sim, (g, g))
        sim.step_dt = sim.step_dt_field
        sim.fgf4_values[...] = np.random.default_rng(0).random((g, g, 1))
        for _ in range(2):
            if fused:
                s = sim.update_concentrations("fgf4_values", np.linspace(0.5, 1.5, n), diffuse=True)
            else:
                sim.update_concentrations("fgf4_values", np.linspace(0.5, 1.5, n), sample=False)
                sim.update_diffusion("fgf4_values")
                s = sim.update_concentrations("fgf4_values", None)
        res.append((sim.fgf4_values.copy(), s.copy()))
    # deposits are fp64 atomics (order-dependent in the last bit): the reference gate for concentrations is 1e-6
    assert np.abs(res[0][0] - res[1][0]).max() <= 1e-12 and np.abs(res[0][1] - res[1][1]).max() <= 1e-12
    # a host edit of the field between calls is seen
    sim.fgf4_values[3, 4, 0] = 1.75
    assert sim.get_concentration("fgf4_values", 0) == sim.fgf4_values[tuple(
        np.ceil(np.floor(2 * loc[0] / 7.07106) / 2).astype(int))]
    sim.locations[0] = [3 * 7.07106, 4 * 7.07106, 0.0]
    assert sim.get_concentration("fgf4_values", 0) == 1.75
