"""The counter-based generator (include/abm_rng.h) and the restated rule methods (oracle/abm_oracle.c orc_rule_*):
the deterministic log / sin / cos against libm, the draws' statistics, independence from cell order and count, and
a model run that behaves like the reference's (cells divide, differentiate after the doxycycline step, the
Boolean network moves, motility forces have the reference's magnitude)."""
import ctypes as C

import numpy as np

from oracle import oracle as orc
from oracle import rules_oracle as ro


def test_deterministic_math_matches_libm():
    L = ro.lib()
    x = np.concatenate([np.random.default_rng(0).random(20000), [1e-300, 1e-12, 0.5, 0.70710678, 0.99999999, 1.0]])
    lg, sn, cs = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    L.orc_rng_math(len(x), x, lg, sn, cs)
    ref = np.log(x)
    assert np.abs((lg - ref) / np.where(ref == 0, 1, ref))[np.abs(ref) > 1e-3].max() < 1e-15     # ~2 ulp
    assert np.abs(lg - ref)[np.abs(ref) <= 1e-3].max() < 1e-18
    u = x - np.floor(x)
    # (numpy's own sin(2*pi*u) carries the rounding of 2*pi*u, ~7e-16 near u = 1)
    assert np.abs(sn - np.sin(2 * np.pi * u)).max() < 2e-15 and np.abs(cs - np.cos(2 * np.pi * u)).max() < 2e-15
    assert np.abs(sn * sn + cs * cs - 1).max() < 1e-15


def test_draws_are_uniform_gaussian_and_independent_of_neighbouring_counters():
    L = ro.lib()
    n = 400_000
    u, g = np.zeros(n), np.zeros(n)
    L.orc_rng_uniform(n, 7, 3, 1, 0, u)
    L.orc_rng_gauss(n, 7, 3, 3, 0, g)
    assert 0 <= u.min() and u.max() < 1
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 1e-3
    hist = np.histogram(u, bins=64, range=(0, 1))[0]
    assert ((hist - n / 64) ** 2 / (n / 64)).sum() < 120                       # chi^2, 63 dof
    assert abs(g.mean()) < 6e-3 and abs(g.var() - 1) < 1e-2
    assert abs(((g - g.mean()) ** 3).mean()) < 2e-2 and abs((g ** 4).mean() - 3) < 6e-2     # skewness, kurtosis
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 5e-3                         # consecutive cell ids
    v = np.zeros(n)
    L.orc_rng_uniform(n, 7, 4, 1, 0, v)                                         # next step, same cells
    assert abs(np.corrcoef(u, v)[0, 1]) < 5e-3
    w = np.zeros(n)
    L.orc_rng_uniform(n, 7, 3, 1, 0, w)
    assert np.array_equal(u, w)                                                 # a pure function of its counters


def make(n=1500, seed=3, size=(500.0, 500.0, 0.0), dox_step=2):
    rng = np.random.default_rng(seed)
    size = np.array(size)
    P = ro.default_params(size, seed=seed, dox_step=dox_step)
    loc = rng.random((n, 3)) * size
    return ro.RulesOracle(loc, 5.0, size, ro.initial_state(n, seed, P), P)


def test_rules_do_not_depend_on_cell_order():
    a, b = make(), make()
    perm = np.random.default_rng(1).permutation(a.number_agents)
    for name in ro.INT_ARRAYS + ("locations", "radii"):
        setattr(b, name, np.ascontiguousarray(getattr(b, name)[perm]))
    for s in (a, b):
        s.step(6)
    # every draw is keyed by persistent cell ids, so the cells that existed before the step end up in the same state
    # wherever they sit in memory.  (Daughters get their ids in memory order of their mothers -- which IS id order in
    # a real run, arrays only ever grow at the end and shrink stably -- so in this shuffled colony they are excluded;
    # they are isolated in the neighbour graph during their first step and influence nobody.)
    n0 = 1500
    ka, kb = a.gid < n0, b.gid < n0
    ord_a, ord_b = np.argsort(a.gid[ka]), np.argsort(b.gid[kb])
    assert np.array_equal(a.gid[ka][ord_a], b.gid[kb][ord_b])
    for name in ("states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK", "div_counters", "diff_counters", "death_counters"):
        assert np.array_equal(getattr(a, name)[ka][ord_a], getattr(b, name)[kb][ord_b]), name
    assert a.number_agents == b.number_agents and a.added == b.added > 0


def test_model_run_behaves_like_the_reference_model():
    sim = make(n=1200, dox_step=3)
    n0 = sim.number_agents
    seen_force = 0.0
    hist = []
    for step in range(1, 41):
        sim.step(step)
        hist.append((sim.number_agents, int(sim.states.sum()), int(sim.GATA6.sum()), sim.added, sim.removed))
        assert len(sim.locations) == sim.number_agents == len(sim.states) == sim.neighbor_graph.vcount()
        assert sim.jkr_graph.vcount() == sim.number_agents
        assert len(np.unique(sim.gid)) == sim.number_agents                     # persistent ids stay unique
    assert sim.number_agents > n0 * 1.3                                         # ~1/72 of the cells divide per step
    assert sum(h[3] for h in hist) > 0.3 * n0
    assert sim.states.sum() > 0 and sim.GATA6.sum() > 0                         # the network fired after dox_step
    assert hist[1][2] == 0                                                      # ... and not before it
    assert not sim.motility_forces.any()                                        # apply_forces leaves them zeroed
    assert np.all(sim.locations >= 0) and np.all(sim.locations[:, :2] <= 500.0)


def test_motility_force_magnitudes():
    sim = make(n=800)
    L, P = ro.lib(), sim.P
    P.step = 9
    sim.get_neighbors("neighbor_graph", 15)
    ptr, cols = sim._csr()
    mot = np.zeros((sim.number_agents, 3))
    L.orc_rule_motility(sim.number_agents, ptr, cols, sim.gid, sim.states, sim.GATA6, sim.NANOG, sim.locations, mot,
                        C.byref(P))
    mag = np.linalg.norm(mot, axis=1)
    crowded = np.diff(ptr) >= 6
    assert np.all(mag[crowded] == 0) and np.all(mag[~crowded] > 0)
    assert mag.max() <= 2e-9 * (1 + 1e-12) and mag[~crowded].min() >= 0.6 * 2e-9 * (1 - 1e-12)   # |0.8 n + 0.2 r| in [0.6, 1]
    assert np.all(mot[:, 2] == 0)
