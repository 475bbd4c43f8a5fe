"""Differentiation-state identity over whole simulations (north-star gate: states identical over 100 steps).

The golden files hold, for every step of the REFERENCE's own CellSimulation.steps() (seeded, one Numba thread),
the discrete state arrays, the agent count and the edge counts of both graphs.  Here the same step loop is run
with the restated host rules (oracle/model.py) on top of
  * the CPU oracle backend  (pins the restated rules and the oracle, no GPU needed), and
  * the B200 backend         (gpu mark: the actual claim -- the kernels drive the model to identical states).

Free-running identity has a horizon: the colony mechanics amplify ulp-level arithmetic differences (libm sqrt/pow
vs BLAS dnrm2 / LLVM pow, summation order) by roughly x10 every 15 steps until a contact at the edge of a
threshold flips (hazard iii).  The horizon is a property of the MODEL, not of an implementation:
oracle/horizon_probe.py runs the reference itself (Numba, one thread) from the recording's seed with one coordinate
of one cell moved by ONE unit in the last place before the first step, and the reference leaves its own recording
after step 83 (five perturbations) / 86 (one) -- tests/golden/horizon_c1_100.json.  No implementation whose
arithmetic is not bit-identical to Numba's LLVM code can stay on the recording longer than that.  So:
  * the free-running tests MEASURE how long each backend stays identical in every discrete array, agent count and
    edge count (the number is printed and, on the GPU box, written to gpurun_out/horizon_b200.json) and require it to
    reach REQUIRED_FRACTION of the reference's own 1-ulp horizon;
  * all 100 steps are asserted with teacher forcing: every step starts from the teacher's state, and the student
    must reproduce its discrete states and graphs exactly and its positions to 1e-12 (the fp64-mode gate).
"""
import json
import os
import random

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as orc
from oracle.model import ColonyRules

from conftest import GOLDEN, ROOT

DISCRETE = ("states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK")
with open(os.path.join(GOLDEN, "horizon_c1_100.json")) as _f:
    REFERENCE_HORIZON = json.load(_f)["horizon_min"]          # 83: the reference vs itself after a 1-ulp nudge
REQUIRED_FRACTION = 0.9


class Diverged(Exception):
    pass


class OracleBacked(ColonyRules, orc.OracleSim):
    def __init__(self, size):
        self.init_constants(size)
        self.multithread = False

    def agent_graph(self, name):
        setattr(self, name, orc.HostGraph(self.number_agents))
        self.graph_names.append(name)


def run_and_compare(sim, g):
    """Free run from the recording's seed.  Returns (steps identical to the recording, what differed first, worst
    position difference while identical)."""
    seed, n0, end_step = int(g["meta"][0]), int(g["meta"][1]), int(g["meta"][2])
    random.seed(seed)
    np.random.seed(seed)
    sim.seed_colony(n0)
    assert np.array_equal(sim.locations, g["locations_0"])           # same RNG streams from the first draw
    worst = [0.0]

    def check(s):
        k = s.current_step
        if s.number_agents != int(g[f"n_{k}"]):
            raise Diverged(k, "agent count")
        for name in DISCRETE:
            if not np.array_equal(getattr(s, name), g[f"{name}_{k}"]):
                raise Diverged(k, name)
        if s.neighbor_graph.ecount() != int(g[f"nbr_edges_{k}"]):
            raise Diverged(k, "neighbour edge count")
        if s.jkr_graph.ecount() != int(g[f"jkr_edges_{k}"]):
            raise Diverged(k, "JKR bond count")
        if f"locations_{k}" in g:
            worst[0] = max(worst[0], np.abs(s.locations - g[f"locations_{k}"]).max())
            assert np.array_equal(s.div_counters, g[f"div_counters_{k}"])

    try:
        sim.run(end_step, check)
    except Diverged as d:
        return d.args[0] - 1, d.args[1], worst[0]
    return end_step, None, worst[0]


def required_steps(end_step):
    return min(end_step, int(REQUIRED_FRACTION * REFERENCE_HORIZON))


def record_horizon(backend, case, horizon, what, drift):
    print(f"\n[horizon] {backend} {case}: identical to the reference recording through step {horizon}"
          f"{'' if what is None else ' (first difference: ' + what + ')'}; reference vs itself after a 1-ulp nudge: "
          f"{REFERENCE_HORIZON}; worst position difference while identical {drift:.2e} um")
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "horizon_b200.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[f"{backend}:{case}"] = {"identical_through_step": horizon, "first_difference": what,
                                     "reference_1ulp_horizon": REFERENCE_HORIZON, "worst_dx_um": drift}
        json.dump(data, open(path, "w"), indent=1)


@pytest.mark.parametrize("case", ["c1_20", "c1_3d_10", "c1_100"])
def test_oracle_backend_reproduces_reference_states(case):
    g = load_golden("model_" + case)
    horizon, what, drift = run_and_compare(OracleBacked(g["meta"][4:7]), g)
    record_horizon("oracle", case, horizon, what, drift)
    assert horizon >= required_steps(int(g["meta"][2])), f"left the recording after step {horizon}: {what}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1_20", "c1_3d_10", "c1_100"])
def test_b200_backend_reproduces_reference_states(case):
    from hipsc_abm_b200.cell_methods import B200CellMethods
    from hipsc_abm_b200.simulation import ColonySimulation

    class B200Backed(ColonyRules, B200CellMethods, ColonySimulation):
        def __init__(self, size):
            ColonySimulation.__init__(self, size)
            self.init_constants(size)

    g = load_golden("model_" + case)
    horizon, what, drift = run_and_compare(B200Backed(g["meta"][4:7]), g)
    record_horizon("b200_fused", case, horizon, what, drift)
    assert horizon >= required_steps(int(g["meta"][2])), f"left the recording after step {horizon}: {what}"


@pytest.mark.gpu
def test_b200_strict_order_backend_horizon():
    """The stage-by-stage path (get_neighbors -> calculate_jkr -> integrate per substep): k_jkr_rows evaluates the
    contact model in the reference's operation order (division, cbrt, sqrt; cell_backend.py:79-109) and sums each
    cell's bonds in ascending partner id, which is the serial reference's += order.  Its horizon against the recording
    is measured next to the fused kernel's; both are bounded by the reference's own 1-ulp horizon."""
    from hipsc_abm_b200.cell_methods import B200CellMethods
    from hipsc_abm_b200.simulation import ColonySimulation

    class Strict(ColonyRules, B200CellMethods, ColonySimulation):
        def __init__(self, size):
            ColonySimulation.__init__(self, size)
            self.init_constants(size)

        def calculate_jkr(self):                  # any override of calculate_jkr selects the unfused loop
            return B200CellMethods.calculate_jkr(self)

    g = load_golden("model_c1_100")
    horizon, what, drift = run_and_compare(Strict(g["meta"][4:7]), g)
    record_horizon("b200_strict_order", "c1_100", horizon, what, drift)
    assert horizon >= required_steps(100), f"left the recording after step {horizon}: {what}"


def copy_state(src, dst):
    """Teacher forcing: give `dst` the teacher's complete state (agent arrays and both graphs)."""
    dst.number_agents, dst.current_step = src.number_agents, src.current_step
    for name in src.agent_array_names:
        setattr(dst, name, getattr(src, name).copy())
    for gname in src.graph_names:
        g_src, g_dst = getattr(src, gname), getattr(dst, gname)
        g_dst.delete_vertices(np.arange(g_dst.vcount()))
        g_dst.add_vertices(src.number_agents)
        g_dst.add_edges(g_src.edge_array())
        g_dst.simplify()


@pytest.mark.gpu
def test_b200_backend_teacher_forced_100_steps():
    """100 steps, every one started from the CPU oracle's state with the same RNG state: discrete states and
    both graphs identical each step, positions within 1e-12 of the colony extent (fp64 gate)."""
    from hipsc_abm_b200.cell_methods import B200CellMethods
    from hipsc_abm_b200.simulation import ColonySimulation

    class B200Backed(ColonyRules, B200CellMethods, ColonySimulation):
        def __init__(self, size):
            ColonySimulation.__init__(self, size)
            self.init_constants(size)

    size = [400.0, 400.0, 0.0]
    random.seed(4321)
    np.random.seed(4321)
    teacher = OracleBacked(size)
    teacher.seed_colony(1000)
    student = B200Backed(size)
    student.agent_array_names = list(teacher.agent_array_names)
    student.agent_graph("neighbor_graph")
    student.agent_graph("jkr_graph")
    worst = 0.0
    for step in range(1, 101):
        copy_state(teacher, student)
        rng = random.getstate()
        teacher.run_one(step)
        after = random.getstate()
        random.setstate(rng)
        student.run_one(step)
        assert random.getstate() == after, f"RNG consumption differs at step {step}"
        assert student.number_agents == teacher.number_agents
        for name in DISCRETE + ("div_counters", "diff_counters", "death_counters"):
            assert np.array_equal(getattr(student, name), getattr(teacher, name)), f"{name} differs at step {step}"
        for gname in ("neighbor_graph", "jkr_graph"):
            assert np.array_equal(getattr(student, gname).edge_array(), orc.canonical_edges(
                getattr(teacher, gname).edge_array())), f"{gname} differs at step {step}"
        err = np.abs(student.locations - teacher.locations).max()
        worst = max(worst, err)
        assert err <= 1e-12 * 400.0, f"positions differ by {err} at step {step}"
    assert teacher.number_agents > 1500 and teacher.states.sum() > 100      # the run actually differentiated
