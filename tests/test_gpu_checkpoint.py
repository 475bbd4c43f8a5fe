"""Checkpoint / resume (reference simulation.py:290-298 temp() pickles the whole Simulation object every
step, :520-536 mode 1 loads it and continues from current_step + 1).  A simulation on the B200 backend is pickled in
the middle of a run -- device-resident graphs, tracked arrays, pinned buffers and all --, reloaded into a fresh object
and continued: it must reach exactly the states, graphs and positions of the run that was never interrupted."""
import pickle
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from hipsc_abm_b200.cell_methods import B200CellMethods
from hipsc_abm_b200.simulation import ColonySimulation
from oracle.model import ColonyRules

ARRAYS = ("states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK", "div_counters", "diff_counters", "death_counters",
          "fds_counters", "radii")


class B200Backed(ColonyRules, B200CellMethods, ColonySimulation):
    def __init__(self, size):
        ColonySimulation.__init__(self, size)
        self.init_constants(size)


@pytest.mark.parametrize("size", [[300.0, 300.0, 0.0], [120.0, 120.0, 120.0]], ids=["2d", "3d"])
def test_pickle_reload_continue_is_identical(size):
    random.seed(99)
    np.random.seed(99)
    sim = B200Backed(size)
    sim.seed_colony(700)
    sim.run(4)
    # ---- temp(): the whole object, protocol -1 (reference simulation.py:294-298), plus the RNG streams the
    # reference leaves to the process (they are not part of its checkpoint either)
    blob = pickle.dumps(sim, protocol=-1)
    rng_state = (random.getstate(), np.random.get_state())
    for step in range(5, 9):
        sim.run_one(step)
    # ---- mode 1: load and continue from current_step + 1 (reference simulation.py:524-536)
    resumed = pickle.loads(blob)
    assert type(resumed.locations) is np.ndarray                 # a checkpoint holds plain arrays ...
    assert "_b200_engine" not in resumed.__dict__ and "_b200_cache" not in resumed.__dict__   # ... and no device handles
    assert resumed.current_step == 4
    random.setstate(rng_state[0])
    np.random.set_state(rng_state[1])
    for step in range(resumed.current_step + 1, 9):
        resumed.run_one(step)
    assert resumed.number_agents == sim.number_agents
    for name in ARRAYS:
        assert np.array_equal(getattr(resumed, name), getattr(sim, name)), name
    assert np.array_equal(resumed.locations, sim.locations)      # bit-identical: same kernels, same inputs
    for g in ("neighbor_graph", "jkr_graph"):
        assert np.array_equal(getattr(resumed, g).edge_array(), getattr(sim, g).edge_array()), g
        assert getattr(resumed, g).max_agents == getattr(sim, g).max_agents
    assert sim.jkr_graph.ecount() > 0 and sim.number_agents >= 700


def test_checkpoint_of_a_morphogen_field_and_method_times():
    sim = B200Backed([200.0, 200.0, 0.0])
    random.seed(1)
    np.random.seed(1)
    sim.seed_colony(300)
    h = 7.07106
    sim.spat_res, sim.spat_res2, sim.diffuse_const, sim.diffuse_dt = h, h * h, 50.0, 0.2 * h * h / 50.0
    sim.max_concentration, sim.degradation = 2.0, 0.1
    g = int(np.ceil(200.0 / h)) + 1
    sim.fgf4_values = np.zeros((g, g, 1))
    sim.run(2)
    keep = sim.step_dt
    sim.step_dt = 4.5 * sim.diffuse_dt
    sim.update_concentrations("fgf4_values", 1.0, diffuse=True)
    sim.step_dt = keep
    copy = pickle.loads(pickle.dumps(sim, protocol=-1))
    assert np.array_equal(copy.fgf4_values, sim.fgf4_values) and type(copy.fgf4_values) is np.ndarray
    assert copy.method_times.keys() == sim.method_times.keys() and "apply_forces" in copy.method_times
    for s in (sim, copy):
        s.step_dt = 4.5 * s.diffuse_dt
        s.out = s.update_concentrations("fgf4_values", 0.5, diffuse=True)
    assert np.abs(copy.fgf4_values - sim.fgf4_values).max() <= 1e-13      # fp64 atomics in the deposit
    assert np.abs(copy.out - sim.out).max() <= 1e-13
