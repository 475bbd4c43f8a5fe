"""The mixins in front of the REAL reference classes (build container only: needs /root/reference, which does not
travel to the GPU box).  INTEGRATION.md says `class FastCells(B200CellMethods, B200Simulation, CellSimulation)`
takes over exactly the hot path; this constructs that class from the reference's own modules (igraph stand-in and
pythonabm alias as in oracle/make_golden.py), checks which class every method resolves to, runs the reference's
agent_initials() and its five rule methods against the DeviceGraph facade, and checks that the hot-path calls refuse
to run without a GPU instead of falling back to anything."""
import contextlib
import io
import os
import random
import sys

import numpy as np
import pytest

REF = "/root/reference/ABM"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from oracle import make_golden as mg          # imports the reference with the shims in place
    import yaml
    work = tmp_path_factory.mktemp("mro")
    os.makedirs(work / "templates")
    general = yaml.safe_load(open(os.path.join(REF, "templates", "general.yaml")))
    experimental = yaml.safe_load(open(os.path.join(REF, "templates", "experimental.yaml")))
    general.update(num_to_start=300, cuda=False, end_step=2, size=[250.0, 250.0, 0.0], output_values=False,
                   output_images=False, record_initial_step=False)
    experimental.update(num_gata6=0, output_tda=False, output_gradients=False, dox_step=1)
    for name, keys in (("general.yaml", general), ("experimental.yaml", experimental)):
        yaml.safe_dump(keys, open(work / "templates" / name, "w"))
    os.makedirs(work / "out" / "fast")
    return mg, work


def make_class(mg):
    from hipsc_abm_b200.cell_methods import B200CellMethods
    from hipsc_abm_b200.simulation import B200Simulation

    if "FastCells" not in globals():
        class FastCells(B200CellMethods, B200Simulation, mg.cell_simulation.CellSimulation):
            pass
        FastCells.__qualname__ = "FastCells"          # module-level name: Simulation.temp() pickles the object
        globals()["FastCells"] = FastCells
    return globals()["FastCells"], B200CellMethods, B200Simulation


def test_method_resolution_order(ref):
    mg, _ = ref
    FastCells, B200CellMethods, B200Simulation = make_class(mg)
    names = [c.__name__ for c in FastCells.__mro__]
    assert names.index("B200CellMethods") < names.index("B200Simulation") < names.index("CellSimulation") < \
        names.index("CellMethods") < names.index("Simulation")
    owner = lambda m: next(c.__name__ for c in FastCells.__mro__ if m in c.__dict__)
    # the hot path is taken over ...
    for m in ("calculate_jkr", "jkr_neighbor_forces", "apply_forces", "update_diffusion", "get_concentration",
              "adjust_morphogens", "update_concentrations"):
        assert owner(m) == "B200CellMethods", m
    for m in ("assign_bins", "get_neighbors", "agent_graph"):
        assert owner(m) == "B200Simulation", m
    # ... and nothing else: rules, outputs, runtime stay the reference's
    for m in ("cell_death", "cell_division", "cell_pathway", "cell_differentiate", "cell_motility", "cell_growth"):
        assert owner(m) == "CellMethods", m
    for m in ("steps", "agent_initials"):
        assert owner(m) == "CellSimulation", m
    for m in ("add_agents", "agent_array", "temp", "step_values", "data", "info", "random_vector", "start"):
        assert owner(m) == "Simulation", m
    # the timing contract (backend.py:199-212): wrapped methods, keyed by name, return value dropped
    for m in ("get_neighbors", "apply_forces", "update_diffusion"):
        assert getattr(FastCells, m).__name__ == m and hasattr(getattr(FastCells, m), "__wrapped__")


def test_reference_rules_run_on_the_device_graph_facade_and_hot_path_needs_a_gpu(ref):
    import torch
    mg, work = ref
    FastCells, _, _ = make_class(mg)
    from hipsc_abm_b200 import AbmError, DeviceGraph
    cwd = os.getcwd()
    os.chdir(work)
    try:
        random.seed(5)
        np.random.seed(5)
        sim = FastCells("fast", str(work / "out") + os.sep)
        sim.agent_initials()                                   # the reference's own, through OUR agent_graph
        assert isinstance(sim.neighbor_graph, DeviceGraph) and isinstance(sim.jkr_graph, DeviceGraph)
        assert sim.neighbor_graph.vcount() == sim.number_agents == 300
        # give the graphs some edges on the host (what a search would have produced) and run the reference's rules
        from scipy.spatial import cKDTree
        pairs = np.array(sorted(cKDTree(sim.locations).query_pairs(15.0)))
        sim.neighbor_graph.add_edges(pairs)
        sim.current_step = 3
        n0 = sim.number_agents
        with contextlib.redirect_stdout(io.StringIO()):
            for step in range(40):                              # enough steps for divisions / deaths to happen
                sim.cell_division()
                sim.cell_death()
                sim.cell_pathway()
                sim.cell_differentiate()
                sim.cell_motility()
        assert sim.number_agents != n0 or sim.states.sum() > 0
        assert sim.neighbor_graph.vcount() == sim.number_agents == sim.jkr_graph.vcount() == len(sim.locations)
        assert np.abs(sim.motility_forces).max() > 0
        if not torch.cuda.is_available():
            for call in (lambda: sim.get_neighbors("neighbor_graph", 15), sim.apply_forces, sim.steps):
                with pytest.raises(AbmError, match="no CPU path"), contextlib.redirect_stdout(io.StringIO()):
                    call()
            assert "get_neighbors" not in sim.method_times      # nothing ran, nothing was timed
    finally:
        os.chdir(cwd)


def test_reference_outputs_and_checkpoint_work_on_tracked_arrays(ref):
    """the reference's own step_values (CSV), step_image (OpenCV), data (CSV), temp (pickle) and
    the mode-1 reload read the agent arrays the hot path hands back -- write-counting ndarray views on pinned memory
    (arrays.TrackedArray).  They must behave exactly like plain arrays there, and a checkpoint must hold plain arrays."""
    import csv
    import pickle
    mg, work = ref
    FastCells, _, _ = make_class(mg)
    from hipsc_abm_b200.arrays import TrackedArray, track
    cwd = os.getcwd()
    os.chdir(work)
    try:
        random.seed(6)
        np.random.seed(6)
        sim = FastCells("fast", str(work / "out") + os.sep)
        sim.agent_initials()
        plain = {name: np.array(sim.__dict__[name], copy=True) for name in sim.agent_array_names}
        for name in sim.agent_array_names:                    # what the path does to the arrays it touches
            if sim.__dict__[name].dtype != object:
                sim.__dict__[name] = track(sim.__dict__[name])
        assert isinstance(sim.locations, TrackedArray)
        sim.current_step, sim.step_start = 1, 0.0
        sim.output_values = sim.output_images = True
        sim.image_quality = 200
        sim.method_times = {"get_neighbors": 0.1, "apply_forces": 0.2}
        with contextlib.redirect_stdout(io.StringIO()):
            sim.step_values(arrays=["locations", "FGF4", "states", "div_counters"])
            sim.step_image()
            sim.data()
            sim.temp()
        rows = list(csv.reader(open(os.path.join(sim.values_path, "fast_values_1.csv"))))
        assert rows[0] == ["locations[0]", "locations[1]", "locations[2]", "FGF4", "states", "div_counters"]
        assert len(rows) == 1 + sim.number_agents
        assert np.allclose(np.array(rows[1:], dtype=float)[:, :3], plain["locations"])
        assert os.path.getsize(os.path.join(sim.images_path, "fast_image_1.png")) > 0
        data = list(csv.reader(open(os.path.join(sim.main_path, "fast_data.csv"))))
        assert data[0][4:6] == ["get_neighbors", "apply_forces"] and int(data[1][1]) == sim.number_agents
        with open(os.path.join(sim.main_path, "fast_temp.pkl"), "rb") as f:
            back = pickle.load(f)                              # mode 1 (simulation.py:524-536)
        assert type(back.locations) is np.ndarray and "_b200_engine" not in back.__dict__
        for name, a in plain.items():
            assert np.array_equal(back.__dict__[name], a), name
        assert back.neighbor_graph.vcount() == sim.number_agents and back.current_step == 1
        # the rules keep working on the reloaded object (plain arrays) and on the live one (tracked arrays) alike
        for s in (sim, back):
            with contextlib.redirect_stdout(io.StringIO()):
                s.cell_division()
                s.cell_motility()
            assert len(s.locations) == s.number_agents
    finally:
        os.chdir(cwd)
