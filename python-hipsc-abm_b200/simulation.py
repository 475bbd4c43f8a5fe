"""Host-side mirror of the reference `Simulation` surface for the hot path
(reference simulation.py:64-288 and backend.py:199-212), backed by libabm_b200.so.

`B200Simulation` is a mixin: put it in front of the reference's classes

    class FastCells(B200CellMethods, B200Simulation, CellSimulation): ...

and `assign_bins`, `get_neighbors`, `agent_graph` (and, from cell_methods.py, `calculate_jkr`,
`apply_forces`, `update_diffusion`, ...) run on the B200 while every rule/output method keeps working on
the same numpy agent arrays.  It can also be used without the reference: `ColonySimulation` below carries
just enough of the base-class API (add_agents / agent_array / agent_graph) to drive the path.
"""
from __future__ import annotations

import time
from functools import wraps

import numpy as np

from .engine import ColonyEngine
from .graph import DeviceGraph


def record_time(function):
    """Same contract as the reference decorator (backend.py:199-212): stores the wall time of the call in
    simulation.method_times[function.__name__] and discards the return value.  The device is synchronised
    before the clock stops, otherwise the asynchronous launches would make the number meaningless."""
    @wraps(function)
    def wrap(simulation, *args, **kwargs):
        start = time.perf_counter()
        function(simulation, *args, **kwargs)
        engine = simulation.__dict__.get("_b200_engine")
        if engine is not None:
            import torch
            torch.cuda.synchronize(engine.device)
        simulation.method_times[function.__name__] = time.perf_counter() - start
    return wrap


class B200Simulation:
    """assign_bins / get_neighbors / agent_graph of the reference Simulation, on the device."""

    # ---- engine plumbing --------------------------------------------------------------------------------
    def b200_engine(self) -> ColonyEngine:
        eng = self.__dict__.get("_b200_engine")
        if eng is None:
            eng = ColonyEngine(self.size, self.__dict__.get("b200_device"))
            self.__dict__["_b200_engine"] = eng
        return eng

    def b200_upload(self, with_forces=False):
        """Phase boundary host -> device: the rule methods replace / edit the numpy agent arrays freely
        (np.delete, np.concatenate, in-place +=; reference cell_methods.py:39-44, 86-105), so the arrays the
        path reads are re-sent at the start of every public hot-path call."""
        eng = self.b200_engine()
        want_jkr = with_forces is True            # with_forces="motility": the fused loop never reads jkr_forces
        eng.upload(self.locations, self.radii,
                   self.motility_forces if with_forces and hasattr(self, "motility_forces") else None,
                   self.jkr_forces if want_jkr and hasattr(self, "jkr_forces") else None)
        return eng

    def __getstate__(self):
        # Simulation.temp() pickles the whole object (reference simulation.py:290-298): drop device handles
        state = dict(self.__dict__)
        state.pop("_b200_engine", None)
        return state

    # ---- reference surface --------------------------------------------------------------------------------
    def agent_graph(self, graph_name):
        """reference simulation.py:181-188 with the igraph store replaced by DeviceGraph."""
        self.__dict__[graph_name] = DeviceGraph(self.number_agents)
        self.graph_names.append(graph_name)

    def assign_bins(self, max_agents, distance):
        """reference simulation.py:190-223.  Returns the reference-shaped (bins, bins_help, bin_locations,
        max_agents) as numpy arrays; capacity follows the reference's rule (double the worst bin)."""
        from . import _lib
        eng = self.b200_upload()
        srt = eng.sort(distance, _lib.SORT_STABLE | _lib.SORT_REFERENCE)
        worst = eng.read_i32(srt.max_count)
        if max_agents < worst:
            max_agents = worst * 2
        bins, bins_help, bin_locations = eng.bins_dense(srt, max_agents)
        return eng.host(bins), eng.host(bins_help), eng.host(bin_locations), int(max_agents)

    @record_time
    def get_neighbors(self, graph_name, distance, clear=True):
        """reference simulation.py:225-288: all pairs within `distance` go into the graph (replacing its
        edges if `clear`, else merged with them)."""
        graph = self.__dict__[graph_name]
        eng = self.b200_upload()
        if graph.n != self.number_agents:
            raise ValueError(f"{graph_name} has {graph.n} vertices for {self.number_agents} agents")
        old = None if clear else graph.device_tensors(eng)
        srt = eng.sort(distance)
        found = eng.search(srt, distance, old)
        # keep the reference's sticky capacities meaningful (backend.py:22-23, simulation.py:221,281)
        worst_bin = eng.read_i32(srt.max_count)
        if graph.max_agents < worst_bin:
            graph.max_agents = worst_bin * 2
        graph.set_device(found, eng)


class ColonySimulation(B200Simulation):
    """Minimal stand-alone base with the reference's agent-array API (simulation.py:110-188), for using
    the hot path without the reference package."""

    def __init__(self, size, device=None):
        self.size = np.asarray(size, dtype=float)
        self.number_agents = 0
        self.agent_array_names = []
        self.graph_names = []
        self.method_times = {}
        if device is not None:
            self.b200_device = device

    def add_agents(self, number, agent_type=None):
        begin = self.number_agents
        self.number_agents += number
        if agent_type is not None:
            if not hasattr(self, "agent_types"):
                self.agent_types = {}
            self.agent_types[agent_type] = (begin, self.number_agents)

    def agent_array(self, array_name, agent_type=None, dtype=float, vector=None, func=None, override=None):
        if override is not None:
            if override.shape[0] != self.number_agents:
                raise Exception("Length of override array does not match number of agents in simulation!")
            self.__dict__[array_name] = override
            self.agent_array_names.append(array_name)
        elif not hasattr(self, array_name):
            self.agent_array_names.append(array_name)
            shape = self.number_agents if vector is None else (self.number_agents, vector)
            self.__dict__[array_name] = (np.empty(shape, dtype=object) if dtype in (str, object)
                                         else np.zeros(shape, dtype=dtype))
        if func is not None:
            begin, end = (0, self.number_agents) if agent_type is None else self.agent_types[agent_type]
            for i in range(begin, end):
                self.__dict__[array_name][i] = func()
