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

    # Arithmetic of the mechanics loop (the dtype flag of SURVEY 8(b)): "f64" = the reference's dtype, forces and
    # positions within 1e-12 of the oracle; "f32" = single precision on bin-relative coordinates, within 1e-5
    # (csrc/jkr32.cu).  The agent arrays keep their float64 type either way.
    b200_dtype = "f64"

    # "tracked": a host array is uploaded only when it changed since the device copy was made (arrays.py);
    # "always": every public call re-uploads what it reads (always correct, also for writes numpy cannot show)
    b200_coherence = "tracked"

    def b200_mark_dirty(self, *names):
        """Tell the path that `names` (all agent arrays / fields if none given) were modified behind numpy's back
        (raw pointers, ctypes, a JIT kernel): their device copies are dropped and re-uploaded at the next use."""
        cache = self.__dict__.setdefault("_b200_cache", {})
        for name in (names or list(cache)):
            cache.pop(name, None)

    def b200_device_array(self, name, shape=None, dtype=np.float64, view=None):
        """Device tensor holding the current contents of the host array `self.<name>`.  Uploads -- and from then on
        tracks the array's in-place writes -- only if the device copy is not known to be current.
        shape: required shape (a scalar / strided array is materialised first); view: callable taking the host array
        to the part that travels (e.g. the [:, :, 0] plane of a morphogen field)."""
        from .arrays import DeviceCopy, track
        eng = self.b200_engine()
        cache = self.__dict__.setdefault("_b200_cache", {})
        a = self.__dict__[name]
        entry = cache.get(name)
        if entry is not None and self.b200_coherence == "tracked" and entry.current_for(a):
            return entry.tensor, False
        host = np.asarray(a)
        if shape is not None and host.shape != tuple(shape):
            host = np.ascontiguousarray(np.broadcast_to(host, shape))     # scalar radius etc.: untracked, re-sent
            tracked = None
        else:
            tracked = track(a)
            if tracked is not a:
                self.__dict__[name] = tracked                    # same memory, now with a write counter
            host = tracked.view(np.ndarray)
        part = host if view is None else view(host)
        tensor = eng.to_device(np.ascontiguousarray(part, dtype=dtype), name)
        if tracked is not None:
            cache[name] = DeviceCopy(tracked, tensor)
        else:
            cache.pop(name, None)
        return tensor, True

    def b200_adopt_host(self, name, array, tensor):
        """`array` (fresh from a device download of `tensor`) becomes self.<name>, known to be coherent."""
        from .arrays import DeviceCopy, track
        tracked = track(array)
        self.__dict__[name] = tracked
        self.__dict__.setdefault("_b200_cache", {})[name] = DeviceCopy(tracked, tensor)
        return tracked

    def b200_upload(self, with_forces=False):
        """Phase boundary host -> device.  The rule methods replace / edit the numpy agent arrays freely
        (np.delete, np.concatenate, in-place +=; reference cell_methods.py:39-44, 86-105); each array the path
        reads is sent again only if it changed since its device copy was made (arrays.py), so an unchanged
        `locations` / `radii` costs nothing at the next call."""
        eng = self.b200_engine()
        eng.use_current_stream()
        eng.mechanics_dtype = self.b200_dtype
        n = int(self.number_agents)
        loc, loc_new = self.b200_device_array("locations", (n, 3))
        rad, rad_new = self.b200_device_array("radii", (n,))
        mot = jkr = None
        if with_forces and hasattr(self, "motility_forces"):
            mot, _ = self.b200_device_array("motility_forces", (n, 3))
        if with_forces is True and hasattr(self, "jkr_forces"):    # "motility": the fused loop handles jkr itself
            jkr, _ = self.b200_device_array("jkr_forces", (n, 3))
        stale_n = eng.n != n
        eng.adopt(n, loc if (loc_new or stale_n or eng.loc is not loc) else None,
                  rad if (rad_new or stale_n or eng.radii is not rad) else None,
                  np.asarray(self.radii).view(np.ndarray) if np.shape(self.radii) == (n,) else None, mot, jkr)
        if mot is None and (eng.mot is None or eng.mot.shape[0] != n):
            eng.mot = eng._zeros((n, 3), __import__("torch").float64)
        if jkr is None and (eng.jkr is None or eng.jkr.shape[0] != n):
            eng.jkr = eng._zeros((n, 3), __import__("torch").float64)
        return eng

    def __getstate__(self):
        # Simulation.temp() pickles the whole object (reference simulation.py:290-298): drop device handles
        state = dict(self.__dict__)
        state.pop("_b200_engine", None)
        state.pop("_b200_cache", None)           # device copies; the arrays themselves pickle as plain ndarrays
        state.pop("_b200_zeroed", None)
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
