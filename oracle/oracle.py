"""CPU oracle for the hiPSC-ABM cell-update hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package never does (tests/test_no_oracle_in_product.py
enforces it).  The arithmetic lives in ``abm_oracle.c`` (one C function per reference JIT function);
this file restates the reference's *host* control flow around those functions:

  =====================  =============================================
  here                   reference (/root/reference/ABM)
  =====================  =============================================
  HostGraph              backend.py:13-28 (igraph.Graph subclass) -- 8-method facade, SURVEY 8(b)
  OracleSim.assign_bins  simulation.py:190-223
  OracleSim.get_neighbors simulation.py:225-288
  OracleSim.calculate_jkr cell_methods.py:342-384
  OracleSim.apply_forces cell_methods.py:386-439
  OracleSim.update_diffusion cell_methods.py:441-468
  OracleSim.get_concentration / adjust_morphogens  cell_methods.py:470-521
  =====================  =============================================

Parity status: pinned by tests/test_oracle_golden.py against fixtures produced from the reference's
own functions (oracle/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libabm_oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile libabm_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "abm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libabm_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_version.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_bin_dims.argtypes = [_f64p, C.c_double, _i64p]
        L.orc_bin_locations.argtypes = [C.c_int64, _f64p, C.c_double, _i64p]
        L.orc_assign_bins.argtypes = [C.c_int64, _i64p, _i64p, _i64p, _i64p, C.c_int64]
        L.orc_assign_bins.restype = C.c_int64
        L.orc_get_neighbors.argtypes = [C.c_int64, _f64p, _i64p, _i64p, _i64p, _i64p, C.c_int64, C.c_double,
                                        _i64p, _u8p, _i64p, C.c_int64]
        L.orc_get_neighbors.restype = C.c_int64
        jkr_args = [C.c_int64, _i64p, _u8p, _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double]
        L.orc_jkr_forces.argtypes = jkr_args
        L.orc_jkr_forces_mt.argtypes = jkr_args
        L.orc_apply_forces.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _f64p, C.c_double, _f64p, C.c_double]
        L.orc_update_diffusion.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                           C.c_double, C.c_double, _f64p]
        L.orc_update_diffusion_field.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                                 C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_update_diffusion_field.restype = C.c_int
        L.orc_update_diffusion3d_field.argtypes = [_f64p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                                   C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_update_diffusion3d_field.restype = C.c_int
        L.orc_adjust_morphogens3d.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _i64p, C.c_double]
        L.orc_get_concentrations.argtypes = [C.c_int64, _f64p, _f64p, _i64p, C.c_double, _f64p]
        L.orc_adjust_morphogens.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _i64p, C.c_double]
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def canonical_edges(edges) -> np.ndarray:
    """(min,max) per row, lexicographically sorted, duplicates and self loops removed."""
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    if len(e) == 0:
        return e
    lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
    keep = lo != hi
    lo, hi = lo[keep], hi[keep]
    if len(lo) == 0:
        return np.zeros((0, 2), dtype=np.int64)
    m = int(hi.max()) + 1
    key = np.unique(lo * m + hi)            # 1-D integer sort: (lo,hi) lexicographic order
    return np.stack([key // m, key % m], axis=1)


class HostGraph:
    """Undirected simple-graph store with the 8 methods the reference uses from igraph.

    Edge rows are kept as (min,max).  ``simplify`` leaves them lexicographically sorted and
    ``neighbors`` is ascending -- the harness *defines* both orders (igraph is not installed, SURVEY
    8(c)); they affect only fp summation order and RNG consumption order, identically for oracle
    and build.
    """

    def __init__(self, n: int = 0):
        self.n = int(n)
        self.edges = np.zeros((0, 2), dtype=np.int64)
        self.max_neighbors = 1
        self.max_agents = 1
        self._adj = None

    # -- mutation -------------------------------------------------------------------------------
    def add_edges(self, es):
        es = np.asarray(es, dtype=np.int64).reshape(-1, 2)
        if len(es):
            rows = np.stack([np.minimum(es[:, 0], es[:, 1]), np.maximum(es[:, 0], es[:, 1])], axis=1)
            self.edges = np.concatenate([self.edges, rows], axis=0)
            self._adj = None

    def delete_edges(self, which=None):
        if which is None:
            self.edges = np.zeros((0, 2), dtype=np.int64)
        else:
            which = np.asarray(which, dtype=np.int64)
            keep = np.ones(len(self.edges), dtype=bool)
            keep[which] = False
            self.edges = self.edges[keep]
        self._adj = None

    def simplify(self):
        self.edges = canonical_edges(self.edges)
        self._adj = None

    def add_vertex(self):
        self.n += 1
        self._adj = None

    def add_vertices(self, k):
        self.n += int(k)
        self._adj = None

    def delete_vertices(self, idx):
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        if len(idx) == 0:
            return
        dead = np.zeros(self.n, dtype=bool)
        dead[idx] = True
        renumber = np.cumsum(~dead) - 1
        if len(self.edges):
            keep = ~(dead[self.edges[:, 0]] | dead[self.edges[:, 1]])
            self.edges = renumber[self.edges[keep]]
        self.n -= int(dead.sum())
        self._adj = None

    # -- queries --------------------------------------------------------------------------------
    def vcount(self):
        return self.n

    def ecount(self):
        return len(self.edges)

    def get_edgelist(self):
        return [tuple(row) for row in self.edges.tolist()]

    def edge_array(self) -> np.ndarray:
        return self.edges

    def _build_adj(self):
        e = self.edges
        both = np.concatenate([e, e[:, ::-1]], axis=0) if len(e) else np.zeros((0, 2), dtype=np.int64)
        order = np.lexsort((both[:, 1], both[:, 0]))
        both = both[order]
        counts = np.bincount(both[:, 0], minlength=self.n) if len(both) else np.zeros(self.n, dtype=np.int64)
        ptr = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(counts, out=ptr[1:])
        self._adj = (ptr, both[:, 1].copy())

    def neighbors(self, i):
        if self._adj is None:
            self._build_adj()
        ptr, cols = self._adj
        return cols[ptr[i]:ptr[i + 1]].tolist()

    def num_neighbors(self, i):
        if self._adj is None:
            self._build_adj()
        ptr, _ = self._adj
        return int(ptr[i + 1] - ptr[i])

    def degrees(self) -> np.ndarray:
        if self._adj is None:
            self._build_adj()
        return np.diff(self._adj[0])


# ---------------------------------------------------------------------------------------------------
# stage functions on plain arrays
# ---------------------------------------------------------------------------------------------------
def assign_bins(locations, size, distance, max_agents):
    """reference simulation.py:190-223 (grow-and-retry until every bin fits)."""
    L = lib()
    locations = np.ascontiguousarray(locations, dtype=np.float64)
    n = len(locations)
    size = np.ascontiguousarray(size, dtype=np.float64)
    dims = np.zeros(3, dtype=np.int64)
    L.orc_bin_dims(size, float(distance), dims)
    bin_locations = np.zeros((n, 3), dtype=np.int64)
    L.orc_bin_locations(n, locations, float(distance), bin_locations)
    max_agents = int(max_agents)
    while True:
        bins_help = np.zeros(tuple(dims), dtype=np.int64)
        bins = np.zeros(tuple(dims) + (max_agents,), dtype=np.int64)
        worst = L.orc_assign_bins(n, bin_locations, bins, bins_help, dims, max_agents)
        if max_agents >= worst:
            break
        max_agents = int(worst) * 2
    return bins, bins_help, bin_locations, max_agents


def neighbor_edges(locations, size, distance, max_agents=1, max_neighbors=1):
    """reference simulation.py:239-284 without the graph: returns (edges[if_edge], max_agents, max_neighbors)."""
    L = lib()
    locations = np.ascontiguousarray(locations, dtype=np.float64)
    n = len(locations)
    bins, bins_help, bin_locations, max_agents = assign_bins(locations, size, distance, max_agents)
    dims = np.array(bins_help.shape, dtype=np.int64)
    max_neighbors = int(max_neighbors)
    while True:
        length = n * max_neighbors
        edges = np.zeros((length, 2), dtype=np.int64)
        if_edge = np.zeros(length, dtype=np.uint8)
        edge_count = np.zeros(n, dtype=np.int64)
        worst = L.orc_get_neighbors(n, locations, bin_locations, bins, bins_help, dims, max_agents,
                                    float(distance), edges, if_edge, edge_count, max_neighbors)
        if max_neighbors >= worst:
            break
        max_neighbors = int(worst) * 2
    return edges[if_edge.astype(bool)], max_agents, max_neighbors


JKR_ADHESION = 0.000107   # reference cell_methods.py:347
JKR_POISSON = 0.5         # reference cell_methods.py:348
JKR_YOUNGS = 1000.0       # reference cell_methods.py:349
STOKES = 10000.0          # reference cell_methods.py:392


def jkr_forces(edges, locations, radii, forces, multithread=False):
    """reference cell_backend.py:66-115; returns delete flags, accumulates into ``forces`` in place."""
    L = lib()
    edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 2)
    delete = np.zeros(len(edges), dtype=np.uint8)
    fn = L.orc_jkr_forces_mt if multithread else L.orc_jkr_forces
    fn(len(edges), edges, delete, np.ascontiguousarray(locations, dtype=np.float64),
       np.ascontiguousarray(radii, dtype=np.float64), forces, JKR_POISSON, JKR_YOUNGS, JKR_ADHESION)
    return delete.astype(bool)


def integrate(jkr_force, motility_force, locations, radii, size, move_dt, stokes=STOKES):
    """reference cell_backend.py:148-172 (in place on ``locations``)."""
    lib().orc_apply_forces(len(locations), np.ascontiguousarray(jkr_force), np.ascontiguousarray(motility_force),
                           locations, np.ascontiguousarray(radii, dtype=np.float64), float(stokes),
                           np.ascontiguousarray(size, dtype=np.float64), float(move_dt))
    return locations


def diffuse_padded(base, steps, diffuse_dt, last_dt, diffuse_const, spat_res2):
    """reference cell_backend.py:175-204 on a ghost-padded 2-D array (in place)."""
    base = np.ascontiguousarray(base, dtype=np.float64)
    scratch = np.zeros_like(base)
    lib().orc_update_diffusion(base, base.shape[0], base.shape[1], int(steps), float(diffuse_dt), float(last_dt),
                               float(diffuse_const), float(spat_res2), scratch)
    return base


def diffuse_field(plane, steps, diffuse_dt, last_dt, diffuse_const, spat_res2, max_concentration, degradation):
    """reference cell_methods.py:451-465 on one (gx,gy) plane; returns the new plane."""
    plane = np.ascontiguousarray(plane, dtype=np.float64).copy()
    rc = lib().orc_update_diffusion_field(plane, plane.shape[0], plane.shape[1], int(steps), float(diffuse_dt),
                                          float(last_dt), float(diffuse_const), float(spat_res2),
                                          float(max_concentration), float(degradation))
    if rc != 0:
        raise MemoryError("oracle diffusion scratch allocation failed")
    return plane


def diffuse_field3d(field, steps, diffuse_dt, last_dt, diffuse_const, spat_res2, max_concentration, degradation):
    """3-D extension of update_diffusion on a (gx,gy,gz) field (PARITY UNPINNED: no reference counterpart; the 2-D
    algorithm of cell_backend.py:175-204 with a third axis, b = 1-6a); returns the new field."""
    field = np.ascontiguousarray(field, dtype=np.float64).copy()
    assert field.ndim == 3
    rc = lib().orc_update_diffusion3d_field(field, field.shape[0], field.shape[1], field.shape[2], int(steps),
                                            float(diffuse_dt), float(last_dt), float(diffuse_const), float(spat_res2),
                                            float(max_concentration), float(degradation))
    if rc != 0:
        raise MemoryError("oracle diffusion scratch allocation failed")
    return field


def adjust_morphogens3d(locations, amounts, gradient, spat_res):
    """adjust_morphogens over the eight corners of the cell's grid cube (3-D extension, PARITY UNPINNED)."""
    assert gradient.flags.c_contiguous and gradient.dtype == np.float64 and gradient.ndim == 3
    amounts = np.ascontiguousarray(np.broadcast_to(np.asarray(amounts, dtype=np.float64), (len(locations),)))
    lib().orc_adjust_morphogens3d(len(locations), np.ascontiguousarray(locations, dtype=np.float64), amounts, gradient,
                                  np.array(gradient.shape, dtype=np.int64), float(spat_res))
    return gradient


def get_concentrations(locations, gradient, spat_res):
    """reference cell_methods.py:470-483 for every cell."""
    gradient = np.ascontiguousarray(gradient, dtype=np.float64)
    out = np.zeros(len(locations), dtype=np.float64)
    lib().orc_get_concentrations(len(locations), np.ascontiguousarray(locations, dtype=np.float64), gradient,
                                 np.array(gradient.shape, dtype=np.int64), float(spat_res), out)
    return out


def adjust_morphogens(locations, amounts, gradient, spat_res):
    """reference cell_methods.py:485-521 for every cell in index order (in place on ``gradient``)."""
    assert gradient.flags.c_contiguous and gradient.dtype == np.float64 and gradient.ndim == 3
    amounts = np.ascontiguousarray(np.broadcast_to(np.asarray(amounts, dtype=np.float64), (len(locations),)))
    lib().orc_adjust_morphogens(len(locations), np.ascontiguousarray(locations, dtype=np.float64), amounts, gradient,
                                np.array(gradient.shape, dtype=np.int64), float(spat_res))
    return gradient


# ---------------------------------------------------------------------------------------------------
# the reference method surface, restated on top of the stage functions
# ---------------------------------------------------------------------------------------------------
class OracleSim:
    """Holds exactly the state the hot path touches and exposes the reference's method names."""

    def __init__(self, locations, radii, size, max_radius=5.0, step_dt=1800.0, move_dt=180.0):
        self.locations = np.ascontiguousarray(locations, dtype=np.float64).copy()
        self.number_agents = len(self.locations)
        self.radii = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, dtype=np.float64),
                                                          (self.number_agents,))).copy()
        self.size = np.asarray(size, dtype=np.float64)
        self.max_radius = max_radius
        self.step_dt = step_dt
        self.move_dt = move_dt
        self.motility_forces = np.zeros((self.number_agents, 3))
        self.jkr_forces = np.zeros((self.number_agents, 3))
        self.neighbor_graph = HostGraph(self.number_agents)
        self.jkr_graph = HostGraph(self.number_agents)
        self.multithread = False

    def assign_bins(self, max_agents, distance):
        return assign_bins(self.locations, self.size, distance, max_agents)

    def get_neighbors(self, graph_name, distance, clear=True):
        graph = getattr(self, graph_name)
        if clear:
            graph.delete_edges(None)
        edges, graph.max_agents, graph.max_neighbors = neighbor_edges(
            self.locations, self.size, distance, graph.max_agents, graph.max_neighbors)
        graph.add_edges(edges)
        if not clear:
            graph.simplify()

    def calculate_jkr(self):
        edges = self.jkr_graph.edge_array()
        if len(edges) > 0:
            delete = jkr_forces(edges, self.locations, self.radii, self.jkr_forces, self.multithread)
            self.jkr_graph.delete_edges(np.arange(len(edges))[delete])

    jkr_neighbor_forces = calculate_jkr

    def apply_forces(self):
        steps, last_dt = divmod(self.step_dt, self.move_dt)
        total_steps = int(steps) + 1
        for step in range(total_steps):
            self.get_neighbors("jkr_graph", 2 * self.max_radius, clear=False)
            self.calculate_jkr()
            move_dt = last_dt if step == total_steps - 1 else self.move_dt
            integrate(self.jkr_forces, self.motility_forces, self.locations, self.radii, self.size, move_dt)
            self.jkr_forces[:, :] = 0
        self.motility_forces[:, :] = 0

    # diffusion attributes are supplied by the caller, as a reference subclass would
    def update_diffusion(self, gradient_name):
        steps, last_dt = divmod(self.step_dt, self.diffuse_dt)
        steps = int(steps) + 1
        field = getattr(self, gradient_name)
        field[:, :, 0] = diffuse_field(field[:, :, 0], steps, self.diffuse_dt, last_dt, self.diffuse_const,
                                       self.spat_res2, self.max_concentration, self.degradation)

    def get_concentration(self, gradient_name, index):
        return float(get_concentrations(self.locations[index:index + 1], getattr(self, gradient_name),
                                        self.spat_res)[0])

    def adjust_morphogens(self, gradient_name, index, amount):
        adjust_morphogens(self.locations[index:index + 1], amount, getattr(self, gradient_name), self.spat_res)

    def update_concentrations(self, gradient_name, amounts=None):
        """Batched a12+a13 (no reference caller): sample every cell first, then deposit in index order."""
        field = getattr(self, gradient_name)
        sampled = get_concentrations(self.locations, field, self.spat_res)
        if amounts is not None:
            adjust_morphogens(self.locations, amounts, field, self.spat_res)
        return sampled
