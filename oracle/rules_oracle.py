"""CPU oracle of the DEVICE-RESIDENT model step (TEST INFRASTRUCTURE ONLY): the reference's step loop
(cell_simulation.py:77-126) with the rule methods restated on the counter-based generator of include/abm_rng.h
(oracle/abm_oracle.c orc_rule_*) around the pinned hot-path oracle (oracle.OracleSim).

PARITY UNPINNED against the reference: the random stream and the update order of cell_pathway differ by design (see
the header of the orc_rule_* block); what is pinned is (a) the hot path underneath, (b) the rules' statistics
(tests/test_rules_oracle.py), and (c) that the CUDA rule kernels reproduce this restatement bit for bit
(tests/test_gpu_rules.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as orc

INT_ARRAYS = ("states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK", "death_counters", "diff_counters", "div_counters",
              "fds_counters", "gid")


class RuleParams(C.Structure):
    _fields_ = [("seed", C.c_int64), ("step", C.c_int64), ("field", C.c_int32), ("dox_step", C.c_int32),
                ("lonely_thresh", C.c_int32), ("guye_move", C.c_int32), ("pluri_div_thresh", C.c_int32),
                ("diff_div_thresh", C.c_int32), ("pluri_to_diff", C.c_int32), ("death_thresh", C.c_int32),
                ("fds_thresh", C.c_int32), ("three_d", C.c_int32), ("max_radius", C.c_double),
                ("min_radius", C.c_double), ("motility_force", C.c_double)]


def default_params(size, seed=0, dox_step=5):
    """cell_simulation.py:33-57 + templates/experimental.yaml as used for the golden runs."""
    return RuleParams(seed=seed, step=0, field=2, dox_step=dox_step, lonely_thresh=2, guye_move=1,
                      pluri_div_thresh=36, diff_div_thresh=72, pluri_to_diff=36, death_thresh=144, fds_thresh=1,
                      three_d=int(size[2] != 0), max_radius=5.0, min_radius=5.0 / 2 ** 0.5, motility_force=2e-9)


_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_bound = False


def lib():
    global _bound
    L = orc.lib()
    if not _bound:
        PP = C.POINTER(RuleParams)
        L.orc_rule_division_mark.argtypes = [C.c_int64, _i32, _i32, _i32, _i32, PP, _u8]
        L.orc_rule_division_vector.argtypes = [C.c_int64, _i32, PP, _f64]
        L.orc_rule_death_mark.argtypes = [C.c_int64, _i32, _i32, _i32, PP, _u8]
        L.orc_rule_pathway.argtypes = [C.c_int64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, PP]
        L.orc_rule_differentiate.argtypes = [C.c_int64, _i32, _i32, _i32, _i32, _i32, PP]
        L.orc_rule_motility.argtypes = [C.c_int64, _i64, _i32, _i32, _i32, _i32, _i32, _f64, _f64, PP]
        L.orc_rng_uniform.argtypes = [C.c_int64] * 5 + [_f64]
        L.orc_rng_gauss.argtypes = [C.c_int64] * 5 + [_f64]
        L.orc_rng_math.argtypes = [C.c_int64, _f64, _f64, _f64, _f64]
        _bound = True
    return L


def initial_state(n, seed, params):
    """A seeded colony in the shape of cell_simulation.py:128-157 (numpy Generator instead of Python's random: the
    initial draw is not part of the per-step contract)."""
    rng = np.random.default_rng(seed)
    f = params.field
    z = lambda: np.zeros(n, dtype=np.int32)
    draw = lambda hi: rng.integers(0, hi, n).astype(np.int32)
    return dict(FGF4=draw(f), FGFR=draw(f), ERK=draw(f), GATA6=z(), NANOG=draw(f), states=z(),
                death_counters=draw(params.death_thresh), diff_counters=draw(params.pluri_to_diff),
                div_counters=draw(params.pluri_div_thresh), fds_counters=draw(params.fds_thresh),
                gid=np.arange(n, dtype=np.int32))


class RulesOracle(orc.OracleSim):
    """One colony: agent arrays + both graphs on the host, stepped with the restated rules and the hot-path oracle."""

    def __init__(self, locations, radii, size, state, params: RuleParams):
        super().__init__(locations, radii, size)
        self.P = params
        for name in INT_ARRAYS:
            setattr(self, name, np.ascontiguousarray(state[name], dtype=np.int32).copy())
        self.next_gid = int(self.gid.max(initial=-1)) + 1
        self.float_arrays = ("locations", "radii", "motility_forces", "jkr_forces")

    def _csr(self):
        e = self.neighbor_graph.edge_array()
        n = self.number_agents
        a = np.concatenate([e[:, 0], e[:, 1]])
        b = np.concatenate([e[:, 1], e[:, 0]])
        order = np.lexsort((b, a))
        ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(a, minlength=n), out=ptr[1:])
        return ptr, np.ascontiguousarray(b[order], dtype=np.int32)

    def _degrees(self):
        return np.diff(self._csr()[0]).astype(np.int32)

    def step(self, step_number):
        L, P = lib(), self.P
        P.step = step_number
        self.get_neighbors("neighbor_graph", 15)                               # cell_simulation.py:90
        # ---- cell_division (cell_methods.py:54-117)
        n = self.number_agents
        flags = np.zeros(n, dtype=np.uint8)
        L.orc_rule_division_mark(n, self.gid, self.states, self.div_counters, self._degrees(), C.byref(P), flags)
        mothers = np.nonzero(flags)[0]
        k = len(mothers)
        if k:
            vec = np.zeros((k, 3))
            L.orc_rule_division_vector(k, np.ascontiguousarray(self.gid[mothers]), C.byref(P), vec)
            for name in INT_ARRAYS + self.float_arrays:
                a = getattr(self, name)
                setattr(self, name, np.concatenate([a, a[mothers]], axis=0))
            daughters = n + np.arange(k)
            self.locations[mothers] += vec
            self.locations[daughters] -= vec
            self.div_counters[mothers] = 0
            self.div_counters[daughters] = 0
            self.gid[daughters] = self.next_gid + np.arange(k, dtype=np.int32)
            self.next_gid += k
            self.number_agents += k
            self.neighbor_graph.add_vertices(k)
            self.jkr_graph.add_vertices(k)
        # ---- cell_death (cell_methods.py:11-52)
        n = self.number_agents
        flags = np.zeros(n, dtype=np.uint8)
        L.orc_rule_death_mark(n, self.states, self.death_counters, self._degrees(), C.byref(P), flags)
        dead = np.nonzero(flags)[0]
        if len(dead):
            for name in INT_ARRAYS + self.float_arrays:
                setattr(self, name, np.ascontiguousarray(np.delete(getattr(self, name), dead, axis=0)))
            self.neighbor_graph.delete_vertices(dead)
            self.jkr_graph.delete_vertices(dead)
            self.number_agents -= len(dead)
        n = self.number_agents
        ptr, cols = self._csr()
        # ---- cell_pathway, cell_differentiate, cell_motility (cell_methods.py:176-340)
        L.orc_rule_pathway(n, ptr, cols, self.gid, self.FGF4.copy(), self.FGF4, self.FGFR, self.ERK, self.GATA6,
                           self.NANOG, self.fds_counters, C.byref(P))
        L.orc_rule_differentiate(n, self.gid, self.GATA6, self.NANOG, self.states, self.diff_counters, C.byref(P))
        self.locations = np.ascontiguousarray(self.locations)
        self.motility_forces = np.ascontiguousarray(self.motility_forces)
        L.orc_rule_motility(n, ptr, cols, self.gid, self.states, self.GATA6, self.NANOG, self.locations,
                            self.motility_forces, C.byref(P))
        self.apply_forces()                                                     # cell_simulation.py:112
        self.added, self.removed = k, len(dead)
