"""A whole model step on the device: the reference's step loop (cell_simulation.py:77-126)

    get_neighbors(15) -> cell_division -> cell_death -> cell_pathway -> cell_differentiate -> cell_motility -> apply_forces

with the agent arrays living in fixed-capacity device tensors, the cell count in device memory and the rule methods as
kernels (csrc/rules.cu).  Nothing travels to the host inside a step and nothing waits for the device, so a step can be
replayed as a CUDA graph; results are read out (and capacities checked) whenever the caller wants them.

The rule kernels draw their randomness from the counter-based contract of include/abm_rng.h and update the FGF4 network
synchronously -- a different realisation of the reference's stochastic model (its Python loops consume one global
Mersenne Twister serially), checked bit for bit against the restated CPU oracle (oracle/rules_oracle.py).  The
seeded-parity path against the ORIGINAL random stream remains the host one: the reference's own rule methods on top of
the B200Simulation / B200CellMethods mixins (tests/test_model_states.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import ColonyEngine, GraphTensors, STOKES

INT_FIELDS = ("gid", "states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK", "death_counters", "diff_counters",
              "div_counters", "fds_counters")


def default_rule_params(size, seed=0, dox_step=5) -> _lib.RuleParams:
    """cell_simulation.py:33-57 and templates/experimental.yaml."""
    return _lib.RuleParams(seed=seed, step=0, field=2, dox_step=dox_step, lonely_thresh=2, guye_move=1,
                           pluri_div_thresh=36, diff_div_thresh=72, pluri_to_diff=36, death_thresh=144, fds_thresh=1,
                           three_d=int(size[2] != 0), max_radius=5.0, min_radius=5.0 / 2 ** 0.5, motility_force=2e-9)


class DeviceColony:
    def __init__(self, size, capacity, params: _lib.RuleParams, device=None, bond_stride=16, search_radius=15.0,
                 contact_radius=10.0, step_dt=1800.0, move_dt=180.0, search_cols_per_cell=14.0):
        self.eng = ColonyEngine(size, device)
        self.dev, self.R, self.stride = self.eng.device, int(capacity), int(bond_stride)
        self.P = params
        self.search_radius, self.contact_radius = float(search_radius), float(contact_radius)
        steps, last_dt = divmod(step_dt, move_dt)                     # cell_methods.py:395-396
        self.dts = [move_dt] * int(steps) + [last_dt]
        f64, i32, dev, R = torch.float64, torch.int32, self.dev, self.R

        def cells():
            d = dict(locations=torch.zeros((R, 3), dtype=f64, device=dev), radii=torch.zeros(R, dtype=f64, device=dev),
                     motility_forces=torch.zeros((R, 3), dtype=f64, device=dev),
                     bond_len=torch.zeros(R, dtype=i32, device=dev),
                     bond_rows=torch.zeros((R, self.stride), dtype=i32, device=dev),
                     nbr_start=torch.zeros(R, dtype=i32, device=dev), nbr_len=torch.zeros(R, dtype=i32, device=dev),
                     n=torch.zeros(1, dtype=i32, device=dev))
            for name in INT_FIELDS:
                d[name] = torch.zeros(R, dtype=i32, device=dev)
            return d
        self.sets = [cells(), cells()]
        self.cur = 0
        self.next_gid = torch.zeros(1, dtype=i32, device=dev)
        self.needed = torch.zeros(2, dtype=i32, device=dev)
        self.index_map = torch.zeros(R, dtype=i32, device=dev)
        self.fgf4_scratch = torch.zeros(R, dtype=i32, device=dev)
        self.ws = torch.empty(int(self.eng.lib.abm_rules_workspace_bytes(R)) + 256, dtype=torch.uint8, device=dev)
        cap = int(search_cols_per_cell * R) + 1024
        self.nbr_cols = torch.empty(cap, dtype=i32, device=dev)
        self.nbr_cursor = torch.zeros(1, dtype=i32, device=dev)
        self.nbr_worst = torch.zeros(1, dtype=i32, device=dev)
        self.bonds = None                     # (len, rows) tensors holding the current JKR bonds
        self.max_radius = float(params.max_radius)
        self.uniform_radius = 0.0             # set by load() when all radii are equal (they are in the shipped model:
        self.steps_done = 0                   # cell_growth is commented out, cell_simulation.py:98)

    # ---- set-up / read-out (host synchronisation) -----------------------------------------------------------------
    def load(self, locations, radii, state: dict):
        """Initial colony: positions, radii and the integer agent arrays of `state` (names as in the reference,
        cell_simulation.py:136-149, plus `gid`)."""
        s = self.sets[self.cur]
        n = int(len(locations))
        if n > self.R:
            raise _lib.AbmError(f"{n} cells for a capacity of {self.R} rows")
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(self.dev, dt)
        s["locations"][:n] = t(locations, torch.float64)
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, dtype=np.float64), (n,)))
        s["radii"][:n] = t(rad, torch.float64)
        self.uniform_radius = float(rad[0]) if n and np.all(rad == rad[0]) else 0.0
        self.max_radius = max(float(rad.max(initial=0.0)), float(self.P.max_radius))
        s["motility_forces"].zero_()
        for name in INT_FIELDS:
            s[name][:n] = t(state[name], torch.int32)
        s["bond_len"].zero_()
        s["n"].fill_(n)
        self.next_gid.fill_(int(np.max(state["gid"], initial=-1)) + 1)
        self.bonds = (s["bond_len"], s["bond_rows"])

    def _arrays(self, s, bonds=None) -> _lib.ColonyArrays:
        blen, rows = bonds if bonds is not None else (s["bond_len"], s["bond_rows"])
        a = _lib.ColonyArrays()
        for name in ("locations", "radii", "motility_forces", "nbr_start", "nbr_len") + INT_FIELDS:
            setattr(a, name, s[name].data_ptr())
        a.bond_len, a.bond_rows = blen.data_ptr(), rows.data_ptr()
        return a

    def host_state(self) -> dict:
        """Host synchronisation: every agent array of the current colony as numpy, plus the JKR bonds as sorted
        (id, id) pairs and the neighbour graph's degrees."""
        s = self.sets[self.cur]
        n = int(s["n"].item())
        out = {name: s[name][:n].cpu().numpy() for name in ("locations", "radii", "motility_forces") + INT_FIELDS}
        out["n"] = n
        blen, rows = self.bonds
        bl = blen[:n].cpu().numpy().astype(np.int64)
        br = rows[:n].cpu().numpy()
        k = np.arange(br.shape[1])[None, :] < bl[:, None]
        a, b = np.repeat(out["gid"].astype(np.int64), bl), br[k].astype(np.int64)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        out["bond_ids"] = np.unique(np.stack([lo, hi], axis=1), axis=0) if len(lo) else np.zeros((0, 2), np.int64)
        return out

    def check(self):
        """Host synchronisation: raise if a fixed-capacity buffer was too small in a step since the last check."""
        need = int(self.needed[0].item())
        if need > self.R:
            raise _lib.AbmError(f"colony arrays too small: needed {need} rows, have {self.R}")
        used = int(self.nbr_worst.item())
        if used > self.nbr_cols.numel():
            raise _lib.AbmError(f"neighbour column buffer too small: needed {used}, have {self.nbr_cols.numel()}")
        b = self.eng._bond_bufs
        if b is not None and int(b["max_len"][0].item()) > b["stride"]:
            raise _lib.AbmError(f"bond rows too narrow: needed {int(b['max_len'][0].item())}, have {b['stride']}")

    # ---- one step, enqueue only -----------------------------------------------------------------------------------
    def enqueue_step(self, step_number: int):
        eng, lib, P, R = self.eng, self.eng.lib, self.P, self.R
        eng.use_current_stream()
        P.step = int(step_number)
        s, o = self.sets[self.cur], self.sets[1 - self.cur]
        before = eng.launches
        # get_neighbors("neighbor_graph", 15)                                             cell_simulation.py:90
        eng.set_device_state(s["locations"], s["radii"], s["motility_forces"], max_radius=self.max_radius,
                             uniform_radius=self.uniform_radius)
        eng.n_device = s["n"]
        eng.global_id = s["gid"]
        srt = eng.sort(self.search_radius)
        eng.enqueue_search(srt, self.search_radius, None, s["nbr_start"], s["nbr_len"], self.nbr_cols, self.nbr_cursor)
        torch.maximum(self.nbr_worst, self.nbr_cursor, out=self.nbr_worst)
        # cell_division, cell_death                                                        cell_simulation.py:94-95
        _lib.check(lib.abm_rule_division(C.byref(self._arrays(s, self.bonds)), R, self.stride, s["n"].data_ptr(),
                                         self.next_gid.data_ptr(), C.byref(P), self.needed.data_ptr(),
                                         self.ws.data_ptr(), self.ws.numel(), eng.stream), "abm_rule_division")
        _lib.check(lib.abm_rule_death(C.byref(self._arrays(s, self.bonds)), C.byref(self._arrays(o)), R, self.stride,
                                      s["n"].data_ptr(), o["n"].data_ptr(), C.byref(P), self.index_map.data_ptr(),
                                      self.ws.data_ptr(), self.ws.numel(), eng.stream), "abm_rule_death")
        # cell_pathway, cell_differentiate, cell_motility                                 cell_simulation.py:96-108
        _lib.check(lib.abm_rule_update(C.byref(self._arrays(o)), R, o["n"].data_ptr(), self.nbr_cols.data_ptr(),
                                       self.index_map.data_ptr(), C.byref(P), self.fgf4_scratch.data_ptr(),
                                       eng.stream), "abm_rule_update")
        eng.launches += 10
        # apply_forces                                                                     cell_simulation.py:112
        eng.set_device_state(o["locations"], o["radii"], o["motility_forces"], max_radius=self.max_radius,
                             uniform_radius=self.uniform_radius)
        eng.n_device = o["n"]
        eng.global_id = o["gid"]
        old = GraphTensors(R, None, self.stride, o["bond_len"], o["bond_rows"], True)
        bonds = eng.enqueue_mechanics(old, self.contact_radius, self.dts, self.stride, STOKES, pair=[0, 1],
                                      zero_flag=False)
        o["motility_forces"].zero_()                       # cell_methods.py:439
        self.bonds = (bonds.row_len, bonds.cols)
        self.cur = 1 - self.cur
        self.steps_done += 1
        self.launches_per_step = eng.launches - before

    def step(self, step_number: int):
        self.enqueue_step(step_number)
