"""Device-resident simulation step: the whole hot path of one reference step enqueued without any host
synchronisation, optionally captured in a CUDA graph.

One step =
    1 x neighbour search at r_search (get_neighbors("neighbor_graph", 15), cell_simulation.py:90)
  + S_m x fused mechanics substep (CellMethods.apply_forces, cell_methods.py:399-436)
  + deposit -> S_d diffusion substeps -> sample (update_diffusion + the per-cell coupling, cell_methods.py:441-521)

Capacities (column buffer of the search, bond-row stride) are fixed while a captured graph is alive; the
kernels report the size they needed and `check()` raises if anything was truncated, so a run is never
silently wrong.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from .engine import ColonyEngine, GraphTensors, STOKES


@dataclass
class StepConfig:
    search_radius: float = 15.0          # cell_simulation.py:90
    contact_radius: float = 10.0         # 2 * max_radius, cell_methods.py:401
    step_dt: float = 1800.0              # cell_simulation.py:34
    move_dt: float = 180.0               # cell_simulation.py:35
    # morphogen field (None -> no diffusion stage)
    grid: Optional[tuple] = None         # (gx, gy): the reference's 2-D field; (gx, gy, gz): the 3-D extension
    spat_res: float = 7.07106
    diffuse_const: float = 50.0
    diffuse_dt: float = 0.24
    diffuse_steps: int = 10
    diffuse_last_dt: Optional[float] = None
    max_concentration: float = 2.0
    degradation: float = 0.1
    deposit_amount: float = 1.0
    bond_stride: int = 16
    search_cols_per_cell: float = 12.0   # column-buffer capacity of the r_search graph, entries per cell
    use_tma: Optional[bool] = None
    # morphogen stage (deposit -> diffuse -> sample) on the positions the step STARTS with -- where the reference
    # calls update_diffusion, before cell_motility/apply_forces (cell_simulation.py:101-112) -- so that it is
    # independent of the step's search + mechanics and runs beside them on a second stream.  False: after the
    # mechanics, on the moved positions, on the same stream.
    overlap_morphogen: bool = False

    def substep_dts(self):
        steps, last_dt = divmod(self.step_dt, self.move_dt)     # cell_methods.py:395-396
        return [self.move_dt] * int(steps) + [last_dt]


class HotPathStep:
    def __init__(self, engine: ColonyEngine, cfg: StepConfig, motility_by_parity=None):
        """motility_by_parity: optional pair of (n,3) float64 device tensors; even steps (counted from the creation
        of this object) then use the first as the motility forces, odd steps the second, instead of engine.mot.
        The reference draws new motility forces every step (cell_methods.py:236-340); a synthetic colony that kept
        one set for hundreds of steps would pile its cells up against the walls, whereas forces that alternate in
        sign keep it statistically stationary (bench.py)."""
        self.eng = engine
        self.cfg = cfg
        self.motility_by_parity = None
        if motility_by_parity is not None:
            a, b = motility_by_parity
            if tuple(a.shape) != (engine.n, 3) or tuple(b.shape) != (engine.n, 3):
                raise ValueError("motility_by_parity: two (n,3) tensors expected")
            self.motility_by_parity = (a, b)
        _ = engine.max_radius, engine.uniform_radius      # evaluated now (device read-back), not inside a capture
        n = engine.n
        dev = engine.device
        self.dts = cfg.substep_dts()
        if len(self.dts) % 2 == 0:
            # the bond rows alternate between two buffers from step to step, which needs an odd substep count
            # (the reference's 1800/180 -> 11); checked here, before anything is enqueued
            raise ValueError(f"HotPathStep needs an odd number of mechanics substeps, got {len(self.dts)} "
                             f"(step_dt={cfg.step_dt}, move_dt={cfg.move_dt}); use engine.mechanics() for this case")
        cap = int(cfg.search_cols_per_cell * max(n, 1)) + 1024
        self.nbr = dict(row_start=torch.empty(max(n, 1), dtype=torch.int32, device=dev),
                        row_len=torch.zeros(max(n, 1), dtype=torch.int32, device=dev),
                        cols=torch.empty(cap, dtype=torch.int32, device=dev),
                        cursor=torch.zeros(1, dtype=torch.int32, device=dev),
                        worst=torch.zeros(1, dtype=torch.int32, device=dev))
        self.neighbor_graph: Optional[GraphTensors] = None
        self.bonds: Optional[GraphTensors] = None
        self.parity = 0
        self.field = self.scratch = self.amounts = self.sampled = None
        self.volume = cfg.grid is not None and len(cfg.grid) == 3     # 3-D field: 7-point stencil, 8-corner deposit
        if cfg.grid is not None:
            self.field = torch.zeros(tuple(int(g) for g in cfg.grid), dtype=torch.float64, device=dev)
            self.scratch = torch.empty_like(self.field)
            self.amounts = torch.full((max(n, 1),), cfg.deposit_amount, dtype=torch.float64, device=dev)
            # one persistent result buffer: every captured graph writes the same tensor (a buffer allocated during a
            # capture would belong to that graph alone)
            self.sampled = torch.zeros(max(n, 1), dtype=torch.float64, device=dev)
            if cfg.diffuse_steps % 2:
                raise ValueError("HotPathStep keeps the field in place: use an even number of diffusion substeps")
        self.graphs = None
        self.launches_per_step = 0
        self._side = self._loc_snap = None
        if cfg.overlap_morphogen:
            self._side = torch.cuda.Stream(device=dev)
            self._loc_snap = torch.empty((max(n, 1), 3), dtype=torch.float64, device=dev)
        # start with an empty bond graph living in row buffer 0
        b = engine._bond_buffers(n, cfg.bond_stride)
        b["len"][0].zero_()
        b["max_len"].zero_()
        self.bonds = GraphTensors(n, None, cfg.bond_stride, b["len"][0], b["cols"][0], True)

    # ---- one step, enqueue only ---------------------------------------------------------------------------
    def enqueue(self):
        eng, cfg = self.eng, self.cfg
        eng.use_current_stream()              # under torch.cuda.graph() this is the capture stream
        before = eng.launches
        overlap = cfg.overlap_morphogen
        if overlap:
            # fork: the neighbour search and the morphogen stage work on a snapshot of the positions and run on a
            # second stream beside the mechanics (all three depend only on the positions the step starts with)
            main = torch.cuda.current_stream(eng.device)
            snap = self._loc_snap[:eng.n]
            snap.copy_(eng.loc)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                eng.use_current_stream()
                self._search(snap)
                if self.field is not None:
                    self._morphogen(snap)
            eng.use_current_stream()
        else:
            self._search(None)
        # mechanics: bonds alternate between row buffers 0 and 1, buffer 2 is the intermediate
        src = self.parity
        if self.motility_by_parity is not None:
            eng.mot = self.motility_by_parity[src]
        pair = [1 - src, 2]
        self.bonds = eng.enqueue_mechanics(self.bonds, cfg.contact_radius, self.dts, cfg.bond_stride, STOKES,
                                           pair=pair, zero_flag=False, carry=True)
        self.parity = 1 - src
        if overlap:
            torch.cuda.current_stream(eng.device).wait_stream(self._side)      # join
        elif self.field is not None:
            self._morphogen(None)
        self.launches_per_step = eng.launches - before

    def _search(self, loc):
        """Neighbour graph (what the rule methods consume) of the positions `loc` (None: the engine's)."""
        eng, cfg = self.eng, self.cfg
        srt = eng.sort(cfg.search_radius, loc=loc)
        self.neighbor_graph = eng.enqueue_search(srt, cfg.search_radius, None, self.nbr["row_start"],
                                                 self.nbr["row_len"], self.nbr["cols"], self.nbr["cursor"], loc=loc)
        torch.maximum(self.nbr["worst"], self.nbr["cursor"], out=self.nbr["worst"])   # sticky overflow record

    def _morphogen(self, loc):
        """deposit -> diffuse -> sample on the engine's current stream (loc=None: the engine's positions)."""
        eng, cfg = self.eng, self.cfg
        eng.deposit(self.field, self.amounts, cfg.spat_res, loc=loc, cube=self.volume)
        last = cfg.diffuse_dt if cfg.diffuse_last_dt is None else cfg.diffuse_last_dt
        diffuse = eng.diffuse3d if self.volume else eng.diffuse
        out = diffuse(self.field, cfg.diffuse_steps, cfg.diffuse_dt, last, cfg.diffuse_const,
                      cfg.spat_res ** 2, cfg.max_concentration, cfg.degradation, use_tma=cfg.use_tma,
                      scratch=self.scratch)
        assert out.data_ptr() == self.field.data_ptr()   # (even substep count, checked in __init__)
        eng.sample(self.field, cfg.spat_res, loc=loc, out=self.sampled)

    def step(self):
        """Eager step (no graph)."""
        self.enqueue()

    # ---- CUDA graphs ------------------------------------------------------------------------------------------
    def capture(self):
        """Capture the two alternating step variants (bond rows in buffer 0 -> 1 and 1 -> 0)."""
        assert self.parity == 0, "capture() must start from an even step"
        torch.cuda.synchronize(self.eng.device)
        # programmatic dependent launches pay off for eager launches; inside a captured graph they were measured
        # slower with the early trigger (profiles/r01b_kernels_c3.md) -- ABM_B200_GRAPH_PDL=1 keeps them for A/B runs
        chained = self.eng.chained_launch
        self.eng.chained_launch = os.environ.get("ABM_B200_GRAPH_PDL", "0") == "1"
        self.graphs = []
        launches = self.eng.launches
        # the mechanics stream outranks the second stream: search + morphogen kernels fill what the mechanics leave idle
        stream = torch.cuda.Stream(device=self.eng.device, priority=-1)
        # a captured step may start from the cells its predecessor handed over (engine.enqueue_mechanics `carry`):
        # remember which hand-over each graph expects and leaves, so that replay() can refuse a broken chain
        self._carry_before, self._carry_after, self._launches = [], [], []
        for _ in range(2):
            g = torch.cuda.CUDAGraph()
            self._carry_before.append(self.eng._carry)
            with torch.cuda.graph(g, stream=stream):
                self.enqueue()
            self.graphs.append(g)
            self._carry_after.append(self.eng._carry)
            self._launches.append(self.launches_per_step)
        self.eng._carry = self._carry_before[0]
        self.eng.chained_launch = chained
        # capturing enqueues nothing on the device: undo the launch accounting, and the state did not advance
        self.eng.use_current_stream()         # leave the capture stream
        self.eng.launches = launches
        self.parity = 0
        b = self.eng._bond_buffers(self.eng.n, self.cfg.bond_stride)
        self.bonds = GraphTensors(self.eng.n, None, self.cfg.bond_stride, b["len"][0], b["cols"][0], True)
        self._replays = 0

    def replay(self):
        k = self._replays & 1
        if self._carry_before[k] is not None and self.eng._carry != self._carry_before[k]:
            raise _lib.AbmError("the cells or the sort workspace changed since the previous step: capture() again")
        self.graphs[k].replay()
        self.eng._carry = self._carry_after[k]
        self._replays += 1
        self.launches_per_step = self._launches[k]
        self.eng.launches += self.launches_per_step

    # ---- capacity check ------------------------------------------------------------------------------------------
    def check(self):
        """Host sync: raise if any fixed-capacity buffer was too small during the steps run so far."""
        eng = self.eng
        used = eng.read_i32(self.nbr["worst"])
        if used > self.nbr["cols"].numel():
            raise _lib.AbmError(f"neighbour column buffer too small: needed {used}, have {self.nbr['cols'].numel()}")
        worst = eng.read_i32(eng._bond_bufs["max_len"])
        if worst > self.cfg.bond_stride:
            raise _lib.AbmError(f"bond rows too narrow: needed {worst}, have {self.cfg.bond_stride}")
        return used, worst

    def current_bonds(self) -> GraphTensors:
        b = self.eng._bond_bufs
        k = (self._replays & 1) if self.graphs is not None else self.parity
        return GraphTensors(self.eng.n, None, self.cfg.bond_stride, b["len"][k], b["cols"][k], True)


class SlabPipeline:
    """The device-side work of ONE rank of a slab-decomposed colony (slab.py): the single-GPU kernels run on
    the rank's owned + halo cells, binned over the slab's own box.  Use with slab.slab_step / loopback_step:
    `compute` is the mechanics callback, and the object itself is the `field_ops` for the morphogen slab."""

    def __init__(self, geom, cfg: StepConfig, device, max_radius=5.0, uniform_radius=0.0, side_stream=True,
                 trailing=True):
        """uniform_radius: the common radius if ALL cells of the colony have the same one (selects the kernel
        variant with the contact constants folded in), else 0.  side_stream: run the neighbour search and the
        morphogen stage of a morphogen_first step on a second stream (off for several slabs emulated on one GPU,
        whose loopback exchange does not synchronise streams).  trailing: the second stream starts its work of a
        step when that step's mechanics have finished, so that it fills the next step's exchange window (pack,
        transfer, host round trip, merge -- when the main stream has little to run) instead of competing with the
        mechanics for the SMs; results of a step (sampled, neighbor_graph, the field) are complete after
        join_side()."""
        self.geom, self.cfg = geom, cfg
        self.eng = ColonyEngine(geom.size, device)
        self.eng.chained_launch = True     # eager launches: chain the mechanics kernels (this engine only)
        self.eng.set_bin_lattice(*geom.local_box())
        self.dts = cfg.substep_dts()
        self.max_radius = float(max_radius)
        self.uniform_radius = float(uniform_radius)
        self.stride = cfg.bond_stride
        self._nbr = None
        self.neighbor_graph = None
        self._sampled_all = self._sampled_local = None
        self._side = torch.cuda.Stream(device=device) if side_stream else None
        self.trailing = bool(trailing)
        self._snap = None
        self._snaps = [None, None]          # two position snapshots: the second stream may still read the older one
        self._side_done = [None, None]      # event after the second stream's work of the step that used snapshot k
        self._turn = 0
        self._amounts = None
        self._defer_search = False

    @property
    def sampled(self):
        """Concentrations sampled by the OWNED cells in the last step (reads the owned count if still pending)."""
        if self._sampled_all is None:
            return None
        self.join_side()
        return self._sampled_all[:self._sampled_local.n_own]

    def join_side(self):
        """Make the current stream wait for everything enqueued on the second stream."""
        if self._side is not None:
            torch.cuda.current_stream(self.eng.device).wait_stream(self._side)

    def compute(self, local):
        """Mechanics of one step on the rank's owned + halo cells (slab.LocalProblem): positions are advanced
        in place, the new bond rows (global ids) are returned.  Nothing here waits for the device: buffer
        overflows are recorded on the device and raised by check(), which the caller runs at its next natural
        synchronisation point."""
        eng, cfg = self.eng, self.cfg
        n = int(local.loc.shape[0])
        eng.set_device_state(local.loc, local.radii, local.mot, max_radius=self.max_radius,
                             uniform_radius=self.uniform_radius)
        eng.global_id = local.gid
        # neighbour graph of the step (rows of owned cells are complete: the halo is deeper than the radius)
        cap = int(cfg.search_cols_per_cell * max(n, 1)) + 1024
        if self._nbr is None or self._nbr["row_start"].numel() < n or self._nbr["cols"].numel() < cap:
            dev = eng.device
            grow = int(1.1 * n) + 1024
            self._nbr = dict(row_start=torch.empty(grow, dtype=torch.int32, device=dev),
                             row_len=torch.zeros(grow, dtype=torch.int32, device=dev),
                             cols=torch.empty(int(1.1 * cap), dtype=torch.int32, device=dev),
                             cursor=torch.zeros(1, dtype=torch.int32, device=dev),
                             worst=torch.zeros(1, dtype=torch.int32, device=dev))
        if not self._defer_search:
            self._do_search(None)
        # mechanics with the bonds carried in from the previous step (fixed-stride rows of global ids)
        stride = int(local.bond_gid.shape[1])
        old = GraphTensors(n, None, stride, local.bond_len, local.bond_gid, True)
        bonds = eng.enqueue_mechanics(old, cfg.contact_radius, self.dts, stride, STOKES, zero_flag=False)
        return bonds.row_len, bonds.cols

    def _do_search(self, loc):
        eng, cfg = self.eng, self.cfg
        srt = eng.sort(cfg.search_radius, loc=loc)
        self.neighbor_graph = eng.enqueue_search(srt, cfg.search_radius, None, self._nbr["row_start"],
                                                 self._nbr["row_len"], self._nbr["cols"], self._nbr["cursor"], loc=loc)
        torch.maximum(self._nbr["worst"], self._nbr["cursor"], out=self._nbr["worst"])

    def search(self, local):
        """The step's neighbour search on the second stream (morphogen_first steps), on the snapshot."""
        if self._defer_search:
            self.eng.use_current_stream()
            self._do_search(self._snap)
            self._defer_search = False

    def check(self):
        """Host sync: raise if a fixed-capacity buffer overflowed in any step since the last check."""
        eng = self.eng
        self.join_side()
        if self._nbr is not None:
            used = eng.read_i32(self._nbr["worst"])
            if used > self._nbr["cols"].numel():
                raise _lib.AbmError(f"neighbour column buffer too small: needed {used}")
        if eng._bond_bufs is not None:
            worst = eng.read_i32(eng._bond_bufs["max_len"])
            if worst > eng._bond_bufs["stride"]:
                raise _lib.AbmError(f"bond rows too narrow: needed {worst}, have {eng._bond_bufs['stride']}")

    # ---- field_ops protocol of slab.rank_step ---------------------------------------------------------------------
    def step_begin(self):
        """Called at the start of a step (rank_step): the buffers this step overwrites (merge outputs, snapshot)
        were last read by the second stream's work of two steps ago -- long finished, but make it formal."""
        ev = self._side_done[self._turn ^ 1]
        if ev is not None:
            torch.cuda.current_stream(self.eng.device).wait_event(ev)

    def begin_side(self, local):
        """morphogen_first: snapshot the positions (the mechanics will move them) and fork the second stream."""
        eng = self.eng
        n = int(local.loc.shape[0])
        main = torch.cuda.current_stream(eng.device)
        self._turn ^= 1
        if self._side_done[self._turn] is not None:          # the second stream's work of two steps ago: it read this
            main.wait_event(self._side_done[self._turn])      # snapshot and that step's cell arrays (long finished)
        if self._snaps[self._turn] is None or self._snaps[self._turn].shape[0] < n:
            self._snaps[self._turn] = torch.empty((int(1.1 * n) + 1024, 3), dtype=torch.float64, device=eng.device)
        self._snap = self._snaps[self._turn]
        self._snap[:n].copy_(local.loc)
        self._snap_n = n
        self._defer_search = self._side is not None          # the search moves to the second stream as well
        if self._side is None:
            return None
        if not self.trailing:
            self._side.wait_stream(main)
        return self._side

    def after_compute(self):
        """Called when the step's mechanics have been enqueued on the main stream (rank_step)."""
        if self._side is not None and self.trailing:
            self._side.wait_stream(torch.cuda.current_stream(self.eng.device))

    def end_side(self):
        self._snap_n = None
        if self._side is not None:
            ev = torch.cuda.Event()
            ev.record(self._side)
            self._side_done[self._turn] = ev
            if not self.trailing:
                torch.cuda.current_stream(self.eng.device).wait_stream(self._side)
        self.eng.use_current_stream()

    def _positions(self, local):
        """(positions, count) the morphogen stage works on: the snapshot inside a morphogen_first step."""
        if getattr(self, "_snap_n", None) is not None:
            return self._snap, self._snap_n
        return local.loc, int(local.loc.shape[0])

    def deposit(self, local, field):
        eng = self.eng
        eng.use_current_stream()
        loc, n = self._positions(local)
        if self._amounts is None or self._amounts.numel() < n:
            self._amounts = torch.full((int(1.1 * n) + 1024,), self.cfg.deposit_amount, dtype=torch.float64,
                                       device=eng.device)
        eng.deposit(field.data, self._amounts, field.h, rows=(field.ext_lo, field.ext_hi), grid_rows=field.gx,
                    loc=loc, n=n)

    def diffuse(self, field):
        cfg = self.cfg
        self.eng.use_current_stream()
        last = cfg.diffuse_dt if cfg.diffuse_last_dt is None else cfg.diffuse_last_dt
        out = self.eng.diffuse(field.data, cfg.diffuse_steps, cfg.diffuse_dt, last, cfg.diffuse_const,
                               field.h ** 2, cfg.max_concentration, cfg.degradation, use_tma=cfg.use_tma)
        if out.data_ptr() != field.data.data_ptr():
            field.data.copy_(out)

    def sample(self, local, field):
        # every local cell samples (no owned count needed on the host yet); only the owned ones -- whose rows are
        # exact -- are handed out by the `sampled` property
        eng = self.eng
        eng.use_current_stream()
        loc, n = self._positions(local)
        self._sampled_all = eng.sample(field.data, field.h, rows=(field.ext_lo, field.ext_hi), grid_rows=field.gx,
                                       loc=loc, n=n)
        self._sampled_local = local


def _ptr_or_0(t):
    return 0 if t is None else t.data_ptr()


class NativeSlabOps:
    """slab.rank_step's select/pack/merge on the device (csrc/slab.cu): two C calls and two small host reads per
    step instead of a few dozen tensor ops."""

    def __init__(self, engine: ColonyEngine, send_fraction=0.12, send_cap=None):
        """send_cap: rows of the fixed-capacity cell messages.  Every rank must use the SAME value (the receiver
        posts a buffer of that size); None = send_fraction * (cells of the first step) + 4096, which agrees across
        ranks only if they start with equal cell counts."""
        self.eng = engine
        self.lib = engine.lib
        self.send_cap = int(send_cap or 0)
        self.send_fraction = send_fraction
        self.worst = torch.zeros(2, dtype=torch.int32, device=engine.device)   # largest counts seen (overflow check)
        self.ws = None
        self._out = None
        self._pin = None

    def _workspace(self, n):
        need = int(self.lib.abm_slab_workspace_bytes(int(n)))
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(int(1.2 * need) + 4096, dtype=torch.uint8, device=self.eng.device)
        return self.ws

    def outgoing(self, state, geom):
        eng, dev = self.eng, self.eng.device
        eng.use_current_stream()
        pend = state.pending()
        if pend is not None:          # last step's local arrays, owned count still on the device: no host sync
            src, n, n_own_dev = pend
        else:
            src, n, n_own_dev = state.fields(), state.n, None
        stride, rec = state.stride, state.record_bytes()
        ws = self._workspace(n)
        if self.send_cap == 0:
            self.send_cap = int(self.send_fraction * n) + 4096
        # full-capacity messages + their valid row counts on the device: nothing here waits for the GPU (a count
        # beyond the capacity truncates the message and is raised by check())
        cap = self.send_cap
        out_l = torch.empty((cap, rec), dtype=torch.uint8, device=dev)
        out_r = torch.empty((cap, rec), dtype=torch.uint8, device=dev)
        counts = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.check(self.lib.abm_slab_pack(
            src["loc"].data_ptr(), src["radii"].data_ptr(), src["mot"].data_ptr(), src["gid"].data_ptr(),
            src["bond_len"].data_ptr(), src["bond_gid"].data_ptr(), n, _ptr_or_0(n_own_dev), stride,
            geom.x_lo + geom.halo,
            geom.x_hi - geom.halo, int(geom.has_left), int(geom.has_right), out_l.data_ptr(), cap,
            out_r.data_ptr(), cap, counts.data_ptr(), ws.data_ptr(), ws.numel(), eng.stream),
            "abm_slab_pack")
        eng.launches += 1
        torch.maximum(self.worst, counts, out=self.worst)
        return out_l, out_r, counts

    def check(self):
        """Host sync: raise if a cell message was ever truncated."""
        worst = int(self.worst.max().item())
        if worst > self.send_cap:
            raise _lib.AbmError(f"cell message capacity too small: needed {worst} rows, have {self.send_cap} "
                                "(NativeSlabOps(send_cap=...))")

    def _out_buffers(self, total, stride):
        """Two alternating sets of output arrays (the merge is out of place: one set holds the step's inputs)."""
        f64, i32, dev = torch.float64, torch.int32, self.eng.device
        if self._out is None or self._out["cap"] < total or self._out["stride"] != stride:
            cap = int(1.15 * total) + 8192
            self._out = dict(cap=cap, stride=stride, turn=0, sets=[
                dict(loc=torch.empty((cap, 3), dtype=f64, device=dev), radii=torch.empty(cap, dtype=f64, device=dev),
                     mot=torch.empty((cap, 3), dtype=f64, device=dev), gid=torch.empty(cap, dtype=i32, device=dev),
                     blen=torch.empty(cap, dtype=i32, device=dev),
                     rows=torch.empty((cap, stride), dtype=i32, device=dev),
                     n_owned=torch.empty(1, dtype=i32, device=dev)) for _ in range(2)])
        self._out["turn"] ^= 1
        return self._out["sets"][self._out["turn"]]

    def merge(self, state, geom, from_l, from_r, counts_in=None):
        from .slab import LocalProblem
        eng = self.eng
        stride = state.stride
        if counts_in is not None:                 # the one host synchronisation of the exchange: the received counts
            pend = state.pending()                # and (if still pending) last step's owned count in one round trip
            if self._pin is None:
                self._pin = torch.empty(3, dtype=torch.int32).pin_memory()
            self._pin[0:2].copy_(counts_in, non_blocking=True)
            if pend is not None:
                self._pin[2:3].copy_(pend[2], non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
            n_l, n_r, n_prev = (int(v) for v in self._pin.tolist())
            n_l, n_r = min(n_l, self.send_cap), min(n_r, self.send_cap)
            if pend is not None:
                state._pending[1]._n_own = n_prev
        else:
            n_l = int(from_l.shape[0]) if from_l is not None else 0
            n_r = int(from_r.shape[0]) if from_r is not None else 0
        n = state.n
        if n_l and not from_l.is_contiguous():
            from_l = from_l.contiguous()
        if n_r and not from_r.is_contiguous():
            from_r = from_r.contiguous()
        total = n + n_l + n_r
        ws = self._workspace(total)
        o = self._out_buffers(total, stride)
        # the inputs must not live in the set that is written now (they do after a reallocation only)
        assert state.loc.data_ptr() != o["loc"].data_ptr()
        _lib.check(self.lib.abm_slab_merge(
            state.loc.data_ptr(), state.radii.data_ptr(), state.mot.data_ptr(), state.gid.data_ptr(),
            state.bond_len.data_ptr(), state.bond_gid.data_ptr(), n,
            from_l.data_ptr() if n_l else 0, n_l, from_r.data_ptr() if n_r else 0, n_r, stride, geom.x_lo, geom.x_hi,
            int(geom.has_left), int(geom.has_right), o["loc"].data_ptr(), o["radii"].data_ptr(), o["mot"].data_ptr(),
            o["gid"].data_ptr(), o["blen"].data_ptr(), o["rows"].data_ptr(), o["n_owned"].data_ptr(), ws.data_ptr(),
            ws.numel(), 0, 0, 0, 0, eng.stream), "abm_slab_merge")
        eng.launches += 4
        # owned count read when first needed
        return LocalProblem(o["n_owned"], o["loc"][:total], o["radii"][:total], o["mot"][:total], o["gid"][:total],
                            o["blen"][:total], o["rows"][:total])
