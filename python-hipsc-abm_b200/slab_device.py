"""Slab decomposition with every count on the device (VERDICT r01 "slab step without a host round trip").

`slab.py` describes the decomposition (geometry, halo depth, the generic rank_step used by the CPU/gloo tests).  This
module is the production path on GPUs.  One rank keeps its cells in FIXED-CAPACITY device arrays and all counts
(owned cells, local cells = owned + halo, message rows) in device memory; every kernel of the path reads its cell
count from there (`n_device` of the C-ABI).  A step therefore needs no host synchronisation at all, its launch
parameters never change, and it is replayed as CUDA graphs:

    main stream                                              side stream
    ---------------------------------------------------     -------------------------------------------------
    [G pack]      cells within `halo` of an interface -> fixed-capacity messages (row count in a trailing header)
    NCCL          one send + one receive per x-neighbour
    [G merge]     owned | halo partition into the other buffer set (migration = ownership by position),
                  snapshot of the positions                   [G deposit] morphogen deposit on the snapshot
                                                              NCCL   morphogen rows (own communicator)
    [G mechanics] 11 x (sort + fused substep) on owned+halo   [G side] absorb rows, r=15 search on the snapshot,
                  cells, no communication                               diffusion substeps, per-cell sampling

The morphogen stage and the neighbour search work on the positions the step STARTS with (the reference's order:
update_diffusion before apply_forces, cell_simulation.py:101-112), which is what makes them independent of the
mechanics.  Capacities are checked after the fact: the kernels record what they needed, `check()` raises.
"""
from __future__ import annotations

from typing import Optional

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import ColonyEngine, GraphTensors, STOKES
from .slab import FieldSlab, SlabGeometry


class DeviceSlab:
    def __init__(self, geom: SlabGeometry, cfg, device, rows: int, send_cap: int, max_radius=5.0, uniform_radius=0.0,
                 field: Optional[FieldSlab] = None):
        """rows: capacity of every per-cell array (owned + halo cells of this rank, with headroom);
        send_cap: rows of a cell message -- the SAME value on every rank (the receiver posts a buffer of that size)."""
        self.geom, self.cfg, self.dev = geom, cfg, torch.device(device)
        self.eng = ColonyEngine(geom.size, device)
        self.eng.set_bin_lattice(*geom.local_box())
        self.R, self.send_cap, self.stride = int(rows), int(send_cap), int(cfg.bond_stride)
        self.rec = 64 + 4 * self.stride
        self.dts = cfg.substep_dts()
        self.max_radius, self.uniform_radius = float(max_radius), float(uniform_radius)
        f64, i32, dev, R = torch.float64, torch.int32, self.dev, self.R

        def cells():
            return dict(loc=torch.zeros((R, 3), dtype=f64, device=dev), radii=torch.zeros(R, dtype=f64, device=dev),
                        mot=torch.zeros((R, 3), dtype=f64, device=dev), gid=torch.zeros(R, dtype=i32, device=dev),
                        blen=torch.zeros(R, dtype=i32, device=dev),
                        rows=torch.zeros((R, self.stride), dtype=i32, device=dev),
                        n_own=torch.zeros(1, dtype=i32, device=dev), n_local=torch.zeros(1, dtype=i32, device=dev))
        self.sets = [cells(), cells()]
        self.cur = 0
        self.bonds = (self.sets[0]["blen"], self.sets[0]["rows"])        # where the current bonds live
        u8 = torch.uint8
        # a message = send_cap records + one trailing header row whose first int32 is the number of valid records: the
        # count travels inside the message, so an exchange is ONE send and ONE receive per neighbour
        self.send = [torch.zeros((self.send_cap + 1, self.rec), dtype=u8, device=dev) for _ in range(2)]
        self.recv = [torch.zeros((self.send_cap + 1, self.rec), dtype=u8, device=dev) for _ in range(2)]
        self.pack_counters = torch.zeros(3, dtype=i32, device=dev)       # kernel scratch, left zeroed by every call
        self.tally = torch.zeros(5, dtype=i32, device=dev)               # likewise (abm_slab_exchange_merge)
        self.worst = torch.zeros(3, dtype=i32, device=dev)               # largest message counts, largest n_local
        # the kernels record the largest substep displacement over the ghost zones only (see allowed_substep_move)
        self.eng.move_window = torch.tensor([geom.x_lo, geom.x_hi], dtype=f64, device=dev)
        need = int(self.eng.lib.abm_slab_workspace_bytes(R + 2 * self.send_cap))
        self.ws = torch.empty(need + 4096, dtype=u8, device=dev)
        self.snap = torch.zeros((R, 3), dtype=f64, device=dev)
        cap = int(cfg.search_cols_per_cell * R) + 1024
        self.nbr = dict(row_start=torch.zeros(R, dtype=i32, device=dev), row_len=torch.zeros(R, dtype=i32, device=dev),
                        cols=torch.empty(cap, dtype=i32, device=dev), cursor=torch.zeros(1, dtype=i32, device=dev),
                        worst=torch.zeros(1, dtype=i32, device=dev))
        self.neighbor_graph = None
        self.field = field
        self.sampled = torch.zeros(R, dtype=f64, device=dev)
        if field is not None:
            self.amounts = torch.full((R,), cfg.deposit_amount, dtype=f64, device=dev)
            k, gy = field.halo_rows, field.gy
            self.rows_out = [torch.zeros((k, gy), dtype=f64, device=dev) for _ in range(2)]
            self.rows_in = [torch.zeros((k, gy), dtype=f64, device=dev) for _ in range(2)]
        self.launches = 0
        self.loops_enqueued = 0          # mechanics loops since the cells were loaded (see enqueue_mechanics)

    # ---- set-up / read-out (these may synchronise) --------------------------------------------------------------
    def load(self, loc, radii, mot, gid):
        """The cells this rank owns at the start (device or host tensors / arrays)."""
        s = self.sets[self.cur]
        n = int(loc.shape[0])
        if n > self.R:
            raise _lib.AbmError(f"{n} cells for a capacity of {self.R} rows")
        t = lambda a, dt: torch.as_tensor(a).to(self.dev, dt)
        s["loc"][:n] = t(loc, torch.float64)
        s["radii"][:n] = t(radii, torch.float64)
        s["mot"][:n] = t(mot, torch.float64)
        s["gid"][:n] = t(gid, torch.int32)
        s["blen"].zero_()
        s["n_own"].fill_(n)
        s["n_local"].fill_(n)
        self.loops_enqueued = 0
        self.bonds = (s["blen"], s["rows"])

    def owned(self):
        """Host synchronisation: the owned cells as tensors (loc, radii, mot, gid, bond_len, bond_rows)."""
        s = self.sets[self.cur]
        n = int(s["n_own"].item())
        blen, rows = self.bonds
        return dict(loc=s["loc"][:n], radii=s["radii"][:n], mot=s["mot"][:n], gid=s["gid"][:n], bond_len=blen[:n],
                    bond_gid=rows[:n], n=n)

    def negate_motility(self):
        self.sets[self.cur]["mot"].neg_()          # all rows: the halo rows behind the owned ones are never sent

    def check(self, halo=True):
        """Host synchronisation: raise if any fixed-capacity buffer was too small in a step since the last check, or
        (halo=True) if the halo depth's assumption on cell motion was violated."""
        w = [int(v) for v in self.worst.tolist()]
        if max(w[0], w[1]) > self.send_cap:
            raise _lib.AbmError(f"cell message capacity too small: needed {max(w[0], w[1])} rows, have {self.send_cap}")
        if w[2] > self.R:
            raise _lib.AbmError(f"cell arrays too small: needed {w[2]} rows, have {self.R}")
        used = int(self.nbr["worst"].item())
        if used > self.nbr["cols"].numel():
            raise _lib.AbmError(f"neighbour column buffer too small: needed {used}")
        b = self.eng._bond_bufs
        if b is not None and int(b["max_len"][0].item()) > b["stride"]:
            raise _lib.AbmError(f"bond rows too narrow: needed {int(b['max_len'][0].item())}, have {b['stride']}")
        if halo:
            self.check_halo()

    def allowed_substep_move(self) -> float:
        """The halo's assumption (slab.required_halo).  The cells a rank holds are all exact when a step starts; a
        ghost within `reach` of the halo's outer edge misses partners, so after the first substep it may be wrong.
        A wrong cell can spoil cells within `reach` of it in the next force evaluation, and these then move by at
        most m: the wrong zone deepens by at most reach + m per substep, m = the largest substep displacement of a
        cell IN THE GHOST ZONES (the kernel's move window).  Owned cells stay exact while substeps x (reach + m)
        (+ one grid pitch when ghosts deposit morphogen) <= halo."""
        reach = self.eng.jkr_reach(self.cfg.contact_radius)
        return (self.geom.halo - (self.field.h if self.field is not None else 0.0)) / len(self.dts) - reach

    def check_halo(self):
        """Host synchronisation: raise if a ghost-zone cell moved further in one substep than the halo depth allows,
        in any step since reset_monitor()."""
        move, allowed = self.eng.max_substep_move(), self.allowed_substep_move()
        if self.geom.world > 1 and move > allowed * (1 + 1e-9):
            raise _lib.AbmError(f"halo too shallow for the observed motion: a ghost-zone cell moved {move:.3f} um in "
                                f"one substep, the halo of {self.geom.halo:.1f} um allows {allowed:.3f} um "
                                f"(slab.required_halo)")

    def reset_monitor(self):
        """Forget the displacement seen so far (e.g. after the violent first steps of a random start, whose results
        are checked otherwise -- tests compare them with the single-domain run)."""
        b = self.eng._bond_bufs
        if b is not None:
            b["max_len"][1:].zero_()

    # ---- the phases of a step: each only ENQUEUES on the current stream ---------------------------------------------
    def enqueue_pack(self):
        eng, g, s = self.eng, self.geom, self.sets[self.cur]
        eng.use_current_stream()
        blen, rows = self.bonds
        _lib.check(eng.lib.abm_slab_exchange_pack(
            s["loc"].data_ptr(), s["radii"].data_ptr(), s["mot"].data_ptr(), s["gid"].data_ptr(), blen.data_ptr(),
            rows.data_ptr(), self.R, s["n_own"].data_ptr(), self.stride, g.x_lo + g.halo, g.x_hi - g.halo,
            int(g.has_left), int(g.has_right), self.send[0].data_ptr(), self.send[1].data_ptr(), self.send_cap,
            self.pack_counters.data_ptr(), self.worst.data_ptr(), eng.stream), "abm_slab_exchange_pack")
        self.launches += 1

    def enqueue_merge(self):
        """Owned | halo partition of (current owned cells + received records) into the other buffer set, and the
        snapshot of the positions the step starts with (two launches)."""
        eng, g = self.eng, self.geom
        eng.use_current_stream()
        s, o = self.sets[self.cur], self.sets[1 - self.cur]
        blen, rows = self.bonds
        _lib.check(eng.lib.abm_slab_exchange_merge(
            s["loc"].data_ptr(), s["radii"].data_ptr(), s["mot"].data_ptr(), s["gid"].data_ptr(), blen.data_ptr(),
            rows.data_ptr(), self.R, s["n_own"].data_ptr(), self.recv[0].data_ptr(), self.recv[1].data_ptr(),
            self.send_cap, self.stride, g.x_lo, g.x_hi, int(g.has_left), int(g.has_right), o["loc"].data_ptr(),
            o["radii"].data_ptr(), o["mot"].data_ptr(), o["gid"].data_ptr(), o["blen"].data_ptr(),
            o["rows"].data_ptr(), self.snap.data_ptr(), self.R, o["n_own"].data_ptr(), o["n_local"].data_ptr(),
            self.tally.data_ptr(), self.worst[2:3].data_ptr(), eng.stream), "abm_slab_exchange_merge")
        self.launches += 2
        self.cur = 1 - self.cur
        self.bonds = (o["blen"], o["rows"])
        eng.set_device_state(o["loc"], o["radii"], o["mot"], max_radius=self.max_radius,
                             uniform_radius=self.uniform_radius)
        eng.n_device = o["n_local"]
        eng.global_id = o["gid"]

    def enqueue_deposit(self):
        """Morphogen deposit of the local cells at the snapshot positions, and the rows the x-neighbours need."""
        if self.field is None:
            return
        eng, f = self.eng, self.field
        eng.use_current_stream()
        eng.n_device = self.sets[self.cur]["n_local"]
        eng.deposit(f.data, self.amounts, f.h, rows=(f.ext_lo, f.ext_hi), grid_rows=f.gx, loc=self.snap, n=self.R)
        self.launches += 1
        if f.has_left:
            self.rows_out[0].copy_(f.owned()[:f.halo_rows])
        if f.has_right:
            own = f.owned()
            self.rows_out[1].copy_(own[own.shape[0] - f.halo_rows:])

    def enqueue_mechanics(self):
        eng, cfg, s = self.eng, self.cfg, self.sets[self.cur]
        eng.use_current_stream()
        before = eng.launches
        old = GraphTensors(self.R, None, self.stride, s["blen"], s["rows"], True)
        # flat world: once every rank has run a loop, every cell a rank holds -- its own, migrants, ghosts -- has been
        # through an integration that clamped its z to 0, so the first substep can run the two-axis kernel too
        eng.flat_z_vouched = self.loops_enqueued >= 1 and len(self.dts) > 0
        bonds = eng.enqueue_mechanics(old, cfg.contact_radius, self.dts, self.stride, STOKES, pair=[0, 1],
                                      zero_flag=False)
        self.loops_enqueued += 1
        self.bonds = (bonds.row_len, bonds.cols)
        self.launches += eng.launches - before

    def enqueue_side(self):
        """Everything that depends only on the step's starting positions (the snapshot): the r = search_radius
        neighbour graph the rule methods consume, and the diffusion + per-cell sampling of the morphogen field
        (whose halo rows have arrived in rows_in)."""
        eng, cfg = self.eng, self.cfg
        eng.use_current_stream()
        eng.n_device = self.sets[self.cur]["n_local"]
        before = eng.launches
        srt = eng.sort(cfg.search_radius, loc=self.snap)
        self.neighbor_graph = eng.enqueue_search(srt, cfg.search_radius, None, self.nbr["row_start"],
                                                 self.nbr["row_len"], self.nbr["cols"], self.nbr["cursor"],
                                                 loc=self.snap)
        torch.maximum(self.nbr["worst"], self.nbr["cursor"], out=self.nbr["worst"])
        if self.field is not None:
            f = self.field
            f.absorb(self.rows_in[0], self.rows_in[1])
            last = cfg.diffuse_dt if cfg.diffuse_last_dt is None else cfg.diffuse_last_dt
            out = eng.diffuse(f.data, cfg.diffuse_steps, cfg.diffuse_dt, last, cfg.diffuse_const, f.h ** 2,
                              cfg.max_concentration, cfg.degradation, use_tma=cfg.use_tma)
            if out.data_ptr() != f.data.data_ptr():
                f.data.copy_(out)
            eng.sample(f.data, f.h, rows=(f.ext_lo, f.ext_hi), grid_rows=f.gx, loc=self.snap, n=self.R,
                       out=self.sampled)
        self.launches += eng.launches - before


class _Phase:
    """One phase of the step as either a direct call or a pair of captured CUDA graphs (one per buffer parity)."""

    def __init__(self, fn):
        self.fn, self.graphs = fn, None

    def run(self, parity):
        if self.graphs is None:
            self.fn()
        else:
            self.graphs[parity].replay()


class SlabStepper:
    """Drives one DeviceSlab (one rank).  exchange_cells / exchange_rows are supplied by the driver below (NCCL) or
    by the loopback emulation (several ranks on one GPU)."""

    def __init__(self, slab: DeviceSlab, side_stream=True):
        self.slab = slab
        # (the pack stays at the START of a step: the host may have given the cells new motility forces since the
        # previous one, and the records carry them)
        self.phases = dict(pack=_Phase(slab.enqueue_pack), merge=_Phase(slab.enqueue_merge),
                           deposit=_Phase(slab.enqueue_deposit), mechanics=_Phase(slab.enqueue_mechanics),
                           side=_Phase(slab.enqueue_side))
        self.side = torch.cuda.Stream(device=slab.dev) if side_stream else None
        self.side_done = None
        self.captured = False
        self.steps = 0
        self.timeline = None          # set to [] to record (name, start event, end event) of every main-stream phase

    def _timed(self, name, fn):
        if self.timeline is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self.timeline.append((name, a, b))
        return out

    def timeline_us(self):
        """Host synchronisation: mean device time per recorded main-stream phase, in microseconds."""
        torch.cuda.synchronize(self.slab.dev)
        acc = {}
        for name, a, b in self.timeline or []:
            acc.setdefault(name, []).append(a.elapsed_time(b) * 1e3)
        return {k: float(np.mean(v)) for k, v in acc.items()}

    def capture(self):
        """Capture the four phases for both buffer parities.  Call after at least two eager steps (every scratch
        buffer then exists) and with an even number of steps done."""
        s = self.slab
        assert s.cur == 0 and s.bonds[0].data_ptr() != s.sets[0]["blen"].data_ptr(), \
            "capture() needs an even number (>= 2) of eager steps first"
        torch.cuda.synchronize(s.dev)
        launches, cur, bonds = s.launches, s.cur, s.bonds
        eng_launches = s.eng.launches
        stream = torch.cuda.Stream(device=s.dev)
        order = self._capture_order()
        graphs = {k: [None, None] for k in order}
        for parity in range(2):
            for name in order:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    self.phases[name].fn()
                graphs[name][parity] = g
        assert s.cur == cur
        s.bonds, s.launches, s.eng.launches = bonds, launches, eng_launches
        s.eng.use_current_stream()
        self._per_step = None
        for name, g in graphs.items():
            self.phases[name].graphs = g
        self.captured = True

    def _capture_order(self):
        """The phases capture() records, in the order a step runs them (each one's Python bookkeeping sets the stage
        for the next)."""
        return ("pack", "merge", "deposit", "mechanics", "side")

    # The bookkeeping a captured phase would have done in Python (which set is current, where the bonds live).
    def _after_merge(self):
        s = self.slab
        if self.captured:
            s.cur = 1 - s.cur
            o = s.sets[s.cur]
            s.bonds = (o["blen"], o["rows"])

    def _after_mechanics(self):
        s = self.slab
        if self.captured:
            b = s.eng._bond_bufs
            k = (len(s.dts) - 1) & 1
            s.bonds = (b["len"][k], b["cols"][k])

    def run(self, exchange_cells, exchange_rows):
        """One step.  exchange_cells() / exchange_rows() enqueue the transfers of this rank on the CURRENT stream
        (cells on the main stream, rows on the side stream)."""
        s = self.slab
        main = torch.cuda.current_stream(s.dev)
        parity = s.cur
        self._front(parity, main, exchange_cells)
        self._after_merge()
        side = self.side or main
        if self.side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            self.phases["deposit"].run(parity)
            if s.field is not None:
                exchange_rows()
        self._timed("mechanics", lambda: self.phases["mechanics"].run(parity))
        self._after_mechanics()
        with torch.cuda.stream(side):
            self.phases["side"].run(parity)
            if self.side is not None:
                self.side_done = torch.cuda.Event()
                self.side_done.record(side)
        s.eng.use_current_stream()
        self.steps += 1

    def _front(self, parity, main, exchange_cells):
        """pack -> cell exchange -> merge on the main stream."""
        self._timed("pack", lambda: self.phases["pack"].run(parity))
        self._timed("exchange_cells", exchange_cells)
        if self.side_done is not None:
            main.wait_event(self.side_done)       # last step's side work read the snapshot the merge overwrites
        self._timed("merge", lambda: self.phases["merge"].run(parity))

    def join(self):
        """Make the current stream wait for the side stream's work of the last step."""
        if self.side_done is not None:
            torch.cuda.current_stream(self.slab.dev).wait_event(self.side_done)


class DistSlab(SlabStepper):
    """One rank over torch.distributed (NCCL): cells and counts in one send/recv group with both x-neighbours, the
    morphogen rows on their own communicator (`rows_group`) so that they never queue behind the cell exchange."""

    def __init__(self, slab: DeviceSlab, rank, world, rows_group=None, side_stream=True, left_group=None):
        """left_group: a second communicator for the cell records that travel towards LOWER ranks.  NCCL runs the
        send/recv pairs of one group in rounds by rank distance (+1, then -1): with both directions in one group an
        interior rank's exchange takes two latency-bound rounds (72 us at 8 GPUs against 37 us at 2).  On their own
        communicator the left-going records travel beside the right-going ones."""
        super().__init__(slab, side_stream)
        import torch.distributed as dist
        self.dist, self.rank, self.world, self.rows_group = dist, rank, world, rows_group
        self.left_group = left_group
        self.bytes_sent = 0
        self.exchanges = 0

    def _post(self, ops):
        for w in self.dist.batch_isend_irecv(ops) if ops else []:
            w.wait()

    def _cells(self):
        s, d = self.slab, self.dist
        lo, hi = self.rank - 1, self.rank + 1
        if self.left_group is None:
            ops = []
            for side, peer in ((0, lo), (1, hi)):
                if 0 <= peer < self.world:
                    ops += [d.P2POp(d.isend, s.send[side], peer), d.P2POp(d.irecv, s.recv[side], peer)]
                    self.bytes_sent += s.send[side].numel()
            self._post(ops)
        else:
            right, left = [], []             # records travelling to rank+1 (default group) / to rank-1 (left_group)
            if hi < self.world:
                right.append(d.P2POp(d.isend, s.send[1], hi))
                left.append(d.P2POp(d.irecv, s.recv[1], hi, group=self.left_group))
                self.bytes_sent += s.send[1].numel()
            if lo >= 0:
                right.append(d.P2POp(d.irecv, s.recv[0], lo))
                left.append(d.P2POp(d.isend, s.send[0], lo, group=self.left_group))
                self.bytes_sent += s.send[0].numel()
            works = (d.batch_isend_irecv(right) if right else []) + (d.batch_isend_irecv(left) if left else [])
            for w in works:
                w.wait()
        self.exchanges += 1

    def _rows(self):
        s, d = self.slab, self.dist
        ops = []
        for side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
            if 0 <= peer < self.world:
                kw = {} if self.rows_group is None else dict(group=self.rows_group)
                ops += [d.P2POp(d.isend, s.rows_out[side], peer, **kw), d.P2POp(d.irecv, s.rows_in[side], peer, **kw)]
                self.bytes_sent += s.rows_out[side].numel() * 8
        self._post(ops)
        self.exchanges += 1

    def step(self):
        self.run(self._cells, self._rows)


class PeerSlab(DistSlab):
    """DistSlab whose CELL exchange goes over NVLink peer memory instead of NCCL send/recv (one node, one process per
    GPU).  At set-up every rank maps its neighbours' receive buffers and flag words through CUDA IPC (handles exported
    by torch's allocator, opened by abm_ipc_open with THIS rank's device current).  Per step ONE launch stores the valid
    records of both messages straight into the neighbours' receive buffers and releases the step number into their
    "data" flag words (abm_peer_exchange); the rank's own stream then waits until both neighbours' step numbers have
    arrived in ITS data words (abm_peer_flags: a poll of local memory), merges, and releases the step number into the
    neighbours' "ack" words -- which a sender polls before it overwrites a buffer at the next step.  No host
    synchronisation, no rendezvous: NCCL needs two latency-bound send/recv rounds for the same exchange (37 us at 2
    GPUs, 72 us at 8).  Once captured, pack -> exchange -> merge -> acknowledgements replay as ONE CUDA graph (the
    step number then comes from a device counter the graph increments itself).  The morphogen rows stay on NCCL (side
    stream, hidden behind the mechanics).  A neighbour that never answers makes a poll give up after `max_wait_s`
    (generous: a neighbour's host may sit in a lazy NCCL communicator set-up for seconds during the first steps) and
    check_peers() raise, instead of hanging the GPU."""

    FLAG_DATA, FLAG_ACK = 0, 2          # flags[FLAG_DATA + side]: step of the last message received from that side;
                                        # flags[FLAG_ACK + side]: last step whose message the neighbour on that side merged

    def __init__(self, slab: DeviceSlab, rank, world, rows_group=None, side_stream=True, max_wait_s=30.0):
        super().__init__(slab, rank, world, rows_group, side_stream)
        dev = slab.dev
        self.flags = torch.zeros(8, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(2, dtype=torch.int32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)     # exchanges done, counted on the device
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.max_wait_ns = int(max_wait_s * 1e9)
        self.peer_step = 0
        torch.cuda.synchronize(dev)
        self.peer = {}                       # side -> {name: address of the neighbour's tensor, mapped into this process}
        self.ok, self.why = True, ""

        def export(t):
            # (device, 64-byte cudaIpcMemHandle_t of the allocation, ..., offset of the storage inside it, ...)
            shared = t.untyped_storage()._share_cuda_()
            handle = bytes(shared[1])
            if len(handle) == 66 and handle[1:2] == b"c":    # torch >= 2.3: version byte, type byte ('c' = cudaMalloc)
                handle = handle[2:]
            if len(handle) != 64:
                raise _lib.AbmError(f"unexpected CUDA IPC handle of {len(handle)} bytes (expandable segments?)")
            return handle, int(shared[3]) + t.storage_offset() * t.element_size()
        try:
            mine = {"recv0": export(slab.recv[0]), "recv1": export(slab.recv[1]), "flags": export(self.flags)}
        except Exception as e:               # (the collective below must still be entered by every rank)
            mine, self.ok, self.why = None, False, repr(e)
        everyone = [None] * world
        self.dist.all_gather_object(everyone, mine)
        try:
            opened = {}                      # an allocation is opened once, whatever number of tensors live in it
            for side, other in ((0, rank - 1), (1, rank + 1)):
                if 0 <= other < world:
                    self.peer[side] = {}
                    for name, (handle, offset) in everyone[other].items():
                        if handle not in opened:
                            base = C.c_void_p()
                            _lib.check(slab.eng.lib.abm_ipc_open(handle, len(handle), C.byref(base)), "abm_ipc_open")
                            opened[handle] = int(base.value)
                        self.peer[side][name] = opened[handle] + offset
        except Exception as e:
            self.ok, self.why = False, repr(e)
        # every rank must take the same path: the peer exchange is used only if ALL ranks mapped their neighbours
        agreed = torch.tensor([1 if self.ok else 0], dtype=torch.int32, device=dev)
        self.dist.all_reduce(agreed, op=self.dist.ReduceOp.MIN)
        self.ok = bool(agreed.item())

    # A flag value is `value` + (*step_dev if on_device): eager steps pass the step number from the host, the captured
    # front graph reads it from the device counter it increments itself (a graph's kernel arguments are frozen).
    def _flag(self, index):
        return self.flags[index:].data_ptr()

    def _exchange(self, k, on_device):
        """Both messages into the neighbours' buffers + their data flags (one launch), then wait for the neighbours'."""
        s, eng = self.slab, self.slab.eng
        eng.use_current_stream()
        step = self.step_dev.data_ptr() if on_device else 0
        args = []
        for side in (0, 1):
            p, there = self.peer.get(side), 1 - side     # my left neighbour receives from ITS right, and vice versa
            args += ([p["recv%d" % there], s.send[side].data_ptr(), p["flags"] + 4 * (self.FLAG_DATA + there),
                      self._flag(self.FLAG_ACK + side)] if p else [0, 0, 0, 0])
        _lib.check(eng.lib.abm_peer_exchange(*args, s.send_cap, s.rec, int(k), step, self.counters.data_ptr(),
                                             self.max_wait_ns, self.timed_out.data_ptr(), eng.stream),
                   "abm_peer_exchange")
        waits = [self._flag(self.FLAG_DATA + side) if side in self.peer else 0 for side in (0, 1)]
        _lib.check(eng.lib.abm_peer_flags(waits[0], waits[1], 0, 0, int(k), step, self.max_wait_ns,
                                          self.timed_out.data_ptr(), eng.stream), "abm_peer_flags")

    def _acks(self, k, on_device):
        """My receive buffers may be overwritten now: the step number into the neighbours' ack words."""
        eng = self.slab.eng
        eng.use_current_stream()
        rel = [self.peer[side]["flags"] + 4 * (self.FLAG_ACK + 1 - side) if side in self.peer else 0 for side in (0, 1)]
        _lib.check(eng.lib.abm_peer_flags(0, 0, rel[0], rel[1], int(k), self.step_dev.data_ptr() if on_device else 0,
                                          self.max_wait_ns, self.timed_out.data_ptr(), eng.stream), "abm_peer_flags")

    def _cells(self):                                            # (eager steps)
        self.peer_step += 1
        self._exchange(self.peer_step, False)
        self._count_exchange()

    def _count_exchange(self):
        for side in self.peer:
            self.bytes_sent += self.slab.send[side].numel()
        self.exchanges += 1

    def _after_merge(self):
        super()._after_merge()
        if not self.captured:
            self._acks(self.peer_step, False)

    # ---- captured: pack -> exchange -> merge -> acknowledgements as ONE graph -------------------------------------
    def _front_fn(self):
        s = self.slab
        self.step_dev.add_(1)
        s.enqueue_pack()
        self._exchange(0, True)
        s.enqueue_merge()
        self._acks(0, True)

    def _capture_order(self):
        self.phases["front"] = _Phase(self._front_fn)
        self.step_dev.fill_(self.peer_step)      # (executed now, not captured: the device counter takes over from here)
        return ("front", "deposit", "mechanics", "side")

    def _front(self, parity, main, exchange_cells):
        if not self.captured:
            return super()._front(parity, main, exchange_cells)
        if self.side_done is not None:
            main.wait_event(self.side_done)       # (long finished: it ran beside the previous step's mechanics)
        self._timed("pack+exchange+merge", lambda: self.phases["front"].run(parity))
        self.peer_step += 1
        self._count_exchange()

    def check_peers(self):
        """Host synchronisation: raise if a wait for a neighbour ever gave up."""
        k = int(self.timed_out.item())
        if k:
            raise _lib.AbmError(f"a neighbour's cell message or acknowledgement had not arrived after "
                                f"{self.max_wait_ns / 1e9:.0f} s (first at exchange {k})")


def loopback_step(steppers):
    """The same step for several ranks emulated on ONE device (tests): the exchanges are device copies between the
    ranks' message buffers.  The ranks advance phase by phase in lock-step, which is what a collective implies."""
    world = len(steppers)
    slabs = [st.slab for st in steppers]
    parities = [s.cur for s in slabs]
    for st, p in zip(steppers, parities):
        st.phases["pack"].run(p)
    for k, s in enumerate(slabs):
        if k > 0:
            s.recv[0].copy_(slabs[k - 1].send[1])
        if k < world - 1:
            s.recv[1].copy_(slabs[k + 1].send[0])
    for st, p in zip(steppers, parities):
        st.phases["merge"].run(p)
        st._after_merge()
        st.phases["deposit"].run(p)
    for k, s in enumerate(slabs):
        if s.field is not None:
            if k > 0:
                s.rows_in[0].copy_(slabs[k - 1].rows_out[1])
            if k < world - 1:
                s.rows_in[1].copy_(slabs[k + 1].rows_out[0])
    for st, p in zip(steppers, parities):
        st.phases["mechanics"].run(p)
        st._after_mechanics()
        st.phases["side"].run(p)
        st.slab.eng.use_current_stream()
        st.steps += 1
