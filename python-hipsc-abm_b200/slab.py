"""1-D slab decomposition of a colony across ranks (one process per GPU), SURVEY 8(e).

The reference has no multi-device path; this is the partitioning the north star asks for, designed around
the B200 pipeline rather than ported from anything:

  * rank k owns the cells with x in [bounds[k], bounds[k+1]);
  * ONCE per simulation step every rank sends each x-neighbour the cells it holds within `halo` of their
    common interface -- on either side of it -- as one packed message (position, radius, motility force,
    global id, JKR bond row).  A receiver keeps the cells that now lie inside its own slab as OWNED (that is the
    migration: ownership is a function of position, evaluated on the same bits by both sides) and the rest as
    read-only halo cells; the sender demotes the cells that left its slab to halo cells for this step;
  * the whole step (neighbour search + the 11 fused mechanics substeps) then runs on owned + halo cells with
    the unmodified single-GPU kernels and NO communication.  A halo cell is simulated redundantly; the error
    that creeps in from the outer edge of the halo travels at most (reach + displacement) per substep, so
    with halo >= substeps * (reach + slack) it never reaches an owned cell;
  * the morphogen rows next to the interface are exchanged once, after the deposit and before the S_d
    diffusion substeps (same argument, one grid row per substep).

So a step costs two neighbour exchanges instead of one per substep; over NVLink they are latency-bound
messages of a few hundred KB.  JKR bond rows are kept by GLOBAL id everywhere (the fused kernel takes the id
table, include/abm_b200.h `global_id`), so nothing is translated when cells change rank.

Everything here is device-agnostic torch tensor plumbing (CPU + gloo in the tests, CUDA + NCCL in bench.py);
the compute callback is the CUDA pipeline (pipeline.SlabPipeline).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Any, Callable, Optional

import numpy as np
import torch


@dataclass
class SlabGeometry:
    size: np.ndarray          # global box
    world: int
    rank: int
    bounds: np.ndarray        # (world+1,) x interfaces, bounds[0] = 0, bounds[-1] = size[0]
    halo: float

    def __post_init__(self):
        # a slab thinner than the halo would need halo cells from its SECOND neighbour, which nobody sends: its
        # owned cells would silently get wrong forces
        if self.world > 1 and float(np.min(np.diff(self.bounds))) < self.halo:
            raise ValueError(f"slab thinner than the halo ({float(np.min(np.diff(self.bounds))):.3f} < "
                             f"{self.halo:.3f} um): use fewer slabs or a shallower halo")

    @staticmethod
    def uniform(size, world, rank, halo):
        size = np.asarray(size, dtype=np.float64)
        return SlabGeometry(size, world, rank, np.linspace(0.0, size[0], world + 1), float(halo))

    @staticmethod
    def balanced(size, world, rank, halo, x_all):
        """Equal-count slabs: the interfaces are the k/world quantiles of the cells' x coordinates (x_all: 1-D
        array-like of ALL cells' x, or of a representative sample; every rank must pass the same values, e.g.
        after one all-gather at set-up).  A slab is never made thinner than the halo (its halo cells must all
        come from its direct neighbours)."""
        size = np.asarray(size, dtype=np.float64)
        x = np.sort(np.asarray(x_all, dtype=np.float64).ravel())
        inner = np.quantile(x, np.arange(1, world) / world) if world > 1 and x.size else np.zeros(0)
        bounds = np.concatenate([[0.0], inner, [size[0]]])
        width = float(halo) * (1.0 + 1e-9)
        for k in range(1, world):                                   # enforce the minimum width, left to right
            bounds[k] = min(max(bounds[k], bounds[k - 1] + width), size[0] - width * (world - k))
        if np.any(np.diff(bounds) < float(halo)):
            raise ValueError("the box is too short for this many slabs at least one halo wide")
        return SlabGeometry(size, world, rank, bounds, float(halo))

    @property
    def x_lo(self):
        return float(self.bounds[self.rank])

    @property
    def x_hi(self):
        return float(self.bounds[self.rank + 1])

    @property
    def has_left(self):
        return self.rank > 0

    @property
    def has_right(self):
        return self.rank < self.world - 1

    def inside(self, x: torch.Tensor) -> torch.Tensor:
        """Ownership predicate: the outermost slabs also own whatever lies beyond the box."""
        m = torch.ones_like(x, dtype=torch.bool)
        if self.has_left:
            m &= x >= self.x_lo
        if self.has_right:
            m &= x < self.x_hi
        return m

    def local_box(self):
        """(origin, extent) of the region the local bins must cover: the slab plus its halos."""
        lo = max(self.x_lo - self.halo, 0.0) if self.has_left else 0.0
        hi = min(self.x_hi + self.halo, self.size[0]) if self.has_right else float(self.size[0])
        return np.array([lo, 0.0, 0.0]), np.array([hi - lo, self.size[1], self.size[2]])


def required_halo(substeps, reach, slack=2.0, spat_res=0.0):
    """Halo depth that keeps owned cells exact for `substeps` communication-free substeps: the influence of a
    wrong halo-edge cell advances by at most reach + slack per substep (slack >= twice the largest
    per-substep displacement; SlabPipeline.check() verifies that assumption against the displacement the mechanics
    kernel records).  spat_res: grid pitch of a carried morphogen field -- halo cells deposit on owned grid rows
    up to one pitch away, so their positions must be exact that much deeper."""
    return float(substeps) * (float(reach) + float(slack)) + float(spat_res)


FIELDS = ("loc", "radii", "mot", "gid", "bond_len", "bond_gid")


class SlabState:
    """Cells owned by one rank, in the kernels' own formats: float64 positions / radii / motility forces,
    int32 global ids, and fixed-stride int32 bond rows holding global ids."""

    def __init__(self, loc, radii, mot, gid, stride=16):
        n = loc.shape[0]
        dev = loc.device
        self._d = dict(loc=loc.contiguous(), radii=radii.contiguous(), mot=mot.contiguous(),
                       gid=gid.to(torch.int32).contiguous(),
                       bond_len=torch.zeros(n, dtype=torch.int32, device=dev),
                       bond_gid=torch.full((n, stride), -1, dtype=torch.int32, device=dev))
        self._pending = None

    # The owned cells of a finished step are the first n_own rows of the step's local arrays; when n_own is still
    # on the device (LocalProblem with a lazy count) the slicing is deferred until somebody needs it, so that the
    # next step's pack can be enqueued without waiting for the GPU.
    def defer(self, full: dict, local: "LocalProblem"):
        self._pending = (full, local)

    def pending(self):
        """(full arrays, rows, device owned-count) if the owned count has not been read yet, else None."""
        if self._pending is None or isinstance(self._pending[1]._n_own, int):
            return None
        full, local = self._pending
        return full, int(full["loc"].shape[0]), local._n_own

    def negate_motility(self):
        """Flip the sign of the owned cells' motility forces in place (bench.py's stationary synthetic colony).
        While the owned count is still on the device the state IS the finished step's local arrays (owned rows
        first, halo rows behind them, which the next pack ignores): those are negated as a whole, so that nothing
        waits for the count."""
        pend = self.pending()
        (pend[0]["mot"] if pend is not None else self._get("mot")).neg_()

    def _settle(self):
        if self._pending is not None:
            full, local = self._pending
            n = local.n_own
            self._d = {f: full[f][:n] for f in FIELDS}
            self._pending = None

    def _get(self, f):
        self._settle()
        return self._d[f]

    def _set(self, f, v):
        self._settle()
        self._d[f] = v

    @property
    def n(self):
        return int(self._get("loc").shape[0])

    @property
    def stride(self):
        d = self._pending[0] if self._pending is not None else self._d
        return int(d["bond_gid"].shape[1])

    def record_bytes(self):
        return 24 + 8 + 24 + 4 + 4 + 4 * self.stride

    def fields(self):
        return {f: getattr(self, f) for f in FIELDS}

    def set_fields(self, d):
        self._pending = None
        self._d = {f: d[f] for f in FIELDS}

    def pack(self, idx) -> torch.Tensor:
        """One byte record per selected cell: loc | radius | motility | gid | bond_len | bond row."""
        k = idx.numel()
        parts = [self.loc[idx].view(torch.uint8).reshape(k, 24), self.radii[idx].view(torch.uint8).reshape(k, 8),
                 self.mot[idx].view(torch.uint8).reshape(k, 24), self.gid[idx].view(torch.uint8).reshape(k, 4),
                 self.bond_len[idx].view(torch.uint8).reshape(k, 4),
                 self.bond_gid[idx].contiguous().view(torch.uint8).reshape(k, 4 * self.stride)]
        return torch.cat(parts, dim=1)


def unpack(msg: torch.Tensor, stride: int):
    k = msg.shape[0]

    def take(a, b, dtype, shape):
        return msg[:, a:b].contiguous().view(dtype).reshape(shape)
    return dict(loc=take(0, 24, torch.float64, (k, 3)), radii=take(24, 32, torch.float64, (k,)),
                mot=take(32, 56, torch.float64, (k, 3)), gid=take(56, 60, torch.int32, (k,)),
                bond_len=take(60, 64, torch.int32, (k,)), bond_gid=take(64, 64 + 4 * stride, torch.int32, (k, stride)))


# ---- exchange fabrics ---------------------------------------------------------------------------------------------
@dataclass
class Exchange:
    """What rank_step asks its driver for: send to_left / to_right (2-D, `width` columns of `dtype`) to the
    x-neighbours and hand back what they sent.
      fixed_rows: both sides know the row count (no size handshake);
      counts:     int32[2] on the device = rows of to_left / to_right that are valid.  The messages are then
                  full-capacity buffers (fixed_rows = capacity) and the counts travel in the same transfer group,
                  so there is neither a handshake nor a host synchronisation; the reply is
                  (from_left, from_right, counts_in)."""
    tag: str
    to_left: Any
    to_right: Any
    width: int
    dtype: Any
    fixed_rows: Optional[int] = None
    counts: Any = None


class DistFabric:
    """Neighbour exchange over torch.distributed (NCCL on GPUs, gloo on CPU): the row counts first, then one
    packed message per neighbour and direction, all transfers of a phase posted together."""

    def __init__(self, rank, world, device, rows_group=None):
        """rows_group: optional second process group (its own communicator and stream) for the fixed-size field
        row exchange, so that it never queues in front of -- or behind -- the cell exchange of the main stream."""
        import torch.distributed as dist
        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        self.rows_group = rows_group
        self.bytes_sent = 0
        self.exchanges = 0
        self._recv = {}

    def _post(self, ops):
        for w in self.dist.batch_isend_irecv(ops) if ops else []:
            w.wait()

    def _p2p(self, fn, tensor, peer, group=None):
        if group is None:
            return self.dist.P2POp(fn, tensor, peer)
        return self.dist.P2POp(fn, tensor, peer, group=group)

    def exchange(self, to_left, to_right, width, dtype=torch.float64, fixed_rows=None, counts=None):
        """Send `to_left` / `to_right` (2-D, `width` columns) to the x-neighbours, return what they sent.
        fixed_rows: both sides know the row count (field rows) -> no size handshake.
        counts: see Exchange (full-capacity messages + their valid row counts in one transfer group)."""
        dist, dev = self.dist, self.device
        left, right = self.rank - 1, self.rank + 1
        has_l, has_r = left >= 0, right < self.world
        if counts is not None:
            cap = int(fixed_rows)
            key = (cap, width, dtype)
            if self._recv.get("key") != key:      # receive buffers are reused: the consumer (merge) is stream-ordered
                self._recv = dict(key=key,        # before the next exchange
                                  left=torch.empty((cap if has_l else 0, width), dtype=dtype, device=dev),
                                  right=torch.empty((cap if has_r else 0, width), dtype=dtype, device=dev),
                                  counts=torch.zeros(2, dtype=counts.dtype, device=dev))
            from_left, from_right, counts_in = self._recv["left"], self._recv["right"], self._recv["counts"]
            ops = []
            if has_l:
                ops += [dist.P2POp(dist.isend, to_left, left), dist.P2POp(dist.irecv, from_left, left),
                        dist.P2POp(dist.isend, counts[0:1], left), dist.P2POp(dist.irecv, counts_in[0:1], left)]
            if has_r:
                ops += [dist.P2POp(dist.isend, to_right, right), dist.P2POp(dist.irecv, from_right, right),
                        dist.P2POp(dist.isend, counts[1:2], right), dist.P2POp(dist.irecv, counts_in[1:2], right)]
            self._post(ops)
            self.bytes_sent += (int(has_l) + int(has_r)) * cap * width * to_left.element_size() if cap else 0
            self.exchanges += 1
            return from_left, from_right, counts_in
        if fixed_rows is None:
            counts_out = torch.tensor([to_left.shape[0] if has_l else 0, to_right.shape[0] if has_r else 0],
                                      dtype=torch.int64, device=dev)
            counts_in = torch.zeros(2, dtype=torch.int64, device=dev)
            ops = []
            if has_l:
                ops += [dist.P2POp(dist.isend, counts_out[0:1], left), dist.P2POp(dist.irecv, counts_in[0:1], left)]
            if has_r:
                ops += [dist.P2POp(dist.isend, counts_out[1:2], right),
                        dist.P2POp(dist.irecv, counts_in[1:2], right)]
            self._post(ops)
            n_l, n_r = (int(v) for v in counts_in.tolist())
        else:
            n_l, n_r = (fixed_rows if has_l else 0), (fixed_rows if has_r else 0)
        from_left = torch.empty((n_l, width), dtype=dtype, device=dev)
        from_right = torch.empty((n_r, width), dtype=dtype, device=dev)
        grp = self.rows_group if fixed_rows is not None else None
        ops = []
        if has_l and to_left.shape[0]:
            ops.append(self._p2p(dist.isend, to_left, left, grp))
        if has_l and n_l:
            ops.append(self._p2p(dist.irecv, from_left, left, grp))
        if has_r and to_right.shape[0]:
            ops.append(self._p2p(dist.isend, to_right, right, grp))
        if has_r and n_r:
            ops.append(self._p2p(dist.irecv, from_right, right, grp))
        self._post(ops)
        self.bytes_sent += sum(t.numel() * t.element_size() for t, ok in ((to_left, has_l), (to_right, has_r)) if ok)
        self.exchanges += 1
        return from_left, from_right


class LoopbackWorld:
    """All ranks in one process (tests / several slabs emulated on one GPU): exchange is a list shuffle."""

    def __init__(self, world):
        self.world = world

    def exchange_all(self, requests):
        """requests[k] = Exchange of rank k -> incoming[k] = (from_left, from_right[, counts_in])."""
        incoming = []
        for k in range(self.world):
            from_left = requests[k - 1].to_right if k > 0 else None
            from_right = requests[k + 1].to_left if k < self.world - 1 else None
            if requests[k].counts is not None:
                c = torch.zeros_like(requests[k].counts)
                if k > 0:
                    c[0] = requests[k - 1].counts[1]
                if k < self.world - 1:
                    c[1] = requests[k + 1].counts[0]
                incoming.append((from_left, from_right, c))
            else:
                incoming.append((from_left, from_right))
        return incoming


# ---- morphogen rows -----------------------------------------------------------------------------------------------------
class FieldSlab:
    """Rows [row_lo,row_hi) of the global (gx,gy) morphogen grid owned by one rank, stored with `halo_rows`
    extra rows on each interior side.  Grid row r sits at x = r*h and belongs to the slab containing it.
    halo_rows must be (diffusion substeps + 1): the stencil invalidates one halo row per substep from the
    outside in, and a cell next to the interface may sample the neighbour's first row."""

    def __init__(self, geom: SlabGeometry, grid, spat_res, halo_rows, device):
        gx, gy = int(grid[0]), int(grid[1])
        h = float(spat_res)
        self.gx, self.gy, self.h, self.halo_rows = gx, gy, h, int(halo_rows)
        self.row_lo = int(np.ceil(geom.x_lo / h)) if geom.has_left else 0
        self.row_hi = min(int(np.ceil(geom.x_hi / h)), gx) if geom.has_right else gx
        self.ext_lo = max(self.row_lo - self.halo_rows, 0) if geom.has_left else 0
        self.ext_hi = min(self.row_hi + self.halo_rows, gx) if geom.has_right else gx
        if self.row_hi - self.row_lo < self.halo_rows:
            raise ValueError("slab narrower than the diffusion halo")
        self.has_left, self.has_right = geom.has_left, geom.has_right
        self.data = torch.zeros((self.ext_hi - self.ext_lo, gy), dtype=torch.float64, device=device)

    def owned(self) -> torch.Tensor:
        return self.data[self.row_lo - self.ext_lo:self.row_hi - self.ext_lo]

    def messages(self):
        o = self.owned()
        to_left = o[:self.halo_rows].clone() if self.has_left else None          # a copy, never a view
        to_right = o[o.shape[0] - self.halo_rows:].clone() if self.has_right else None
        return to_left, to_right

    def absorb(self, from_left, from_right):
        if self.has_left:
            k = self.row_lo - self.ext_lo
            self.data[:k] = from_left[from_left.shape[0] - k:]
        if self.has_right:
            k = self.ext_hi - self.row_hi
            self.data[self.data.shape[0] - k:] = from_right[:k]


# ---- one step of one rank ---------------------------------------------------------------------------------------------------
class LocalProblem:
    """Owned (first n_own) + halo cells of one rank for one step; what the single-GPU pipeline runs on.
    n_own may be given as a one-element device tensor: it is then read (one host synchronisation) only when
    somebody asks for it -- at the end of the step, after the whole step has been enqueued."""

    def __init__(self, n_own, loc, radii, mot, gid, bond_len, bond_gid):
        self._n_own = n_own
        self.loc, self.radii, self.mot = loc, radii, mot
        self.gid = gid                  # int32 (n_local,)
        self.bond_len = bond_len        # int32 (n_local,)
        self.bond_gid = bond_gid        # int32 (n_local, stride) global ids

    @property
    def n_own(self) -> int:
        if not isinstance(self._n_own, int):
            self._n_own = int(self._n_own.item())
        return self._n_own


for _f in FIELDS:
    setattr(SlabState, _f, property(lambda self, f=_f: self._get(f), lambda self, v, f=_f: self._set(f, v)))


class TorchSlabOps:
    """Selecting / packing / merging with plain torch tensor ops: runs anywhere (the CPU + gloo tests use it) but
    costs a few dozen small kernels per step; on a GPU use pipeline.NativeSlabOps (csrc/slab.cu)."""

    def outgoing(self, state: SlabState, geom: SlabGeometry):
        x = state.loc[:, 0]
        empty = torch.zeros(0, dtype=torch.int64, device=x.device)
        idx_l = torch.nonzero(x < geom.x_lo + geom.halo).flatten() if geom.has_left else empty
        idx_r = torch.nonzero(x >= geom.x_hi - geom.halo).flatten() if geom.has_right else empty
        return (state.pack(idx_l) if geom.has_left else None, state.pack(idx_r) if geom.has_right else None)

    def merge(self, state: SlabState, geom: SlabGeometry, from_l, from_r) -> LocalProblem:
        stride = state.stride
        x = state.loc[:, 0]
        parts = [m for m in (from_l, from_r) if m is not None and m.shape[0]]
        recv = unpack(torch.cat(parts, dim=0), stride) if parts else None
        stay = geom.inside(x)
        if recv is not None:
            mine = geom.inside(recv["loc"][:, 0])
            n_leave, n_arrive = (int(v) for v in torch.stack([(~stay).sum(), mine.sum()]).tolist())
        else:
            mine, n_leave, n_arrive = None, int((~stay).sum()), 0
        own = state.fields()
        if n_leave == 0 and n_arrive == 0:
            groups = [own] + ([recv] if recv is not None else [])
            n_own = state.n
        else:
            # migration: arrivals become owned, departures stay around as halo cells for this step
            keep, gone = torch.nonzero(stay).flatten(), torch.nonzero(~stay).flatten()
            groups = [{f: own[f][keep] for f in FIELDS}]
            n_own = keep.numel() + n_arrive
            if recv is not None:
                arr, rest = torch.nonzero(mine).flatten(), torch.nonzero(~mine).flatten()
                groups += [{f: recv[f][arr] for f in FIELDS}, {f: recv[f][rest] for f in FIELDS}]
            groups.append({f: own[f][gone] for f in FIELDS})
        cat = {f: (torch.cat([g[f] for g in groups], dim=0) if len(groups) > 1 else groups[0][f]) for f in FIELDS}
        return LocalProblem(n_own, cat["loc"], cat["radii"], cat["mot"], cat["gid"], cat["bond_len"],
                            cat["bond_gid"])


def rank_step(state: SlabState, geom: SlabGeometry, compute: Callable[[LocalProblem], tuple],
              field: Optional[FieldSlab] = None, field_ops=None, ops=None, morphogen_first=False):
    """One simulation step of one rank, written as a generator: it yields
    (tag, (to_left, to_right), width, dtype, fixed_rows) whenever it needs a neighbour exchange and is sent
    (from_left, from_right) back.  `compute(local)` runs the mechanics in place on local.loc and returns the new
    (bond_len, bond_gid) of all local cells.  `field_ops` (if a morphogen field is carried) provides
    deposit(local, field), diffuse(field) and sample(local, field).  `ops` selects/packs/merges the cells
    (TorchSlabOps by default, pipeline.NativeSlabOps on a GPU).

    morphogen_first: the morphogen stage works on the positions the step starts with (after the exchange) --
    the reference's order, update_diffusion before apply_forces (cell_simulation.py:101-112).  It is then
    independent of the mechanics: field_ops.begin_side(local) may return a second stream on which the stage
    (including its row exchange) runs beside them, field_ops.end_side() joins it."""
    ops = ops or TorchSlabOps()
    if hasattr(field_ops, "step_begin"):
        field_ops.step_begin()
    out = ops.outgoing(state, geom)
    if len(out) == 3:                                         # full-capacity buffers + device counts, no handshake
        reply = yield Exchange("cells", out[0], out[1], state.record_bytes(), torch.uint8,
                               fixed_rows=ops.send_cap, counts=out[2])
        local = ops.merge(state, geom, reply[0], reply[1], reply[2])
    else:
        from_l, from_r = yield Exchange("cells", out[0], out[1], state.record_bytes(), torch.uint8)
        local = ops.merge(state, geom, from_l, from_r)

    if field is not None and morphogen_first:
        begin = getattr(field_ops, "begin_side", None)
        side = begin(local) if begin else None                # snapshot of the positions, fork point
        if side is not None:
            blen, bgid = compute(local)                       # the mechanics are enqueued first: the GPU starts on
            if hasattr(field_ops, "after_compute"):
                field_ops.after_compute()
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):   # them while the host
            field_ops.deposit(local, field)                   # feeds the second stream.  owned + halo cells deposit
            to_l, to_r = field.messages()
            from_l, from_r = yield Exchange("rows", to_l, to_r, field.gy, torch.float64, fixed_rows=field.halo_rows)
            if side is not None and hasattr(field_ops, "search"):
                field_ops.search(local)                       # the neighbour search (same snapshot) while the rows
            field.absorb(from_l, from_r)                      # travel
            field_ops.diffuse(field)
            field_ops.sample(local, field)
        if side is None:
            blen, bgid = compute(local)
        if begin:
            field_ops.end_side()
        field = None                                          # done
    else:
        blen, bgid = compute(local)
    if field is not None:
        field_ops.deposit(local, field)                       # owned + halo cells at their new positions
        to_l, to_r = field.messages()
        from_l, from_r = yield Exchange("rows", to_l, to_r, field.gy, torch.float64, fixed_rows=field.halo_rows)
        field.absorb(from_l, from_r)
        field_ops.diffuse(field)                              # S_d substeps, no communication
        field_ops.sample(local, field)
    full = dict(loc=local.loc, radii=local.radii, mot=local.mot, gid=local.gid, bond_len=blen, bond_gid=bgid)
    if isinstance(local._n_own, int):
        n = local.n_own
        state.set_fields({f: full[f][:n] for f in FIELDS})
    else:
        state.defer(full, local)                              # owned count still on the device: slice when needed
    return local


def slab_step(state: SlabState, geom: SlabGeometry, fabric: DistFabric, compute, field=None, field_ops=None,
              ops=None, morphogen_first=False):
    """Drive rank_step over torch.distributed."""
    gen = rank_step(state, geom, compute, field, field_ops, ops, morphogen_first)
    incoming = None
    try:
        while True:
            req = gen.send(incoming) if incoming is not None else next(gen)
            incoming = fabric.exchange(req.to_left, req.to_right, req.width, req.dtype, req.fixed_rows, req.counts)
    except StopIteration as done:
        return done.value


def loopback_step(states, geoms, computes, fields=None, field_ops=None, ops=None, morphogen_first=False):
    """The same step for all ranks emulated in one process (tests; several slabs on one GPU)."""
    world = LoopbackWorld(len(states))
    if callable(computes):
        computes = [computes] * len(states)
    fields = fields or [None] * len(states)
    field_ops = field_ops or [None] * len(states)
    ops = ops or [None] * len(states)
    gens = [rank_step(s, g, c, f, o, p, morphogen_first)
            for s, g, c, f, o, p in zip(states, geoms, computes, fields, field_ops, ops)]
    requests = [next(g) for g in gens]
    results = [None] * len(gens)
    while any(r is not None for r in requests):
        assert all(r is not None for r in requests), "ranks fell out of lock-step"
        incoming = world.exchange_all(requests)
        nxt = []
        for k, (g, inc) in enumerate(zip(gens, incoming)):
            try:
                nxt.append(g.send(inc))
            except StopIteration as done:
                results[k] = done.value
                nxt.append(None)
        requests = nxt
    return results
