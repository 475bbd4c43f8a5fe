"""1-D slab decomposition of a colony across ranks (one process per GPU), SURVEY 8(e).

The reference has no multi-device path; this is the partitioning the north star asks for, designed around
the B200 pipeline rather than ported from anything:

  * rank k owns the cells with x in [bounds[k], bounds[k+1]);
  * ONCE per simulation step it receives, from each x-neighbour, a halo of the neighbour's cells (position,
    radius, motility force, global id and JKR bond row) lying within `halo` of the interface;
  * the whole step (neighbour search + the 11 fused mechanics substeps) then runs on owned + halo cells with
    the unmodified single-GPU kernels and NO communication: a halo cell is simulated redundantly, and the
    error that creeps in from the outer edge of the halo travels at most (reach + displacement) per substep,
    so with halo >= substeps * (reach + slack) it never reaches an owned cell;
  * after the step the halo is dropped, and owned cells that crossed an interface migrate with their state.

So a step costs three neighbour exchanges (halo, migration, and the morphogen rows) instead of one per
substep; over NVLink they are latency-bound messages of a few hundred KB.

Everything here is device-agnostic torch tensor plumbing (it runs on CPU with gloo in the tests and on CUDA
with NCCL in bench.py); the compute callback is the CUDA pipeline.
Bond rows cross rank boundaries as GLOBAL ids and are translated to local indices for the kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch


@dataclass
class SlabGeometry:
    size: np.ndarray          # global box
    world: int
    rank: int
    bounds: np.ndarray        # (world+1,) x interfaces, bounds[0] = 0, bounds[-1] = size[0]
    halo: float

    @staticmethod
    def uniform(size, world, rank, halo):
        size = np.asarray(size, dtype=np.float64)
        return SlabGeometry(size, world, rank, np.linspace(0.0, size[0], world + 1), float(halo))

    @property
    def x_lo(self):
        return float(self.bounds[self.rank])

    @property
    def x_hi(self):
        return float(self.bounds[self.rank + 1])

    def local_box(self):
        """(origin, extent) of the region the local bins must cover: the slab plus its halos."""
        lo = max(self.x_lo - self.halo, 0.0) if self.rank > 0 else 0.0
        hi = min(self.x_hi + self.halo, self.size[0]) if self.rank < self.world - 1 else float(self.size[0])
        origin = np.array([lo, 0.0, 0.0])
        extent = np.array([hi - lo, self.size[1], self.size[2]])
        return origin, extent


def required_halo(substeps, reach, slack=2.0):
    """Halo depth that keeps owned cells exact for `substeps` communication-free substeps: the influence of a
    wrong halo-edge cell advances by at most reach + slack per substep (slack >= twice the largest
    per-substep displacement)."""
    return float(substeps) * (float(reach) + float(slack))


class SlabState:
    """Cells owned by one rank.  Bond rows hold global ids (-1 padded)."""

    FIELDS = ("loc", "radii", "mot", "gid", "bond_len", "bond_gid")

    def __init__(self, loc, radii, mot, gid, stride=16):
        n = loc.shape[0]
        self.loc, self.radii, self.mot, self.gid = loc, radii, mot, gid
        self.bond_len = torch.zeros(n, dtype=torch.int64, device=loc.device)
        self.bond_gid = torch.full((n, stride), -1, dtype=torch.int64, device=loc.device)

    @property
    def n(self):
        return int(self.loc.shape[0])

    @property
    def stride(self):
        return int(self.bond_gid.shape[1])

    def grow_stride(self, stride):
        if stride > self.stride:
            pad = torch.full((self.n, stride - self.stride), -1, dtype=torch.int64, device=self.loc.device)
            self.bond_gid = torch.cat([self.bond_gid, pad], dim=1)

    # ---- packing: one float64 matrix per message (ids are exact in float64 below 2^53) -----------------------
    def pack(self, mask) -> torch.Tensor:
        idx = torch.nonzero(mask, as_tuple=False).flatten()
        cols = [self.loc[idx], self.radii[idx, None], self.mot[idx], self.gid[idx, None].double(),
                self.bond_len[idx, None].double(), self.bond_gid[idx].double()]
        return torch.cat(cols, dim=1).contiguous()

    @staticmethod
    def unpack(msg: torch.Tensor, stride):
        if msg.numel() == 0:
            msg = msg.reshape(0, 9 + stride)
        return dict(loc=msg[:, 0:3].contiguous(), radii=msg[:, 3].contiguous(), mot=msg[:, 4:7].contiguous(),
                    gid=msg[:, 7].long(), bond_len=msg[:, 8].long(), bond_gid=msg[:, 9:9 + stride].long())

    def width(self):
        return 9 + self.stride

    def keep(self, mask):
        idx = torch.nonzero(mask, as_tuple=False).flatten()
        for f in self.FIELDS:
            setattr(self, f, getattr(self, f)[idx].contiguous())

    def append(self, part):
        for f in self.FIELDS:
            setattr(self, f, torch.cat([getattr(self, f), part[f]], dim=0))


# ---- exchange fabrics ---------------------------------------------------------------------------------------------
class DistFabric:
    """Neighbour exchange over torch.distributed (NCCL on GPUs, gloo on CPU): sizes first, then one packed
    float64 message per neighbour and direction, all four transfers posted together."""

    def __init__(self, rank, world, device):
        import torch.distributed as dist
        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        self.bytes_sent = 0
        self.exchanges = 0

    def exchange(self, to_left: Optional[torch.Tensor], to_right: Optional[torch.Tensor], width: int):
        dist = self.dist
        left, right = self.rank - 1, self.rank + 1
        has_l, has_r = left >= 0, right < self.world
        counts_out = torch.tensor([to_left.shape[0] if has_l else 0, to_right.shape[0] if has_r else 0],
                                  dtype=torch.int64, device=self.device)
        counts_in = torch.zeros(2, dtype=torch.int64, device=self.device)
        ops = []
        if has_l:
            ops += [dist.P2POp(dist.isend, counts_out[0:1], left), dist.P2POp(dist.irecv, counts_in[0:1], left)]
        if has_r:
            ops += [dist.P2POp(dist.isend, counts_out[1:2], right), dist.P2POp(dist.irecv, counts_in[1:2], right)]
        for w in dist.batch_isend_irecv(ops) if ops else []:
            w.wait()
        n_l, n_r = (int(v) for v in counts_in.tolist())
        from_left = torch.empty((n_l, width), dtype=torch.float64, device=self.device)
        from_right = torch.empty((n_r, width), dtype=torch.float64, device=self.device)
        ops = []
        if has_l:
            if to_left.shape[0]:
                ops.append(dist.P2POp(dist.isend, to_left, left))
            if n_l:
                ops.append(dist.P2POp(dist.irecv, from_left, left))
        if has_r:
            if to_right.shape[0]:
                ops.append(dist.P2POp(dist.isend, to_right, right))
            if n_r:
                ops.append(dist.P2POp(dist.irecv, from_right, right))
        for w in dist.batch_isend_irecv(ops) if ops else []:
            w.wait()
        self.bytes_sent += sum(t.numel() * 8 for t, ok in ((to_left, has_l), (to_right, has_r)) if ok)
        self.exchanges += 1
        return from_left, from_right


class LoopbackWorld:
    """All ranks in one process (tests / single-GPU emulation): exchange is a list shuffle."""

    def __init__(self, world):
        self.world = world

    def exchange_all(self, outgoing):
        """outgoing[k] = (to_left, to_right) -> incoming[k] = (from_left, from_right)."""
        incoming = []
        for k in range(self.world):
            from_left = outgoing[k - 1][1] if k > 0 else None
            from_right = outgoing[k + 1][0] if k < self.world - 1 else None
            incoming.append((from_left, from_right))
        return incoming


# ---- per-step operations of one rank -----------------------------------------------------------------------------------
@dataclass
class LocalProblem:
    """Owned + halo cells of one rank for one step; what the single-GPU pipeline runs on."""
    n_own: int
    loc: torch.Tensor
    radii: torch.Tensor
    mot: torch.Tensor
    gid: torch.Tensor
    bond_len: torch.Tensor      # int32 (n_local,)
    bond_cols: torch.Tensor     # int32 (n_local, stride) local indices


def halo_messages(state: SlabState, geom: SlabGeometry):
    x = state.loc[:, 0]
    to_left = state.pack(x < geom.x_lo + geom.halo) if geom.rank > 0 else None
    to_right = state.pack(x >= geom.x_hi - geom.halo) if geom.rank < geom.world - 1 else None
    return to_left, to_right


def assemble_local(state: SlabState, from_left, from_right) -> LocalProblem:
    parts = [dict(loc=state.loc, radii=state.radii, mot=state.mot, gid=state.gid, bond_len=state.bond_len,
                  bond_gid=state.bond_gid)]
    for msg in (from_left, from_right):
        if msg is not None and msg.shape[0]:
            parts.append(SlabState.unpack(msg, state.stride))
    cat = {f: torch.cat([p[f] for p in parts], dim=0) for f in SlabState.FIELDS}
    gid = cat["gid"]
    n_local, stride = gid.shape[0], state.stride
    # global id -> local index (absent partners -> -1, their bonds are dropped for this step)
    order = torch.argsort(gid)
    sorted_gid = gid[order]
    want = cat["bond_gid"]
    pos = torch.searchsorted(sorted_gid, want.clamp(min=0).reshape(-1)).reshape(n_local, stride)
    pos = pos.clamp(max=max(n_local - 1, 0))
    hit = (sorted_gid[pos] == want) & (want >= 0) if n_local else torch.zeros_like(want, dtype=torch.bool)
    k = torch.arange(stride, device=gid.device)[None, :]
    hit &= k < cat["bond_len"][:, None]
    local = torch.where(hit, order[pos], torch.full_like(want, -1))
    # compact each row: valid entries first (stable), count them
    rank_key = torch.argsort((~hit).to(torch.int8), dim=1, stable=True)
    cols = torch.gather(local, 1, rank_key)
    return LocalProblem(state.n, cat["loc"].contiguous(), cat["radii"].contiguous(), cat["mot"].contiguous(), gid,
                        hit.sum(dim=1).to(torch.int32), cols.to(torch.int32).contiguous())


def adopt_results(state: SlabState, local: LocalProblem, new_loc, bond_len, bond_cols):
    """Copy the owned cells' results of a step back: positions, and bond rows translated to global ids."""
    n = local.n_own
    state.loc = new_loc[:n].contiguous()
    stride = bond_cols.shape[1]
    state.grow_stride(stride)
    blen = bond_len[:n].long()
    cols = bond_cols[:n].long()
    k = torch.arange(stride, device=cols.device)[None, :]
    valid = k < blen[:, None]
    cols = torch.where(valid, cols, torch.zeros_like(cols))       # entries past the row length are undefined
    g = local.gid[cols]
    state.bond_gid = torch.where(valid, g, torch.full_like(g, -1))
    if state.bond_gid.shape[1] < state.stride:
        state.grow_stride(state.stride)
    state.bond_len = blen


def migration_messages(state: SlabState, geom: SlabGeometry):
    x = state.loc[:, 0]
    go_left = (x < geom.x_lo) if geom.rank > 0 else torch.zeros_like(x, dtype=torch.bool)
    go_right = (x >= geom.x_hi) if geom.rank < geom.world - 1 else torch.zeros_like(x, dtype=torch.bool)
    to_left = state.pack(go_left) if geom.rank > 0 else None
    to_right = state.pack(go_right) if geom.rank < geom.world - 1 else None
    state.keep(~(go_left | go_right))
    return to_left, to_right


def absorb_migrants(state: SlabState, from_left, from_right):
    for msg in (from_left, from_right):
        if msg is not None and msg.shape[0]:
            state.append(SlabState.unpack(msg, state.stride))


def halo_violation(state: SlabState, geom: SlabGeometry):
    """Largest distance an owned cell sits outside its slab (device scalar).  The halo argument needs it to
    stay below the slack the halo was sized with; callers check it once per step."""
    x = state.loc[:, 0]
    worst = torch.zeros((), dtype=torch.float64, device=x.device)
    if x.numel() == 0:
        return worst
    if geom.rank > 0:
        worst = torch.maximum(worst, (geom.x_lo - x).max())
    if geom.rank < geom.world - 1:
        worst = torch.maximum(worst, (x - geom.x_hi).max())
    return worst


class FieldSlab:
    """Rows [row_lo,row_hi) of the global (gx,gy) morphogen grid owned by one rank, stored with `halo_rows`
    extra rows on each interior side.  Grid row r sits at x = r*h and belongs to the slab containing it.
    halo_rows must be (diffusion substeps + 1): the stencil invalidates one halo row per substep from the
    outside in, and a cell next to the interface may sample the neighbour's first row."""

    def __init__(self, geom: SlabGeometry, grid, spat_res, halo_rows, device):
        gx, gy = int(grid[0]), int(grid[1])
        h = float(spat_res)
        self.gx, self.gy, self.h, self.halo_rows = gx, gy, h, int(halo_rows)
        self.row_lo = int(np.ceil(geom.x_lo / h)) if geom.rank > 0 else 0
        self.row_hi = min(int(np.ceil(geom.x_hi / h)), gx) if geom.rank < geom.world - 1 else gx
        self.ext_lo = max(self.row_lo - self.halo_rows, 0) if geom.rank > 0 else 0
        self.ext_hi = min(self.row_hi + self.halo_rows, gx) if geom.rank < geom.world - 1 else gx
        if self.row_hi - self.row_lo < self.halo_rows:
            raise ValueError("slab narrower than the diffusion halo")
        self.has_left, self.has_right = geom.rank > 0, geom.rank < geom.world - 1
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


def rank_step(state: SlabState, geom: SlabGeometry, compute: Callable[[LocalProblem], tuple],
              field: Optional[FieldSlab] = None, field_ops=None):
    """One simulation step of one rank, written as a generator: it yields (tag, (to_left, to_right), width)
    whenever it needs a neighbour exchange and is sent (from_left, from_right) back.  `compute(local)` runs the
    mechanics on the local problem and returns (new_loc, bond_len, bond_cols).  `field_ops`, if a morphogen
    field is carried, is an object with deposit(local, field), diffuse(field) and sample(local, field)."""
    from_l, from_r = yield ("halo", halo_messages(state, geom), state.width())
    local = assemble_local(state, from_l, from_r)
    new_loc, blen, bcols = compute(local)
    local.loc = new_loc
    if field is not None:
        field_ops.deposit(local, field)                       # owned + halo cells at their new positions
        from_l, from_r = yield ("rows", field.messages(), field.gy)
        field.absorb(from_l, from_r)
        field_ops.diffuse(field)                              # S_d substeps, no communication
        field_ops.sample(local, field)
    adopt_results(state, local, new_loc, blen, bcols)
    from_l, from_r = yield ("migrate", migration_messages(state, geom), state.width())
    absorb_migrants(state, from_l, from_r)
    return local


def slab_step(state: SlabState, geom: SlabGeometry, fabric: DistFabric, compute, field=None, field_ops=None):
    """Drive rank_step over torch.distributed."""
    gen = rank_step(state, geom, compute, field, field_ops)
    incoming = None
    try:
        while True:
            tag, (to_l, to_r), width = gen.send(incoming) if incoming is not None else next(gen)
            incoming = fabric.exchange(to_l, to_r, width)
    except StopIteration as done:
        return done.value


def loopback_step(states, geoms, computes, fields=None, field_ops=None):
    """The same step for all ranks emulated in one process (tests; several slabs on one GPU)."""
    world = LoopbackWorld(len(states))
    if callable(computes):
        computes = [computes] * len(states)
    fields = fields or [None] * len(states)
    field_ops = field_ops or [None] * len(states)
    gens = [rank_step(s, g, c, f, o) for s, g, c, f, o in zip(states, geoms, computes, fields, field_ops)]
    requests = [next(g) for g in gens]
    results = [None] * len(gens)
    while any(r is not None for r in requests):
        incoming = world.exchange_all([r[1] for r in requests])
        nxt = []
        for k, (g, inc) in enumerate(zip(gens, incoming)):
            try:
                nxt.append(g.send(inc))
            except StopIteration as done:
                results[k] = done.value
                nxt.append(None)
        requests = nxt
        if any(r is None for r in requests):
            assert all(r is None for r in requests), "ranks fell out of lock-step"
    return results
