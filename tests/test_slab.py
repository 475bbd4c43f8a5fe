"""Slab decomposition: host-side logic on CPU.  The compute callback is the CPU oracle, so these
tests check the partitioning itself -- halo selection, bond-row translation through global ids, migration --
by requiring an N-slab run to reproduce the single-domain run.  The last test does it over torch.distributed
(gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch

from hipsc_abm_b200 import slab
from oracle import oracle as orc

SIZE = np.array([640.0, 180.0, 0.0])
STRIDE = 24


def make_cells(n=1150, seed=3):
    rng = np.random.default_rng(seed)
    loc = rng.random((n, 3)) * SIZE
    mot = np.zeros((n, 3))
    th = rng.random(n) * 2 * np.pi
    mot[:, 0], mot[:, 1] = 2e-9 * np.cos(th), 2e-9 * np.sin(th)
    return loc, np.full(n, 5.0), mot


def oracle_compute(local):
    """The hot-path step of one rank on its owned + halo cells, with the CPU oracle.  Bond rows hold global ids:
    translate to local indices for the oracle's graph and back (partners that are not local are dropped)."""
    n = local.loc.shape[0]
    gid = local.gid.numpy().astype(np.int64)
    lookup = {int(g): i for i, g in enumerate(gid)}
    sim = orc.OracleSim(local.loc.numpy(), local.radii.numpy(), SIZE)
    sim.motility_forces = local.mot.numpy().copy()
    blen, bgid = local.bond_len.numpy(), local.bond_gid.numpy()
    edges = [(i, lookup[int(g)]) for i in range(n) for g in bgid[i, :blen[i]] if int(g) in lookup]
    sim.jkr_graph.add_edges(np.array(edges, dtype=np.int64).reshape(-1, 2))
    sim.jkr_graph.simplify()
    sim.apply_forces()
    deg = sim.jkr_graph.degrees()
    assert deg.max(initial=0) <= STRIDE
    rows = np.full((n, STRIDE), -1, dtype=np.int32)
    for i in range(n):
        nb = sim.jkr_graph.neighbors(i)
        rows[i, :len(nb)] = gid[nb]
    local.loc.copy_(torch.from_numpy(sim.locations))                 # in place, like the device pipeline
    return torch.from_numpy(deg.astype(np.int32)), torch.from_numpy(rows)


def single_domain(loc, radii, mot, steps):
    sim = orc.OracleSim(loc, radii, SIZE)
    for _ in range(steps):
        sim.motility_forces = mot.copy()
        sim.apply_forces()
    return sim.locations, sim.jkr_graph.edge_array()


def split(loc, radii, mot, world, halo):
    geoms = [slab.SlabGeometry.uniform(SIZE, world, r, halo) for r in range(world)]
    gid = np.arange(len(loc))
    states = []
    for g in geoms:
        m = (loc[:, 0] >= g.x_lo) & ((loc[:, 0] < g.x_hi) | (g.rank == world - 1))
        states.append(slab.SlabState(torch.from_numpy(loc[m].copy()), torch.from_numpy(radii[m].copy()),
                                     torch.from_numpy(mot[m].copy()), torch.from_numpy(gid[m].copy()), STRIDE))
    return geoms, states


def collect(states):
    gid = torch.cat([s.gid for s in states]).numpy()
    loc = torch.cat([s.loc for s in states]).numpy()
    order = np.argsort(gid)
    edges = []
    for s in states:
        k = np.arange(s.stride)[None, :] < s.bond_len.numpy()[:, None]
        rows = np.repeat(s.gid.numpy().astype(np.int64), s.bond_len.numpy())
        edges.append(np.stack([rows, s.bond_gid.numpy()[k].astype(np.int64)], axis=1))
    return gid[order], loc[order], orc.canonical_edges(np.concatenate(edges))


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_reproduce_single_domain(world):
    loc, radii, mot = make_cells()
    steps = 3
    halo = slab.required_halo(11, 10.3100107, slack=5.0)
    assert halo < SIZE[0] / world
    ref_loc, ref_edges = single_domain(loc, radii, mot, steps)
    geoms, states = split(loc, radii, mot, world, halo)
    moved = 0
    for _ in range(steps):
        before = [set(s.gid.tolist()) for s in states]
        slab.loopback_step(states, geoms, oracle_compute)
        moved += sum(len(set(s.gid.tolist()) - b) for s, b in zip(states, before))
    # ownership follows position at the start of every step: after one more (empty-force) exchange all owned
    # cells sit inside their slab
    gid, got_loc, got_edges = collect(states)
    assert np.array_equal(gid, np.arange(len(loc)))                 # nobody lost or duplicated
    assert moved > 0                                                 # migration was exercised
    assert np.array_equal(got_edges, ref_edges)                      # bonds identical across the interfaces
    assert np.abs(got_loc - ref_loc).max() <= 1e-9                   # fp sum order differs with local indexing


def test_pack_unpack_roundtrip_and_ownership():
    loc, radii, mot = make_cells(300, seed=5)
    geoms, states = split(loc, radii, mot, 2, 60.0)
    s0, s1 = states
    s0.bond_gid[3, :2] = torch.tensor([7, 11], dtype=torch.int32)
    s0.bond_len[3] = 2
    idx = torch.tensor([3, 0, 5])
    msg = s0.pack(idx)
    assert msg.dtype == torch.uint8 and msg.shape == (3, s0.record_bytes())
    back = slab.unpack(msg, s0.stride)
    for f in slab.FIELDS:
        assert torch.equal(back[f], getattr(s0, f)[idx])
    # ownership is a pure function of x, complementary on the two sides of an interface
    x = torch.tensor([geoms[0].x_hi - 1e-9, geoms[0].x_hi, geoms[0].x_hi + 1e-9, -5.0, SIZE[0] + 5.0])
    a, b = geoms[0].inside(x), geoms[1].inside(x)
    assert torch.equal(a ^ b, torch.ones_like(a)) and a.tolist() == [True, False, False, True, False]
    origin, extent = geoms[1].local_box()
    assert origin[0] == geoms[1].x_lo - 60.0 and abs(origin[0] + extent[0] - SIZE[0]) < 1e-12


def _worker(rank, world, port, loc, radii, mot, steps, halo, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    geoms, states = split(loc, radii, mot, world, halo)
    state, geom = states[rank], geoms[rank]
    fabric = slab.DistFabric(rank, world, torch.device("cpu"))
    for _ in range(steps):
        slab.slab_step(state, geom, fabric, oracle_compute)
    gathered = [None] * world
    k = np.arange(state.stride)[None, :] < state.bond_len.numpy()[:, None]
    edges = np.stack([np.repeat(state.gid.numpy().astype(np.int64), state.bond_len.numpy()),
                      state.bond_gid.numpy()[k].astype(np.int64)], axis=1)
    dist.all_gather_object(gathered, (state.gid.numpy(), state.loc.numpy(), edges, fabric.exchanges))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    loc, radii, mot = make_cells(900, seed=8)
    steps = 2
    halo = slab.required_halo(11, 10.3100107, slack=5.0)
    ref_loc, ref_edges = single_domain(loc, radii, mot, steps)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, loc, radii, mot, steps, halo, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gid = np.concatenate([g[0] for g in gathered])
    got = np.concatenate([g[1] for g in gathered])[np.argsort(gid)]
    edges = orc.canonical_edges(np.concatenate([g[2] for g in gathered]))
    assert np.array_equal(np.sort(gid), np.arange(len(loc)))
    assert np.array_equal(edges, ref_edges)
    assert np.abs(got - ref_loc).max() <= 1e-9
    assert gathered[0][3] == steps                                   # ONE cell exchange per step (migration rides on it)


class OracleFieldOps:
    """field_ops of slab.rank_step with the CPU oracle (global-grid arithmetic, then cut to the slab's rows)."""
    H, D, SD = 7.0, 50.0, 6

    def __init__(self):
        self.sampled = None
        self.dt = 0.2 * self.H ** 2 / self.D

    def deposit(self, local, field):
        full = np.zeros((field.gx, field.gy, 1))
        orc.adjust_morphogens(local.loc.numpy(), 1.0, full, self.H)
        field.data += torch.from_numpy(full[field.ext_lo:field.ext_hi, :, 0])

    def diffuse(self, field):
        field.data[:] = torch.from_numpy(orc.diffuse_field(field.data.numpy(), self.SD, self.dt, self.dt, self.D,
                                                           self.H ** 2, 2.0, 0.1))

    def sample(self, local, field):
        full = np.zeros((field.gx, field.gy, 1))
        full[field.ext_lo:field.ext_hi, :, 0] = field.data.numpy()
        self.sampled = orc.get_concentrations(local.loc.numpy()[:local.n_own], full, self.H)


@pytest.mark.parametrize("first", [False, True], ids=["morphogen_after", "morphogen_first"])
def test_field_slabs_reproduce_single_domain(first):
    world, steps = 3, 2
    loc, radii, mot = make_cells()
    ops0 = OracleFieldOps()
    gx, gy = int(np.ceil(SIZE[0] / ops0.H)) + 1, int(np.ceil(SIZE[1] / ops0.H)) + 1
    # single domain: mechanics -> deposit -> diffuse -> sample, twice (morphogen_first: the morphogen stage
    # before the mechanics, the reference's order)
    sim = orc.OracleSim(loc, radii, SIZE)
    full = np.zeros((gx, gy, 1))

    def morphogen():
        orc.adjust_morphogens(sim.locations, 1.0, full, ops0.H)
        full[:, :, 0] = orc.diffuse_field(full[:, :, 0], ops0.SD, ops0.dt, ops0.dt, ops0.D, ops0.H ** 2, 2.0, 0.1)
        return orc.get_concentrations(sim.locations, full, ops0.H)
    for _ in range(steps):
        if first:
            ref_sampled = morphogen()
        sim.motility_forces = mot.copy()
        sim.apply_forces()
        if not first:
            ref_sampled = morphogen()
    halo = slab.required_halo(11, 10.3100107, slack=5.0) + ops0.H
    geoms, states = split(loc, radii, mot, world, halo)
    # one more halo row than substeps: a cell next to the interface samples the neighbour's first row
    fields = [slab.FieldSlab(g, (gx, gy), ops0.H, ops0.SD + 1, torch.device("cpu")) for g in geoms]
    ops = [OracleFieldOps() for _ in geoms]
    assert [f.row_lo for f in fields[1:]] == [f.row_hi for f in fields[:-1]] and fields[-1].row_hi == gx
    for _ in range(steps):
        slab.loopback_step(states, geoms, oracle_compute, fields, ops, morphogen_first=first)
    got = np.concatenate([f.owned().numpy() for f in fields], axis=0)
    assert got.shape == (gx, gy)
    assert np.abs(got - full[:, :, 0]).max() <= 1e-12                 # deposits are summed in a different order
    # sampled values of the owned cells (taken before migration) match the single-domain values by global id
    # (cells that migrated this step are sampled by their previous owner, so match through positions instead)
    assert all(o.sampled is not None and len(o.sampled) > 0 for o in ops)
    assert abs(sum(o.sampled.sum() for o in ops) - ref_sampled.sum()) <= 1e-9


class PaddedOps(slab.TorchSlabOps):
    """The protocol of pipeline.NativeSlabOps on the CPU: full-capacity messages with their valid row counts in a
    tensor (no size handshake), the owned count of a merged step left in a tensor until somebody asks for it, and
    the next pack fed from the not yet sliced local arrays."""
    send_cap = 400

    def outgoing(self, state, geom):
        pend = state.pending()
        if pend is not None:                                   # owned count still "on the device"
            full, rows, n_own = pend
            view = slab.SlabState.__new__(slab.SlabState)
            view._pending, view._d = None, {f: full[f][:int(n_own.item())] for f in slab.FIELDS}
            msgs = super().outgoing(view, geom)
        else:
            msgs = super().outgoing(state, geom)
        rec = state.record_bytes()
        out, counts = [], torch.zeros(2, dtype=torch.int32)
        for k, m in enumerate(msgs):
            buf = torch.full((self.send_cap, rec), 255, dtype=torch.uint8)     # garbage beyond the count
            if m is not None:
                assert m.shape[0] <= self.send_cap
                buf[:m.shape[0]] = m
                counts[k] = m.shape[0]
            out.append(buf)
        return out[0], out[1], counts

    def merge(self, state, geom, from_l, from_r, counts_in=None):
        n_l, n_r = (int(v) for v in counts_in.tolist())
        local = super().merge(state, geom, from_l[:n_l] if from_l is not None else None,
                              from_r[:n_r] if from_r is not None else None)
        return slab.LocalProblem(torch.tensor([local.n_own]), local.loc, local.radii, local.mot, local.gid,
                                 local.bond_len, local.bond_gid)


def test_fixed_capacity_messages_and_deferred_owned_count():
    world, steps = 3, 3
    loc, radii, mot = make_cells()
    halo = slab.required_halo(11, 10.3100107, slack=5.0)
    ref_loc, ref_edges = single_domain(loc, radii, mot, steps)
    geoms, states = split(loc, radii, mot, world, halo)
    ops = [PaddedOps() for _ in geoms]
    for _ in range(steps):
        slab.loopback_step(states, geoms, oracle_compute, ops=ops)
        assert all(s.pending() is not None for s in states)        # nothing forced the owned count yet
    gid, got_loc, got_edges = collect(states)                      # reading the state settles it
    assert all(s.pending() is None for s in states)
    assert np.array_equal(gid, np.arange(len(loc)))
    assert np.array_equal(got_edges, ref_edges)
    assert np.abs(got_loc - ref_loc).max() <= 1e-9


def test_alternating_motility_sign_on_a_deferred_state():
    """bench.py flips the sign of the motility forces after every step (stationary synthetic colony) through
    SlabState.negate_motility(), which must not force the deferred owned count and must follow migrating cells."""
    world, steps = 3, 4
    loc, radii, mot = make_cells()
    halo = slab.required_halo(11, 10.3100107, slack=5.0)
    sim = orc.OracleSim(loc, radii, SIZE)
    for k in range(steps):
        sim.motility_forces = mot.copy() if k % 2 == 0 else -mot
        sim.apply_forces()
    geoms, states = split(loc, radii, mot, world, halo)
    ops = [PaddedOps() for _ in geoms]
    for _ in range(steps):
        slab.loopback_step(states, geoms, oracle_compute, ops=ops)
        for s in states:
            s.negate_motility()
        assert all(s.pending() is not None for s in states)        # the flip did not read the owned count
    gid, got_loc, got_edges = collect(states)
    assert np.array_equal(gid, np.arange(len(loc)))
    assert np.array_equal(got_edges, sim.jkr_graph.edge_array())
    assert np.abs(got_loc - sim.locations).max() <= 1e-9
    # settled state: the flip works on the owned rows themselves
    before = states[0].mot.clone()
    states[0].negate_motility()
    assert torch.equal(states[0].mot, -before)


def test_balanced_slabs_follow_the_cell_density():
    rng = np.random.default_rng(11)
    size = np.array([1200.0, 300.0, 0.0])
    # three clusters of very different populations along x
    x = np.concatenate([rng.normal(250, 110, 3000), rng.normal(600, 60, 600), rng.normal(1000, 40, 1400)]).clip(0, 1200)
    halo = 40.0
    geoms = [slab.SlabGeometry.balanced(size, 4, r, halo, x) for r in range(4)]
    b = geoms[0].bounds
    assert all(np.array_equal(g.bounds, b) for g in geoms) and b[0] == 0.0 and b[-1] == 1200.0
    assert np.all(np.diff(b) >= halo)
    counts = np.histogram(x, bins=b)[0]
    uniform = np.histogram(x, bins=np.linspace(0, 1200, 5))[0]
    assert counts.max() - counts.min() < 0.35 * len(x) / 4 < uniform.max() - uniform.min()
    # every cell has exactly one owner under the position predicate
    xt = torch.from_numpy(x)
    assert int(sum(g.inside(xt).sum() for g in geoms)) == len(x)
    with pytest.raises(ValueError):
        slab.SlabGeometry.balanced(size, 40, 0, halo, x)


def test_balanced_slabs_reproduce_single_domain():
    world, steps = 3, 2
    loc, radii, mot = make_cells()
    loc = loc.copy()
    loc[:, 0] = SIZE[0] * (loc[:, 0] / SIZE[0]) ** 1.6            # denser towards x = 0
    halo = slab.required_halo(11, 10.3100107, slack=5.0)
    ref_loc, ref_edges = single_domain(loc, radii, mot, steps)
    geoms = [slab.SlabGeometry.balanced(SIZE, world, r, halo, loc[:, 0]) for r in range(world)]
    gid = np.arange(len(loc))
    states = []
    for g in geoms:
        m = g.inside(torch.from_numpy(loc[:, 0])).numpy()
        states.append(slab.SlabState(torch.from_numpy(loc[m].copy()), torch.from_numpy(radii[m].copy()),
                                     torch.from_numpy(mot[m].copy()), torch.from_numpy(gid[m].copy()), STRIDE))
    uni = np.histogram(loc[:, 0], bins=np.linspace(0, SIZE[0], world + 1))[0]
    assert max(s.n for s in states) - min(s.n for s in states) < uni.max() - uni.min()   # as far as the halo allows
    for _ in range(steps):
        slab.loopback_step(states, geoms, oracle_compute)
    got_gid, got_loc, got_edges = collect(states)
    assert np.array_equal(got_gid, gid)
    assert np.array_equal(got_edges, ref_edges)
    assert np.abs(got_loc - ref_loc).max() <= 1e-9
