"""Empirical check of the halo depth used by bench.py: several slabs emulated on one GPU against the single-domain
run on a bench-like colony (uniform random start, constant 2 nN motility), after it has relaxed for a few steps.
usage: check_halo.py <cells> <world> <slack> <steps>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hipsc_abm_b200 import slab
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep, NativeSlabOps, SlabPipeline, StepConfig

n, world, slack, steps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
wl = bench.make_workload("c3", seed=5, n_override=n)
size, dev = wl["size"], torch.device("cuda")
cfg = StepConfig(grid=None, bond_stride=32, search_cols_per_cell=16.0)
eng = ColonyEngine(size); eng.upload(wl["loc"], wl["radii"], wl["mot"])
one = HotPathStep(eng, cfg)
relax = 6
for _ in range(relax):
    one.step()
one.check()
loc0 = eng.locations_host().copy()
b = one.current_bonds()
rows0, len0 = b.cols[:n].clone(), b.row_len[:n].clone()
for _ in range(steps):
    one.step()
one.check()
ref_loc = eng.locations_host()
rb = one.current_bonds()
def edges(gid, row_len, cols):
    gid, row_len = gid.cpu().numpy().astype(np.int64), row_len.cpu().numpy().astype(np.int64)
    cols = cols.cpu().numpy().astype(np.int64)
    k = np.arange(cols.shape[1])[None, :] < row_len[:, None]
    e = np.stack([np.repeat(gid, row_len), cols[k]], axis=1)
    lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
    return np.unique(lo * (n + 1) + hi)
ref_edges = edges(torch.arange(n), rb.row_len[:n], rb.cols[:n])

halo = slab.required_halo(11, eng.jkr_reach(10.0), slack=slack)
geoms = [slab.SlabGeometry.uniform(size, world, r, halo) for r in range(world)]
states, pipes = [], []
for g in geoms:
    m = (loc0[:, 0] >= g.x_lo) & ((loc0[:, 0] < g.x_hi) | (g.rank == world - 1))
    idx = np.nonzero(m)[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    st = slab.SlabState(t(loc0[idx]), t(wl["radii"][idx]), t(wl["mot"][idx]), t(idx), 32)
    st.bond_len, st.bond_gid = len0[idx].clone(), rows0[idx].clone()
    states.append(st)
    pipes.append(SlabPipeline(g, cfg, dev, uniform_radius=5.0, side_stream=False))
ops = [NativeSlabOps(p.eng, send_cap=n) for p in pipes]
for _ in range(steps):
    slab.loopback_step(states, geoms, [p.compute for p in pipes], None, None, ops)
for p, o in zip(pipes, ops):
    p.check(); o.check()
gid = torch.cat([s.gid for s in states]).cpu().numpy()
got = torch.cat([s.loc for s in states]).cpu().numpy()[np.argsort(gid)]
got_edges = np.unique(np.concatenate([edges(s.gid, s.bond_len, s.bond_gid) for s in states]))
d = np.abs(got - ref_loc).max(axis=1)
only_ref, only_got = np.setdiff1d(ref_edges, got_edges), np.setdiff1d(got_edges, ref_edges)
if len(only_ref) or len(only_got):
    show = lambda k: [(int(v // (n + 1)), int(v % (n + 1))) for v in k[:5]]
    print("bonds only in the single-domain run:", len(only_ref), show(only_ref), " only in the slabs:", len(only_got),
          show(only_got))
    for lo, hi in show(only_ref)[:3] + show(only_got)[:3]:
        print("   pair", lo, hi, "distance", float(np.linalg.norm(ref_loc[lo] - ref_loc[hi])), "x", ref_loc[lo][0], ref_loc[hi][0])
print(f"cells {n} slabs {world} slack {slack} halo {halo:.1f} um, {steps} steps after {relax} relaxing steps: "
      f"max |dx| {d.max():.3e} um, cells off by > 1e-9: {(d > 1e-9).sum()}, "
      f"bond sets equal: {np.array_equal(got_edges, ref_edges)} ({len(ref_edges)} bonds)")
