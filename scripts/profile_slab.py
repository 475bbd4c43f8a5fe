"""Host-side profile of the slab step: two slabs emulated on one GPU (loopback), cProfile of rank logic."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hipsc_abm_b200 import slab
from hipsc_abm_b200.pipeline import SlabPipeline, StepConfig

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
world = 2
dev = torch.device("cuda")
wls = [bench.make_workload(name, seed=r) for r in range(world)]
spec, n = wls[0]["spec"], wls[0]["n"]
cfg = StepConfig(grid=wls[0]["grid"], spat_res=spec["h"] or 1.0, diffuse_dt=spec["diffuse_dt"] or 1.0,
                 diffuse_steps=spec["sd"], bond_stride=16, search_cols_per_cell=12.0)
side = wls[0]["size"][0]
gsize = wls[0]["size"].copy(); gsize[0] = side * world
halo = slab.required_halo(11, 10.3100107, slack=4.0) + spec["h"]
geoms = [slab.SlabGeometry.uniform(gsize, world, r, halo) for r in range(world)]
pipes = [SlabPipeline(g, cfg, dev, uniform_radius=5.0, side_stream=False) for g in geoms]
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
states = []
for r, (wl, g) in enumerate(zip(wls, geoms)):
    loc = wl["loc"].copy(); loc[:, 0] += g.x_lo
    states.append(slab.SlabState(t(loc), t(wl["radii"]), t(wl["mot"]), t(np.arange(n) + r * n), 16))
gx = int(np.ceil(gsize[0] / spec["h"])) + 1
fields = [slab.FieldSlab(g, (gx, wls[0]["grid"][1]), spec["h"], spec["sd"] + 1, dev) for g in geoms]
from hipsc_abm_b200.pipeline import NativeSlabOps
ops = [NativeSlabOps(p.eng) for p in pipes] if os.environ.get("NATIVE", "1") == "1" else None
def step():
    slab.loopback_step(states, geoms, [p.compute for p in pipes], fields, pipes, ops)
for _ in range(3):
    step()
torch.cuda.synchronize()
if os.environ.get("ONLY_STEPS"):
    for _ in range(int(os.environ["ONLY_STEPS"])):
        step()
    torch.cuda.synchronize()
    print("n_local", [int(p.eng.n) for p in pipes])
    sys.exit(0)
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("ms per loopback step (2 slabs, serial):", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

# ---- coarse device-time breakdown (synchronising timers) ----
import collections
acc = collections.defaultdict(float)
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] += time.perf_counter() - t
        return r
    return w
for p in pipes:
    p.compute = timed("compute", p.compute)
    p.deposit = timed("deposit", p.deposit)
    p.diffuse = timed("diffuse", p.diffuse)
    p.sample = timed("sample", p.sample)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("sync-timed total ms/step:", tot / 5 * 1e3, {k: round(v / 5 * 1e3, 3) for k, v in acc.items()},
      "plumbing:", round((tot - sum(acc.values())) / 5 * 1e3, 3))
# single-domain eager reference for the same cell count
from hipsc_abm_b200.engine import ColonyEngine
from hipsc_abm_b200.pipeline import HotPathStep
wl = wls[0]
eng = ColonyEngine(wl["size"]); eng.upload(wl["loc"], wl["radii"], wl["mot"])
one = HotPathStep(eng, cfg)
for _ in range(3): one.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): one.step()
torch.cuda.synchronize(); print("single-domain eager ms/step:", (time.perf_counter() - t0) / 5 * 1e3)
