"""Phase timing of the slab step over NCCL (run under torchrun with 2+ ranks); synchronising timers."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from hipsc_abm_b200 import slab
from hipsc_abm_b200.pipeline import NativeSlabOps, SlabPipeline, StepConfig

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = bench.make_workload(name, seed=rank); spec, n = wl["spec"], wl["n"]
cfg = StepConfig(grid=wl["grid"], spat_res=spec["h"], diffuse_dt=spec["diffuse_dt"], diffuse_steps=spec["sd"])
gsize = wl["size"].copy(); gsize[0] *= world
halo = slab.required_halo(11, 10.3100107, slack=4.0) + spec["h"]
geom = slab.SlabGeometry.uniform(gsize, world, rank, halo)
first = len(sys.argv) > 2 and sys.argv[2] == "first"
pipe = SlabPipeline(geom, cfg, dev, uniform_radius=5.0)
loc = wl["loc"].copy(); loc[:, 0] += geom.x_lo
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
state = slab.SlabState(t(loc), t(wl["radii"]), t(wl["mot"]), t(np.arange(n) + rank * n), 16)
gx = int(np.ceil(gsize[0] / spec["h"])) + 1
field = slab.FieldSlab(geom, (gx, wl["grid"][1]), spec["h"], spec["sd"] + 1, dev)
fabric = slab.DistFabric(rank, world, dev); ops = NativeSlabOps(pipe.eng, send_cap=int(0.12 * n) + 4096)
acc = collections.defaultdict(float)
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        acc[name] += time.perf_counter() - t0; return r
    return w
def step(): slab.slab_step(state, geom, fabric, pipe.compute, field, pipe, ops, morphogen_first=first)
for _ in range(3): step()
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); free = (time.perf_counter() - t0) / 10
# host-side cost of each phase (no synchronisation: what the CPU spends enqueueing)
host = collections.defaultdict(float)
def hosted(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); host[name] += time.perf_counter() - t0; return r
    return w
saved = dict(outgoing=ops.outgoing, merge=ops.merge, exchange=fabric.exchange, compute=pipe.compute,
             deposit=pipe.deposit, diffuse=pipe.diffuse, sample=pipe.sample)
ops.outgoing = hosted("pack", ops.outgoing); ops.merge = hosted("merge(+count sync)", ops.merge)
fabric.exchange = hosted("exchange", fabric.exchange); pipe.compute = hosted("compute", pipe.compute)
pipe.deposit = hosted("deposit", pipe.deposit); pipe.diffuse = hosted("diffuse", pipe.diffuse)
pipe.sample = hosted("sample", pipe.sample)
dist.barrier(); t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); tot_h = (time.perf_counter() - t0) / 10
if rank == 0:
    print("host ms/step %.3f" % (tot_h * 1e3), {k: round(v / 10 * 1e3, 3) for k, v in host.items()},
          "other(+n_own sync) %.3f" % ((tot_h - sum(host.values()) / 10) * 1e3))
ops.outgoing, ops.merge, fabric.exchange, pipe.compute = saved["outgoing"], saved["merge"], saved["exchange"], saved["compute"]
pipe.deposit, pipe.diffuse, pipe.sample = saved["deposit"], saved["diffuse"], saved["sample"]
ops.outgoing = timed("pack", ops.outgoing); ops.merge = timed("merge", ops.merge)
fabric.exchange = timed("exchange", fabric.exchange)
pipe.compute = timed("compute", pipe.compute); pipe.deposit = timed("deposit", pipe.deposit)
pipe.diffuse = timed("diffuse", pipe.diffuse); pipe.sample = timed("sample", pipe.sample)
dist.barrier(); t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 10
if rank == 0:
    print("free-running ms/step %.3f; sync-timed %.3f" % (free * 1e3, tot * 1e3),
          {k: round(v / 10 * 1e3, 3) for k, v in acc.items()}, "other %.3f" % ((tot - sum(acc.values()) / 10) * 1e3))
dist.barrier(); dist.destroy_process_group()
