"""Event timeline of the free-running slab step over NCCL (torchrun, 2+ ranks): where the main stream waits.
Events are recorded on the main stream after each phase's work was ENQUEUED, so the gaps show device time."""
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
wl = bench.make_workload("c3", seed=rank); spec, n = wl["spec"], wl["n"]
cfg = StepConfig(grid=wl["grid"], spat_res=spec["h"], diffuse_dt=spec["diffuse_dt"], diffuse_steps=spec["sd"])
gsize = wl["size"].copy(); gsize[0] *= world
halo = slab.required_halo(11, 10.3100107, slack=4.0) + spec["h"]
geom = slab.SlabGeometry.uniform(gsize, world, rank, halo)
pipe = SlabPipeline(geom, cfg, dev, uniform_radius=5.0, trailing=os.environ.get("TRAILING", "1") == "1")
loc = wl["loc"].copy(); loc[:, 0] += geom.x_lo
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
state = slab.SlabState(t(loc), t(wl["radii"]), t(wl["mot"]), t(np.arange(n) + rank * n), 16)
gx = int(np.ceil(gsize[0] / spec["h"])) + 1
field = slab.FieldSlab(geom, (gx, wl["grid"][1]), spec["h"], spec["sd"] + 1, dev)
fabric = slab.DistFabric(rank, world, dev, rows_group=dist.new_group(list(range(world)))); ops = NativeSlabOps(pipe.eng, send_cap=int(0.05 * n) + 4096)
main = torch.cuda.Stream(device=dev, priority=-1); torch.cuda.set_stream(main)
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.current_stream(dev)); marks.append((name, e, time.perf_counter()))
def wrap(name, fn):
    def w(*a, **k):
        r = fn(*a, **k); mark(name); return r
    return w
def step(): slab.slab_step(state, geom, fabric, pipe.compute, field, pipe, ops, morphogen_first=True)
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
ops.outgoing = wrap("pack", ops.outgoing); ops.merge = wrap("merge", ops.merge)
fabric.exchange = wrap("exchange", fabric.exchange); pipe.compute = wrap("compute", pipe.compute)
pipe.sample = wrap("side_done", pipe.sample)
K = 8
for _ in range(K):
    mark("start"); step()
mark("end"); torch.cuda.synchronize()
if rank == 0:
    t0e, t0h = marks[0][1], marks[0][2]
    rows = [(nm, t0e.elapsed_time(e), (h - t0h) * 1e3) for nm, e, h in marks]
    per = len(rows) // K
    print("phase: device-ms (since run start) | host-ms; steps 3..4")
    for nm, d, h in rows[3 * per:5 * per + 1]:
        print(f"  {nm:10s} dev {d:8.3f}  host {h:8.3f}")
dist.barrier(); dist.destroy_process_group()
