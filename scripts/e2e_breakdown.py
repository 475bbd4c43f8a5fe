"""Where the end-to-end (host numpy arrays) step goes: wall time per public call of the mixin API, bytes each call
moves, and the PCIe rates of this box for the same transfer sizes (GPU box).
   python scripts/e2e_breakdown.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from hipsc_abm_b200.cell_methods import B200CellMethods
from hipsc_abm_b200.simulation import ColonySimulation

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = bench.make_workload(name); spec = wl["spec"]; n = wl["n"]
dev = torch.device("cuda:0")


def rate(nbytes, pinned_numpy=False):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    devt = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for label, f in (("h2d", lambda: devt.copy_(host, non_blocking=True)), ("d2h", lambda: host.copy_(devt, non_blocking=True))):
        f(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            f()
        torch.cuda.synchronize()
        out[label] = round(nbytes * 10 / (time.perf_counter() - t) / 1e9, 1)
    return out


print("PCIe GB/s, pinned torch buffers:", {f"{mb}MB": rate(mb << 20) for mb in (4, 12, 64)})


class Cells(B200CellMethods, ColonySimulation):
    pass


sim = Cells(wl["size"], dev)
sim.add_agents(n)
sim.agent_array("locations", override=wl["loc"].copy())
sim.agent_array("radii", override=wl["radii"].copy())
sim.agent_array("motility_forces", vector=3)
sim.agent_array("jkr_forces", vector=3)
sim.agent_graph("neighbor_graph")
sim.agent_graph("jkr_graph")
sim.max_radius, sim.step_dt, sim.move_dt = 5.0, 1800, 180
h = spec["h"]
sim.spat_res, sim.spat_res2, sim.diffuse_const = h, h * h, 50.0
sim.diffuse_dt = spec["diffuse_dt"]
sim.max_concentration, sim.degradation = 2.0, 0.1
sim.fgf4_values = np.zeros(wl["grid"] + (1,))
eng = sim.b200_engine()
acc = {}


def timed(label, f):
    h0, d0 = eng.h2d_bytes, eng.d2h_bytes
    torch.cuda.synchronize()
    t = time.perf_counter()
    f()
    torch.cuda.synchronize()
    a = acc.setdefault(label, [0.0, 0, 0])
    a[0] += time.perf_counter() - t
    a[1] += eng.h2d_bytes - h0
    a[2] += eng.d2h_bytes - d0


def step():
    timed("get_neighbors(15)", lambda: sim.get_neighbors("neighbor_graph", 15))
    timed("(input fill: motility_forces[:] = ...)", lambda: sim.motility_forces.__setitem__(slice(None), wl["mot"]))
    timed("apply_forces", sim.apply_forces)
    keep = sim.step_dt
    sim.step_dt = spec["diffuse_dt"] * (spec["sd"] - 1) + 0.5 * spec["diffuse_dt"]
    timed("update_concentrations(deposit)", lambda: sim.update_concentrations("fgf4_values", 1.0, sample=False))
    timed("update_diffusion", lambda: sim.update_diffusion("fgf4_values"))
    timed("update_concentrations(sample)", lambda: sim.update_concentrations("fgf4_values", None))
    sim.step_dt = keep


for _ in range(3):
    step()
acc.clear()
K = 10
for _ in range(K):
    step()
print("| call | ms | H2D MB | D2H MB |\n|---|---|---|---|")
for k, (t, hb, db) in acc.items():
    print(f"| {k} | {t / K * 1e3:.2f} | {hb / K / 1e6:.1f} | {db / K / 1e6:.1f} |")
print("total ms", round(sum(v[0] for v in acc.values()) / K * 1e3, 2))

if os.environ.get("E2E_PROFILE"):                      # host-side profile of the public calls
    import cProfile, pstats
    pr = cProfile.Profile()
    for _ in range(10):
        sim.get_neighbors("neighbor_graph", 15)
        sim.motility_forces[:] = wl["mot"]
        torch.cuda.synchronize()
        pr.enable()
        sim.apply_forces()
        pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
