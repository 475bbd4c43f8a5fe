"""cProfile of the reference-facing API step (host numpy arrays in/out) on a bench workload: where the e2e time goes."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import hipsc_abm_b200  # noqa
from hipsc_abm_b200.simulation import ColonySimulation
from hipsc_abm_b200.cell_methods import B200CellMethods

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = bench.make_workload(name); spec = wl["spec"]; n = wl["n"]
class Cells(B200CellMethods, ColonySimulation):
    pass
sim = Cells(wl["size"], torch.device("cuda"))
sim.add_agents(n)
sim.agent_array("locations", override=wl["loc"].copy()); sim.agent_array("radii", override=wl["radii"].copy())
sim.agent_array("motility_forces", vector=3); sim.agent_array("jkr_forces", vector=3)
sim.agent_graph("neighbor_graph"); sim.agent_graph("jkr_graph")
sim.max_radius, sim.step_dt, sim.move_dt = 5.0, 1800, 180
h = spec["h"]
sim.spat_res, sim.spat_res2, sim.diffuse_const = h, h * h, 50.0
sim.diffuse_dt = spec["diffuse_dt"]; sim.max_concentration, sim.degradation = 2.0, 0.1
sim.fgf4_values = np.zeros(wl["grid"] + (1,))
def api_step():
    sim.get_neighbors("neighbor_graph", 15)
    sim.motility_forces[:] = wl["mot"]
    sim.apply_forces()
    keep = sim.step_dt
    sim.step_dt = spec["diffuse_dt"] * (spec["sd"] - 1) + 0.5 * spec["diffuse_dt"]
    sim.update_concentrations("fgf4_values", 1.0, sample=False)
    sim.update_diffusion("fgf4_values")
    sim.update_concentrations("fgf4_values", None)
    sim.step_dt = keep
for _ in range(3): api_step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): api_step()
torch.cuda.synchronize(); print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
print({k: round(v, 4) for k, v in getattr(sim, "method_times", {}).items()})
pr = cProfile.Profile(); pr.enable()
for _ in range(5): api_step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
