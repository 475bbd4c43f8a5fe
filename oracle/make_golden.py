"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported read-only from
/root/reference/ABM) on seeded inputs.  Run in the build container only:

    NUMBA_NUM_THREADS=1 python oracle/make_golden.py

The reference cannot travel to the GPU box, so the vectors are committed; this script is the
provenance.  Recipe = SURVEY.md Appendix B: an ``igraph`` stand-in (oracle.HostGraph, which defines
edge order = sorted (min,max) and neighbors = ascending), ``pythonabm`` aliased to the in-tree
backend/simulation modules, CWD with templates/, explicit seeding, single Numba thread.
"""
import os

os.environ.setdefault("NUMBA_NUM_THREADS", "1")   # the parallel force kernel races

import contextlib
import io
import random
import shutil
import sys
import tempfile
import types

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/ABM"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.oracle import HostGraph, canonical_edges  # noqa: E402


def import_reference():
    """Import the reference modules with the igraph / pythonabm shims in place."""
    ig = types.ModuleType("igraph")
    ig.Graph = HostGraph
    sys.modules["igraph"] = ig
    sys.path.insert(0, REF)
    import backend  # noqa
    import simulation  # noqa
    pk = types.ModuleType("pythonabm")
    pk.backend, pk.simulation = backend, simulation
    sys.modules["pythonabm"] = pk
    sys.modules["pythonabm.backend"] = backend
    sys.modules["pythonabm.simulation"] = simulation
    import cell_backend  # noqa
    import cell_methods  # noqa
    import cell_simulation  # noqa
    return backend, simulation, cell_backend, cell_methods, cell_simulation


backend, simulation, cell_backend, cell_methods, cell_simulation = import_reference()
import numba  # noqa: E402

assert numba.get_num_threads() == 1, "golden vectors must be generated single-threaded"


class Stub(cell_methods.CellMethods, simulation.Simulation):
    """A reference Simulation that skips __init__ (no templates needed) and holds only hot-path state."""

    def __init__(self, locations, radii, size, max_radius=5.0):
        self.locations = np.array(locations, dtype=float)
        self.number_agents = len(self.locations)
        self.radii = np.array(radii, dtype=float)
        self.size = np.array(size, dtype=float)
        self.cuda = False
        self.method_times = {}
        self.max_radius = max_radius
        self.step_dt, self.move_dt = 1800, 180
        self.motility_forces = np.zeros((self.number_agents, 3))
        self.jkr_forces = np.zeros((self.number_agents, 3))
        self.neighbor_graph = backend.Graph(self.number_agents)
        self.jkr_graph = backend.Graph(self.number_agents)


def colony(seed, n, size, kind="uniform"):
    rng = np.random.default_rng(seed)
    size = np.array(size, dtype=float)
    loc = rng.random((n, 3)) * size
    radii = np.full(n, 5.0)
    if kind == "ragged":
        # mixed radii, a few cells pushed just outside the box (division displacement),
        # exact duplicates (mag == 0) and a dense clump so bins/rows overflow their first capacity
        radii = rng.uniform(5 / 2 ** 0.5, 5.0, n)
        k = max(4, n // 50)
        loc[:k, 0] = -rng.random(k) * 1.4
        loc[k:2 * k, 1] = size[1] + rng.random(k) * 1.4
        loc[2 * k:3 * k] = loc[3 * k:4 * k]
        loc[4 * k:6 * k, :2] = size[:2] * 0.5 + rng.normal(0, 4.0, (2 * k, 2))
        if size[2] > 0:
            loc[4 * k:6 * k, 2] = size[2] * 0.5 + rng.normal(0, 4.0, 2 * k)
    return loc, radii, size


def kernel_case(name, seed, n, size, kind):
    loc, radii, size = colony(seed, n, size, kind)
    out = {"locations": loc, "radii": radii, "size": size}
    sim = Stub(loc, radii, size)
    # a1/a2: assign_bins (reference simulation.py:190) at r=15, starting capacity 1
    bins, bins_help, bin_locations, max_agents = sim.assign_bins(1, 15)
    out.update(bins15=bins, bins_help15=bins_help, bin_locations15=bin_locations, max_agents15=max_agents)
    # a3/a4: get_neighbors clear=True at r=15 and r=10
    for r in (15, 10):
        g = backend.Graph(sim.number_agents)
        sim.probe = g
        sim.get_neighbors("probe", r)
        out[f"edges{r}"] = canonical_edges(g.edges)
        out[f"max_neighbors{r}"] = g.max_neighbors
        out[f"max_agents{r}"] = g.max_agents
    # a6/a7: one calculate_jkr on the r=10 contact graph
    sim.get_neighbors("jkr_graph", 10, clear=False)
    out["jkr_edges_in"] = sim.jkr_graph.edges.copy()
    sim.calculate_jkr()
    out["jkr_forces"] = sim.jkr_forces.copy()
    out["jkr_edges_out"] = sim.jkr_graph.edges.copy()
    # a9: apply_forces_cpu with dt=180 and dt=0 (reference cell_backend.py:148)
    rng = np.random.default_rng(seed + 100)
    mot = rng.normal(0, 2e-9, (sim.number_agents, 3))
    if size[2] == 0:
        mot[:, 2] = 0
    out["motility"] = mot
    for dt in (180, 0):
        new = cell_backend.apply_forces_cpu(sim.number_agents, sim.jkr_forces.copy(), mot, loc.copy(), radii,
                                            10000, size, dt)
        out[f"integrated_dt{dt}"] = new
    # a8: the whole 11-substep apply_forces() from the same start (fresh sim, hysteresis exercised)
    sim2 = Stub(loc, radii, size)
    sim2.motility_forces = mot.copy()
    sim2.apply_forces()
    out["apply_locations"] = sim2.locations.copy()
    out["apply_jkr_edges"] = sim2.jkr_graph.edges.copy()
    # ... and a second call, so that edges persisting from the first call matter
    sim2.motility_forces = -mot
    sim2.apply_forces()
    out["apply2_locations"] = sim2.locations.copy()
    out["apply2_jkr_edges"] = sim2.jkr_graph.edges.copy()
    np.savez_compressed(os.path.join(OUT, f"kernels_{name}.npz"), **out)
    print(name, "n", n, "edges15", len(out["edges15"]), "edges10", len(out["edges10"]), "jkr_out",
          len(out["jkr_edges_out"]), "apply edges", len(out["apply_jkr_edges"]))


def diffusion_case(name, seed, gx, gy, steps, a_coef, last_frac, n_cells, h, exact):
    rng = np.random.default_rng(seed)
    D = 50.0
    dt = a_coef * h * h / D
    out = {}
    # a11: raw JIT on a ghost-padded array
    base = np.pad(rng.random((gx, gy)) * 3 - 0.5, 1)
    out["base_in"] = base.copy()
    out["base_out"] = cell_backend.update_diffusion_jit(base.copy(), steps, dt, dt * last_frac, D, h * h)
    out["jit_args"] = np.array([steps, dt, dt * last_frac, D, h * h])
    # a10: the method (divmod substep count, clamp, degrade)
    # exact: size is an exact multiple of h, so cells sitting on the upper boundary have corner
    # points outside the grid (the bounds check of cell_methods.py:506)
    eps = 0.0 if exact else 1e-9
    size = np.array([(gx - 1) * h - eps, (gy - 1) * h - eps, 0.0])
    loc = rng.random((n_cells, 3)) * size
    loc[: n_cells // 20, 0] = size[0]           # cells on the upper boundary -> corner points out of bounds
    loc[n_cells // 20: n_cells // 10, 1] = 0.0
    sim = Stub(loc, np.full(n_cells, 5.0), size)
    sim.spat_res, sim.spat_res2, sim.diffuse_const = h, h * h, D
    sim.diffuse_dt = dt
    sim.step_dt = dt * (steps - 1) + dt * last_frac
    sim.max_concentration, sim.degradation = 2, 0.1
    sim.gradient_size = np.ceil(size / h).astype(int) + 1
    assert tuple(sim.gradient_size) == (gx, gy, 1), sim.gradient_size
    sim.fgf4_values = (rng.random((gx, gy, 1)) * 3 - 0.5)
    out["field_in"] = sim.fgf4_values.copy()
    out["method_args"] = np.array([sim.step_dt, sim.diffuse_dt, D, h, 2, 0.1])
    sim.update_diffusion("fgf4_values")
    out["field_out"] = sim.fgf4_values.copy()
    # a12/a13 per cell, sample-all-then-deposit-all (the batched contract of update_concentrations)
    out["locations"] = loc
    out["size"] = size
    out["sampled"] = np.array([sim.get_concentration("fgf4_values", i) for i in range(n_cells)])
    amounts = rng.random(n_cells)
    out["amounts"] = amounts
    for i in range(n_cells):
        sim.adjust_morphogens("fgf4_values", i, amounts[i])
    out["field_deposited"] = sim.fgf4_values.copy()
    np.savez_compressed(os.path.join(OUT, f"diffusion_{name}.npz"), **out)
    print(name, "grid", gx, gy, "steps", steps, "sum in/out", out["field_in"].sum(), out["field_out"].sum())


def full_model(name, seed, n, size, end_step, snap_every):
    """C1: the reference's own CellSimulation.steps() (cell_simulation.py:77-126), seeded."""
    work = tempfile.mkdtemp(prefix="golden_")
    try:
        os.makedirs(os.path.join(work, "templates"))
        with open(os.path.join(REF, "templates", "general.yaml")) as f:
            general = yaml.safe_load(f)
        with open(os.path.join(REF, "templates", "experimental.yaml")) as f:
            experimental = yaml.safe_load(f)
        general.update(num_to_start=n, cuda=False, end_step=end_step, size=list(size), output_values=False,
                       output_images=False, record_initial_step=False)
        experimental.update(num_gata6=0, output_tda=False, output_gradients=False, dox_step=5)
        for fname, keys in (("general.yaml", general), ("experimental.yaml", experimental)):
            with open(os.path.join(work, "templates", fname), "w") as f:
                yaml.safe_dump(keys, f)
        outdir = os.path.join(work, "out") + os.sep
        os.makedirs(os.path.join(outdir, name))
        snaps = {}

        class Recorder(cell_simulation.CellSimulation):
            def temp(self):          # skip the per-step pickle
                pass

            def data(self):          # last call of every step (cell_simulation.py:123)
                s = self.current_step
                snaps[f"n_{s}"] = self.number_agents
                for arr in ("states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK"):
                    snaps[f"{arr}_{s}"] = self.__dict__[arr].astype(np.int8)
                snaps[f"nbr_edges_{s}"] = self.neighbor_graph.ecount()
                snaps[f"jkr_edges_{s}"] = self.jkr_graph.ecount()
                if s % snap_every == 0 or s == 1 or s == self.end_step:
                    snaps[f"locations_{s}"] = self.locations.copy()
                    snaps[f"div_counters_{s}"] = self.div_counters.copy()

        cwd = os.getcwd()
        os.chdir(work)
        try:
            random.seed(seed)
            np.random.seed(seed)
            sim = Recorder(name, outdir)
            sim.agent_initials()
            snaps["locations_0"] = sim.locations.copy()
            with contextlib.redirect_stdout(io.StringIO()):
                sim.steps()
        finally:
            os.chdir(cwd)
        snaps["meta"] = np.array([seed, n, end_step, snap_every] + list(size), dtype=float)
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **snaps)
        print(name, "final agents", sim.number_agents, "sum states", int(sim.states.sum()),
              "nbr edges", sim.neighbor_graph.ecount(), "jkr edges", sim.jkr_graph.ecount())
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    kernel_case("2d_uniform", 0, 1500, [387.3, 387.3, 0], "uniform")
    kernel_case("2d_ragged", 1, 1200, [300, 260, 0], "ragged")
    kernel_case("3d_uniform", 2, 1500, [114.5, 114.5, 114.5], "uniform")
    kernel_case("3d_ragged", 3, 1000, [90, 100, 80], "ragged")
    kernel_case("tiny", 4, 3, [40, 40, 0], "uniform")
    diffusion_case("small", 5, 23, 31, 7, 0.24, 0.37, 400, 7.07106, False)
    diffusion_case("odd", 6, 64, 49, 12, 0.2, 0.0, 300, 7.0, True)
    full_model("c1_20", 1234, 1000, [400, 400, 0], 20, 5)
    full_model("c1_100", 1234, 1000, [400, 400, 0], 100, 10)
    full_model("c1_3d_10", 77, 600, [150, 150, 150], 10, 5)
