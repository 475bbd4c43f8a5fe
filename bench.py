#!/usr/bin/env python
"""bench.py -- cell-updates/sec of the neighbor+force+diffusion step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

Default workload c3 = BASELINE.json configs[2] (500,000-cell 2-D colony + FGF4 diffusion): the README's headline
scale and the only config the metric is quoted on at both 1 and 8 B200; c2 (50k cells, ~5 MB of state) cannot
load a B200 -- every kernel there is launch-latency bound.

One "step" = one reference simulation step of the hot path on a synthetic seeded colony
(SURVEY 8d): 1 x search(r=15) + 11 x [sort + fused search(r=10)/JKR/integrate substep] + deposit +
S_d diffusion substeps + sample.  Prints ONE JSON line (rank 0).

  value     whole-job throughput with everything resident in HBM (CUDA-graph replay, device-timed)
  e2e       same metric through the reference-facing Python API with HOST numpy arrays (H2D/D2H inside)
  roofline  dominant kernel (k_jkr_substep): SURVEY 8(d) algorithmic bytes of the stages one launch fuses / event-timed
            launch duration after an L2 flush (the smaller compulsory-byte figure of the fused kernel is reported too)
  cpu_baseline / --impl reference : the oracle port (oracle/abm_oracle.c, OpenMP) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-updates/sec (neighbor+force+diffusion step)"
MOTILITY_NOTE = ("2 nN per cell in a fixed random direction, sign alternating from step to step (the colony stays "
                 "statistically stationary for any number of steps)")
UNIT = "cell-updates/s"

WORKLOADS = {
    # name: (cells, dim, density, grid spec)          SURVEY 8(d) / BASELINE.json configs[1..4]
    "c2": dict(n=50_000, dim=2, density=0.01, h=7.07106, diffuse_dt=0.24, sd=10,
               desc="configs[1]: 50,000-cell 2D colony, JKR adhesion + FGF4 diffusion (318^2 grid, 10 substeps)"),
    "c3": dict(n=500_000, dim=2, density=0.01, h=7.07106, diffuse_dt=0.24, sd=10,
               desc="configs[2]: 500,000-cell 2D colony + FGF4 diffusion (1002^2 grid, 10 substeps)"),
    "c4": dict(n=4_000_000, dim=3, density=0.001, h=7.07106, diffuse_dt=0.15 * 7.07106 ** 2 / 50.0, sd=10,
               desc="configs[3]: 4,000,000-cell 3D colony + FGF4 diffusion on a 3D grid (226^3, 10 substeps; the 3-D "
                    "stencil is this build's extension, the reference's diffusion is 2-D only)"),
    "c5": dict(n=100_000, dim=2, density=0.01, h=1.5441, diffuse_dt=0.011445, sd=10,
               desc="configs[4]: 100,000 cells on a fine 2048^2 morphogen grid, 10 diffusion substeps"),
}
W = 8  # bytes per real: the path computes in float64, the reference's dtype


def make_workload(name, seed=0, n_override=None):
    spec = dict(WORKLOADS[name])
    n = int(n_override or spec["n"])
    rng = np.random.default_rng(seed)
    if spec["dim"] == 2:
        side = (n / spec["density"]) ** 0.5
        size = np.array([side, side, 0.0])
    else:
        side = (n / spec["density"]) ** (1.0 / 3.0)
        size = np.array([side, side, side])
    loc = rng.random((n, 3)) * size
    theta = rng.random(n) * 2 * np.pi
    mot = np.zeros((n, 3))
    if spec["dim"] == 2:
        mot[:, 0], mot[:, 1] = np.cos(theta), np.sin(theta)
    else:
        phi = rng.random(n) * 2 * np.pi
        mot[:, 0], mot[:, 1], mot[:, 2] = np.cos(phi) * np.cos(theta), np.cos(phi) * np.sin(theta), np.sin(phi)
    mot *= 2e-9                                     # motility force, reference cell_methods.py:252
    # The reference draws a new direction for (most) cells every step.  The bench keeps ONE direction per cell and
    # alternates its sign from step to step (every arm: B200, slabs, end-to-end API, CPU): with a constant force
    # the cells travel 3.8 um per step, pile up against the walls and each other, and after ~190 steps a cell has
    # more than 16 bonds (checked with the oracle); with the alternating sign bonds and neighbours per cell settle
    # (3.9 and 5.7 on average, at most 6 and 9) and stay there.
    grid = None
    if spec["h"]:
        g = int(np.ceil(side / spec["h"])) + 1      # gradient_size, reference cell_simulation.py:67
        g += g % 2                                  # even row length (TMA row pitch multiple of 16 B)
        grid = (g, g) if spec["dim"] == 2 else (g, g, g)
    return dict(name=name, n=n, size=size, loc=loc, radii=np.full(n, 5.0), mot=mot, grid=grid, spec=spec)


# ---- clocks -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.rows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- byte accounting -----------------------------------------------------------------------------
def step_bytes(n, b15, b10, e15, e10, ej, g, sm, sd, w_mech=W):
    """SURVEY 8(d) bytes of one step.  w_mech: bytes per real inside the mechanics substeps (4 in fp32 mode; the
    r=15 search, the diffusion and the per-cell coupling always compute in float64)."""
    search15 = 7 * W * n + 4 * n + 8 * b15 + 4 * W * n + 4 * b15 + 8 * e15
    w = w_mech
    substep = (7 * w * n + 4 * n + 8 * b10) + (4 * w * n + 4 * b10 + 8 * e10) + 8 * (2 * ej + e10) + \
              (9 * ej + 4 * w * n) + 6 * w * n
    return search15 + sm * substep + sd * 2 * W * g + 5 * W * n


def substep_kernel_bytes(n, b10, ej):
    """Compulsory traffic of ONE k_jkr_substep launch (DESIGN.md 'kernels'): read sorted {x,y,z,r} 4wN,
    sorted_id+sorted_bin 8N, bin offsets 4B, motility 3wN; write locations 3wN, bond rows 2*4*E_j, row_len 4N."""
    return 10 * W * n + 12 * n + 4 * b10 + 8 * ej


def substep_algorithmic_bytes(n, b10, e10, ej, w=W):
    """SURVEY 8(d)'s algorithmic bytes of the stages ONE k_jkr_substep launch replaces: search(r=10)
    4wN+4B+8E, merge/dedupe 8(2E_j+E), force 9E_j+4wN, integrate 6wN (the sort's 7wN+4N+8B belongs to the sort
    kernels)."""
    return (4 * w * n + 4 * b10 + 8 * e10) + 8 * (2 * ej + e10) + (9 * ej + 4 * w * n) + 6 * w * n


def sort_algorithmic_bytes(n, b10, w=W):
    """SURVEY 8(d) bin/sort: 7wN + 4N + 8B."""
    return 7 * w * n + 4 * n + 8 * b10


def host_field(wl):
    """The morphogen field as the reference holds it: (gx, gy, 1) for the 2-D colony (cell_simulation.py:67-69), a
    full (gx, gy, gz) volume for the 3-D extension."""
    return np.zeros(wl["grid"] + (1,)) if len(wl["grid"]) == 2 else np.zeros(wl["grid"])


def oracle_morphogen_stage(orc, sim, wl, last_dt=None):
    """deposit -> diffusion -> per-cell sampling on the CPU oracle; returns the sampled concentrations.
    last_dt: dt of the last diffusion substep (default: a full diffuse_dt, what HotPathStep runs; the reference-facing
    update_diffusion takes it from divmod(step_dt, diffuse_dt), cell_methods.py:447-448)."""
    spec = wl["spec"]
    args = (spec["sd"], sim.diffuse_dt, sim.diffuse_dt if last_dt is None else last_dt, sim.diffuse_const,
            sim.spat_res2, sim.max_concentration, sim.degradation)
    if len(wl["grid"]) == 3:
        orc.adjust_morphogens3d(sim.locations, 1.0, sim.fgf4_values, sim.spat_res)
        sim.fgf4_values = orc.diffuse_field3d(sim.fgf4_values, *args)
    else:
        orc.adjust_morphogens(sim.locations, 1.0, sim.fgf4_values, sim.spat_res)
        sim.fgf4_values[:, :, 0] = orc.diffuse_field(sim.fgf4_values[:, :, 0], *args)
    return orc.get_concentrations(sim.locations, sim.fgf4_values, sim.spat_res)


# ---- the CPU arm (oracle port) ----------------------------------------------------------------------------------
def oracle_step_fn(wl, threads):
    from oracle import oracle as orc
    orc.set_threads(threads)
    sim = orc.OracleSim(wl["loc"], wl["radii"], wl["size"])
    sim.multithread = threads > 1
    spec = wl["spec"]
    if wl["grid"]:
        h = spec["h"]
        sim.spat_res, sim.spat_res2, sim.diffuse_const = h, h * h, 50.0
        sim.diffuse_dt = spec["diffuse_dt"]
        sim.max_concentration, sim.degradation = 2.0, 0.1
        sim.fgf4_values = host_field(wl)

    count = [0]

    def step():
        sim.get_neighbors("neighbor_graph", 15)
        sim.motility_forces[:] = wl["mot"] if count[0] % 2 == 0 else -wl["mot"]
        count[0] += 1
        sim.apply_forces()
        if wl["grid"]:
            oracle_morphogen_stage(orc, sim, wl)
    return step


def host_cores():
    """Host threads the CPU arm uses: the cores this process may run on.  Set explicitly on the OpenMP runtime --
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would silently make the reference arm a
    single-threaded one at N > 1."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# cell-steps one CPU leg may cost (16 cores run ~0.15 M cell-steps/s in 2-D, ~0.06 M in 3-D): the full 500,000-cell
# colony of c3 fits `--steps 20 --warmup 5` (about 85 s); only c4's 4 M cells are sampled
CPU_BUDGET_CELL_STEPS = {2: 14_000_000, 3: 5_000_000}


def cpu_sample_cells(name, total_steps):
    spec = WORKLOADS[name]
    return int(min(spec["n"], max(100_000, CPU_BUDGET_CELL_STEPS[spec["dim"]] // max(total_steps, 1))))


def workload_label(name, cells=None, world=1, cpu_arm=False):
    """config.workload.  The CPU arm runs the stated colony whenever it fits its time budget; when it has to run a
    smaller sample of it (c4, or the N-GPU colony at N > 1), the label says so -- the two arms' configs then differ,
    as they should."""
    spec = WORKLOADS[name]
    label = f"{name}: {spec['desc']}" if world == 1 else f"{name} x {world}: {spec['desc']} per GPU, one colony"
    if cpu_arm and (cells != spec["n"] or world > 1):
        label += f" -- CPU arm on a {cells}-cell sample of it (same density, geometry per cell and step recipe)"
    return label


def common_config(name, wl, sm, world=1, cpu_arm=False, scaling="weak"):
    """The part of `config` that defines the workload: identical, key for key, in the B200 line and in the
    `--impl reference` line whenever both run the same colony."""
    spec = WORKLOADS[name]
    if world > 1 and scaling == "strong":
        return {"workload": f"{name}: {spec['desc']}, one colony slab-decomposed over {world} GPUs",
                "cells_per_gpu": wl["n"] // world, "dim": spec["dim"], "mechanics_substeps": sm,
                "diffusion_substeps": spec["sd"], "grid": list(wl["grid"]) if wl["grid"] else None,
                "motility": MOTILITY_NOTE,
                "l2": "B200 arm: flushed between timed steps (256 MiB memset); CPU arm: not applicable"}
    return {"workload": workload_label(name, wl["n"], world, cpu_arm), "cells_per_gpu": wl["n"], "dim": spec["dim"],
            "mechanics_substeps": sm, "diffusion_substeps": spec["sd"],
            "grid": list(wl["grid"]) if wl["grid"] else None, "motility": MOTILITY_NOTE,
            "l2": "B200 arm: flushed between timed steps (256 MiB memset); CPU arm: not applicable"}


def run_cpu(name, steps, warmup, threads=None, cells=None):
    from oracle import oracle as orc
    cells = cells or cpu_sample_cells(name, steps + warmup)
    wl = make_workload(name, seed=0, n_override=cells)
    threads = threads or host_cores()
    step = oracle_step_fn(wl, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    full = cells == WORKLOADS[name]["n"]
    sample = ((f"the whole workload {name} ({cells} cells)" if full else
               f"{cells} cells of workload {name} (same density and step recipe)") +
              f", {steps} steps after {warmup} warm-up, oracle/abm_oracle.c with OpenMP on {threads} threads")
    return wl["n"] * steps / dt, dt / steps, threads, sample, wl


def parity_check(name, dev, steps=5, cells=100_000, dtype="f64"):
    """The colony bench.py times is not compared with anything while it is timed; this runs the same step recipe on
    a `cells`-cell sample of the workload through BOTH arms (the B200 path through the reference-facing API, the CPU
    oracle with its serial force pass) for a few steps and reports how far they are apart.  The gates are PER STEP
    (north star: "per-step forces and positions"): every step starts from the oracle's positions (teacher forcing;
    the bond and neighbour graphs must be identical anyway), because the colony amplifies last-bit differences by
    about x10 every 15 substeps (tests/test_model_states.py); the free-running drift is reported beside it."""
    from oracle import oracle as orc
    from hipsc_abm_b200.cell_methods import B200CellMethods
    from hipsc_abm_b200.simulation import ColonySimulation
    wl = make_workload(name, seed=7, n_override=min(cells, WORKLOADS[name]["n"]))
    spec, n = wl["spec"], wl["n"]

    class Cells(B200CellMethods, ColonySimulation):
        pass
    sim = Cells(wl["size"], dev)
    sim.b200_dtype = dtype
    sim.add_agents(n)
    sim.agent_array("locations", override=wl["loc"].copy())
    sim.agent_array("radii", override=wl["radii"].copy())
    sim.agent_array("motility_forces", vector=3)
    sim.agent_array("jkr_forces", vector=3)
    sim.agent_graph("neighbor_graph")
    sim.agent_graph("jkr_graph")
    sim.max_radius, sim.step_dt, sim.move_dt = 5.0, 1800, 180
    orc.set_threads(host_cores())
    ref = orc.OracleSim(wl["loc"], wl["radii"], wl["size"])
    ref.multithread = False                       # serial force pass: the reference's parallel one races
    if wl["grid"]:
        h = spec["h"]
        for o in (sim, ref):
            o.spat_res, o.spat_res2, o.diffuse_const, o.diffuse_dt = h, h * h, 50.0, spec["diffuse_dt"]
            o.max_concentration, o.degradation = 2.0, 0.1
            o.fgf4_values = host_field(wl)
        sim.b200_diffusion_3d = len(wl["grid"]) == 3
    free = Cells(wl["size"], dev)                 # the same colony, never reset to the oracle (free-running drift)
    free.b200_dtype = dtype
    free.add_agents(n)
    free.agent_array("locations", override=wl["loc"].copy())
    free.agent_array("radii", override=wl["radii"].copy())
    free.agent_array("motility_forces", vector=3)
    free.agent_array("jkr_forces", vector=3)
    free.agent_graph("jkr_graph")
    free.max_radius, free.step_dt, free.move_dt = 5.0, 1800, 180
    out = {"cells": n, "steps": steps, "neighbor_edges_equal": True, "jkr_bonds_equal": True, "max_abs_dx_um": 0.0,
           "rel_dx_per_step": [],
           "max_abs_dfield": 0.0 if wl["grid"] else None, "max_abs_dsample": 0.0 if wl["grid"] else None}
    # The synthetic colony starts from uniformly random positions, i.e. with a few hundred cell pairs almost on top of
    # each other: their first separation is ill-conditioned (direction = difference of two nearly equal positions,
    # times a repulsion of micrometres per substep), so ANY two fp64 evaluation orders differ there by more than the
    # gate.  The first `relax` steps let those pairs separate; they are compared like the others (edges must be
    # identical) and listed in rel_dx_per_step, but the position gate is taken over the steps after them.
    relax = 2
    out["relax_steps_before_position_gate"] = relax
    for k in range(relax + steps):
        mot = wl["mot"] if k % 2 == 0 else -wl["mot"]
        sim.get_neighbors("neighbor_graph", 15)
        ref.get_neighbors("neighbor_graph", 15)
        out["neighbor_edges_equal"] &= bool(np.array_equal(sim.neighbor_graph.edge_array(),
                                                           orc.canonical_edges(ref.neighbor_graph.edge_array())))
        sim.motility_forces[:] = mot
        ref.motility_forces[:] = mot
        sim.apply_forces()
        ref.apply_forces()
        free.motility_forces[:] = mot
        free.apply_forces()
        same = bool(np.array_equal(sim.jkr_graph.edge_array(), ref.jkr_graph.edge_array()))
        out["jkr_bonds_equal"] &= same
        if not same and dtype == "f32":           # fp32: pairs at a threshold may differ; teacher-force the bonds too
            a = set(map(tuple, sim.jkr_graph.edge_array().tolist()))
            b = set(map(tuple, ref.jkr_graph.edge_array().tolist()))
            out["jkr_bonds_differing"] = out.get("jkr_bonds_differing", 0) + len(a ^ b)
            sim.jkr_graph.delete_edges(None)
            sim.jkr_graph.add_edges(ref.jkr_graph.edge_array())
        dx = float(np.abs(sim.locations - ref.locations).max())
        out["rel_dx_per_step"].append(dx / float(np.abs(ref.locations).max()))
        if k >= relax:
            out["max_abs_dx_um"] = max(out["max_abs_dx_um"], dx)
        sim.locations = ref.locations.copy()      # teacher forcing: the next step starts from the oracle's positions
        if wl["grid"]:
            keep = sim.step_dt
            # update_diffusion runs divmod(step_dt, diffuse_dt)[0] + 1 substeps, the last one with the remainder as its
            # dt (cell_methods.py:447-448): sd substeps, the last of them half a diffuse_dt long -- on both arms
            sim.step_dt = spec["diffuse_dt"] * (spec["sd"] - 1) + 0.5 * spec["diffuse_dt"]
            q, last_dt = divmod(sim.step_dt, spec["diffuse_dt"])
            assert int(q) + 1 == spec["sd"]
            sim.update_concentrations("fgf4_values", 1.0, sample=False)
            sim.update_diffusion("fgf4_values")
            got = sim.update_concentrations("fgf4_values", None)
            sim.step_dt = keep
            want = oracle_morphogen_stage(orc, ref, wl, last_dt=last_dt)
            out["max_abs_dfield"] = max(out["max_abs_dfield"], float(np.abs(sim.fgf4_values - ref.fgf4_values).max()))
            out["max_abs_dsample"] = max(out["max_abs_dsample"], float(np.abs(np.asarray(got) - want).max()))
    scale = float(np.abs(ref.locations).max())
    out["max_rel_dx"] = out["max_abs_dx_um"] / scale
    out["free_running_abs_dx_um_after_all_steps"] = float(np.abs(free.locations - ref.locations).max())
    if dtype == "f32":
        out["gates"] = ("per step: neighbour edges identical; positions 1e-5 relative (fp32 mode); JKR bonds identical "
                        "up to pairs at a threshold (counted); field 1e-6")
        out["ok"] = bool(out["neighbor_edges_equal"] and out["max_rel_dx"] <= 1e-5 and
                         out.get("jkr_bonds_differing", 0) <= 1e-4 * max(len(ref.jkr_graph.edge_array()), 1) * steps and
                         (out["max_abs_dfield"] is None or out["max_abs_dfield"] <= 1e-6))
    else:
        out["gates"] = "per step: edges identical; positions 1e-12 relative (fp64 mode); field 1e-6"
        out["ok"] = bool(out["neighbor_edges_equal"] and out["jkr_bonds_equal"] and out["max_rel_dx"] <= 1e-12 and
                         (out["max_abs_dfield"] is None or out["max_abs_dfield"] <= 1e-6))
    return out


# ---- the B200 arm -------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import hipsc_abm_b200  # noqa: F401
    from hipsc_abm_b200.engine import ColonyEngine, GraphTensors
    from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
    from hipsc_abm_b200.simulation import ColonySimulation
    from hipsc_abm_b200.cell_methods import B200CellMethods

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"           # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # weak scaling: one workload-sized colony per GPU; strong scaling: every rank builds the same colony and keeps its slab
    wl = make_workload(args.workload, seed=rank if (world == 1 or args.scaling == "weak") else 0)
    spec = wl["spec"]
    n = wl["n"]
    cfg = StepConfig(grid=wl["grid"], spat_res=spec["h"] or 1.0, diffuse_dt=spec["diffuse_dt"] or 1.0,
                     diffuse_steps=spec["sd"], bond_stride=args.bond_stride or (16 if spec["dim"] == 2 else 32),
                     search_cols_per_cell=12.0 if spec["dim"] == 2 else 24.0,
                     overlap_morphogen=not args.no_overlap)
    if world > 1:
        line = run_b200_slabs(args, rank, world, dev, dist, barrier, wl, cfg)
        dist.barrier()
        dist.destroy_process_group()
        return line
    eng = ColonyEngine(wl["size"], dev)
    eng.mechanics_dtype = args.dtype
    w_mech = 4 if args.dtype == "f32" else W
    eng.upload(wl["loc"], wl["radii"], wl["mot"])
    # the motility forces alternate in sign from step to step (see make_workload): the colony then stays statistically
    # stationary, so the timed steps do the same work whatever --steps / --warmup say
    pipe = HotPathStep(eng, cfg, motility_by_parity=(eng.mot, -eng.mot))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    # warm-up (eager), then capture and a couple of replays
    for _ in range(2):
        pipe.step()
    pipe.check()
    if pipe.parity:
        pipe.step()
    pipe.capture()
    sampler = ClockSampler(local_rank)
    if not args.no_clocks:
        sampler.start()                   # sampled from the warm-up replays to the end of the timed steps
    # nvidia-smi takes a while to print its first row.  Keep the GPU busy meanwhile, but not with steps of the colony:
    # the motility forces are constant, so seconds of extra steps would pile the cells up against the walls
    t_wait = time.time()
    while not sampler.rows and sampler.proc is not None and time.time() - t_wait < 3.0:
        flush.zero_()
        torch.cuda.synchronize(dev)
    for _ in range(max(args.warmup, 3) + 40):           # 40 extra replays: a few 20 ms clock samples under load
        pipe.replay()
    pipe.check()
    if pipe._replays & 1:
        pipe.replay()

    # ---- value: K graph-replayed steps, each bracketed by events, L2 flushed between steps --------------------
    launches0 = eng.launches
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        pipe.replay()
        b.record()
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_s = sum(step_ms) / 1e3
    launches = eng.launches - launches0
    used_cols, worst_row = pipe.check()
    if dist is not None:
        t = torch.tensor([total_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_s = float(t.item())
    value = n * world * args.steps / total_s

    # ---- graph statistics for the byte model -----------------------------------------------------------------------
    e15 = used_cols // 2
    bonds = pipe.current_bonds()
    ej = int(bonds.row_len[:n].sum().item()) // 2
    b15 = int(np.prod(eng.bin_dims(cfg.search_radius)))
    reach = eng.jkr_reach(cfg.contact_radius)
    b10 = int(np.prod(eng.bin_dims(reach)))
    g = int(np.prod(wl["grid"])) if wl["grid"] else 0
    sm = len(pipe.dts)
    bytes_step = step_bytes(n, b15, b10, e15, ej, ej, g, sm, spec["sd"], w_mech)

    # ---- roofline of the dominant kernel: event-timed single launches, L2 flushed before each -----------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    eng.use_current_stream()
    srt = eng.sort(reach)
    b = eng._bond_bufs
    old = pipe.current_bonds()
    spare = [i for i in range(3) if b["cols"][i].data_ptr() != old.cols.data_ptr()][0]
    loc_save = eng.loc.clone()
    lib, _l = eng.lib, hipsc_abm_b200._lib

    # a flat world's substeps run the two-axis kernel (every z is exactly 0 after the first substep: jkr_plane.cuh)
    plane = spec["dim"] == 2 and eng.plane_kernel
    sub_flags = _l.SUBSTEP_FLAT_Z if plane else 0

    def launch_substep():
        _l.check(lib.abm_jkr_substep(
            srt.bin_start.data_ptr(), srt.sorted_id.data_ptr(), srt.sorted_pos.data_ptr(), srt.sorted_bin.data_ptr(),
            n, _l.i3(srt.dims), cfg.contact_radius, reach, 0, old.row_stride, old.row_len.data_ptr(),
            old.cols.data_ptr(), b["len"][spare].data_ptr(), b["cols"][spare].data_ptr(), cfg.bond_stride,
            b["max_len"].data_ptr(), sub_flags, eng.mot.data_ptr(), eng.loc.data_ptr(), 0.5, 1000.0, 0.000107, 10000.0,
            _l.d3(wl["size"]), 180.0, 0, eng.uniform_radius, None, None, 0, 0, 0, None, eng.stream), "abm_jkr_substep")

    def time_launches(cold, reps=20):
        out = []
        for _ in range(reps):
            if cold:
                flush.zero_()
            else:
                launch_substep()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            launch_substep()
            z.record()
            torch.cuda.synchronize(dev)
            out.append(a.elapsed_time(z) * 1e3)
        return float(np.median(out))

    if args.dtype == "f32":
        # the fp32 substep kernel has no stand-alone entry point: the dominant "kernel" of the fp32 line is the
        # mechanics loop, 11 x [scan + placement + k_jkr_substep32], timed as a whole and divided by the substeps
        def launch_substep():
            eng.enqueue_mechanics(old, cfg.contact_radius, pipe.dts, cfg.bond_stride, zero_flag=False)
    for _ in range(3):
        launch_substep()
    cold_us = time_launches(True)
    hot_us = time_launches(False)
    eng.loc.copy_(loc_save)

    # ---- the same kernel timed INSIDE a running step: every k_jkr_substep launch of a mechanics loop bracketed by CUDA
    # events on the launching stream (abm_mechanics_loop_timed).  As in the timed region, the L2 is flushed before the
    # step and the loop starts from the cells its predecessor handed over; the first launch of a step finds a cold
    # L2, the other ten find what scan / placement / the previous substep left there.
    in_step = None
    if args.dtype != "f32":
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * sm)]
        for e in evs:
            e.record()                              # (creates the handles)
        per_launch = []
        bonds_save = (old.row_len.clone(), old.cols.clone())
        for _ in range(6):
            eng.loc.copy_(loc_save)
            eng._carry = None
            old.row_len.copy_(bonds_save[0])        # (the timed loop below may write the row buffer `old` lives in)
            old.cols.copy_(bonds_save[1])
            prev = eng.enqueue_mechanics(old, cfg.contact_radius, pipe.dts, cfg.bond_stride, zero_flag=False, carry=True)
            flush.zero_()
            eng.substep_events = evs
            try:
                eng.enqueue_mechanics(prev, cfg.contact_radius, pipe.dts, cfg.bond_stride, zero_flag=False, carry=True)
            finally:
                eng.substep_events = None
            torch.cuda.synchronize(dev)
            per_launch.append([evs[2 * k].elapsed_time(evs[2 * k + 1]) * 1e3 for k in range(sm)])
        eng.loc.copy_(loc_save)
        old.row_len.copy_(bonds_save[0])
        old.cols.copy_(bonds_save[1])
        eng._carry = None
        per_launch = np.array(per_launch)
        in_step = {"mean_us": float(per_launch.mean()), "first_launch_of_step_us": float(per_launch[:, 0].mean()),
                   "later_launches_us": float(per_launch[:, 1:].mean()), "steps": int(per_launch.shape[0]),
                   "launches": int(per_launch.size)}
    kbytes = substep_kernel_bytes(n, b10, ej)
    abytes = substep_algorithmic_bytes(n, b10, ej, ej)
    kernel_name, bytes_model = ("k_jkr_substep_plane<uniform radius>" if plane else "k_jkr_substep<uniform radius>"), (
        "SURVEY 8(d) algorithmic bytes of the stages the launch fuses: search(r=10) + merge/dedupe + force + integrate")
    if args.dtype == "f32":
        cold_us, hot_us = cold_us / sm, hot_us / sm
        kbytes = None
        abytes = substep_algorithmic_bytes(n, b10, ej, ej, 4) + sort_algorithmic_bytes(n, b10, 4)
        kernel_name = "mechanics loop per substep: k_scan_apply + k_rescatter32 + k_jkr_substep32 (fp32, bin-relative)"
        bytes_model = "SURVEY 8(d) algorithmic bytes with w = 4: bin/sort + search(r=10) + merge/dedupe + force + integrate"
    # roofline.achieved: algorithmic bytes over the kernel's AVERAGE launch duration in a running step (events around
    # every launch of the loop); the stand-alone launch after an L2 flush (every byte from HBM) is reported beside it
    kernel_us = in_step["mean_us"] if in_step else cold_us
    achieved = abytes / (kernel_us * 1e-6) / 1e9
    achieved_c = (kbytes or 0) / (kernel_us * 1e-6) / 1e9
    traffic = l2_hit = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        prof = json.load(open(tpath))
        traffic = prof.get(args.workload, {}).get("k_jkr_substep")
        l2_hit = prof.get("l2_hit_pct", {}).get(args.workload, {}).get("k_jkr_substep")
    nominal = 8000.0                                                     # B200 HBM3e nominal, SURVEY 8(d)
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "peak_source": peak_src, "frac_of_nominal_8000": round(achieved / nominal, 4),
                "l2_hit_pct_ncu": l2_hit, "bytes_per_launch": abytes,
                "bytes_model": bytes_model,
                "compulsory_bytes_per_launch": kbytes, "achieved_on_compulsory_bytes": round(achieved_c, 1),
                "frac_on_compulsory_bytes": round(achieved_c / peak, 4),
                "launch_us": round(kernel_us, 2),
                "launch_us_source": ("mean over every launch of 6 event-timed mechanics loops (abm_mechanics_loop_timed), L2 "
                                     "flushed before each loop" if in_step else "stand-alone launch after an L2 flush"),
                "launch_us_in_step": in_step,
                "launch_us_cold_l2": round(cold_us, 2), "launch_us_warm_l2": round(hot_us, 2),
                "frac_stand_alone_cold_l2": round(abytes / (cold_us * 1e-6) / 1e9 / peak, 4),
                "launches_per_step": sm,
                "step_algorithmic_bytes": bytes_step,
                "step_frac": round(bytes_step / (total_s / args.steps) / 1e9 / peak, 4),
                "step_frac_of_nominal_8000": round(bytes_step / (total_s / args.steps) / 1e9 / nominal, 4)}

    # ---- per-stage device times of one eager step (events around each stage, same stream, warm L2) -------------------
    def timed_stage(fn, reps=5):
        out = []
        for _ in range(reps):
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            z.record()
            torch.cuda.synchronize(dev)
            out.append(a.elapsed_time(z) * 1e3)
        return round(float(np.median(out)), 1)

    eng.use_current_stream()
    loc_keep = eng.loc.clone()
    stage_bonds = pipe.current_bonds()

    def stage_mechanics():
        eng.enqueue_mechanics(stage_bonds, cfg.contact_radius, pipe.dts, cfg.bond_stride, zero_flag=False)
    stages = {"search_r15_with_sort_us": timed_stage(lambda: pipe._search(None)),
              "mechanics_11_substeps_us": timed_stage(stage_mechanics)}
    eng.loc.copy_(loc_keep)
    if wl["grid"]:
        field_keep = pipe.field.clone()
        stages["morphogen_deposit_diffuse_sample_us"] = timed_stage(lambda: pipe._morphogen(None))
        pipe.field.copy_(field_keep)
    stages["note"] = "each stage alone on one stream (eager launches); the graph replay overlaps search + morphogen with the mechanics"

    # ---- e2e: the reference-facing API with host numpy arrays ----------------------------------------------------------
    e2e = None
    if rank == 0 or world > 1:
        class Cells(B200CellMethods, ColonySimulation):
            pass
        sim = Cells(wl["size"], dev)
        sim.b200_dtype = args.dtype
        sim.add_agents(n)
        sim.agent_array("locations", override=wl["loc"].copy())
        sim.agent_array("radii", override=wl["radii"].copy())
        sim.agent_array("motility_forces", vector=3)
        sim.agent_array("jkr_forces", vector=3)
        sim.agent_graph("neighbor_graph")
        sim.agent_graph("jkr_graph")
        sim.max_radius, sim.step_dt, sim.move_dt = 5.0, 1800, 180
        if wl["grid"]:
            h = spec["h"]
            sim.spat_res, sim.spat_res2, sim.diffuse_const = h, h * h, 50.0
            sim.diffuse_dt = spec["diffuse_dt"]
            sim.step_dt_diffusion = spec["diffuse_dt"] * spec["sd"]
            sim.max_concentration, sim.degradation = 2.0, 0.1
            sim.fgf4_values = host_field(wl)
            sim.b200_diffusion_3d = len(wl["grid"]) == 3

        mot_by_parity = (wl["mot"], -wl["mot"])
        api_count = [0]

        def api_step(fused_field=True):
            sim.get_neighbors("neighbor_graph", 15)          # (returns with the device idle: record_time synchronises)
            # stands in for cell_motility (a rule method outside the path): apply_forces leaves motility_forces
            # zeroed, so the step's input has to be written again.  Host-only work between two public calls; it is
            # INSIDE the e2e clock (and reported separately as input_fill_ms_per_step)
            t = time.perf_counter()
            sim.motility_forces[:] = mot_by_parity[api_count[0] % 2]
            fill = time.perf_counter() - t
            api_count[0] += 1
            sim.apply_forces()
            if wl["grid"]:
                # S_d substeps exactly: step_dt = (S_d-1)*dt + dt (divmod gives S_d-1 full substeps + remainder)
                keep = sim.step_dt
                sim.step_dt = spec["diffuse_dt"] * (spec["sd"] - 1) + 0.5 * spec["diffuse_dt"]
                if fused_field:
                    # deposit -> diffusion -> per-cell sampling as one public call on a device-resident field
                    sim.update_concentrations("fgf4_values", 1.0, diffuse=True)
                else:
                    sim.update_concentrations("fgf4_values", 1.0, sample=False)
                    sim.update_diffusion("fgf4_values")
                    sim.update_concentrations("fgf4_values", None)
                sim.step_dt = keep
            return fill
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            api_step()
        eng2 = sim.b200_engine()
        h2d0, d2h0, l0 = eng2.h2d_bytes, eng2.d2h_bytes, eng2.launches
        barrier()
        t0 = time.perf_counter()
        fill_s = 0.0
        for _ in range(e_steps):
            fill_s += api_step()
        torch.cuda.synchronize(dev)
        e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_s = float(t.item())
        h2d_step, d2h_step = (eng2.h2d_bytes - h2d0) // e_steps, (eng2.d2h_bytes - d2h0) // e_steps
        three = None
        if wl["grid"] and dist is None:
            for _ in range(2):
                api_step(False)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                api_step(False)
            torch.cuda.synchronize(dev)
            three = (time.perf_counter() - t0) / e_steps * 1e3
        e2e = {"value": n * world * e_steps / e_s, "unit": UNIT,
               "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step, "steps": e_steps,
               "ms_per_step": e_s / e_steps * 1e3,
               "input_fill_ms_per_step": fill_s / e_steps * 1e3,
               "input_fill": "host write of the step's motility forces (stand-in for cell_motility), inside the clock",
               "api": "Simulation.get_neighbors + CellMethods.apply_forces + update_concentrations(amounts, "
                      "diffuse=True) [deposit -> update_diffusion -> sample in one call]; host numpy arrays in and out, "
                      "an array is uploaded only if it changed since its device copy was made (arrays.py)",
               "ms_per_step_with_three_morphogen_calls": three}

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        parity = parity_check(args.workload, dev, dtype=args.dtype)
        v, spp, threads, sample, _ = run_cpu(args.workload, 2, 1)
        v1, spp1, _, sample1, _ = run_cpu(args.workload, 1, 1, threads=1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "s_per_step": spp, "single_thread": {"value": v1, "s_per_step": spp1, "cores": 1,
                                                     "sample": sample1}}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": common_config(args.workload, wl, sm),
        "method": {"edges_r15": e15, "jkr_bonds": ej,
                   "morphogen_stage": ("reference order (before the mechanics, on the step's starting positions), "
                                       "on a second stream beside search + mechanics") if cfg.overlap_morphogen and
                   wl["grid"] else "after the mechanics, same stream",
                   "timing": "CUDA events per step around a CUDA-graph replay, summed; max over ranks",
                   "parallelism": "single GPU"},
        "parity_check": parity,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "stages": stages,
        "cpu_baseline": cpu,
    }


def field_wanted(wl):
    return bool(wl["grid"])


def run_b200_slabs(args, rank, world, dev, dist, barrier, wl, cfg):
    """N > 1: ONE colony of world x workload cells, slab-decomposed along x (hipsc_abm_b200/slab_device.py): per step
    one cell-halo exchange (which also carries the migrating cells) and one morphogen-row exchange over NCCL
    send/recv; the whole step on owned+halo cells without further communication.  Every count lives on the device:
    no host synchronisation inside a step, the phases between the exchanges are replayed as CUDA graphs."""
    import torch
    from hipsc_abm_b200 import slab
    from hipsc_abm_b200.engine import ColonyEngine
    from hipsc_abm_b200.slab_device import DeviceSlab, DistSlab, PeerSlab

    spec, n = wl["spec"], wl["n"]
    side = wl["size"][0]
    gsize = wl["size"].copy()
    weak = args.scaling == "weak"
    if weak:
        gsize[0] = side * world                 # weak scaling: one workload-sized slab per GPU
        n_rank = n
    else:
        n_rank = n // world                     # strong scaling: the workload's colony cut into `world` slabs
    probe = ColonyEngine(gsize, dev)
    probe.max_radius = 5.0
    reach = probe.jkr_reach(cfg.contact_radius)
    halo = slab.required_halo(len(cfg.substep_dts()), reach, slack=4.0, spat_res=spec["h"] or 0.0)
    geom = slab.SlabGeometry.uniform(gsize, world, rank, halo)
    width = gsize[0] / world
    if weak:
        loc = wl["loc"].copy()
        loc[:, 0] += geom.x_lo
        radii, mot = wl["radii"], wl["mot"]
        gid = np.arange(n, dtype=np.int64) + rank * n
    else:
        m = (wl["loc"][:, 0] >= geom.x_lo) & ((wl["loc"][:, 0] < geom.x_hi) | (rank == world - 1))
        loc, radii, mot, gid = wl["loc"][m], wl["radii"][m], wl["mot"][m], np.nonzero(m)[0]
        n_rank = int(m.sum())
    field = None
    if wl["grid"] and len(wl["grid"]) == 2:     # the slab path carries the reference's 2-D field only
        gx = int(np.ceil(gsize[0] / spec["h"])) + 1
        field = slab.FieldSlab(geom, (gx, wl["grid"][1]), spec["h"], spec["sd"] + 1, dev)
    # capacities, the same on every rank: the expected cells within `halo` of an interface, with headroom
    expect = n_rank if weak else n / world
    send_cap = int(1.15 * expect * halo / width) + 1024      # (an overflow is detected: DeviceSlab.check() raises)
    rows = int(1.15 * expect * (1 + 2 * halo / width)) + 8192
    ds = DeviceSlab(geom, cfg, dev, rows=rows, send_cap=send_cap, max_radius=5.0, uniform_radius=5.0, field=field)
    ds.eng.mechanics_dtype = args.dtype
    ds.load(loc, radii, mot, gid)
    # the field rows travel on their own communicator: they are exchanged on the second stream and must not queue in
    # front of the next step's cell exchange
    rows_group = dist.new_group(list(range(world))) if field is not None else None
    left_group = dist.new_group(list(range(world))) if world > 2 and not os.environ.get("ABM_B200_ONE_CELL_GROUP") else None
    # cell records over NVLink peer memory (PeerSlab) unless ABM_B200_CELL_EXCHANGE=nccl or the CUDA-IPC mapping of the
    # neighbours' buffers fails on any rank (then: NCCL send/recv, the two directions on two communicators)
    rank_step, exchange = None, "nccl send/recv"
    if os.environ.get("ABM_B200_CELL_EXCHANGE", "peer") == "peer":
        rank_step = PeerSlab(ds, rank, world, rows_group=rows_group)
        if rank_step.ok:
            exchange = "peer-memory copies over NVLink (CUDA IPC) + flag words"
        else:
            exchange = "nccl send/recv (peer mapping failed: %s)" % (rank_step.why or "on another rank")
            rank_step = None
    if rank_step is None:
        rank_step = DistSlab(ds, rank, world, rows_group=rows_group, left_group=left_group)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # the step's main stream outranks the second stream (search + morphogen stage), whose kernels then fill what the
    # mechanics leave idle instead of competing with them
    main_stream = torch.cuda.Stream(device=dev, priority=-1)
    main_stream.wait_stream(torch.cuda.current_stream(dev))
    torch.cuda.set_stream(main_stream)

    def step():
        rank_step.step()
        ds.negate_motility()              # alternating motility sign (make_workload)

    for _ in range(2 * ((max(args.warmup, 3) + 1) // 2)):       # an even number of eager steps, then capture
        step()
    ds.check(halo=False)
    # the uniformly random start is violent (a few overlapping pairs fly apart by tens of um in the first substeps);
    # the halo assumption (DeviceSlab.allowed_substep_move) is verified over the timed steps and reported for these
    warmup_move = ds.eng.max_substep_move()
    if not args.no_graphs:
        rank_step.capture()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for _ in range(40):                   # a few clock samples under load (a fixed count: ranks stay in lock-step)
        step()
    launches0, bytes0, ex0 = ds.launches, rank_step.bytes_sent, rank_step.exchanges
    timeline = None
    if args.timeline:                     # a few extra (untimed) steps with events around the main-stream phases
        rank_step.timeline = []
        for _ in range(10):
            step()
        timeline = {k: round(v, 1) for k, v in rank_step.timeline_us().items()}
        rank_step.timeline = None
    ds.reset_monitor()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k, (a, b) in enumerate(evs):
        flush.zero_()
        a.record()
        step()
        if k == len(evs) - 1:
            rank_step.join()              # the last step's second-stream work belongs to the timed region
        b.record()
    barrier()
    clocks = sampler.stop()
    ds.check()                            # capacities, and the halo assumption over the timed steps
    if hasattr(rank_step, "check_peers"):
        rank_step.check_peers()
    timed_move = ds.eng.max_substep_move()
    total_s = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    tt = torch.tensor([total_s], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_s = float(tt.item())
    n_own = int(ds.sets[ds.cur]["n_own"].item())
    cells = torch.tensor([n_own], device=dev)
    dist.all_reduce(cells)
    n_total = int(cells.item())
    per_step_launches = (ds.launches - launches0) // max(args.steps, 1) if not rank_step.captured else None

    # e2e: the same step with the host in the loop -- the rule methods' inputs (motility forces of the owned
    # cells) are uploaded from pinned memory and the positions + sampled concentrations downloaded, every step
    e_steps = max(3, min(args.steps, 10))
    pin_mot = torch.zeros((ds.R, 3), dtype=torch.float64).pin_memory()
    pin_mot[:n_own].copy_(ds.sets[ds.cur]["mot"][:n_own])     # realistic values (rows beyond the count: no force)
    h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    for it in range(e_steps):
        cur = ds.sets[ds.cur]
        k = int(cur["n_own"].item())
        cur["mot"][:k].copy_(pin_mot[:k], non_blocking=True)
        if it % 2:
            cur["mot"][:k].neg_()                             # the alternating sign of make_workload
        rank_step.step()
        rank_step.join()
        now = ds.sets[ds.cur]
        k2 = int(now["n_own"].item())
        out_loc = now["loc"][:k2].cpu()
        out_c = ds.sampled[:k2].cpu() if field is not None else None
        h2d += k * 24
        d2h += out_loc.numel() * 8 + (out_c.numel() * 8 if out_c is not None else 0) + 8
    torch.cuda.synchronize(dev)
    e_s = time.perf_counter() - t0
    tt = torch.tensor([e_s], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e_s = float(tt.item())
    if rank != 0:
        return None
    value = n_total * args.steps / total_s
    config = common_config(args.workload, wl, len(cfg.substep_dts()), world, scaling=args.scaling)
    if wl["grid"] and field is None:
        config.update(diffusion_substeps=0, grid=None,
                      field_note="the slab path carries the reference's 2-D field only: this colony's 3-D field stage "
                                 "is not part of the multi-GPU step")
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config,
        "method": {"cells_total": n_total, "halo_um": halo, "rows_capacity": rows, "message_capacity": send_cap,
                   "halo_check_rank0": {"allowed_ghost_zone_move_per_substep_um": round(ds.allowed_substep_move(), 3),
                                        "largest_in_timed_steps_um": round(timed_move, 3),
                                        "largest_in_first_warmup_steps_um": round(warmup_move, 3)},
                   "morphogen_stage": "reference order (before the mechanics), with its row exchange on a second "
                                      "stream beside search + mechanics" if field is not None else "none",
                   "exchanges_per_step": (rank_step.exchanges - ex0) / max(args.steps + e_steps, 1),
                   "bytes_sent_per_step_rank0": (rank_step.bytes_sent - bytes0) // max(args.steps + e_steps, 1),
                   "timing": "CUDA events per step, summed; max over ranks",
                   "host_synchronisations_per_step": 0, "main_stream_phases_us_rank0": timeline,
                   "cuda_graphs": bool(rank_step.captured),
                   "cell_exchange": exchange,
                   "parallelism": f"slab{world}: 1-D slab decomposition along x; per step one cell-halo exchange "
                                  "(migration rides on it; fixed-capacity messages, counts on the device) and one "
                                  "morphogen-row exchange over NCCL send/recv"},
        "clocks": clocks,
        "e2e": {"value": n_total * e_steps / e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // e_steps,
                "d2h_bytes_per_step": d2h // e_steps, "steps": e_steps, "ms_per_step": e_s / e_steps * 1e3,
                "api": "slab_device.DistSlab.step with host-resident motility forces in, positions + "
                       "concentrations out"},
        "gpu_launches": (per_step_launches if per_step_launches is not None else ds_launches_per_step(ds)) * args.steps,
        "roofline": None, "cpu_baseline": None,
    }


def ds_launches_per_step(ds):
    """Kernels of this library inside one captured slab step (the graphs replay what an eager step enqueues):
    pack 1 + merge 4 + deposit 1 + mechanics 3 x substeps + 2 + sort/search 5 + diffusion + sample."""
    sm = len(ds.dts)
    n = 1 + 4 + 3 * sm + 2 + 5
    if ds.field is not None:
        n += 1 + int(ds.eng.lib.abm_diffusion_launches(ds.cfg.diffuse_steps, ds.eng.diffusion_fuse)) + 1
    return n


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Numba reference
    itself cannot travel to the GPU box), all host threads, on the same workload recipe."""
    if rank != 0:
        return None
    v, spp, threads, sample, wl = run_cpu(args.workload, args.steps, args.warmup)
    sm = int(1800 // 180) + 1
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": common_config(args.workload, wl, sm, world, cpu_arm=True),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS),
                    help="c3 = BASELINE configs[2], the 500k-cell headline colony quoted at 1 and 8 B200 (default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="arithmetic of the mechanics loop: f64 = the reference's dtype (default, the headline); f32 = "
                         "single precision on bin-relative coordinates (north-star gate 1e-5)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one workload-sized slab per GPU (default); strong = the workload's colony cut "
                         "into N slabs (BASELINE configs[3]: --workload c4 --scaling strong)")
    ap.add_argument("--no-graphs", action="store_true", help="N > 1: eager launches instead of CUDA-graph replays")
    ap.add_argument("--timeline", action="store_true", help="N > 1: device time of every main-stream phase of a step")
    ap.add_argument("--no-clocks", action="store_true",
                    help="no nvidia-smi sampling (profiler runs: keeps the launch sequence deterministic)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="run deposit/diffusion/sample after the mechanics on one stream instead of beside them")
    ap.add_argument("--bond-stride", type=int, default=0, help="fixed width of the JKR bond rows (0 = 16 / 32)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun (one rank per GPU); running 1 GPU", file=sys.stderr)
        try:
            line = run_b200(args, rank, world, local_rank)
        except Exception as err:
            # a captured step has fixed buffer sizes; a colony that needs wider bond rows is re-run once with them
            if world == 1 and "bond rows too narrow" in str(err) and not args.bond_stride:
                args.bond_stride = 32 if WORKLOADS[args.workload]["dim"] == 2 else 64
                print(f"bench.py: {err}; re-running with --bond-stride {args.bond_stride}", file=sys.stderr)
                line = run_b200(args, rank, world, local_rank)
            else:
                raise
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
