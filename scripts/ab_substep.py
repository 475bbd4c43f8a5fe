"""A/B timing of k_jkr_substep variants: builds variant libraries (local, needs nvcc) or times one (GPU box).
  build:  python scripts/ab_substep.py build
  time:   ABM_B200_LIB=<so> python scripts/ab_substep.py time <workload>
  run:    python scripts/ab_substep.py run <workload> [variant ...]      (GPU box, ONE gpurun call: for every variant
          library in build_ab/ -- or the named ones -- the GPU parity tests, then the timing; variants whose name
          starts with "no_" are timing ablations with wrong results and skip the tests)"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {                      # -D switches of csrc/jkr.cu (see the comment at ABM_JKR_EXP there)
    "base": [],
    "regs64": ["ABM_JKR_MINBLOCKS=8"], "regs80": ["ABM_JKR_MINBLOCKS=6"], "regs96": ["ABM_JKR_MINBLOCKS=5"],
    "cta64": ["ABM_JKR_THREADS=64", "ABM_JKR_MINBLOCKS=14"], "cta256": ["ABM_JKR_THREADS=256", "ABM_JKR_MINBLOCKS=3"],
    "hits12": ["ABM_JKR_MAXHITS=12"],
    "early_motility": ["ABM_JKR_EXP=1"], "early_old_row": ["ABM_JKR_EXP=2"],
    "no_contact_loop": ["ABM_JKR_EXP=4", "ABM_ALLOW_WRONG_RESULTS"],            # timing ablations
    "no_walk_no_contact": ["ABM_JKR_EXP=12", "ABM_ALLOW_WRONG_RESULTS"],
    "no_gather": ["ABM_JKR_EXP=64", "ABM_ALLOW_WRONG_RESULTS"],                 # partner data from shared memory (2-D)
    "no_pdl": ["ABM_PDL=0"],
    "pdl_late": ["ABM_PDL_LATE=1"],        # dependent launch triggered after the contact loop (run with ABM_B200_GRAPH_PDL=1)
    "trim_math": ["ABM_JKR_EXP=32"],       # Horner polynomial + rsqrt without special cases (written without a GPU)
    "slot_rows_trim_math": ["ABM_JKR_EXP=48"],
    "slot_rows_pdl_late": ["ABM_JKR_EXP=16", "ABM_PDL_LATE=1"],
    "slot_rows": ["ABM_JKR_EXP=16"],
    # the two-axis kernel of flat worlds (csrc/jkr_plane.cuh); AB_PLANE=0 in the environment times the three-axis one
    "plane_9cta": ["ABM_PLANE_MINBLOCKS=9"], "plane_hits12": ["ABM_JKR_MAXHITS=12"],
    # tile of the diffusion kernels (csrc/diffusion.cu); timed by the "diffuse_10_us" entry below
    "diff_16x128": ["ABM_DIFF_TILE_I=16"], "diff_32x64": ["ABM_DIFF_TILE_J=64"],
    "diff_16x64": ["ABM_DIFF_TILE_I=16", "ABM_DIFF_TILE_J=64"], "diff_64x128": ["ABM_DIFF_TILE_I=64"],
}
if sys.argv[1] == "build":
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "python-hipsc-abm_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    os.makedirs(os.path.join(ROOT, "build_ab"), exist_ok=True)
    for name, defs in VARIANTS.items():
        print(name, b.build(defines=defs, out=os.path.join(ROOT, "build_ab", f"libabm_{name}.so")))
elif sys.argv[1] == "run":
    import glob, subprocess
    names = sys.argv[3:] or sorted(os.path.basename(f)[len("libabm_"):-3] for f in glob.glob(os.path.join(ROOT, "build_ab", "libabm_*.so")))
    for name in ["default"] + names:
        env = dict(os.environ)
        if name.endswith("+pdl"):                 # "<variant>+pdl": programmatic dependent launches inside the graph too
            env["ABM_B200_GRAPH_PDL"] = "1"
            name = name[:-4]
        if name != "default":
            env["ABM_B200_LIB"] = os.path.join(ROOT, "build_ab", f"libabm_{name}.so")
        if not name.startswith("no_"):
            t = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q"], cwd=ROOT, env=env,
                               capture_output=True, text=True)
            print(name, "tests:", (t.stdout.strip().splitlines() or ["?"])[-1], flush=True)
            if t.returncode != 0:
                print(t.stdout[-1500:])
                continue
        subprocess.run([sys.executable, os.path.abspath(__file__), "time", sys.argv[2]], cwd=ROOT, env=env)
else:
    import numpy as np, torch, bench
    from hipsc_abm_b200 import _lib as _l
    from hipsc_abm_b200.engine import ColonyEngine
    from hipsc_abm_b200.pipeline import HotPathStep, StepConfig
    name = sys.argv[2]
    wl = bench.make_workload(name); spec = wl["spec"]; n = wl["n"]
    full = os.environ.get("AB_FULL")          # AB_FULL=1: with the morphogen stage, AB_FULL=overlap: on a side stream
    cfg = StepConfig(grid=wl["grid"] if full else None, spat_res=spec["h"] or 1.0, diffuse_dt=spec["diffuse_dt"] or 1.0,
                     diffuse_steps=spec["sd"], overlap_morphogen=(full == "overlap"), bond_stride=16 if spec["dim"] == 2 else 32, search_cols_per_cell=12.0 if spec["dim"] == 2 else 24.0)
    eng = ColonyEngine(wl["size"]); eng.upload(wl["loc"], wl["radii"], wl["mot"])
    eng.plane_kernel = os.environ.get("AB_PLANE", "1") != "0"
    pipe = HotPathStep(eng, cfg)
    for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
        pipe.step()
    pipe.check()
    reach = eng.jkr_reach(10.0); srt = eng.sort(reach); b = eng._bond_bufs; old = pipe.current_bonds()
    spare = [i for i in range(3) if b["cols"][i].data_ptr() != old.cols.data_ptr()][0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # AB_PLANE=0: the three-axis kernel also in a flat world (single launch: no ABM_SUBSTEP_FLAT_Z; loop: ABM_LOOP_NO_PLANE)
    sub_flags = _l.SUBSTEP_FLAT_Z if (spec["dim"] == 2 and eng.plane_kernel) else 0
    def launch():
        _l.check(eng.lib.abm_jkr_substep(
            srt.bin_start.data_ptr(), srt.sorted_id.data_ptr(), srt.sorted_pos.data_ptr(), srt.sorted_bin.data_ptr(),
            n, _l.i3(srt.dims), 10.0, reach, 0, old.row_stride, old.row_len.data_ptr(), old.cols.data_ptr(),
            b["len"][spare].data_ptr(), b["cols"][spare].data_ptr(), cfg.bond_stride, b["max_len"].data_ptr(), sub_flags,
            eng.mot.data_ptr(), eng.loc.data_ptr(), 0.5, 1000.0, 0.000107, 10000.0, _l.d3(wl["size"]), 180.0, 0,
            eng.uniform_radius, None, None, 0, 0, 0, None, eng.stream), "abm_jkr_substep")
    res = {}
    for cold in (True, False):
        ts = []
        for _ in range(15):
            flush.zero_() if cold else launch()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); launch(); z.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(z) * 1e3)
        res["cold" if cold else "warm"] = float(np.median(ts))
    # whole step enqueued eagerly (what the slab pipeline does), L2 flushed before each
    ts = []
    for _ in range(12):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.step(); z.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(z) * 1e3)
    res["eager_step_us"] = float(np.median(ts))
    # whole step as a CUDA-graph replay (no diffusion grid), L2 flushed before each
    if pipe.parity:
        pipe.step()
    pipe.capture()
    for _ in range(6):
        pipe.replay()
    ts = []
    for _ in range(20):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.replay(); z.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(z) * 1e3)
    res["step_us"] = float(np.median(ts))
    pipe.check()
    # ten diffusion substeps alone on the workload's grid (two fused launches by default), L2 flushed before each
    if wl["grid"]:
        field = torch.rand(wl["grid"], dtype=torch.float64, device="cuda")
        scratch = torch.empty_like(field)
        ts = []
        for _ in range(20):
            flush.zero_()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.diffuse(field, 10, spec["diffuse_dt"], spec["diffuse_dt"], 50.0, spec["h"] ** 2, 2.0, 0.1, scratch=scratch)
            z.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(z) * 1e3)
        res["diffuse_10_us"] = float(np.median(ts))
    print(("plane " if sub_flags else "3axis ") + os.path.basename(os.environ.get("ABM_B200_LIB", "default")) + ("+pdl" if os.environ.get("ABM_B200_GRAPH_PDL") == "1" else ""), name, {k: round(v, 1) for k, v in res.items()})
