"""bench.py pieces that run without a GPU: the synthetic workload, the byte model and the reference (CPU) arm's
JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_follows_the_survey_recipe():
    wl = bench.make_workload("c2")
    n, size = wl["n"], wl["size"]
    assert n == 50_000 and size[2] == 0.0
    assert abs(n / (size[0] * size[1]) - 0.01) < 1e-12                       # cells per um^2
    assert np.all(wl["loc"] >= 0) and np.all(wl["loc"][:, :2] <= size[:2])
    assert np.allclose(np.linalg.norm(wl["mot"], axis=1), 2e-9)              # reference cell_methods.py:252
    assert np.all(wl["mot"][:, 2] == 0) and np.all(wl["radii"] == 5.0)
    assert wl["grid"][0] == wl["grid"][1] and wl["grid"][1] % 2 == 0
    w4 = bench.make_workload("c4", n_override=1000)
    assert w4["size"][2] > 0 and len(w4["grid"]) == 3      # the 3-D colony carries a 3-D field (this build's extension)
    assert np.array_equal(bench.make_workload("c2", seed=1)["loc"], bench.make_workload("c2", seed=1)["loc"])


def test_byte_model_is_survey_8d():
    n, b15, b10, e15, e10, ej, g, sm, sd = 1000, 300, 700, 3500, 1570, 1570, 4096, 11, 10
    w = 8
    search15 = 7 * w * n + 4 * n + 8 * b15 + 4 * w * n + 4 * b15 + 8 * e15
    substep = (7 * w * n + 4 * n + 8 * b10) + (4 * w * n + 4 * b10 + 8 * e10) + 8 * (2 * ej + e10) + \
              (9 * ej + 4 * w * n) + 6 * w * n
    assert bench.step_bytes(n, b15, b10, e15, e10, ej, g, sm, sd) == search15 + sm * substep + sd * 2 * w * g + 5 * w * n


def test_cpu_arm_alternates_the_motility_sign():
    wl = bench.make_workload("c2", n_override=400)
    seen = []
    step = bench.oracle_step_fn(wl, 1)
    sim = step.__closure__ and [c.cell_contents for c in step.__closure__ if hasattr(c.cell_contents, "motility_forces")][0]
    orig = type(sim).apply_forces

    def spy(self):
        seen.append(self.motility_forces.copy())
        return orig(self)
    type(sim).apply_forces = spy
    try:
        for _ in range(3):
            step()
    finally:
        type(sim).apply_forces = orig
    assert np.array_equal(seen[0], wl["mot"]) and np.array_equal(seen[1], -wl["mot"]) and np.array_equal(seen[2], wl["mot"])


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("c2")
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cpu_arm_runs_the_stated_colony_when_it_fits_the_budget():
    # c3 at the driver's --steps 20 --warmup 5: the whole 500,000-cell colony; c4's 4 M cells are sampled and say so
    assert bench.cpu_sample_cells("c3", 25) == 500_000
    assert bench.cpu_sample_cells("c2", 25) == 50_000
    n4 = bench.cpu_sample_cells("c4", 25)
    assert 100_000 <= n4 < 4_000_000
    assert "sample" not in bench.workload_label("c3", 500_000, cpu_arm=True)
    assert f"{n4}-cell sample" in bench.workload_label("c4", n4, cpu_arm=True)
    assert "sample" in bench.workload_label("c3", 500_000, world=8, cpu_arm=True)     # one GPU's share of the colony
    # both arms print the same workload-defining config when they run the same colony
    wl = bench.make_workload("c2")
    assert bench.common_config("c2", wl, 11) == bench.common_config("c2", wl, 11, cpu_arm=True)


def test_reference_arm_ignores_torchrun_thread_cap():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=600,
                         env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == bench.host_cores()
    assert f"on {bench.host_cores()} threads" in line["cpu_baseline"]["sample"]


def test_ab_script_variants_name_existing_build_switches():
    """scripts/ab_substep.py builds variant libraries with -D switches; every switch it names must still exist in csrc/
    (a renamed or removed experiment would silently build the default kernel and be "measured")."""
    import ast
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, "scripts", "ab_substep.py")).read())
    variants = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "VARIANTS":
            variants = ast.literal_eval(node.value)
    assert variants, "VARIANTS not found"
    sources = "".join(open(f).read() for f in glob.glob(os.path.join(root, "python-hipsc-abm_b200", "csrc", "*")))
    for name, defines in variants.items():
        for d in defines:
            macro = d.split("=")[0]
            assert re.search(r"\b%s\b" % re.escape(macro), sources), f"variant {name}: {macro} is not used in csrc/"
