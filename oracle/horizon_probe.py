"""How long can ANY implementation stay on the reference's recorded trajectory?

Runs the REFERENCE's own CellSimulation.steps() (imported read-only from /root/reference/ABM, one Numba thread,
the recipe of oracle/make_golden.py:full_model) on the seeded C1 colony of tests/golden/model_c1_100.npz, once
unperturbed (must reproduce the recording step for step) and then with ONE coordinate of ONE cell moved by one unit
in the last place (np.nextafter) before the first step, and reports the first step at which any discrete state
array, the agent count or an edge count differs from the recording.  That step is the horizon beyond which
"states identical" is not a property of an implementation but of bit-identical arithmetic: a 1-ulp difference --
far below the north star's own 1e-12 position gate -- is enough to leave the recording.

Build container only (needs /root/reference):   NUMBA_NUM_THREADS=1 python oracle/horizon_probe.py
Writes tests/golden/horizon_c1_100.json (committed; tests/test_model_states.py reads it).
"""
import contextlib
import io
import json
import os
import random
import shutil
import sys
import tempfile

os.environ.setdefault("NUMBA_NUM_THREADS", "1")
import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_golden as mg  # noqa: E402  (imports the reference with the igraph / pythonabm shims)

DISCRETE = ("states", "GATA6", "NANOG", "FGF4", "FGFR", "ERK")


def run(perturb=None, name="c1_100"):
    g = np.load(os.path.join(mg.OUT, f"model_{name}.npz"))
    seed, n, end_step = int(g["meta"][0]), int(g["meta"][1]), int(g["meta"][2])
    size = [float(v) for v in g["meta"][4:7]]
    work = tempfile.mkdtemp(prefix="horizon_")
    first = {"step": None, "what": None}
    drift = {}
    try:
        os.makedirs(os.path.join(work, "templates"))
        with open(os.path.join(mg.REF, "templates", "general.yaml")) as f:
            general = yaml.safe_load(f)
        with open(os.path.join(mg.REF, "templates", "experimental.yaml")) as f:
            experimental = yaml.safe_load(f)
        general.update(num_to_start=n, cuda=False, end_step=end_step, size=size, output_values=False,
                       output_images=False, record_initial_step=False)
        experimental.update(num_gata6=0, output_tda=False, output_gradients=False, dox_step=5)
        for fname, keys in (("general.yaml", general), ("experimental.yaml", experimental)):
            with open(os.path.join(work, "templates", fname), "w") as f:
                yaml.safe_dump(keys, f)
        outdir = os.path.join(work, "out") + os.sep
        os.makedirs(os.path.join(outdir, name))

        class Probe(mg.cell_simulation.CellSimulation):
            def temp(self):
                pass

            def data(self):
                s = self.current_step
                if first["step"] is not None:
                    return
                what = None
                if self.number_agents != int(g[f"n_{s}"]):
                    what = "agent count"
                else:
                    for arr in DISCRETE:
                        if not np.array_equal(self.__dict__[arr].astype(np.int8), g[f"{arr}_{s}"]):
                            what = arr
                            break
                    if what is None and self.neighbor_graph.ecount() != int(g[f"nbr_edges_{s}"]):
                        what = "neighbour edge count"
                    if what is None and self.jkr_graph.ecount() != int(g[f"jkr_edges_{s}"]):
                        what = "JKR bond count"
                if what is not None:
                    first["step"], first["what"] = s, what
                elif f"locations_{s}" in g.files:
                    drift[s] = float(np.abs(self.locations - g[f"locations_{s}"]).max())

        cwd = os.getcwd()
        os.chdir(work)
        try:
            random.seed(seed)
            np.random.seed(seed)
            sim = Probe(name, outdir)
            sim.agent_initials()
            if perturb is not None:
                cell, axis, direction = perturb
                sim.locations[cell, axis] = np.nextafter(sim.locations[cell, axis], direction * np.inf)
            with contextlib.redirect_stdout(io.StringIO()):
                sim.steps()
        finally:
            os.chdir(cwd)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return first["step"], first["what"], drift, end_step


if __name__ == "__main__":
    out = {"recording": "tests/golden/model_c1_100.npz", "perturbation": "np.nextafter on one coordinate of one cell "
           "before the first step (1 ulp, ~5e-14 um)", "runs": []}
    step, what, drift, end_step = run(None)
    assert step is None, f"the unperturbed reference run must reproduce its own recording, differs at {step}: {what}"
    print("unperturbed: identical through step", end_step)
    for perturb in [(0, 0, +1), (0, 0, -1), (123, 1, +1), (500, 0, +1), (777, 1, -1), (999, 0, +1)]:
        step, what, drift, _ = run(perturb)
        horizon = (step - 1) if step is not None else end_step
        out["runs"].append({"cell": perturb[0], "axis": perturb[1], "direction": perturb[2],
                            "identical_through_step": horizon, "first_difference": what,
                            "position_drift_um": {str(k): v for k, v in sorted(drift.items())}})
        print(perturb, "identical through step", horizon, "first difference:", what, "drift", drift)
    hs = [r["identical_through_step"] for r in out["runs"]]
    out["horizon_min"], out["horizon_max"], out["end_step"] = min(hs), max(hs), end_step
    with open(os.path.join(mg.OUT, "horizon_c1_100.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("horizons", hs)
