"""Debug probe: run the TMA diffusion path alone (used under compute-sanitizer on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hipsc_abm_b200.engine import ColonyEngine
from oracle import oracle as orc
gx, gy, steps = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (318, 318, 2)
rng = np.random.default_rng(0)
field = rng.random((gx, gy)) * 3 - 0.5
h, D = 7.07106, 50.0
dt = 0.24 * h * h / D
eng = ColonyEngine([1, 1, 0])
out = eng.diffuse(torch.from_numpy(field.copy()).cuda(), steps, dt, dt * 0.3, D, h * h, 2.0, 0.1, use_tma=True)
torch.cuda.synchronize()
ref = orc.diffuse_field(field, steps, dt, dt * 0.3, D, h * h, 2.0, 0.1)
print("tma diffusion equal:", np.array_equal(out.cpu().numpy(), ref))
