"""Launches for ncu captures of the other kernels: `wide [N']` (one object, 1.1e6 samples by default) or `global`
(1480 galaxies x 11000 samples, production_kernel<0, false>)."""
import sys, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
which = sys.argv[1]
if which == "wide":
    G, N, cfg = 1, (int(sys.argv[2]) if len(sys.argv) > 2 else 1100000), 4
else:
    G, N, cfg = 1480, 11000, 3
meas, err = synthetic_table(max(G, 8), config=cfg)
m = torch.from_numpy(meas[:G].copy()).cuda(); e = torch.from_numpy(err[:G].copy()).cuda()
ws = ops.ProductionWorkspace(G, N, sc.T_ALL, m.device)
for _ in range(3):
    ops.forward_production(m, e, N, sc.T_ALL, workspace=ws)
torch.cuda.synchronize()
print("done", which)
