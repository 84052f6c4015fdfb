"""Per-stage cycles of the selection of a typical column.  Build the library with
-DMCZ_PHASE_TIMING -DMCZ_COLTIME_DUMP=2 (range+H, plan, G) or =3 (range, finish, total)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G = 2960
meas, err = synthetic_table(125000, config=4, start=0, stop=G)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, 2200, sc.T_ALL, m.device)
ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws); torch.cuda.synchronize()
pct, nv, st, tr, ret, Z = ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws)
torch.cuda.synchronize()
t = pct.cpu().numpy().astype(np.float64); st = st.cpu().numpy()
ok = st == 0
print(sys.argv[1:], "mean over OK columns:", [round(float(t[:, :, i][ok].mean())) for i in range(3)],
      "median:", [round(float(np.median(t[:, :, i][ok]))) for i in range(3)])
