"""Per-stage cycles of the selection of a typical column.  Build the library with
-DMCZ_PHASE_TIMING -DMCZ_COLTIME_DUMP=2 (range+H, plan, G) or =3 (range, finish, total).
   python tools/dbg_stages.py <label> [N' [galaxies [md]]]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2200
G = int(sys.argv[3]) if len(sys.argv) > 3 else 2960
if N <= 2304:
    meas, err = synthetic_table(125000, config=4, start=0, stop=G)
else:
    meas, err = synthetic_table(G, config=3)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
TOK = sc.parse_md(sys.argv[4])[0] if len(sys.argv) > 4 else sc.T_ALL
ws = ops.ProductionWorkspace(G, N, TOK, m.device)
ops.forward_production(m, e, N, TOK, workspace=ws); torch.cuda.synchronize()
pct, nv, st, tr, ret, Z = ops.forward_production(m, e, N, TOK, workspace=ws)
torch.cuda.synchronize()
t = pct.cpu().numpy().astype(np.float64); st = st.cpu().numpy()
ok = st == 0
print(sys.argv[1:2], N, "mean over OK columns:", [round(float(t[:, :, i][ok].mean())) for i in range(3)],
      "median:", [round(float(np.median(t[:, :, i][ok]))) for i in range(3)])
for i in range(3):
    v = t[:, :, i][ok]
    print("  stage %d: p10 %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (i, *np.percentile(v, [10, 50, 90, 99]), v.max()))
