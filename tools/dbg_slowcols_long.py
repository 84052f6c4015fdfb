"""Long columns (N' = 11000): selection time per column by key and by path flags.
Library built with -DMCZ_PHASE_TIMING -DMCZ_COLTIME_DUMP=3 (pct = range, finish, total cycles; flags in nvalid >> 16:
bits 0-2 crowded ranges, 8 tie merged by rank, 16 repeated value in the prefix, bits 5-6 number of ranges)."""
import sys, numpy as np, torch, collections
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G, N = 1480, 11000
meas, err = synthetic_table(G, config=3)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, N, sc.T_ALL, m.device)
ops.forward_production(m, e, N, sc.T_ALL, workspace=ws); torch.cuda.synchronize()
pct, nv, st, tr, ret, Z = ops.forward_production(m, e, N, sc.T_ALL, workspace=ws)
torch.cuda.synchronize()
t = pct.cpu().numpy().astype(np.float64); st = st.cpu().numpy(); nv = nv.cpu().numpy()
tot = t[:, :, 2]
live = st != 1
print("per galaxy: mean of the slowest column %.0f, mean over columns %.0f" % (np.where(live, tot, 0).max(axis=1).mean(), tot[live].mean()))
worst = np.where(live, tot, 0).argmax(axis=1)
print("slowest column by key:", collections.Counter(sc.KEYS[j] for j in worst).most_common(8))
for j, k in enumerate(sc.KEYS):
    if live[:, j].any():
        v = tot[:, j][live[:, j]]
        print("%-10s n %4d mean %7.0f p50 %7.0f p90 %7.0f max %7.0f  status ok %.2f" % (k, v.size, v.mean(), np.median(v), np.percentile(v, 90), v.max(), (st[:, j] == 0).mean()))
fl = (nv >> 16)[live & (st == 0)]
tt = tot[live & (st == 0)]
for f, c in collections.Counter(fl.tolist()).most_common(12):
    v = tt[fl == f]
    print("flags %#5x: %6d columns, mean %7.0f p90 %7.0f" % (f, c, v.mean(), np.percentile(v, 90)))
v = tot[live & (st != 0)]
if v.size:
    print("status != ok: %d columns, mean %.0f" % (v.size, v.mean()))
cr = live & (st == 0) & (((nv >> 16) & 7) != 0)
if cr.any():
    print("columns with a crowded range: %d; mean of the three dumped values %s (COLTIME_DUMP=2: range + pass H, plan, gather; =3: range, finish, total)"
          % (cr.sum(), [round(float(t[:, :, i][cr].mean())) for i in range(3)]))
    print("  by key:", collections.Counter(sc.KEYS[j] for j in np.nonzero(cr)[1]).most_common(6))
