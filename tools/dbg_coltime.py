"""Per-galaxy, per-column selection cycles (library built with -DMCZ_PHASE_TIMING -DMCZ_COLTIME_DUMP:
the p84 output slot then carries the column's cycle count)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G = 14800
meas, err = synthetic_table(125000, config=4, start=0, stop=G)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, 2200, sc.T_ALL, m.device)
ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws); torch.cuda.synchronize()
pct, nv, st, tr, ret, Z = ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws)
torch.cuda.synchronize()
t = pct[:, :, 2].cpu().numpy().astype(np.float64); st = st.cpu().numpy(); nv = nv.cpu().numpy()
act = [j for j, k in enumerate(sc.KEYS) if (st[:, j] != 1).any() and k != 'M08_R23']
T = t[:, act]
mx = T.max(1); am = T.argmax(1)
print("wall(max col) mean %.0f  median %.0f  p90 %.0f  p99 %.0f ; mean of col means %.0f" % (mx.mean(), np.median(mx), np.percentile(mx, 90), np.percentile(mx, 99), T.mean()))
ok = st[:, act] == 0
print("mean over OK columns %.0f ; over empty %.0f" % (T[ok].mean(), T[~ok].mean() if (~ok).any() else 0))
print("second largest mean %.0f" % np.sort(T, 1)[:, -2].mean())
for jj, j in enumerate(act):
    sel = am == jj
    print("%-10s last in %.3f of galaxies; its time then mean %.0f; overall mean %.0f p50 %.0f p90 %.0f p99 %.0f" % (
        sc.KEYS[j], sel.mean(), T[sel, jj].mean() if sel.any() else 0, T[:, jj].mean(), np.median(T[:, jj]), np.percentile(T[:, jj], 90), np.percentile(T[:, jj], 99)))
# histogram of wall
h, edges = np.histogram(mx, bins=[0, 28000, 30000, 32000, 34000, 36000, 38000, 40000, 45000, 50000, 60000, 1e9])
print("wall histogram", list(zip(edges[:-1].astype(int), h)))
# classification of slow columns (nvalid high bits carry flags)
fl = (nv[:, act] >> 16)
slow = T > 40000
print("slow columns: %d of %d" % (slow.sum(), slow.size))
import collections
c = collections.Counter()
for f_, jj in zip(fl[slow], np.nonzero(slow)[1]):
    c[(sc.KEYS[act[jj]], int(f_) & 7, (int(f_) >> 3) & 1, (int(f_) >> 4) & 1, (int(f_) >> 5) & 7, int(f_) >> 8)] += 1
print("key, crowded mask, tie merged, has_tie, nr, rounds : count")
for k, v in sorted(c.items(), key=lambda kv: -kv[1])[:40]: print(k, v)
fast = ~slow & (st[:, act] == 0)
c2 = collections.Counter()
for f_ in fl[fast]: c2[(int(f_) & 7, (int(f_) >> 3) & 1, (int(f_) >> 4) & 1, int(f_) >> 8)] += 1
print("fast columns by (crowded, tie merged, has_tie, rounds):", sorted(c2.items(), key=lambda kv: -kv[1])[:12])
# examples: for a few slow columns, print value stats
Zc = Z.cpu().numpy() if Z is not None else None
