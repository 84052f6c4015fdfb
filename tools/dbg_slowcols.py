"""Why are some columns slow?  (library built with -DMCZ_PHASE_TIMING -DMCZ_COLTIME_DUMP)"""
import sys, numpy as np, torch, collections
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G = 1480
meas, err = synthetic_table(125000, config=4, start=0, stop=G)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
pct, nv, st, tr, ret, Z = ops.forward_production(m, e, 2200, sc.T_ALL, return_samples=True)
torch.cuda.synchronize()
t = pct[:, :, 2].cpu().numpy().astype(np.float64); st = st.cpu().numpy(); nv = nv.cpu().numpy(); Z = Z.cpu().numpy()
slow = (t > 40000) & (st == 0)
print("slow", slow.sum())
shown = collections.Counter()
for g, j in zip(*np.nonzero(slow)):
    key = sc.KEYS[j]
    if shown[key] >= 3: continue
    shown[key] += 1
    x = Z[g, j]; fin = np.isfinite(x); xf = x[fin]
    p = x[:64]; p = np.sort(p[np.isfinite(p)]); mm = len(p)
    qlo = p[int(0.06 * (mm - 1))]; qhi = p[mm - 1 - int(0.06 * (mm - 1))]
    w = 1.5 * (qhi - qlo); L = qlo - w; H = qhi + w
    f = np.clip(((xf - L) * (1024 / (H - L))), 0, 1023).astype(int)
    h = np.bincount(f, minlength=1024)
    vals, cnts = np.unique(xf, return_counts=True)
    q = np.percentile(xf, [16, 50, 84])
    fq = np.clip(((q - L) * (1024 / (H - L))), 0, 1023).astype(int)
    print("%-10s g %d flags %x n %d pilot m %d range [%.4f, %.4f] data [%.4f, %.4f] pcts %s bins %s occupancy %s maxbin %d end bins %d %d top repeats %s" % (
        key, g, nv[g, j] >> 16, len(xf), mm, L, H, xf.min(), xf.max(), np.round(q, 4), fq, h[fq], h.max(), h[0], h[1023],
        list(zip(np.round(vals[np.argsort(-cnts)[:3]], 4), np.sort(cnts)[::-1][:3]))))
