"""Experiment: production-kernel KK04 trip counts and per-sample values against the oracle on the
kernel's OWN samples (mcz_debug_production_samples), for a few thousand synthetic galaxies.
Writes gpurun_out/trip_parity_<tag>.json.  Usage: python tools/trip_parity.py G N md tag"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pymcz_b200 import _scales as sc, ops  # noqa: E402
from pymcz_b200.synth import synthetic_table  # noqa: E402
from tests._helpers import kernel_samples, oracle_on_samples  # noqa: E402

G, n = int(sys.argv[1]), int(sys.argv[2])
mds = sys.argv[3] if len(sys.argv) > 3 else "all"
tag = sys.argv[4] if len(sys.argv) > 4 else "run"
seed, goff = 0x1234ABCD5678, 1000
tok, _, _ = sc.parse_md(mds)
meas, err = synthetic_table(G, config=4)
t0 = time.time()
out = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), n, tok, sc.DUST_ON,
                             seed, goff, False, False, True)
torch.cuda.synchronize()
pct, nvalid, status, trips, ret, Z = [o.cpu().numpy() if o is not None else None for o in out]
flux, noise = kernel_samples(meas, err, n, seed, goff, tok)
t1 = time.time()
Zo, po, trips_o, conv = oracle_on_samples(flux, noise, mds, True)
t2 = time.time()
same = np.all(trips == trips_o, axis=1)
rep = {"G": G, "n": n, "mds": mds, "gpu_s": t1 - t0, "oracle_s": t2 - t1,
       "trip_mismatch_galaxies": int((~same).sum()), "present_equal": bool(np.array_equal(status != sc.ST_ABSENT, po)),
       "mismatches": []}
for g in np.nonzero(~same)[0]:
    c1, c2 = conv[g]
    rep["mismatches"].append({"g": int(g), "gpu": trips[g].tolist(), "oracle": trips_o[g].tolist(),
                              "conv_n2ha_over_tol": [c / 1e-3 for c in c1[:12]] + ["..."] + [c / 1e-3 for c in c1[-3:]],
                              "conv_r23_over_tol": [c / 1e-4 for c in c2[:12]] + ["..."] + [c / 1e-4 for c in c2[-3:]]})
Zd = Z.astype(np.float64)
nan_mism = np.isnan(Zd) != np.isnan(Zo)
both = np.isfinite(Zd) & np.isfinite(Zo)
d = np.abs(np.where(both, Zd - Zo, 0.0))
rep["mask_flips_total"] = int(nan_mism.sum())
rep["mask_flips_in_trip_equal_galaxies"] = int(nan_mism[same].sum())
rep["mask_flips_by_key"] = {sc.KEYS[j]: int(nan_mism[:, j].sum()) for j in range(sc.NKEYS) if nan_mism[:, j].any()}
rep["values_over_1e-3_total"] = int((d > 1e-3).sum())
rep["values_over_1e-3_in_trip_equal_galaxies"] = int((d[same] > 1e-3).sum())
rep["values_over_1e-3_by_key"] = {sc.KEYS[j]: int((d[:, j] > 1e-3).sum()) for j in range(sc.NKEYS) if (d[:, j] > 1e-3).any()}
rep["max_abs_diff_by_key"] = {sc.KEYS[j]: float(d[:, j].max()) for j in range(sc.NKEYS) if both[:, j].any()}
rep["median_abs_diff"] = float(np.median(d[both]))
big = np.argwhere(d > 1e-3)
rep["big_examples"] = [{"g": int(a), "key": sc.KEYS[b], "s": int(c), "gpu": float(Zd[a, b, c]), "oracle": float(Zo[a, b, c])}
                       for a, b, c in big[:40]]
# distribution of the oracle's trip counts and how close its deciding sums come to the tolerance
margins = []
for g in range(G):
    for cv, tol in ((conv[g][0], 1e-3), (conv[g][1], 1e-4)):
        for c in cv:
            if c == c and c > 0:
                margins.append(abs(np.log(c / tol)))
margins = np.array(margins)
rep["decisions"] = int(margins.size)
rep["decisions_within"] = {str(r): int((margins < r).sum()) for r in (1e-4, 1e-3, 3e-3, 1e-2, 3e-2, 1e-1)}
rep["oracle_trip_hist_n2ha"] = np.bincount(np.minimum(trips_o[:, 0], 100), minlength=101).tolist()
rep["oracle_trip_hist_r23"] = np.bincount(np.minimum(trips_o[:, 1], 100), minlength=101).tolist()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "trip_parity_%s.json" % tag), "w"), indent=1)
print(json.dumps({k: v for k, v in rep.items() if k not in ("mismatches", "big_examples", "oracle_trip_hist_n2ha", "oracle_trip_hist_r23")}, indent=1))
for m in rep["mismatches"][:30]:
    print(m["g"], m["gpu"], m["oracle"])
