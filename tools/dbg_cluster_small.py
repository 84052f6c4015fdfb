"""Smallest run of the cluster kernel (a few galaxies, one launch): launch statistics (bit 63 of word 7 =
a protocol wait timed out, bits 56.. = where) and agreement with the global-scratch kernel."""
import os, sys, ctypes, time, torch
sys.path.insert(0, '.')
import numpy as np
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L = _lib.lib()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 11000
md = sys.argv[3] if len(sys.argv) > 3 else "all"
tok = sc.parse_md(md)[0]
meas, err = synthetic_table(max(G, 8), config=3)
m = torch.from_numpy(meas[:G].copy()).cuda(); e = torch.from_numpy(err[:G].copy()).cuda()
res = {}
for mode in ("1", "0"):
    os.environ["MCZ_CLUSTER"] = mode
    ws = ops.ProductionWorkspace(G, N, tok, m.device)
    t0 = time.time()
    ops.forward_production(m, e, N, tok, workspace=ws); torch.cuda.synchronize()
    dt = time.time() - t0
    st = (ctypes.c_ulonglong * 8)()
    L.mcz_production_last_stats(ws.scratch.data_ptr(), st, None)
    st = list(st)
    print("MCZ_CLUSTER=%s %s: %.3f s, stats %s poison %d where %d" % (mode, L.mcz_production_kernel_name(G, N, tok).decode(), dt,
          [x & ((1 << 56) - 1) for x in st], st[7] >> 63, (st[7] >> 56) & 63), flush=True)
    res[mode] = [t.cpu().numpy().copy() for t in (ws.pct, ws.nvalid, ws.status, ws.trips)]
a, b = res["1"], res["0"]
print("trips equal:", np.array_equal(a[3], b[3]), " status equal:", np.array_equal(a[2], b[2]), " nvalid equal:", np.array_equal(a[1], b[1]))
same = np.all(a[3] == b[3], axis=1)
d = np.abs(a[0][same] - b[0][same]); d = d[np.isfinite(d)]
print("max |pct diff| where trips agree:", d.max() if d.size else None, " nan pattern equal:", np.array_equal(np.isnan(a[0][same]), np.isnan(b[0][same])))
