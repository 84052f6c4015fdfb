"""Single-object runs on one GPU: per-galaxy kernel (MCZ_WIDE=0) against the cooperative wide kernel."""
import os, sys, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
meas, err = synthetic_table(8, config=4)
m = torch.from_numpy(meas[:1].copy()).cuda(); e = torch.from_numpy(err[:1].copy()).cuda()
for N in (22000, 110000, 1100000, 11000000):
    row = []
    for wide in ("0", "1"):
        os.environ["MCZ_WIDE"] = wide
        ws = ops.ProductionWorkspace(1, N, sc.T_ALL, m.device)
        ops.forward_production(m, e, N, sc.T_ALL, workspace=ws); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            ops.forward_production(m, e, N, sc.T_ALL, workspace=ws)
        b.record(); torch.cuda.synchronize()
        row.append(a.elapsed_time(b) / 3)
    print("one object, N' = %8d: per-galaxy kernel %9.3f ms, wide kernel %7.3f ms (%.2e galaxy-samples/s)" % (N, row[0], row[1], N / row[1] * 1e3))
