"""Wide (cooperative, samples tiled across blocks) kernel against the per-galaxy kernel on the same
inputs: MCZ_WIDE=1 forces the wide path, MCZ_WIDE=0 forbids it."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
tok = sc.T_ALL if len(sys.argv) < 4 else sc.parse_md(sys.argv[3])[0]
meas, err = synthetic_table(40, config=4)
meas, err = meas[:G].copy(), err[:G].copy()
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()

def run(wide):
    os.environ["MCZ_WIDE"] = "1" if wide else "0"
    out = ops.forward_production(m, e, N, tok, sc.DUST_ON, 0xB200)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ops.forward_production(m, e, N, tok, sc.DUST_ON, 0xB200)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return [o.cpu().numpy() if o is not None else None for o in out], dt

ref, t_ref = run(False)
wid, t_wid = run(True)
same_trips = np.all(ref[3] == wid[3], axis=1)
ok = np.array_equal(ref[2], wid[2]) and np.array_equal(ref[4], wid[4])
nv_ok = np.array_equal(ref[1][same_trips], wid[1][same_trips])
pct_ok = np.array_equal(ref[0][same_trips].view(np.int32), wid[0][same_trips].view(np.int32))
print("G=%d N'=%d: status/ret equal %s, trips equal for %d/%d galaxies, nvalid equal %s, percentiles bit-equal %s; per-galaxy kernel %.2f ms, wide kernel %.2f ms" % (
    G, N, ok, same_trips.sum(), G, nv_ok, pct_ok, t_ref * 1e3, t_wid * 1e3))
if not pct_ok:
    d = np.nonzero((ref[0][same_trips].view(np.int32) != wid[0][same_trips].view(np.int32)).any(axis=2))
    print("first mismatches (galaxy, key):", list(zip(*d))[:8])
    for gi, ki in list(zip(*d))[:4]:
        print(sc.KEYS[ki], ref[0][same_trips][gi, ki], wid[0][same_trips][gi, ki], ref[1][same_trips][gi, ki], wid[1][same_trips][gi, ki])
sys.exit(0 if (ok and nv_ok and pct_ok) else 1)
