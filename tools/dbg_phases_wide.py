"""Phase cycles of the wide kernel (library built with -DMCZ_PHASE_TIMING): python tools/dbg_phases_wide.py <N'>"""
import sys, ctypes, torch
sys.path.insert(0, '.')
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L = _lib.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1100000
G = 1
meas, err = synthetic_table(8, config=4)
m = torch.from_numpy(meas[:G].copy()).cuda(); e = torch.from_numpy(err[:G].copy()).cuda()
ws = ops.ProductionWorkspace(G, N, sc.T_ALL, m.device)
out = (ctypes.c_ulonglong * 8)()
for _ in range(3):
    ops.forward_production(m, e, N, sc.T_ALL, workspace=ws)
torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); ops.forward_production(m, e, N, sc.T_ALL, workspace=ws); b.record(); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
o = list(out)
names = ["objects", "reset", "A", "KK04 batches", "C", "extremes", "radix rounds", "results"]
tot = sum(o[1:])
print("%s N'=%d: %.3f ms; trips %s" % (L.mcz_production_kernel_name(G, N, sc.T_ALL).decode(), N, a.elapsed_time(b), ws.trips.cpu().numpy().tolist()))
print("  " + "  ".join("%s %.1f us" % (names[i], o[i] / 1965.0) for i in range(1, 8)) + "  | sum %.1f us" % (tot / 1965.0))
