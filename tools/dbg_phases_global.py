"""Per-phase cycles of production_kernel<0> (N' = 11000, columns in global scratch); library built with -DMCZ_PHASE_TIMING."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L = _lib.lib()
G, N = 1480, 11000
tok = sc.T_ALL if len(sys.argv) < 2 else sc.T_KD02
meas, err = synthetic_table(10000, config=3, start=0, stop=G)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, N, tok, m.device)
out = (ctypes.c_ulonglong * 8)()
ops.forward_production(m, e, N, tok, workspace=ws); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); ops.forward_production(m, e, N, tok, workspace=ws); b.record(); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
o = list(out); g = o[0]
print("galaxies %d, kernel %.2f ms; cycles per galaxy: A %.0f  B+C %.0f  D own %.0f  D wall %.0f  total %.0f" % (
    g, a.elapsed_time(b), o[1] / g, o[2] / g, o[3] / g, o[4] / g, (o[1] + o[2] + o[4]) / g))
