"""Per-phase cycle breakdown of the production kernel (library built with -DMCZ_PHASE_TIMING)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L = _lib.lib()
G = 14800
meas, err = synthetic_table(125000, config=4, start=0, stop=G)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
out = (ctypes.c_ulonglong * 8)()
ws = ops.ProductionWorkspace(G, 2200, sc.T_ALL, m.device)
ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
cc = (ctypes.c_ulonglong * 74)()
L.mcz_debug_col_cycles(cc, 1)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws); b.record(); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
o = list(out)
g = o[0]
print("galaxies %d, kernel %.2f ms" % (g, a.elapsed_time(b)))
print("cycles per galaxy: A %.0f  B+C %.0f  D own(mean over warps) %.0f  D wall %.0f  total %.0f" % (o[1]/g, o[2]/g, o[3]/g, o[4]/g, (o[1]+o[2]+o[4])/g))
L.mcz_debug_col_cycles(cc, 1)
cc = list(cc)
print("B prologue %.0f  KK04 loops %.0f  phase C %.0f" % (o[5]/g, o[6]/g, o[7]/g))
print("sample-parallel stages (cycles per galaxy): D0 %.0f  D1 %.0f  D2 %.0f  D3 %.0f" % tuple(c / g for c in cc[:4]))
for j, k in enumerate(sc.KEYS):
    if cc[j]:
        print("%-10s mean %.0f max %d" % (k, cc[j] / g, cc[37 + j]))
