"""Per-phase cycles of the cluster kernel (library built with -DMCZ_PHASE_TIMING)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L = _lib.lib()
G, N = 2000, 11000
md = sys.argv[1] if len(sys.argv) > 1 else "all"
tok = sc.parse_md(md)[0]
meas, err = synthetic_table(G, config=3)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
out = (ctypes.c_ulonglong * 8)()
ws = ops.ProductionWorkspace(G, N, tok, m.device)
ops.forward_production(m, e, N, tok, workspace=ws); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); ops.forward_production(m, e, N, tok, workspace=ws); b.record(); torch.cuda.synchronize()
L.mcz_debug_phase_cycles(out, 1)
o = list(out); g = max(o[0], 1)
print("kernel %s: galaxies(x blocks) %d, %.2f ms" % (L.mcz_production_kernel_name(G, N, tok).decode(), g, a.elapsed_time(b)))
print("cycles per galaxy-block: A %.0f  B+C %.0f  D own %.0f  D wall %.0f  total %.0f | B prologue %.0f loops %.0f C %.0f" % (
    o[1]/g, o[2]/g, o[3]/g, o[4]/g, (o[1]+o[2]+o[4])/g, o[5]/g, o[6]/g, o[7]/g))
cc = (ctypes.c_ulonglong * 74)()
L.mcz_debug_col_cycles(cc, 1)
cc = list(cc)
names = ["range", "passH", "push", "wait1", "plan", "passG", "wait2", "finish", "wholecol"]
for role, off in (("owner", 0), ("peer", 16)):
    n = max(cc[off + 15], 1)
    print("%s columns %d: " % (role, n) + "  ".join("%s %.0f" % (names[i], cc[off + i] / n) for i in range(9)) +
          "  | total %.0f" % (sum(cc[off:off + 9]) / n))
