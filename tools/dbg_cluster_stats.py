import sys, ctypes, torch
sys.path.insert(0, '.')
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L = _lib.lib()
for md, G, N in (("all", 2000, 11000), ("KD02", 2000, 11000), ("all", 2000, 4608)):
    tok = sc.parse_md(md)[0]
    meas, err = synthetic_table(G, config=3)
    m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
    ws = ops.ProductionWorkspace(G, N, tok, m.device)
    ops.forward_production(m, e, N, tok, workspace=ws); torch.cuda.synchronize()
    st = (ctypes.c_ulonglong * 8)()
    L.mcz_production_last_stats(ws.scratch.data_ptr(), st, None)
    print(md, N, L.mcz_production_kernel_name(G, N, tok).decode(), list(st))
