"""How many galaxies of a bench table end their KK04 loops at the 100-trip cap (per loop), and how many
trips the launch executed: python tools/count_caps.py <config 3|4|5> [galaxies]"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
cfg = int(sys.argv[1])
shape = {3: (10_000, 11000, "all"), 4: (100_000, 2200, "all"), 5: (10_000, 11000, "KD02")}[cfg]
G = int(sys.argv[2]) if len(sys.argv) > 2 else shape[0]
N, md = shape[1], shape[2]
tok = sc.parse_md(md)[0]
meas, err = synthetic_table(G, config=cfg)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, N, tok, m.device)
out = ops.forward_production(m, e, N, tok, sc.DUST_ON, seed=0xB200, workspace=ws)
torch.cuda.synchronize()
tr = out[3].cpu().numpy()
print("config %d, %d galaxies x %d: N2Ha loop at the cap %d (%.2f %%), R23 loop at the cap %d (%.2f %%); mean trips %.2f + %.2f" % (
    cfg, G, N, (tr[:, 0] >= 100).sum(), 100.0 * (tr[:, 0] >= 100).mean(), (tr[:, 1] >= 100).sum(), 100.0 * (tr[:, 1] >= 100).mean(),
    tr[:, 0].mean(), tr[:, 1].mean()))
