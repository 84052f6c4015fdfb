"""One launch of the production kernel family on a bench.py launch shape, for ncu:
   python tools/prof_bench_shape.py <config 3|4|5> [galaxies]
(the same synthetic table, tokens and seed as bench.py at N=1)."""
import sys, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
cfg = int(sys.argv[1])
shape = {3: (10_000, 11000, "all"), 4: (1_000_000, 2200, "all"), 5: (100_000, 11000, "KD02")}[cfg]
G = int(sys.argv[2]) if len(sys.argv) > 2 else shape[0]
N, md = shape[1], shape[2]
tok = sc.parse_md(md)[0]
meas, err = synthetic_table(G, config=cfg)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, N, tok, m.device)
for _ in range(2):
    ops.forward_production(m, e, N, tok, sc.DUST_ON, seed=0xB200, workspace=ws)
torch.cuda.synchronize()
print("done", cfg, G, N, md)
