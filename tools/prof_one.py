"""One production launch over 14 800 galaxies x 2200 samples (100 per SM), for ncu:
   ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 \
       -o gpurun_out/prof python tools/prof_one.py"""
import sys, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G = 14800
meas, err = synthetic_table(125000, config=4, start=0, stop=G)
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
ws = ops.ProductionWorkspace(G, 2200, sc.T_ALL, m.device)
for _ in range(3):
    ops.forward_production(m, e, 2200, sc.T_ALL, workspace=ws)
torch.cuda.synchronize()
print("done")
