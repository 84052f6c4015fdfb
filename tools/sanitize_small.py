"""Small production + replay launches for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table
G = 320
meas, err = synthetic_table(125000, config=4, start=0, stop=G)
meas[5, 1] = np.nan; err[5, 1] = np.nan          # a galaxy without Hb
err[7] = 0.0                                     # constant columns
m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
for toks, n in ((sc.T_ALL, 2200), (sc.T_ALL | sc.T_M08ALL | sc.T_P05 | sc.T_C01, 700), (sc.T_KD02, 5000)):
    out = ops.forward_production(m, e, n, toks, return_samples=(n == 700))
    torch.cuda.synchronize()
    print("tokens %x n %d ok, median KD02comb of galaxy 0: %.4f" % (toks, n, float(out[0][0, 34, 0])))
