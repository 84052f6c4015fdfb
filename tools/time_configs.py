"""Kernel-only throughput of the other BASELINE configs (parity-test cases, not bench lines):
cfg3 G=10^4 x N'=11000 md=all and cfg5 G=10^5 x N'=11000 md=KD02 use production_kernel<0> (columns in
an L2-resident global scratch); cfg4-shaped md=KD02 at N'=2200 uses production_kernel<4>."""
import sys, torch
sys.path.insert(0, '.')
from pymcz_b200 import _scales as sc, ops
from pymcz_b200.synth import synthetic_table

def run(name, config, G, n_draw, tok, reps=3):
    meas, err = synthetic_table(G, config=config)
    m = torch.from_numpy(meas).cuda(); e = torch.from_numpy(err).cuda()
    ws = ops.ProductionWorkspace(G, n_draw, tok, m.device)
    ops.forward_production(m, e, n_draw, tok, workspace=ws); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.forward_production(m, e, n_draw, tok, workspace=ws)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-34s G=%d N'=%d: %.2f ms per launch, %.3g galaxy-samples/s" % (name, G, n_draw, ms, G * n_draw / ms * 1e3))

run("cfg3 md=all (global variant)", 3, 10000, 11000, sc.T_ALL)
run("cfg5 md=KD02 (global variant)", 5, 20000, 11000, sc.T_KD02)
run("md=KD02 at N'=2200 (shared variant)", 4, 125000, 2200, sc.T_KD02)
run("cfg2-like N'=22000 md=all", 3, 2000, 22000, sc.T_ALL)
