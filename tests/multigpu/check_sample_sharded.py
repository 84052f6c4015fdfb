"""Under torchrun: one object's samples sharded over the ranks (dist.production_table_sample_sharded)
against the single-GPU result for the same seed: status, trips, n and percentiles identical.
   python -m torch.distributed.run --nproc-per-node 2 tests/multigpu/check_sample_sharded.py [N']"""
import os, sys, time, numpy as np, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pymcz_b200 import _scales as sc, ops, dist as mdist
from pymcz_b200.synth import synthetic_table
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
td.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300001
meas, err = synthetic_table(40, config=4)
meas, err = meas[:5].copy(), err[:5].copy()
meas[3, 1] = np.nan; err[3, 1] = np.nan            # one measurement without Hb
tok = sc.T_ALL
# single-GPU reference on every rank (the wide kernel alone, or the per-galaxy kernel for small N')
ref = ops.forward_production(torch.from_numpy(meas).to(dev), torch.from_numpy(err).to(dev), N, tok, sc.DUST_ON, 0xB200)
torch.cuda.synchronize()
ref = [o.cpu().numpy() for o in ref[:5]]
m_d, e_d = torch.from_numpy(meas).to(dev), torch.from_numpy(err).to(dev)
t0 = time.perf_counter()
ops.forward_production(m_d, e_d, N, tok, sc.DUST_ON, 0xB200); torch.cuda.synchronize()
t_one = time.perf_counter() - t0
out = mdist.production_table_sample_sharded(meas, err, N, tok, sc.DUST_ON, 0xB200)
dt = out["kernel_ms"]
same = np.all(ref[3] == out["trips"], axis=1)
ok = (np.array_equal(ref[2], out["status"]) and np.array_equal(ref[4], out["ret"]) and same.mean() >= 0.8 and
      np.array_equal(ref[1][same], out["nvalid"][same]) and
      np.allclose(ref[0][same], out["pct"][same], rtol=0, atol=4e-6, equal_nan=True))
# the drop-in host entry takes this route by itself when there are fewer objects than GPUs
if N >= 262144:
    tab = mdist.production_table(meas[:1], err[:1], N, tok, sc.DUST_ON, 0xB200)
    ok = ok and np.array_equal(tab["status"], ref[2][:1]) and (
        not same[0] or np.allclose(tab["pct"], ref[0][:1], rtol=0, atol=4e-6, equal_nan=True))
flag = torch.tensor([1 if ok else 0], device=dev); td.all_reduce(flag, op=td.ReduceOp.MIN)
if rank == 0:
    print("sample-sharded over %d GPUs == single GPU: %s (N'=%d, trips equal for %d/%d objects; kernel %.2f ms sharded, %.2f ms on one GPU)" % (
        world, bool(flag.item()), N, same.sum(), len(same), dt, t_one * 1e3), flush=True)
td.destroy_process_group()
sys.exit(0 if flag.item() else 1)
