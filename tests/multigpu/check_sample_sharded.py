"""Under torchrun: one object's samples sharded over the ranks (dist.production_table_sample_sharded)
   python -m torch.distributed.run --nproc-per-node 2 tests/multigpu/check_sample_sharded.py [N']
Checked on every rank:
  * against the ORACLE on the kernel's own samples (read back from the device): present table, trip
    counts, n per scale and the 50/16/84th percentiles of the oracle's float64 columns;
  * against the one-GPU run of the same kernel: bit-identical tables (the cross-rank reductions are
    integer sums and histogram merges, so the number of ranks cannot change a bit)."""
import os, sys, time, numpy as np, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pymcz_b200 import _scales as sc, ops, dist as mdist
from pymcz_b200.synth import synthetic_table
from tests._helpers import kernel_samples, oracle_on_samples
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
td.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300001
meas, err = synthetic_table(40, config=4)
meas, err = meas[:5].copy(), err[:5].copy()
meas[3, 1] = np.nan; err[3, 1] = np.nan            # one measurement without Hb
tok = sc.T_ALL
m_d, e_d = torch.from_numpy(meas).to(dev), torch.from_numpy(err).to(dev)
# one-GPU run of the same launch shape on every rank (the wide kernel for N' >= 32768)
one = ops.forward_production(m_d, e_d, N, tok, sc.DUST_ON, 0xB200)
torch.cuda.synchronize()
t0 = time.perf_counter()
ops.forward_production(m_d, e_d, N, tok, sc.DUST_ON, 0xB200); torch.cuda.synchronize()
t_one = time.perf_counter() - t0
one = [o.cpu().numpy() for o in one[:5]]
out = mdist.production_table_sample_sharded(meas, err, N, tok, sc.DUST_ON, 0xB200)
dt = out["kernel_ms"]
bit = (np.array_equal(one[0].view(np.int32), out["pct"].view(np.int32)) and np.array_equal(one[1], out["nvalid"]) and
       np.array_equal(one[2], out["status"]) and np.array_equal(one[3], out["trips"]) and np.array_equal(one[4], out["ret"]))
# the oracle on the kernel's own samples (rank 0 computes, everyone gets the verdict)
ok = bit
msg = ""
if rank == 0:
    flux, noise = kernel_samples(meas, err, N, 0xB200, 0, tok)
    Zo, po, trips_o, _ = oracle_on_samples(flux, noise, "all", True, procs=5)
    ok = ok and np.array_equal(out["status"] != sc.ST_ABSENT, po) and np.array_equal(out["trips"], trips_o)
    worst = 0.0
    for g in range(meas.shape[0]):
        for j in range(sc.NKEYS):
            col = Zo[g, j][np.isfinite(Zo[g, j])]
            if out["status"][g, j] == sc.ST_OK:
                want = np.percentile(col, [50, 16, 84])
                worst = max(worst, float(np.max(np.abs(out["pct"][g, j] - want))))
                ok = ok and abs(int(out["nvalid"][g, j]) - col.size) <= 2      # (samples within float rounding of a mask edge)
    ok = ok and worst < 2e-4
    msg = "trips %s vs oracle %s; worst |percentile - oracle percentile| = %.2e" % (out["trips"].tolist(), trips_o.tolist(), worst)
# the drop-in host entry takes this route by itself when there are fewer objects than GPUs
if N >= 262144:
    tab = mdist.production_table(meas[:1], err[:1], N, tok, sc.DUST_ON, 0xB200)
    ok = ok and np.array_equal(tab["status"], one[2][:1]) and np.array_equal(tab["pct"].view(np.int32), one[0][:1].view(np.int32))
flag = torch.tensor([1 if ok else 0], device=dev); td.all_reduce(flag, op=td.ReduceOp.MIN)
if rank == 0:
    print("sample-sharded over %d GPUs == oracle and == single GPU (bitwise %s): %s (N'=%d; %s; kernel %.2f ms sharded, %.2f ms on one GPU)" % (
        world, bit, bool(flag.item()), N, msg, dt, t_one * 1e3), flush=True)
td.destroy_process_group()
sys.exit(0 if flag.item() else 1)
