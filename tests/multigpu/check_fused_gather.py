"""Under torchrun: the kernel-fused gather (dist.GatheredTables) against NCCL all_gather of the
local tables, bit for bit, plus the time of both:
   python -m torch.distributed.run --nproc-per-node 2 tests/multigpu/check_fused_gather.py [galaxies_per_gpu]"""
import os, sys, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pymcz_b200 import _scales as sc, ops, dist as mdist
from pymcz_b200.synth import synthetic_table, shard_bounds
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
td.init_process_group("nccl", device_id=dev)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
G = per * world + 7                     # uneven shards
lo, hi = shard_bounds(G, world, rank)
meas, err = synthetic_table(G, config=4, start=lo, stop=hi)
m = torch.from_numpy(meas).to(dev); e = torch.from_numpy(err).to(dev)
N, tok = 2200, sc.T_ALL
ws = ops.ProductionWorkspace(hi - lo, N, tok, dev)
gt = mdist.GatheredTables(G, dev)
for t in (gt.pct, gt.nvalid, gt.status):
    t.fill_(-7)                          # poison: every row must be overwritten
td.barrier(); torch.cuda.synchronize()

def fused():
    ops.forward_production(m, e, N, tok, sc.DUST_ON, 0xB200, galaxy_offset=lo, workspace=ws, gather=gt)
    gt.wait()

def nccl():
    ops.forward_production(m, e, N, tok, sc.DUST_ON, 0xB200, galaxy_offset=lo, workspace=ws)
    return mdist.gather_tables(dict(pct=ws.pct, nvalid=ws.nvalid, status=ws.status), G)

fused(); torch.cuda.synchronize()
ref = nccl(); torch.cuda.synchronize()
ok = (torch.equal(gt.nvalid, ref["nvalid"]) and torch.equal(gt.status, ref["status"]) and
      torch.equal(gt.pct.view(torch.int32), ref["pct"].view(torch.int32)))
flag = torch.tensor([1 if ok else 0], device=dev); td.all_reduce(flag, op=td.ReduceOp.MIN)

def timeit(fn, n=5):
    td.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev); td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())

tf, tn = timeit(fused), timeit(nccl)
# the drop-in host entry (what mcz.run calls) takes the fused route under NCCL
import numpy as np
full_m, full_e = synthetic_table(G, config=4)
tab = mdist.production_table(full_m, full_e, N, tok, sc.DUST_ON, 0xB200)
same = (np.array_equal(tab["nvalid"], ref["nvalid"].cpu().numpy()) and
        np.array_equal(tab["pct"].view(np.int32), ref["pct"].cpu().numpy().view(np.int32)) and
        tab["trips"].shape == (G, 2))
flag2 = torch.tensor([1 if same else 0], device=dev); td.all_reduce(flag2, op=td.ReduceOp.MIN)
flag = torch.minimum(flag, flag2)
if rank == 0:
    print("fused gather == all_gather: %s ; ms per step fused %.3f, kernel + NCCL all_gather %.3f (G=%d over %d GPUs)" % (
        bool(flag.item()), tf, tn, G, world), flush=True)
td.destroy_process_group()
sys.exit(0 if flag.item() else 1)
