"""Under torchrun: kernel time of ONE object whose samples are sharded over the ranks
   python -m torch.distributed.run --nproc-per-node W tools/time_wide_sharded.py [N' ...]
(dist.production_table_sample_sharded; the first call of a size is a warm-up)."""
import os, sys, numpy as np, torch, torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymcz_b200 import _scales as sc, dist as mdist
from pymcz_b200.synth import synthetic_table
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
td.init_process_group("nccl", device_id=dev)
meas, err = synthetic_table(8, config=4)
meas, err = meas[:1].copy(), err[:1].copy()
for N in [int(a) for a in sys.argv[1:]] or [1100000, 11000000]:
    ms = []
    for it in range(4):
        out = mdist.production_table_sample_sharded(meas, err, N, sc.T_ALL, sc.DUST_ON, 0xB200, device=dev)
        t = torch.tensor([out["kernel_ms"]], device=dev); td.all_reduce(t, op=td.ReduceOp.MAX)
        ms.append(float(t.item()))
    if rank == 0:
        print("one object, N' = %9d over %d GPUs: %.3f ms (max over ranks, best of 3 after warm-up; all: %s) = %.2e galaxy-samples/s; trips %s"
              % (N, world, min(ms[1:]), ["%.3f" % m for m in ms], N / min(ms[1:]) * 1e3, out["trips"].tolist()), flush=True)
td.destroy_process_group()
