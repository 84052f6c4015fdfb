"""Probe: torch symmetric memory under torchrun (peer pointers for the fused gather)."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1024, dtype=torch.float32, device=dev)
t.fill_(float(rank))
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "world", h.world_size, flush=True)
dist.barrier(); torch.cuda.synchronize()
peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0].item()), flush=True)
dist.barrier()
dist.destroy_process_group()
