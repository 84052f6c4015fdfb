"""Galaxy sharding over the GPUs of one node.

Every galaxy is independent (the reference's loop body has no cross-galaxy state,
/root/reference/pyMCZ/mcz.py:575), so ranks take contiguous galaxy ranges, the Philox
sub-sequence is the GLOBAL galaxy index (results do not depend on the number of ranks) and the
only exchange of the whole path is one all-gather of the per-rank result tables -- the analogue
of the reference's Pool.map return (mcz.py:553-569).  One process per GPU under torchrun;
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as td

from .synth import shard_bounds


def initialized():
    return td.is_available() and td.is_initialized()


def rank():
    return td.get_rank() if initialized() else 0


def world_size():
    return td.get_world_size() if initialized() else 1


def all_gather_rows(local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """Concatenate row-sharded tensors (rank r holds counts[r] leading rows) on every rank.

    Shards may be uneven: each rank pads to max(counts) rows, one all_gather, then the padding is
    dropped.  Works with any backend (tensors stay on their device)."""
    world = len(counts)
    if world == 1:
        return local
    cmax = max(counts)
    pad_shape = (cmax,) + tuple(local.shape[1:])
    send = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    recv = torch.empty((world,) + pad_shape, dtype=local.dtype, device=local.device)
    td.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group) if recv.is_cuda else \
        td.all_gather(list(recv.unbind(0)), send, group=group)
    return torch.cat([recv[r, : counts[r]] for r in range(world)], dim=0)


def gather_tables(tables: dict, G: int, group=None) -> dict:
    """all-gather every tensor of ``tables`` (row-sharded by shard_bounds(G, world, rank))."""
    world = world_size()
    counts = [shard_bounds(G, world, r)[1] - shard_bounds(G, world, r)[0] for r in range(world)]
    return {k: (all_gather_rows(v, counts, group) if v is not None else None) for k, v in tables.items()}


def production_table(meas, err, n_draw, tokens, dust, seed, correlated=False, deterministic=False,
                     return_samples=False, device=None):
    """Run the fused production kernel over a whole host table, sharded over the ranks of the
    default process group (or on one GPU when torch.distributed is not initialised).

    meas, err: float32 [G, 11] host arrays (every rank passes the same table).
    Returns a dict of host numpy arrays with the FULL tables on every rank:
    pct [G,37,3] f32, nvalid [G,37] i32, status [G,37] i8, trips [G,2] i32, ret [G] i32,
    Z [G,37,n_draw] f32 or None.
    """
    from . import ops
    if not torch.cuda.is_available():
        raise RuntimeError("pymcz_b200 needs a CUDA device: there is no CPU fallback")
    meas = np.ascontiguousarray(meas, dtype=np.float32)
    err = np.ascontiguousarray(err, dtype=np.float32)
    G = meas.shape[0]
    world, r = world_size(), rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    lo, hi = shard_bounds(G, world, r)
    m = torch.from_numpy(meas[lo:hi]).to(device)
    e = torch.from_numpy(err[lo:hi]).to(device)
    if hi > lo:
        pct, nvalid, status, trips, ret, Z = ops.forward_production(
            m, e, n_draw, tokens, dust, seed, galaxy_offset=lo, correlated=correlated,
            deterministic=deterministic, return_samples=return_samples)
    else:
        from ._scales import NKEYS
        pct = torch.empty((0, NKEYS, 3), dtype=torch.float32, device=device)
        nvalid = torch.empty((0, NKEYS), dtype=torch.int32, device=device)
        status = torch.empty((0, NKEYS), dtype=torch.int8, device=device)
        trips = torch.empty((0, 2), dtype=torch.int32, device=device)
        ret = torch.empty((0,), dtype=torch.int32, device=device)
        Z = torch.empty((0, NKEYS, n_draw), dtype=torch.float32, device=device) if return_samples else None
    out = gather_tables(dict(pct=pct, nvalid=nvalid, status=status, trips=trips, ret=ret, Z=Z), G)
    torch.cuda.synchronize(device)
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
