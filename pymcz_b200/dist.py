"""Galaxy sharding over the GPUs of one node.

Every galaxy is independent (the reference's loop body has no cross-galaxy state,
/root/reference/pyMCZ/mcz.py:575), so ranks take contiguous galaxy ranges, the Philox
sub-sequence is the GLOBAL galaxy index (results do not depend on the number of ranks) and the
only exchange of the whole path is the gather of the per-rank result tables -- the analogue
of the reference's Pool.map return (mcz.py:553-569).  On NCCL it is fused into the kernel
(GatheredTables: every rank's kernel writes its rows into all ranks' tables over NVLink peer
mappings); the all-gather form remains for the per-sample output and for gloo.  One process per GPU under torchrun;
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


class GatheredTables:
    """Full-size result tables (pct [G,37,3] f32, nvalid [G,37] i32, status [G,37] i8) on every rank,
    allocated as torch symmetric memory so that every rank holds NVLink peer mappings of all of
    them.  ops.forward_production(..., gather=self) makes the production kernel write each result
    row into all of them as it is produced: the gather of the reference's Pool.map results
    (mcz.py:553-569) is fused into the kernel and no collective moves data.  wait() is the one
    barrier after which the tables are complete everywhere.  Needs an initialised NCCL group."""

    def __init__(self, G_total: int, device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from ._scales import NKEYS
        from .synth import shard_bounds as _sb
        self.group = group if group is not None else td.group.WORLD
        self.world = td.get_world_size(self.group)
        self.rank = td.get_rank(self.group)
        if self.world > 8:
            raise ValueError("GatheredTables: at most 8 ranks (one node)")
        self.G_total = int(G_total)
        self.row0 = _sb(self.G_total, self.world, self.rank)[0]
        rows = max(self.G_total, 1)
        self.pct = symm.empty((rows, NKEYS, 3), dtype=torch.float32, device=device)
        self.nvalid = symm.empty((rows, NKEYS), dtype=torch.int32, device=device)
        self.status = symm.empty((rows, NKEYS), dtype=torch.int8, device=device)
        self._handles = [symm.rendezvous(t, self.group) for t in (self.pct, self.nvalid, self.status)]
        arr = ctypes.c_void_p * self.world
        self.pct_ptrs = arr(*[int(p) for p in self._handles[0].buffer_ptrs])
        self.nvalid_ptrs = arr(*[int(p) for p in self._handles[1].buffer_ptrs])
        self.status_ptrs = arr(*[int(p) for p in self._handles[2].buffer_ptrs])
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)

    def wait(self):
        """Every rank's launch has finished, so every row of the local tables has been written.
        (A 4-byte all-reduce enqueued after the kernel: it cannot complete on any rank before all
        ranks have reached it.)"""
        td.all_reduce(self._flag, group=self.group)


def production_table_sample_sharded(meas, err, n_draw, tokens, dust, seed, correlated=False, group=None,
                                    device=None):
    """Single-object runs: the samples of every galaxy sharded over the ranks.

    Every rank passes the same small table; rank r evaluates samples [N' r / W, N' (r+1) / W) of each
    galaxy and the reductions over the whole sample vector (has* flags, KK04 loop means per trip,
    order statistics) go through NVLink inside the kernel.  Returns the same dict of host arrays as
    production_table (without Z) on every rank; results do not depend on the number of ranks."""
    import ctypes
    import torch.distributed._symmetric_memory as symm
    from . import _lib
    from ._scales import NKEYS
    L = _lib.lib()
    group = group if group is not None else td.group.WORLD
    world, r = td.get_world_size(group), td.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    meas = torch.from_numpy(np.ascontiguousarray(meas, dtype=np.float32)).to(device)
    err = torch.from_numpy(np.ascontiguousarray(err, dtype=np.float32)).to(device)
    G = int(meas.shape[0])
    shared = symm.empty(int(L.mcz_samples_shared_bytes(int(tokens))), dtype=torch.uint8, device=device)
    shared.zero_()
    handle = symm.rendezvous(shared, group)
    ptrs = (ctypes.c_void_p * world)(*[int(p) for p in handle.buffer_ptrs])
    scratch = torch.empty(int(L.mcz_samples_scratch_bytes(int(n_draw), int(tokens))), dtype=torch.uint8, device=device)
    pct = torch.empty((G, NKEYS, 3), dtype=torch.float32, device=device)
    nvalid = torch.empty((G, NKEYS), dtype=torch.int32, device=device)
    status = torch.empty((G, NKEYS), dtype=torch.int8, device=device)
    trips = torch.empty((G, 2), dtype=torch.int32, device=device)
    ret = torch.empty((G,), dtype=torch.int32, device=device)
    torch.cuda.synchronize(device)
    td.barrier(group)                       # every rank's shared region is zeroed before any kernel starts
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    with torch.cuda.device(device):
        rc = L.mcz_forward_production_samples(
            meas.data_ptr(), err.data_ptr(), G, int(n_draw), int(tokens), int(dust), int(seed), 0,
            1 if correlated else 0, 0, world, r, ptrs, pct.data_ptr(), nvalid.data_ptr(), status.data_ptr(),
            trips.data_ptr(), ret.data_ptr(), scratch.data_ptr(), scratch.numel(),
            torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "mcz_forward_production_samples")
    ev1.record()
    torch.cuda.synchronize(device)
    td.barrier(group)
    return dict(pct=pct.cpu().numpy(), nvalid=nvalid.cpu().numpy(), status=status.cpu().numpy(),
                trips=trips.cpu().numpy(), ret=ret.cpu().numpy(), Z=None, kernel_ms=ev0.elapsed_time(ev1))


class _HostPath:
    """Per-device state of production_table: the C-ABI host-buffer context (mcz_ctx_create) and
    pinned result tables, both reused from call to call (allocating 640 MB of pinned memory costs
    more than the whole 10^6-galaxy run)."""

    def __init__(self, device):
        from . import _lib
        self.device = torch.device(device)
        self.L = _lib.lib()
        self.ctx = self.L.mcz_ctx_create(self.device.index if self.device.index is not None else torch.cuda.current_device())
        if not self.ctx:
            raise RuntimeError("mcz_ctx_create failed: " + self.L.mcz_last_error().decode())
        self.cap = 0
        self.out = None
        self.gathered = {}              # G_total -> GatheredTables (symmetric-memory rendezvous is slow: keep)
        self.copy_stream = None

    def tables(self, G):
        from ._scales import NKEYS
        if G > self.cap:
            self.out = dict(pct=torch.empty((G, NKEYS, 3), dtype=torch.float32).pin_memory(),
                            nvalid=torch.empty((G, NKEYS), dtype=torch.int32).pin_memory(),
                            status=torch.empty((G, NKEYS), dtype=torch.int8).pin_memory(),
                            trips=torch.empty((G, 2), dtype=torch.int32).pin_memory(),
                            ret=torch.empty((G,), dtype=torch.int32).pin_memory())
            self.cap = G
        return {k: v[:G] for k, v in self.out.items()}

    def close(self):
        if self.ctx:
            self.L.mcz_ctx_destroy(self.ctx)
            self.ctx = None


_host_paths = {}


def _host_path(device):
    key = str(device)
    if key not in _host_paths:
        _host_paths[key] = _HostPath(device)
    return _host_paths[key]


def release_host_paths():
    """Free the cached contexts / pinned tables of production_table."""
    for hp in _host_paths.values():
        hp.close()
    _host_paths.clear()


def _have_symm() -> bool:
    import importlib
    try:
        importlib.import_module("torch.distributed._symmetric_memory")
        return True
    except ImportError:
        return False


def _agree(ok: bool, device) -> bool:
    """True iff `ok` holds on every rank (the ranks must take the same path through a collective)."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    return bool(int(flag.item()))


def production_table(meas, err, n_draw, tokens, dust, seed, correlated=False, deterministic=False,
                     return_samples=False, device=None, copy=True):
    """Run the fused production kernel over a whole host table, sharded over the ranks of the
    default process group (or on one GPU when torch.distributed is not initialised).  This is the
    call mcz.run makes (the reference's loop over measurements, mcz.py:545-607).

    meas, err: float32 [G, 11] host arrays (every rank passes the same table); pinned memory makes
    the uploads asynchronous.
    Returns a dict of host numpy arrays with the FULL tables on every rank:
    pct [G,37,3] f32, nvalid [G,37] i32, status [G,37] i8, trips [G,2] i32, ret [G] i32,
    Z [G,37,n_draw] f32 or None.  With copy=False the tables are views of pinned buffers that the
    next call on this device overwrites.

    One GPU: the C ABI's host-buffer entry point (mcz_ctx_run_host: upload, up to 8 chunked
    launches, result tables copied back while the next chunk computes).  Several GPUs (NCCL): every
    rank uploads its shard, the kernel writes its result rows into every rank's full tables over
    NVLink (GatheredTables), one barrier, and every rank copies the full tables to its host.
    """
    from . import ops
    from ._scales import NKEYS
    if not torch.cuda.is_available():
        raise RuntimeError("pymcz_b200 needs a CUDA device: there is no CPU fallback")
    meas = np.ascontiguousarray(meas, dtype=np.float32)
    err = np.ascontiguousarray(err, dtype=np.float32)
    G = meas.shape[0]
    world, r = world_size(), rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    done = (lambda t: {k: (np.array(v) if (copy and v is not None) else v) for k, v in t.items()})

    if world == 1 and not return_samples:
        hp = _host_path(device)
        t = hp.tables(max(G, 1))
        rc = hp.L.mcz_ctx_run_host(hp.ctx, meas.ctypes.data, err.ctypes.data, G, int(n_draw), int(tokens), int(dust),
                                   int(seed), 0, 1 if correlated else 0, 1 if deterministic else 0,
                                   t["pct"].data_ptr(), t["nvalid"].data_ptr(), t["status"].data_ptr(),
                                   t["trips"].data_ptr(), t["ret"].data_ptr())
        from . import _lib
        _lib.check(rc, "mcz_ctx_run_host")
        out = {k: v[:G].numpy() for k, v in t.items()}
        out["Z"] = None
        return done(out)

    nccl = world > 1 and td.get_backend() == "nccl"
    if nccl and G < world and n_draw >= 262144 and not return_samples and not deterministic:
        # single-object run: fewer galaxies than GPUs, very many samples each -> shard the samples
        if _agree(_have_symm(), device):
            out = production_table_sample_sharded(meas, err, n_draw, tokens, dust, seed, correlated=correlated,
                                                  device=device)
            out.pop("kernel_ms", None)
            return out
    lo, hi = shard_bounds(G, world, r)
    m = torch.from_numpy(meas[lo:hi]).to(device, non_blocking=True)
    e = torch.from_numpy(err[lo:hi]).to(device, non_blocking=True)
    gt = None
    if nccl and not return_samples:
        # the gather is fused into the kernel: result rows go straight into every rank's tables.
        # Whether the peer mappings exist is decided COLLECTIVELY: all ranks take the same path.
        hp = _host_path(device)
        gt = hp.gathered.get(G)
        if gt is None:
            if _agree(_have_symm(), device):
                try:
                    gt = GatheredTables(G, device)
                except RuntimeError as exc:            # peer mapping refused on this rank
                    gt = None
                    import warnings
                    warnings.warn("fused gather unavailable on rank %d (%s): NCCL all-gathers instead" % (r, exc))
                if not _agree(gt is not None, device):
                    gt = None
                else:
                    hp.gathered[G] = gt
    if gt is not None:
        # (cached tables: no rank's kernel may write into a peer's table while that peer still
        # copies the previous call's rows out -- the same 4-byte all-reduce, ahead of the kernel)
        gt.wait()
        hp = _host_path(device)
        t = hp.tables(max(G, 1))
        # The shard runs as up to 4 launches; after each one (and the barrier that says every rank's
        # launch of that chunk has landed its rows everywhere) the rows of that chunk -- of EVERY
        # rank's shard -- go down to the host on a second stream while the next chunk computes.
        per = hi - lo
        K = 4 if G >= 4 * 16384 * world else 1
        if hp.copy_stream is None:
            hp.copy_stream = torch.cuda.Stream(device)
        trips_l = torch.empty((per, 2), dtype=torch.int32, device=device)
        ret_l = torch.empty((per,), dtype=torch.int32, device=device)
        ws_cache = {}
        for k in range(K):
            a, b = lo + per * k // K, lo + per * (k + 1) // K
            if b > a:
                ws = ws_cache.get(b - a)
                if ws is None:
                    ws = ws_cache[b - a] = ops.ProductionWorkspace(b - a, n_draw, tokens, device)
                _, _, _, tr, rt, _ = ops.forward_production(
                    m[a - lo:b - lo], e[a - lo:b - lo], n_draw, tokens, dust, seed, galaxy_offset=a,
                    correlated=correlated, deterministic=deterministic, gather=gt, row0=a, workspace=ws)
                trips_l[a - lo:b - lo].copy_(tr)
                ret_l[a - lo:b - lo].copy_(rt)
            gt.wait()
            done_ev = torch.cuda.Event()
            done_ev.record()
            with torch.cuda.stream(hp.copy_stream):
                hp.copy_stream.wait_event(done_ev)
                for q in range(world):
                    qlo, qhi = shard_bounds(G, world, q)
                    ra, rb = qlo + (qhi - qlo) * k // K, qlo + (qhi - qlo) * (k + 1) // K
                    if rb > ra:
                        for name, src in (("pct", gt.pct), ("nvalid", gt.nvalid), ("status", gt.status)):
                            t[name][ra:rb].copy_(src[ra:rb], non_blocking=True)
        small = gather_tables(dict(trips=trips_l, ret=ret_l), G)
        t["trips"][:G].copy_(small["trips"][:G], non_blocking=True)
        t["ret"][:G].copy_(small["ret"][:G], non_blocking=True)
        torch.cuda.synchronize(device)
        out = {k: v[:G].numpy() for k, v in t.items()}
        out["Z"] = None
        return done(out)
    if hi > lo:
        pct, nvalid, status, trips, ret, Z = ops.forward_production(
            m, e, n_draw, tokens, dust, seed, galaxy_offset=lo, correlated=correlated,
            deterministic=deterministic, return_samples=return_samples)
    else:
        pct = torch.empty((0, NKEYS, 3), dtype=torch.float32, device=device)
        nvalid = torch.empty((0, NKEYS), dtype=torch.int32, device=device)
        status = torch.empty((0, NKEYS), dtype=torch.int8, device=device)
        trips = torch.empty((0, 2), dtype=torch.int32, device=device)
        ret = torch.empty((0,), dtype=torch.int32, device=device)
        Z = torch.empty((0, NKEYS, n_draw), dtype=torch.float32, device=device) if return_samples else None
    out = gather_tables(dict(pct=pct, nvalid=nvalid, status=status, trips=trips, ret=ret, Z=Z), G)
    torch.cuda.synchronize(device)
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
