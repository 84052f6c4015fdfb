"""Tensor-level entry points over the C ABI: torch owns device memory and streams, the kernels
do the work.  These are the three ops of SURVEY.md section 8(b).

* ``forward_replay``       -- metallicity.calculation on caller-supplied samples
                              (/root/reference/pyMCZ/metallicity.py:86-257)
* ``segmented_percentiles``-- savehist's statistics (/root/reference/pyMCZ/mcz.py:273-301)
* ``forward_production``   -- sampling + calculation + percentiles fused
                              (mcz.py:509-519, 436-442, 580-589; metallicity.py:86-257; mcz.py:273-301)
"""
from __future__ import annotations

import torch

from . import _lib
from ._scales import NKEYS, NLINES, DUST_ON


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (pymcz_b200 has no CPU path)" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


AUX_NAMES = ("logN2O2", "logN2S2", "N2S2", "R2", "R3", "R23", "Z_init_guess", "logq", "N2O2_root_hi", "N2O2_root_lo")


def forward_replay(flux, noise, tokens, dust=DUST_ON, precision="f64", return_aux=False):
    """flux [11,G,N] f64 (cuda), noise [5,G,N] f64 or None ->
    (Z [G,37,N] f64, present [G] int64 bit masks, trips [G,2] i32, ret [G] i32)
    and, with return_aux, aux [G,10,N] f64: the intermediate quantities AUX_NAMES the reference keeps
    as attributes of the diagnostics object (metscales.py:393-440, 468-489, 1112)."""
    L = _lib.lib()
    _need_cuda(flux, "flux")
    if flux.dtype != torch.float64 or flux.dim() != 3 or flux.shape[0] != NLINES:
        raise ValueError("flux must be float64 [11, G, N]")
    G, N = int(flux.shape[1]), int(flux.shape[2])
    if noise is not None:
        _need_cuda(noise, "noise")
        if noise.dtype != torch.float64 or tuple(noise.shape) != (5, G, N):
            raise ValueError("noise must be float64 [5, G, N]")
    prec = {"f64": 0, "f32": 1}[precision]
    dev = flux.device
    with torch.cuda.device(dev):
        Z = torch.empty((G, NKEYS, N), dtype=torch.float64, device=dev)
        present = torch.zeros((G,), dtype=torch.int64, device=dev)
        trips = torch.zeros((G, 2), dtype=torch.int32, device=dev)
        ret = torch.zeros((G,), dtype=torch.int32, device=dev)
        nbytes = L.mcz_replay_scratch_bytes(G, N, prec)
        scratch = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)
        aux = torch.empty((G, len(AUX_NAMES), N), dtype=torch.float64, device=dev) if return_aux else None
        rc = L.mcz_forward_replay_aux(flux.data_ptr(), noise.data_ptr() if noise is not None else None,
                                      G, N, int(tokens), int(dust), prec, Z.data_ptr(), present.data_ptr(),
                                      trips.data_ptr(), ret.data_ptr(), aux.data_ptr() if aux is not None else None,
                                      scratch.data_ptr(), nbytes, _stream_ptr(dev))
        _lib.check(rc, "mcz_forward_replay_aux")
        # scratch must outlive the kernel: tie its lifetime to the stream
        scratch.record_stream(torch.cuda.current_stream(dev))
    if return_aux:
        return Z, present, trips, ret, aux
    return Z, present, trips, ret


def segmented_percentiles(Z, present=None):
    """Z [..., N] f64 (cuda); present optional bool/uint8 [...]. ->
    (pct [..., 3] f64, nvalid [...] i32, status [...] i8)."""
    L = _lib.lib()
    _need_cuda(Z, "Z")
    if Z.dtype != torch.float64:
        raise ValueError("Z must be float64")
    lead = tuple(Z.shape[:-1])
    N = int(Z.shape[-1])
    rows = 1
    for d in lead:
        rows *= int(d)
    dev = Z.device
    pres = None
    if present is not None:
        pres = present.to(device=dev, dtype=torch.uint8).contiguous()
        if pres.numel() != rows:
            raise ValueError("present must have one entry per row")
    with torch.cuda.device(dev):
        pct = torch.empty(lead + (3,), dtype=torch.float64, device=dev)
        nvalid = torch.empty(lead, dtype=torch.int32, device=dev)
        status = torch.empty(lead, dtype=torch.int8, device=dev)
        rc = L.mcz_segmented_percentiles(Z.data_ptr(), pres.data_ptr() if pres is not None else None,
                                         rows, N, pct.data_ptr(), nvalid.data_ptr(), status.data_ptr(),
                                         _stream_ptr(dev))
        _lib.check(rc, "mcz_segmented_percentiles")
    return pct, nvalid, status


class ProductionWorkspace:
    """Reusable output tables + scratch for ``forward_production`` (no allocation per call)."""

    def __init__(self, G, n_draw, tokens, device, return_samples=False):
        L = _lib.lib()
        self.G, self.n_draw, self.tokens = int(G), int(n_draw), int(tokens)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            d = self.device
            self.pct = torch.empty((G, NKEYS, 3), dtype=torch.float32, device=d)
            self.nvalid = torch.empty((G, NKEYS), dtype=torch.int32, device=d)
            self.status = torch.empty((G, NKEYS), dtype=torch.int8, device=d)
            self.trips = torch.empty((G, 2), dtype=torch.int32, device=d)
            self.ret = torch.empty((G,), dtype=torch.int32, device=d)
            self.Z = (torch.empty((G, NKEYS, n_draw), dtype=torch.float32, device=d)
                      if return_samples else None)
            self.scratch_bytes = L.mcz_production_scratch_bytes(G, n_draw, self.tokens)
            self.scratch = torch.empty((max(self.scratch_bytes, 256),), dtype=torch.uint8, device=d)


def forward_production(meas, err, n_draw, tokens, dust=DUST_ON, seed=0xB200, galaxy_offset=0,
                       correlated=False, deterministic=False, return_samples=False, workspace=None,
                       gather=None, row0=None):
    """meas, err [G,11] f32 (cuda) -> (pct [G,37,3] f32, nvalid [G,37] i32, status [G,37] i8,
    trips [G,2] i32, ret [G] i32, Z [G,37,n_draw] f32 or None).  Enqueues on the current stream.

    gather: a dist.GatheredTables.  The kernel then writes this rank's result rows straight into
    every rank's full tables (rows gather.row0 ...) over NVLink peer mappings instead of the
    workspace's local pct/nvalid/status, which are returned untouched; call gather.wait() before
    reading the tables.  row0: first row of this launch in the gathered tables (default: the rank's
    first row; a shard run as several launches passes each chunk's own)."""
    L = _lib.lib()
    _need_cuda(meas, "meas")
    _need_cuda(err, "err")
    if meas.dtype != torch.float32 or err.dtype != torch.float32:
        raise ValueError("meas/err must be float32 (the reference reader parses fluxes as 'f')")
    if meas.dim() != 2 or meas.shape[1] != NLINES or tuple(err.shape) != tuple(meas.shape):
        raise ValueError("meas/err must be [G, 11]")
    G = int(meas.shape[0])
    dev = meas.device
    ws = workspace
    if ws is None:
        ws = ProductionWorkspace(G, n_draw, tokens, dev, return_samples)
    elif (ws.G, ws.n_draw, ws.tokens) != (G, int(n_draw), int(tokens)) or ws.device != dev:
        raise ValueError("workspace does not match this call")
    if gather is not None:
        if return_samples:
            raise ValueError("return_samples is not available with gather")
        with torch.cuda.device(dev):
            rc = L.mcz_forward_production_gather(
                meas.data_ptr(), err.data_ptr(), G, int(n_draw), int(tokens), int(dust), int(seed),
                int(galaxy_offset), 1 if correlated else 0, 1 if deterministic else 0,
                gather.world, int(gather.row0 if row0 is None else row0), gather.pct_ptrs, gather.nvalid_ptrs,
                gather.status_ptrs,
                ws.trips.data_ptr(), ws.ret.data_ptr(), ws.scratch.data_ptr(), ws.scratch_bytes + 0,
                _stream_ptr(dev))
            _lib.check(rc, "mcz_forward_production_gather")
            if workspace is None:
                ws.scratch.record_stream(torch.cuda.current_stream(dev))
        return ws.pct, ws.nvalid, ws.status, ws.trips, ws.ret, None
    with torch.cuda.device(dev):
        rc = L.mcz_forward_production(
            meas.data_ptr(), err.data_ptr(), G, int(n_draw), int(tokens), int(dust), int(seed),
            int(galaxy_offset), 1 if correlated else 0, 1 if deterministic else 0,
            ws.pct.data_ptr(), ws.nvalid.data_ptr(), ws.status.data_ptr(), ws.trips.data_ptr(),
            ws.ret.data_ptr(), ws.Z.data_ptr() if ws.Z is not None else None,
            ws.scratch.data_ptr(), ws.scratch_bytes + 0, _stream_ptr(dev))
        _lib.check(rc, "mcz_forward_production")
        if workspace is None:
            ws.scratch.record_stream(torch.cuda.current_stream(dev))
    return ws.pct, ws.nvalid, ws.status, ws.trips, ws.ret, ws.Z


def present_to_bool(present):
    """[G] int64 bit masks -> [G,37] bool tensor."""
    bits = torch.arange(NKEYS, device=present.device, dtype=torch.int64)
    return ((present.unsqueeze(-1) >> bits) & 1).bool()


def histogram_assist(Z, bin_numbers):
    """Z [..., N] f64 (cuda) distributions; bin_numbers: iterable of ints m in [1, 8192] ->
    (stats [..., 6] f64 = n, min, max, mean, std, sum of the finite values;
     counts: list with one int32 tensor [..., m] per requested m, equal to
     np.histogram(data, np.linspace(min(data), max(data), m + 1))[0] of each row's finite values).
    Replaces the data-parallel inside of savehist's binning (/root/reference/pyMCZ/mcz.py:90-99, 289)."""
    import ctypes
    L = _lib.lib()
    _need_cuda(Z, "Z")
    if Z.dtype != torch.float64:
        raise ValueError("Z must be float64")
    lead = tuple(Z.shape[:-1])
    N = int(Z.shape[-1])
    rows = 1
    for d in lead:
        rows *= int(d)
    ms = [int(m) for m in bin_numbers]
    total = sum(ms)
    dev = Z.device
    with torch.cuda.device(dev):
        stats = torch.empty((rows, 6), dtype=torch.float64, device=dev)
        counts = torch.empty((rows, max(total, 1)), dtype=torch.int32, device=dev)
        scratch = torch.empty((max(len(ms), 1),), dtype=torch.int32, device=dev)
        arr = (ctypes.c_int * max(len(ms), 1))(*ms)
        rc = L.mcz_histogram_assist(Z.data_ptr(), rows, N, arr, len(ms), stats.data_ptr(), counts.data_ptr(),
                                    scratch.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "mcz_histogram_assist")
        scratch.record_stream(torch.cuda.current_stream(dev))
    out, o = [], 0
    for m in ms:
        out.append(counts[:, o:o + m].reshape(lead + (m,)))
        o += m
    return stats.reshape(lead + (6,)), out


def ks_2samp(a, b):
    """Two-sample Kolmogorov-Smirnov statistic of two 1-D float64 CUDA tensors (finite values, any
    order): tensor [3] = D (two-sided), D+, D- exactly as scipy.stats.ks_2samp forms them
    (/root/reference/pyMCZ/testcompleteness.py:111).  The p-value is host work."""
    L = _lib.lib()
    _need_cuda(a, "a")
    _need_cuda(b, "b")
    if a.dtype != torch.float64 or b.dtype != torch.float64 or a.dim() != 1 or b.dim() != 1:
        raise ValueError("a, b must be 1-D float64")
    dev = a.device
    with torch.cuda.device(dev):
        out = torch.empty((3,), dtype=torch.float64, device=dev)
        scratch = torch.empty((2,), dtype=torch.int64, device=dev)
        rc = L.mcz_ks_2samp(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), out.data_ptr(), scratch.data_ptr(),
                            _stream_ptr(dev))
        _lib.check(rc, "mcz_ks_2samp")
        scratch.record_stream(torch.cuda.current_stream(dev))
    return out
