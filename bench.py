#!/usr/bin/env python
"""Benchmark of the pyMCZ Monte-Carlo metallicity hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code

Metric (BASELINE.json): galaxy-samples/s, all scales (--md all), dust-corrected.

Workloads (BASELINE.json `configs`), chosen with --config:
  4 (default)  synthetic 10^6 HII-region table x nsample=2000 (N'=2200), --md all
  3            synthetic 10^4 galaxies x nsample=10^4 (N'=11000), --md all
  5            KD02comb-only stress: 10^5 galaxies x nsample=10^4 (N'=11000), --md KD02
Scaling is STRONG by default: the configuration's whole table at every N, sharded over galaxies
(N=1 runs the full 10^6 x 2200 table on one GPU).  --scaling weak gives every GPU 1/8 of the table.

A "step" is one pass of the hot path over the whole table: on every rank one fused kernel launch
from device-resident meas/err shards to device-resident percentile tables, plus, for N>1, the
gather of the per-rank tables, which is fused into the kernel (every rank's kernel writes its
result rows into all ranks' full tables over NVLink; --gather nccl uses NCCL all-gathers instead).
`e2e` is the same step through pymcz_b200.dist.production_table -- the call mcz.run makes -- with
HOST tables in and out: pinned upload, kernel(s), gather, pinned download of the full tables on
every rank (on one GPU this is the C ABI's mcz_ctx_run_host).
At N>1 the run also VERIFIES the exchange: the gathered tables must equal an NCCL all-gather of
the ranks' local tables bit for bit, and a single-object run with its samples sharded over the
ranks must reproduce the one-GPU result bit for bit.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "galaxy-samples/sec (all scales, dust-corrected)"
UNIT = "galaxy-samples/s"

CONFIGS = {
    3: dict(G=10_000, nsample=10_000, md="all",
            name="BASELINE config 3: synthetic 10^4 galaxies x nsample=10^4 (N'=11000), all 11 lines, --md all, dust on"),
    4: dict(G=1_000_000, nsample=2_000, md="all",
            name="BASELINE config 4: synthetic 10^6 HII-region table (SURVEY 8d generator, seed 1004) x nsample=2000 "
                 "(N'=2200), --md all, dust on"),
    5: dict(G=100_000, nsample=10_000, md="KD02",
            name="BASELINE config 5: KD02comb-only stress, synthetic 10^5 galaxies x nsample=10^4 (N'=11000), --md KD02, dust on"),
}


def n_draw(nsample):
    return int(nsample + 0.1 * nsample) if nsample > 1 else nsample


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own modules (oracle/_ref) on host cores
# ---------------------------------------------------------------------------------------------
def _ref_kind():
    from oracle import ref_harness
    return "reference" if ref_harness.available() else "port"


def _cpu_worker(args):
    """One measurement, as the reference's Pool worker calc() does (mcz.py:432-454): correlated
    perturbation, metallicity.calculation with the per-sample np.roots loops, savehist statistics.
    Runs the UNMODIFIED reference modules staged in oracle/_ref when present, else the oracle port."""
    meas_row, err_row, nsample, md, seed, kind = args
    rs = np.random.RandomState(seed)
    if kind == "reference":
        from oracle import ref_harness
        _, ret, table = ref_harness.run_measurement(meas_row, err_row, nsample, mds=md, dust_corr=True, rs=rs)
        return len(table)
    from oracle import mcz_oracle as orc
    out = orc.run_table(meas_row[None], err_row[None], nsample, mds=md, dust_corr=True,
                        mode="correlated", roots_mode="loop", rs=rs, normal=rs.normal)
    return float(np.nansum(out["pct"]))


def cpu_reference_run(G, cfg, config, nps, pool=None, seed0=0, kind=None):
    """Time G measurements over a Pool(nps) like the reference's --multiproc path (mcz.py:545-554)."""
    import multiprocessing as mp
    from pymcz_b200.synth import synthetic_table
    kind = kind or _ref_kind()
    meas, err = synthetic_table(max(G, 1), config=config)
    jobs = [(meas[i], err[i], cfg["nsample"], cfg["md"], seed0 + i, kind) for i in range(G)]
    own = pool is None
    if own and nps > 1:
        pool = mp.get_context("fork").Pool(nps)
    t0 = time.perf_counter()
    if nps > 1:
        pool.map(_cpu_worker, jobs, chunksize=1)
    else:
        for j in jobs:
            _cpu_worker(j)
    dt = time.perf_counter() - t0
    if own and nps > 1:
        pool.close()
        pool.join()
    return G * n_draw(cfg["nsample"]) / dt, dt


def reference_pool_size():
    return min((os.cpu_count() or 2) - 1 or 1, 10)       # mcz.py:48, 545


def _ref_describe(kind):
    return ("unmodified reference modules (oracle/_ref/pyMCZ: metallicity.calculation, metscales.diagnostics, "
            "savehist statistics)" if kind == "reference" else "oracle port (oracle/mcz_oracle.py, per-sample np.roots loop)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cfg = CONFIGS[args.config]
    kind = _ref_kind()
    nps = reference_pool_size()
    N = n_draw(cfg["nsample"])
    # a bounded sample of the table per step: ~8 measurements per worker at N'=2200, fewer at 11000
    per_worker = 8 if N <= 4096 else 2
    G_step = per_worker * nps
    pool = mp.get_context("fork").Pool(nps) if nps > 1 else None
    for w in range(args.warmup):
        cpu_reference_run(nps, cfg, args.config, nps, pool, seed0=1000 * w, kind=kind)
    t_tot = 0.0
    for k in range(args.steps):
        _, dt = cpu_reference_run(G_step, cfg, args.config, nps, pool, seed0=7000 + 1000 * k, kind=kind)
        t_tot += dt
    if pool is not None:
        pool.close()
        pool.join()
    value = args.steps * G_step * N / t_tot
    sample = "%d measurements x N'=%d per step, Pool(%d) one measurement per task (mcz.py:545-554), %s" % (
        G_step, N, nps, _ref_describe(kind))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": cfg["name"] + "; bounded sample of %d measurements per step" % G_step,
                   "config": args.config, "md": cfg["md"], "nsample": cfg["nsample"], "n_draw": N,
                   "sampling": "correlated (the reference's --multiproc worker, mcz.py:442)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nps, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi via NVML) during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _bits_equal(a, b):
    """Bitwise equality of two tensors of the same dtype/shape (NaN payloads included)."""
    import torch
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == torch.float32:
        a, b = a.view(torch.int32), b.view(torch.int32)
    return bool(torch.equal(a, b))


def _profile_traffic(kernel_name, G, N, md):
    """dram__bytes_read+write per launch from the committed ncu --set full capture of THIS kernel
    on THIS launch shape (profiles/r2_*.json written by tools/ncu_summary.py); None if there is no
    capture that matches -- the number is never rescaled from another shape."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_*.json"))):
        try:
            prof = json.load(open(path))
        except Exception:
            continue
        if (prof.get("kernel_base") == kernel_name and prof.get("galaxies") == G and prof.get("n_draw") == N
                and prof.get("md") == md and "dram_bytes_read" in prof):
            return prof["dram_bytes_read"] + prof["dram_bytes_write"], {
                "file": os.path.relpath(path, ROOT), "commit": prof.get("commit"), "galaxies": G, "n_draw": N}
    return None, None


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pymcz_b200 import _lib, _scales as sc, ops
    from pymcz_b200 import dist as mdist
    from pymcz_b200.synth import synthetic_table, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    cfg = CONFIGS[args.config]
    G_total = cfg["G"] if args.galaxies is None else args.galaxies
    if args.scaling == "weak":
        G_total = max(G_total // 8, 1) * world
    lo, hi = shard_bounds(G_total, world, rank)
    full_m, full_e = synthetic_table(G_total, config=args.config)
    meas_h, err_h = np.ascontiguousarray(full_m[lo:hi]), np.ascontiguousarray(full_e[lo:hi])
    G = hi - lo
    N = n_draw(cfg["nsample"])
    tok, _, _ = sc.parse_md(cfg["md"])
    meas = torch.from_numpy(meas_h).to(dev)
    err = torch.from_numpy(err_h).to(dev)
    ws = ops.ProductionWorkspace(G, N, tok, dev)
    # The path's only exchange is the gather of the per-rank result tables.  Default:
    # fused into the kernel -- every rank's kernel writes its rows straight into all ranks' full
    # tables over NVLink peer mappings (pymcz_b200.dist.GatheredTables), then one 4-byte
    # all-reduce as the barrier.  --gather nccl: kernel, then three NCCL all-gathers.
    gath = fused = None
    if world > 1:
        if args.gather == "fused":
            try:
                fused = mdist.GatheredTables(G_total, dev)
            except Exception as exc:        # no peer mappings on this box: measure the NCCL form and say so
                sys.stderr.write("bench.py: fused gather unavailable (%s); using NCCL all-gathers\n" % (exc,))
            ok = torch.tensor([1 if fused is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if not int(ok.item()):
                fused, args.gather = None, "nccl"
        if fused is None:
            counts = [shard_bounds(G_total, world, r)[1] - shard_bounds(G_total, world, r)[0] for r in range(world)]
            if len(set(counts)) != 1:
                raise SystemExit("bench.py --gather nccl needs a table that divides evenly over the ranks")
            gath = dict(pct=torch.empty((G_total, sc.NKEYS, 3), dtype=torch.float32, device=dev),
                        nvalid=torch.empty((G_total, sc.NKEYS), dtype=torch.int32, device=dev),
                        status=torch.empty((G_total, sc.NKEYS), dtype=torch.int8, device=dev))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        ops.forward_production(meas, err, N, tok, sc.DUST_ON, seed=0xB200, galaxy_offset=lo, workspace=ws,
                               gather=fused)
        if fused is not None:
            fused.wait()
        elif world > 1:
            dist.all_gather_into_tensor(gath["pct"], ws.pct)
            dist.all_gather_into_tensor(gath["nvalid"], ws.nvalid)
            dist.all_gather_into_tensor(gath["status"], ws.status)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.mcz_launch_count()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)                       # L2 flush between timed iterations (not timed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        evs.append((a, b))
    barrier()
    launches = L.mcz_launch_count() - launches0
    clocks = sampler.stop()
    t_ms = sum(a.elapsed_time(b) for a, b in evs)
    # kernel-only duration (same stream, CUDA events around the launch alone)
    kev = []
    for _ in range(min(args.steps, 5)):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.forward_production(meas, err, N, tok, sc.DUST_ON, seed=0xB200, galaxy_offset=lo, workspace=ws)
        b.record()
        kev.append((a, b))
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    mean_trips = float(ws.trips.float().sum(dim=1).mean().item())
    # trips the kernel actually executed (a provably divergent N2Ha loop is cut short and reported
    # with the reference's outcome of 100 trips): the roofline is credited with EXECUTED work only.
    # The statistics of the last launch live in the header of its scratch buffer.
    import ctypes
    tc = (ctypes.c_ulonglong * 8)()
    _lib.check(L.mcz_production_last_stats(ws.scratch.data_ptr(), tc, torch.cuda.current_stream(dev).cuda_stream),
               "mcz_production_last_stats")
    exec_trips = (tc[1] + tc[2]) / max(tc[0], 1)
    early_exits = int(tc[3]) + int(tc[4])

    # ---- verification of the exchange (N > 1): after timing, not timed
    verified = {}
    if world > 1:
        counts = [shard_bounds(G_total, world, r)[1] - shard_bounds(G_total, world, r)[0] for r in range(world)]
        if fused is not None:
            # local tables of this rank (plain launch) -> NCCL all-gather -> must equal the tables the
            # kernels wrote into each other's memory, bit for bit, on every rank
            ops.forward_production(meas, err, N, tok, sc.DUST_ON, seed=0xB200, galaxy_offset=lo, workspace=ws)
            ref = {k: mdist.all_gather_rows(getattr(ws, k), counts) for k in ("pct", "nvalid", "status")}
            ops.forward_production(meas, err, N, tok, sc.DUST_ON, seed=0xB200, galaxy_offset=lo, workspace=ws, gather=fused)
            fused.wait()
            torch.cuda.synchronize()
            okg = all(_bits_equal(getattr(fused, k)[:G_total], ref[k]) for k in ("pct", "nvalid", "status"))
            flag = torch.tensor([1 if okg else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            verified["gather_verified"] = bool(int(flag.item()))
            del ref
        # single-object run, samples sharded over the ranks: every rank must end with
        # the tables a one-GPU run of the same kernel produces (integer reductions: bit-identical)
        try:
            so_m, so_e = synthetic_table(3, config=4)
            so_n = 400_000
            sh = mdist.production_table_sample_sharded(so_m, so_e, so_n, sc.T_ALL, sc.DUST_ON, 0xB200, device=dev)
            one = ops.forward_production(torch.from_numpy(so_m).to(dev), torch.from_numpy(so_e).to(dev), so_n,
                                         sc.T_ALL, sc.DUST_ON, seed=0xB200)
            torch.cuda.synchronize()
            oks = (np.array_equal(sh["pct"].view(np.int32), one[0].cpu().numpy().view(np.int32)) and
                   np.array_equal(sh["nvalid"], one[1].cpu().numpy()) and
                   np.array_equal(sh["status"], one[2].cpu().numpy()) and
                   np.array_equal(sh["trips"], one[3].cpu().numpy()))
            flag = torch.tensor([1 if oks else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            verified["sample_sharded_verified"] = bool(int(flag.item()))
            verified["sample_sharded_ms"] = sh["kernel_ms"]
        except Exception as exc:
            verified["sample_sharded_verified"] = False
            verified["sample_sharded_error"] = repr(exc)[:200]

    # ---- e2e: the call mcz.run makes (pymcz_b200.dist.production_table), pinned host tables in and out
    hm = torch.from_numpy(full_m).pin_memory()
    he = torch.from_numpy(full_e).pin_memory()
    hm_np, he_np = hm.numpy(), he.numpy()

    def e2e_step():
        return mdist.production_table(hm_np, he_np, N, tok, sc.DUST_ON, 0xB200, copy=False)

    for _ in range(max(min(args.warmup, 3), 1)):
        out = e2e_step()
    barrier()
    e2e_n = max(min(args.steps, 10), 1)
    t0 = time.perf_counter()
    for _ in range(e2e_n):
        out = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_matches = True
    if world == 1:      # the host-path tables are the device-path tables
        e2e_matches = bool(np.array_equal(out["pct"].view(np.int32), ws.pct.cpu().numpy().view(np.int32)) and
                           np.array_equal(out["nvalid"], ws.nvalid.cpu().numpy()))
    h2d = int(meas_h.nbytes + err_h.nbytes)
    d2h = int(sum(out[k].nbytes for k in ("pct", "nvalid", "status", "trips", "ret")))
    mdist.release_host_paths()

    # max over ranks
    tt = torch.tensor([t_ms, e2e_s, k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, e2e_s, k_ms = [float(x) for x in tt.tolist()]
    units_per_step = G_total * N
    value = units_per_step * args.steps / (t_ms * 1e-3)
    e2e_value = units_per_step * e2e_n / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        f_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        # definitional work per galaxy-sample:
        #   md=all:  W_alu = 1120 + 35 T, W_sfu = 62 + T;   md=KD02:  W_alu = 520 + 35 T, W_sfu = 39 + T
        base_alu, base_sfu = (1120.0, 62.0) if cfg["md"] == "all" else (520.0, 39.0)
        w_alu, w_sfu = base_alu + 35.0 * exec_trips, base_sfu + exec_trips
        w_alu_ref = base_alu + 35.0 * mean_trips
        kernel_base = L.mcz_production_kernel_name(G, N, tok).decode()
        traffic, traffic_src = _profile_traffic(kernel_base, G, N, cfg["md"])
        per_gpu_gs = (G * N) / (k_ms * 1e-3)
        peak_alu = sms * 128 * f_mhz * 1e6
        peak_sfu = sms * 16 * f_mhz * 1e6
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bytes_per_galaxy = 2 * 11 * 4 + 37 * (3 * 4 + 4 + 1) + 2 * 4 + 4
        roofline = {
            "bound": "alu_issue",
            "achieved": per_gpu_gs * w_alu / 1e12, "peak": peak_alu / 1e12, "unit": "Tlane-op/s",
            "frac": per_gpu_gs * w_alu / peak_alu,
            "traffic": traffic, "traffic_source": traffic_src,
            "traffic_note": ("ncu dram__bytes_read+write of one launch of this kernel on this launch shape"
                             if traffic is not None else "no ncu --set full capture of this launch shape is committed"),
            "kernel": kernel_base, "kernel_ms": k_ms,
            "mean_trips_reference": mean_trips, "mean_trips_executed": exec_trips, "early_exits_per_launch": early_exits,
            "frac_if_credited_with_reference_trips": per_gpu_gs * w_alu_ref / peak_alu,
            "w_alu_per_gs": w_alu, "w_sfu_per_gs": w_sfu, "sfu_frac": per_gpu_gs * w_sfu / peak_sfu,
            "sm_mhz_used": f_mhz, "sms": sms,
            "hbm": {"algorithmic_bytes_per_launch": bytes_per_galaxy * G,
                    "achieved_gbs": bytes_per_galaxy * G / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "frac": bytes_per_galaxy * G / (k_ms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "measured" if "hbm_gbs" in peaks else "fallback"},
            "note": "SURVEY 8(d): the fused path moves ~0.33 B per galaxy-sample, so the binding roof is "
                    "SM issue (128 lanes/SM/clk at the SM clock measured during the run), not HBM",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            nps = reference_pool_size()
            kind = _ref_kind()
            # size the sample for ~15 s from a one-measurement probe
            cpu_reference_run(1, cfg, args.config, 1, kind=kind)            # (imports, first-call costs)
            _, dt1 = cpu_reference_run(1, cfg, args.config, 1, kind=kind)
            Gc = int(max(nps, min(64 * nps, round(15.0 / max(dt1, 1e-3) * nps))))
            v, dt = cpu_reference_run(Gc, cfg, args.config, nps, kind=kind)
            cpu = {"value": v, "unit": UNIT, "cores": nps, "kind": kind,
                   "sample": "%d measurements x N'=%d of the same table, Pool(%d) one measurement per task, correlated "
                             "sampling, %s (%.1f s)" % (Gc, N, nps, _ref_describe(kind), dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"] + ("; the whole table at every N (strong scaling)" if args.scaling == "strong"
                                                   else "; weak scaling, 1/8 of the table per GPU"),
                       "config": args.config, "galaxies_total": G_total, "galaxies_per_gpu": G,
                       "nsample": cfg["nsample"], "n_draw": N, "md": cfg["md"], "dust": True,
                       "sampling": "independent per line (Philox4x32-10, seed 0xB200)",
                       "parallelism": ("galaxy-sharded x%d; result tables gathered by %s" % (
                           world, "the kernel itself (rows written into every rank's table over NVLink peer "
                           "mappings, then a 4-byte all-reduce)" if args.gather == "fused" else "three NCCL all-gathers")
                           if world > 1 else "single GPU"),
                       "l2": "256 MiB buffer written between timed steps (not timed); per-step CUDA events"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_n, "tables_match_device_path": e2e_matches,
                    "api": "pymcz_b200.dist.production_table (the call mcz.run makes): pinned host tables in, full host "
                           "tables out on every rank" + ("; mcz_ctx_run_host underneath" if world == 1 else "; gather included")},
            "roofline": roofline,
        }
        line.update(verified)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration: 4 (default, the metric's), 3 or 5")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the whole table at every N; weak: 1/8 of it per GPU")
    ap.add_argument("--galaxies", type=int, default=None, help="override the configuration's table size (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: result tables gathered by the kernel's own peer writes (default) or by NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
