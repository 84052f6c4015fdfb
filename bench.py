#!/usr/bin/env python
"""Benchmark of the pyMCZ Monte-Carlo metallicity hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Metric (BASELINE.json): galaxy-samples/s, all scales (--md all), dust-corrected.
Workload: BASELINE config 4 -- the synthetic 10^6 HII-region table x nsample=2000 (N'=2200) of
SURVEY.md section 8(d), sharded over galaxies.  Scaling is WEAK: every GPU gets 125 000 galaxies
(at 8 GPUs that is the full 10^6 table), so N=1 runs the first 1/8 of the table.

A "step" is one pass of the hot path over the rank's shard: one fused kernel launch from
device-resident meas/err tables to device-resident percentile tables; for N>1 the gather of the
per-rank tables is part of the step and is fused into the kernel (every rank's kernel writes its
result rows into all ranks' full tables over NVLink; --gather nccl uses NCCL all-gathers instead).
`e2e` is the same step through the C-ABI host-buffer entry point (mcz_ctx_run_host): pinned host
tables in, H2D, kernel, D2H, host tables out.
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

GALAXIES_PER_GPU = 125_000
NSAMPLE = 2000
METRIC = "galaxy-samples/sec (all scales, dust-corrected)"
UNIT = "galaxy-samples/s"


def n_draw(nsample):
    return int(nsample + 0.1 * nsample) if nsample > 1 else nsample


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference algorithm on host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One galaxy, as the reference's Pool worker calc() does (mcz.py:432-454): correlated
    perturbation, metallicity.calculation with the per-sample np.roots loop, percentiles."""
    meas_row, err_row, nsample, seed = args
    from oracle import mcz_oracle as orc
    rs = np.random.RandomState(seed)
    out = orc.run_table(meas_row[None], err_row[None], nsample, mds="all", dust_corr=True,
                        mode="correlated", roots_mode="loop", rs=rs, normal=rs.normal)
    return float(np.nansum(out["pct"]))


def cpu_reference_run(G, nsample, nps, pool=None, seed0=0):
    """Time G galaxies over a Pool(nps) like the reference's --multiproc path (mcz.py:545-554)."""
    import multiprocessing as mp
    from pymcz_b200.synth import synthetic_table
    meas, err = synthetic_table(max(G, 1), config=4)
    jobs = [(meas[i], err[i], nsample, seed0 + i) for i in range(G)]
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
    return G * n_draw(nsample) / dt, dt


def reference_pool_size():
    return min((os.cpu_count() or 2) - 1 or 1, 10)       # mcz.py:48, 545


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    nps = reference_pool_size()
    G_step = 8 * nps
    pool = mp.get_context("fork").Pool(nps) if nps > 1 else None
    for w in range(args.warmup):
        cpu_reference_run(G_step, NSAMPLE, nps, pool, seed0=1000 * w)
    t_tot = 0.0
    for k in range(args.steps):
        _, dt = cpu_reference_run(G_step, NSAMPLE, nps, pool, seed0=7000 + 1000 * k)
        t_tot += dt
    if pool is not None:
        pool.close()
        pool.join()
    value = args.steps * G_step * n_draw(NSAMPLE) / t_tot
    sample = "%d galaxies x N'=%d per step, Pool(%d), per-sample np.roots loop" % (G_step, n_draw(NSAMPLE), nps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": "BASELINE config 4 shape (synthetic HII-region table, nsample=2000, N'=2200, "
                               "--md all, dust on); bounded sample of %d galaxies per step" % G_step,
                   "md": "all", "nsample": NSAMPLE, "n_draw": n_draw(NSAMPLE), "sampling": "correlated (mcz.py:442)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nps, "kind": "port", "sample": sample},
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


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pymcz_b200 import _lib, _scales as sc, ops
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

    G_total = args.galaxies_per_gpu * world
    lo, hi = shard_bounds(G_total, world, rank)
    meas_h, err_h = synthetic_table(G_total, config=4, start=lo, stop=hi)
    G = hi - lo
    N = n_draw(args.nsample)
    tok = sc.T_ALL
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
                from pymcz_b200.dist import GatheredTables
                fused = GatheredTables(G_total, dev)
            except Exception as exc:        # no peer mappings on this box: measure the NCCL form and say so
                sys.stderr.write("bench.py: fused gather unavailable (%s); using NCCL all-gathers\n" % (exc,))
                args.gather = "nccl"
            ok = torch.tensor([1 if fused is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if not int(ok.item()):
                fused, args.gather = None, "nccl"
        if fused is None:
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
    # with the reference's outcome of 100 trips): the roofline is credited with EXECUTED work only
    import ctypes
    tc = (ctypes.c_ulonglong * 4)()
    L.mcz_debug_trip_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.mcz_debug_trip_counters(tc, 1)
    ops.forward_production(meas, err, N, tok, sc.DUST_ON, seed=0xB200, galaxy_offset=lo, workspace=ws)
    torch.cuda.synchronize()
    L.mcz_debug_trip_counters(tc, 1)
    exec_trips = (tc[1] + tc[2]) / max(tc[0], 1)
    early_exits = int(tc[3])

    # ---- e2e through the C-ABI host-buffer call (pinned host tables)
    ctx = L.mcz_ctx_create(local_rank)
    if not ctx:
        raise SystemExit("mcz_ctx_create failed: " + L.mcz_last_error().decode())
    hm = torch.from_numpy(meas_h).pin_memory()
    he = torch.from_numpy(err_h).pin_memory()
    hp = torch.empty((G, sc.NKEYS, 3), dtype=torch.float32).pin_memory()
    hn = torch.empty((G, sc.NKEYS), dtype=torch.int32).pin_memory()
    hs = torch.empty((G, sc.NKEYS), dtype=torch.int8).pin_memory()
    ht = torch.empty((G, 2), dtype=torch.int32).pin_memory()
    hr = torch.empty((G,), dtype=torch.int32).pin_memory()

    def e2e_step():
        rc = L.mcz_ctx_run_host(ctx, hm.data_ptr(), he.data_ptr(), G, N, tok, sc.DUST_ON, 0xB200, lo, 0, 0,
                                hp.data_ptr(), hn.data_ptr(), hs.data_ptr(), ht.data_ptr(), hr.data_ptr())
        if rc:
            raise SystemExit("mcz_ctx_run_host: " + L.mcz_last_error().decode())

    for _ in range(max(min(args.warmup, 3), 1)):
        e2e_step()
    barrier()
    e2e_n = max(min(args.steps, 10), 1)
    t0 = time.perf_counter()
    for _ in range(e2e_n):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    L.mcz_ctx_destroy(ctx)
    h2d = int(hm.numel() * 4 + he.numel() * 4)
    d2h = int(hp.numel() * 4 + hn.numel() * 4 + hs.numel() + ht.numel() * 4 + hr.numel() * 4)

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
        # definitional work per galaxy-sample, md=all: W_alu = 1120 + 35 T, W_sfu = 62 + T
        w_alu, w_sfu = 1120.0 + 35.0 * exec_trips, 62.0 + exec_trips
        w_alu_ref = 1120.0 + 35.0 * mean_trips
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1_production_kernel.json")))
        except Exception:
            pass
        traffic = None
        if prof.get("galaxies"):
            traffic = (prof["dram_bytes_read"] + prof["dram_bytes_write"]) / prof["galaxies"] * G
        per_gpu_gs = (G * N) / (k_ms * 1e-3)
        peak_alu = sms * 128 * f_mhz * 1e6
        peak_sfu = sms * 16 * f_mhz * 1e6
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bytes_per_galaxy = 2 * 11 * 4 + 37 * (3 * 4 + 4 + 1) + 2 * 4 + 4
        roofline = {
            "bound": "alu_issue",
            "achieved": per_gpu_gs * w_alu / 1e12, "peak": peak_alu / 1e12, "unit": "Tlane-op/s",
            "frac": per_gpu_gs * w_alu / peak_alu,
            "traffic": traffic,
            "traffic_note": "ncu dram__bytes_read+write of one launch over %s galaxies (profiles/), scaled per galaxy to "
                            "this launch; the ~640 B/galaxy result tables stay in the 126 MB L2" % prof.get("galaxies"),
            "kernel": "production_kernel<4>", "kernel_ms": k_ms,
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
            # size the sample for ~15 s from a one-galaxy probe
            _, dt1 = cpu_reference_run(1, NSAMPLE, 1)
            Gc = int(max(nps, min(64 * nps, round(15.0 / max(dt1, 1e-3) * nps))))
            v, dt = cpu_reference_run(Gc, NSAMPLE, nps)
            cpu = {"value": v, "unit": UNIT, "cores": nps, "kind": "port",
                   "sample": "%d galaxies x N'=%d of the same table, Pool(%d) one galaxy per task, correlated "
                             "sampling, per-sample np.roots loop as the reference (%.1f s)" % (Gc, n_draw(NSAMPLE), nps, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 4: synthetic 10^6 HII-region table (SURVEY 8d generator, seed 1004) "
                                   "x nsample=2000 (N'=2200), --md all, dust on; weak scaling, %d galaxies per GPU "
                                   "(N=8 is the full 10^6 table)" % args.galaxies_per_gpu,
                       "galaxies_total": G_total, "galaxies_per_gpu": args.galaxies_per_gpu,
                       "nsample": args.nsample, "n_draw": N, "md": "all", "dust": True,
                       "sampling": "independent per line (Philox4x32-10, seed 0xB200)",
                       "parallelism": ("galaxy-sharded x%d; result tables gathered by %s" % (
                           world, "the kernel itself (rows written into every rank's table over NVLink peer "
                           "mappings, then a 4-byte all-reduce)" if args.gather == "fused" else "three NCCL all-gathers")
                           if world > 1 else "single GPU"),
                       "l2": "256 MiB buffer written between timed steps (not timed); per-step CUDA events"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_n, "api": "mcz_ctx_run_host (C ABI, pinned host tables)"},
            "roofline": roofline,
        }
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
    ap.add_argument("--galaxies-per-gpu", type=int, default=GALAXIES_PER_GPU)
    ap.add_argument("--nsample", type=int, default=NSAMPLE)
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
