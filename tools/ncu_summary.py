"""Summarise an `ncu --set full` report of the production kernel into profiles/:
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep <galaxies in the profiled launch> <tag>
writes profiles/<tag>_production_kernel.json (read by bench.py for roofline.traffic),
profiles/<tag>_production_kernel.txt (key counters) and profiles/<tag>_source_hotspots.txt."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    # per-pipe instruction issue, % of the pipe's peak while the SM is active (the XU pipe is the SFU: MUFU)
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
]


def main(rep, galaxies, tag, n_draw=2200, md="all", kernel_base=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[2]
    vals = {}
    for i, h in enumerate(hdr):
        if h in WANT or h == "Kernel Name":
            vals[h] = (row[i], units[i])

    def num(k):
        v, u = vals[k]
        f = float(v.replace(",", ""))
        mult = {"Kbyte": 1e3, "Kbyte/block": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        return f * mult
    out = {
        "kernel": vals["Kernel Name"][0], "galaxies": int(galaxies), "n_draw": n_draw,
        "duration_s_under_ncu": num("gpu__time_duration.sum"),
        "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
        "registers_per_thread": int(float(vals["launch__registers_per_thread"][0])),
        "dynamic_smem_bytes": num("launch__shared_mem_per_block_dynamic"),
        "issue_active_pct": float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "warp_instructions": float(vals["smsp__inst_executed.sum"][0]),
        "thread_instructions_per_galaxy_sample": (float(vals["smsp__inst_executed.sum"][0]) *
                                                  float(vals["smsp__thread_inst_executed_per_inst_executed.ratio"][0]) /
                                                  (int(galaxies) * n_draw)),
        "issue_slots_per_galaxy_sample": float(vals["smsp__inst_executed.sum"][0]) * 32 / (int(galaxies) * n_draw),
        "source": os.path.basename(rep),
        # what bench.py matches a capture against: the kernel launch_production() picks, the launch shape,
        # the tokens, and the commit of the profiled build
        "kernel_base": kernel_base, "md": md,
        "commit": os.environ.get("MCZ_COMMIT") or subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"],
                                                                  capture_output=True, text=True).stdout.strip(),
        "pipe_pct_of_peak_active": {k.split("pipe_")[1].split(".")[0]: float(vals[k][0].replace(",", ""))
                                    for k in vals if k.startswith("sm__inst_executed_pipe_") and k.endswith("sustained_active")},
    }
    outdir = os.environ.get("MCZ_PROFILE_DIR") or os.path.join(ROOT, "profiles")      # (on the GPU box: under gpurun_out/)
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "%s_production_kernel.json" % tag), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(outdir, "%s_production_kernel.txt" % tag), "w") as f:
        f.write("ncu --set full --clock-control none, one launch of %s over %s galaxies x %d samples\n" % (out["kernel"], galaxies, n_draw))
        for k in WANT:
            if k in vals:
                f.write("%-75s %s %s\n" % (k, vals[k][0], vals[k][1]))
        f.write("\nderived: %.0f issue slots (warp-instructions x 32) per galaxy-sample; %.0f thread instructions per galaxy-sample\n"
                % (out["issue_slots_per_galaxy_sample"], out["thread_instructions_per_galaxy_sample"]))
        f.write("derived: DRAM traffic %.1f B per galaxy (algorithmic: 88 B in + 641 B out; the result tables stay in L2)\n"
                % ((out["dram_bytes_read"] + out["dram_bytes_write"]) / int(galaxies)))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", "ncu_src_%s.csv" % tag)
    open(tmp, "w").write(src)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp, "40"], capture_output=True, text=True).stdout
    with open(os.path.join(outdir, "%s_source_hotspots.txt" % tag), "w") as f:
        f.write("per source file / line: share of warp-instructions executed and of warp-stall samples (ncu source page)\n")
        f.write(hot)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 2200,
         sys.argv[5] if len(sys.argv) > 5 else "all", sys.argv[6] if len(sys.argv) > 6 else None)
