"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel:
   python tools/ncu_launches.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<tag>_launches_summary.txt"""
import csv
import sys
from collections import defaultdict


def main(path, cmd):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    durs = defaultdict(list)
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
        tot[r[ik][:70]][0] += 1
        tot[r[ik][:70]][1] += v
        durs[r[ik][:70]].append(v)
    allns = sum(v[1] for v in tot.values())
    print("ncu --metrics gpu__time_duration.sum --clock-control none: %s" % cmd)
    print("kernel, launches, total ns, share, min / median / max ms per launch")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        d = sorted(durs[k])
        print("%s, %d, %.0f, %.4f, %.3f / %.3f / %.3f" % (k, n, ns, ns / allns, d[0] / 1e6, d[len(d) // 2] / 1e6, d[-1] / 1e6))
    prod = sorted(v for k, d in durs.items() if "production_kernel" in k for v in d)
    if prod:
        # the launches over the whole table (the timed `value` steps and the verification run) are the
        # long ones; a step of the e2e leg (mcz_ctx_run_host) runs the table as 8 chunk launches
        whole = [v for v in prod if v > 0.5 * prod[-1]]
        chunks = [v for v in prod if v <= 0.5 * prod[-1]]
        print("production_kernel launches over the whole table: %d, mean %.3f ms (cold caches, serialised under ncu)"
              % (len(whole), sum(whole) / len(whole) / 1e6))
        if chunks:
            print("chunk launches of the e2e leg (8 per step): %d, mean %.3f ms, %.3f ms per 8"
                  % (len(chunks), sum(chunks) / len(chunks) / 1e6, 8 * sum(chunks) / len(chunks) / 1e6))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
