"""Kernel-only timing of library variants (tools/build_variant.sh) on the bench workload shape:
   python tools/time_variants.py [--g 59200] [--n 2200] [--md all] lib1.so lib2.so ...
Each library runs in its own process (MCZ_B200_LIB); prints ms per launch, galaxy-samples/s and a
checksum of the result tables (variants that must not change results have to agree on it)."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(g, n, md):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from pymcz_b200 import _scales as sc, ops
    from pymcz_b200.synth import synthetic_table
    cfg = 4 if n <= 2304 else (5 if md == "KD02" else 3)
    meas, err = synthetic_table(g, config=cfg)
    tok = sc.parse_md(md)[0]
    m, e = torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda()
    ws = ops.ProductionWorkspace(g, n, tok, m.device)
    for _ in range(2):
        ops.forward_production(m, e, n, tok, workspace=ws)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.forward_production(m, e, n, tok, workspace=ws)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    h = hashlib.sha1()
    for t in (ws.pct, ws.nvalid, ws.status, ws.trips):
        h.update(t.cpu().numpy().tobytes())
    print(json.dumps({"ms": float(np.median(ts)), "min_ms": float(min(ts)), "gs_per_s": g * n / (np.median(ts) * 1e-3),
                      "sum": h.hexdigest()[:12]}))


if __name__ == "__main__":
    args = sys.argv[1:]
    g, n, md = 59200, 2200, "all"
    if args and args[0] == "--child":
        child(int(args[1]), int(args[2]), args[3])
        sys.exit(0)
    libs = []
    i = 0
    while i < len(args):
        if args[i] == "--g":
            g = int(args[i + 1]); i += 2
        elif args[i] == "--n":
            n = int(args[i + 1]); i += 2
        elif args[i] == "--md":
            md = args[i + 1]; i += 2
        else:
            libs.append(args[i]); i += 1
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["MCZ_B200_LIB"] = os.path.abspath(lib)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(g), str(n), md], env=env,
                             capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ("ERR " + out.stderr[-400:])
        print("%-40s %s" % (os.path.basename(lib), line), flush=True)
