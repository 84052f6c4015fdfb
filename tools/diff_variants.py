"""Compare the per-sample output and tables of two library variants on the same launch:
   python tools/diff_variants.py libA.so libB.so [G] [N] [md]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(out, g, n, md):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from pymcz_b200 import _scales as sc, ops
    from pymcz_b200.synth import synthetic_table
    meas, err = synthetic_table(g, config=4)
    tok = sc.parse_md(md)[0]
    r = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), n, tok, return_samples=True)
    torch.cuda.synchronize()
    np.savez(out, pct=r[0].cpu().numpy(), nvalid=r[1].cpu().numpy(), status=r[2].cpu().numpy(), trips=r[3].cpu().numpy(),
             Z=r[5].cpu().numpy())


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
        sys.exit(0)
    import numpy as np
    sys.path.insert(0, ROOT)
    from pymcz_b200 import _scales as sc
    a, b = sys.argv[1], sys.argv[2]
    g = sys.argv[3] if len(sys.argv) > 3 else "300"
    n = sys.argv[4] if len(sys.argv) > 4 else "2200"
    md = sys.argv[5] if len(sys.argv) > 5 else "all"
    outs = []
    for i, lib in enumerate((a, b)):
        env = dict(os.environ)
        if lib != "default":
            env["MCZ_B200_LIB"] = os.path.abspath(lib)
        out = "/tmp/diffvar_%d.npz" % i
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", out, g, n, md], env=env)
        outs.append(np.load(out))
    A, B = outs
    print("trips equal:", np.array_equal(A["trips"], B["trips"]), "status equal:", np.array_equal(A["status"], B["status"]),
          "nvalid equal:", np.array_equal(A["nvalid"], B["nvalid"]))
    za, zb = A["Z"], B["Z"]
    for j, k in enumerate(sc.KEYS):
        x, y = za[:, j], zb[:, j]
        neq = (x.view(np.int32) != y.view(np.int32)) & ~(np.isnan(x) & np.isnan(y))
        if neq.any():
            both = np.isfinite(x) & np.isfinite(y)
            print("%-10s differing samples %d of %d, max |d| %.3g, mask differences %d" % (
                k, int(neq.sum()), x.size, float(np.max(np.abs(x[both] - y[both]))), int((np.isnan(x) != np.isnan(y)).sum())))
    pa, pb = A["pct"], B["pct"]
    neq = (pa.view(np.int32) != pb.view(np.int32)) & ~(np.isnan(pa) & np.isnan(pb))
    print("percentile entries differing:", int(neq.sum()), "max", float(np.nanmax(np.abs(pa - pb))) if neq.any() else 0.0)
