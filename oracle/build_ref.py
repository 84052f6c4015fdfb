"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, staged so that they
travel to the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY, not product).

    python oracle/build_ref.py          # needs /root/reference (the build container)

The reference is pure Python: "building" it is copying pyMCZ/metallicity.py, pyMCZ/metscales.py and
pyMCZ/mcz.py byte for byte from /root/reference into oracle/_ref/pyMCZ/ (git-ignored, NOT
gpurun-ignored, exactly like a compiled reference .so would be).  Nothing is edited; the py2-isms
on the path are bridged from the outside by oracle/ref_harness.py (a dict with iterkeys(); mcz.py
is never imported -- its py3-valid line ranges are exec'd, as oracle/gen_golden.py does).
A manifest with the sha256 of every staged file is written next to them so that a run can state
exactly what it timed.
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference/pyMCZ"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "pyMCZ")
FILES = ["metallicity.py", "metscales.py", "mcz.py"]


def build(quiet=False):
    if not os.path.isdir(REF):
        if not quiet:
            print("oracle/build_ref.py: %s not present (GPU box): using the staged copy, if any" % REF)
        return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
        manifest[f] = hashlib.sha256(open(os.path.join(OUT, f), "rb").read()).hexdigest()
    json.dump({"source": REF, "sha256": manifest}, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    if not quiet:
        print("oracle/_ref/pyMCZ: staged %s" % ", ".join(FILES))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
