"""Which source lines the local-memory (spill / stack) instructions of a kernel belong to:
   python tools/sass_local_lines.py <object.o> <mangled kernel name substring> [top]
(extracts the cubin, runs nvdisasm -g, counts STL / LDL per `//## File ..., line N` annotation)."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main(obj, name, top=30):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    cur, inside = None, False
    cnt = collections.Counter()
    for l in txt.splitlines():
        if l.startswith("//--------------------- .text."):
            inside = name in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.search(r'\b(STL|LDL)\b', l) or re.search(r'\b(STL|LDL)\.', l):
            cnt[cur] += 1
    print("local-memory instructions:", sum(cnt.values()))
    for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
        print(k, v)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
