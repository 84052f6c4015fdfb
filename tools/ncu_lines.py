"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump per source file / line:
instructions executed and stall samples.  Usage: python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys
from collections import defaultdict


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    cur_file = None
    per_line = defaultdict(lambda: [0, 0, ""])
    per_file = defaultdict(lambda: [0, 0])
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            i_inst = hdr.index('Instructions Executed')
            i_samp = hdr.index('# Samples')
            continue
        if r[0] == 'Function Name' or hdr is None:
            continue
        if r[0] != '':          # a source line row (aggregated over its SASS)
            try:
                inst = int(r[i_inst]); samp = int(r[i_samp])
            except ValueError:
                continue
            key = (cur_file, int(r[0]))
            per_line[key][0] += inst
            per_line[key][1] += samp
            per_line[key][2] = r[1].strip()[:90]
            per_file[cur_file][0] += inst
            per_file[cur_file][1] += samp
    tot_i = sum(v[0] for v in per_file.values())
    tot_s = sum(v[1] for v in per_file.values())
    print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
    for f, (i, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print("%-28s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * i / tot_i, 100.0 * s / tot_s))
    print()
    for (f, ln), (i, s, src) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-22s:%4d inst %5.2f%% samp %5.2f%%  %s" % (f, ln, 100.0 * i / tot_i, 100.0 * s / tot_s, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
