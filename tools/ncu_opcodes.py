"""Per (source file, line) and per opcode: warp-instructions executed, from an ncu
`--page source --csv --print-source cuda,sass` dump.  Usage:
   python tools/ncu_opcodes.py dump.csv <units (galaxy-samples in the launch)> [opcode-prefix ...]
Without opcode prefixes: the opcode histogram (issue slots per unit).  With prefixes: the source
lines that execute the most instructions of those opcodes."""
import collections
import csv
import re
import sys


def main(path, units, want):
    rows = list(csv.reader(open(path)))
    cur_file, cur_line, cur_src = None, None, ""
    i_inst = None
    by_op = collections.Counter()
    by_line = collections.Counter()
    src_of = {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            i_inst = r.index('Instructions Executed')
            continue
        if r[0] == 'Function Name' or i_inst is None:
            continue
        if r[0] != '':
            cur_line, cur_src = int(r[0]), r[1].strip()[:80]
            continue
        if len(r) <= i_inst or r[2] == '...':
            continue
        try:
            n = int(r[i_inst])
        except ValueError:
            continue
        s = re.sub(r'^@!?U?P\d+\s+', '', r[3].strip())
        op = s.split()[0] if s else '?'
        by_op[op] += n
        if any(op.startswith(w) for w in want):
            by_line[(cur_file, cur_line)] += n
            src_of[(cur_file, cur_line)] = cur_src
    tot = sum(by_op.values())
    if not want:
        short = collections.Counter()
        for op, n in by_op.items():
            short['.'.join(op.split('.')[:2])] += n
        for op, n in short.most_common(60):
            print("%-16s %6.2f%%  %6.1f slots/unit" % (op, 100.0 * n / tot, n * 32.0 / units))
    else:
        sel = sum(by_line.values())
        print("selected opcodes: %.2f%% of all instructions, %.1f slots/unit" % (100.0 * sel / tot, sel * 32.0 / units))
        for (f, ln), n in by_line.most_common(40):
            print("%-20s:%4d %6.1f slots/unit  %s" % (f, ln, n * 32.0 / units, src_of[(f, ln)]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), sys.argv[3:])
