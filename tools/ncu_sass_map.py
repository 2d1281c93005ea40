"""Address-ordered view of an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass): stall samples and
executed instructions per stretch of SASS, with the source lines that dominate each stretch. Separates the several
inlined copies of a template (which share source lines) by where they sit in the binary."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
BIN = int(sys.argv[2]) if len(sys.argv) > 2 else 200
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
ix = {}
for i, n in enumerate(hdr):
    ix.setdefault(n, i)


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


inst = []
cur_line, cur_src = None, ""
for r in rows[h + 1:]:
    if r and r[0].isdigit():
        cur_line, cur_src = int(r[0]), r[1].strip()
    elif r and len(r) > 3 and r[2].startswith("0x"):
        inst.append((int(r[2], 16), cur_line, cur_src, r[3].strip(), num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]])))
inst.sort()
tot_s = sum(i[4] for i in inst) or 1
tot_i = sum(i[5] for i in inst) or 1
print("instructions", len(inst), "samples", tot_s, "executed", tot_i)
for k in range(0, len(inst), BIN):
    seg = inst[k:k + BIN]
    s = sum(i[4] for i in seg)
    e = sum(i[5] for i in seg)
    by = Counter()
    for i in seg:
        by[i[1]] += i[4]
    top = ", ".join("%d:%d" % (ln, c) for ln, c in by.most_common(4) if c)
    lines = sorted(set(i[1] for i in seg))
    print("%6d-%6d  samples %5.1f%%  exec %5.1f%%  lines %d..%d  top %s" % (k, k + len(seg), 100.0 * s / tot_s, 100.0 * e / tot_i, lines[0], lines[-1], top))
