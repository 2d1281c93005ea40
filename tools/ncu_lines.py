"""Per-source-line stall summary of an ncu report: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
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


lines = [r for r in rows[h + 1:] if r and r[0].isdigit()]
tot = sum(num(r[ix["# Samples"]]) for r in lines)
names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(num(r[ix[n]]) for r in lines) for n in names}
print("total samples", tot, {k: "%.1f%%" % (100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]})
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for r in sorted(lines, key=lambda r: -num(r[ix["# Samples"]]))[:N]:
    st = {n: num(r[ix[n]]) for n in names}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(r[0].rjust(4), r[ix["# Samples"]].rjust(6), "%5.1f%%" % (100.0 * num(r[ix["# Samples"]]) / tot),
          " ".join("%s=%d" % (k[6:], v) for k, v in top).ljust(30), r[1].strip()[:110])
