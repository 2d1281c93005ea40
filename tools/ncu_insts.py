"""Per-source-line instruction counts of an ncu report: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_insts.py [N]
(the kernel is issue-bound when instructions / (SMs * 4 schedulers) approaches the cycle count)"""
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
tot = sum(num(r[ix["Instructions Executed"]]) for r in lines)
thr = sum(num(r[ix["Thread Instructions Executed"]]) for r in lines)
print("warp instructions %d, thread instructions %d, avg active threads per instruction %.1f" % (tot, thr, thr / max(tot, 1)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
acc = 0
for r in sorted(lines, key=lambda r: -num(r[ix["Instructions Executed"]]))[:N]:
    n = num(r[ix["Instructions Executed"]])
    acc += n
    print(r[0].rjust(5), str(n).rjust(10), "%5.1f%%" % (100.0 * n / tot), "cum %5.1f%%" % (100.0 * acc / tot), "thr/inst %4.1f" % (num(r[ix["Thread Instructions Executed"]]) / max(n, 1)), r[1].strip()[:120])
