"""Reduce an ncu launch list (csv with dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per kernel_construct launch)
to the per-launch DRAM traffic bench.py reports as roofline.traffic. usage: ncu_traffic.py <csv> <batches> [out.json]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, nb = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02", "ncu_construct_traffic_%dM.json" % nb)
rows = [r for r in csv.reader(open(path)) if r]
h = next(i for i, r in enumerate(rows) if "Metric Name" in r)
hdr = rows[h]
iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
launches = {}
for r in rows[h + 1:]:
    if len(r) <= iv or "kernel_construct" not in r[ik]:
        continue
    launches.setdefault(r[iid], {})[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
L = [v for v in launches.values() if "dram__bytes_read.sum" in v]
rd = sum(v["dram__bytes_read.sum"] for v in L)
wr = sum(v["dram__bytes_write.sum"] for v in L)
us = sum(v.get("gpu__time_duration.sum", 0.0) for v in L)
res = {"workload": "terrain_synth_%dM, one pass from reset" % nb, "launches": len(L), "dram_bytes_read": rd, "dram_bytes_written": wr,
       "dram_bytes_per_launch": round((rd + wr) / max(len(L), 1)), "dram_bytes_per_point": round((rd + wr) / (nb * 1e6), 2),
       "kernel_us_under_ncu": round(us, 1),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kernel_construct python tools/traffic_run.py %d (%s)" % (nb, os.path.basename(path))}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
