"""Workload for the ncu DRAM-traffic launch list of bench.py's N = 1 workload: one pass of the 350 M-point stream from reset.
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kernel_construct \
      --csv --log-file gpurun_out/traffic_350M.csv python tools/traffic_run.py 350
then  python tools/ncu_traffic.py gpurun_out/traffic_350M.csv 350   writes profiles/r02/ncu_construct_traffic_350M.json"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simlod_b200 import SimLOD, data  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 350
n = NB * 1_000_000
sim = SimLOD(1920, 1080, persistent_bytes=max(4 << 30, NB * (40 << 20)))
sim.set_box((0, 0, 0), data.TERRAIN_EXTENT)
dptr = sim.device_alloc(n * 16)
sim.generate(sim.GEN_TERRAIN, dptr, n, 0, n, 7)
sim.reset()
kms, tms = sim.insert_device(dptr, n)
st = sim.stats()
print("inserted", st.numPoints, "launches", sim.launch_info()["launches"], "kernel ms (under ncu)", kms, flush=True)
sim.close()
