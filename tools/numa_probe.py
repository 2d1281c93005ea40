"""Developer tool: host->device copy bandwidth from page-locked memory allocated on each NUMA node."""
import glob, os, re, time
import torch
torch.cuda.init()
dev = torch.device("cuda", 0)
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
print("gpu0 pci bus", bus)
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.join(os.path.dirname(p), "class")).read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print(os.path.dirname(p).split("/")[-1], "numa_node", open(p).read().strip())
    except OSError:
        pass
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
dst = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
all_cpus = os.sched_getaffinity(0)
for nd in nodes:
    txt = open(os.path.join(nd, "cpulist")).read().strip()
    cpus = set()
    for part in txt.split(","):
        a, _, b = part.partition("-")
        cpus |= set(range(int(a), int(b or a) + 1))
    cpus &= all_cpus
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    src = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    print(os.path.basename(nd), "cpus", txt, "H2D %.1f GB/s" % (8 * (1 << 28) / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
    del src
os.sched_setaffinity(0, all_cpus)
