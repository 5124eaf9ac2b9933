"""Host->device copy bandwidth of the bench's 8 MB latent batches, with the process (and therefore
its pinned allocations) bound to each NUMA node in turn.  Diagnostic for bench.py's e2e leg."""
import glob
import os
import subprocess
import sys
import time

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def measure(tag):
    dev = torch.device("cuda", 0)
    h = [torch.randn(64, 32, 32, 32).pin_memory() for _ in range(4)]
    d = [torch.empty(64, 32, 32, 32, device=dev) for _ in range(2)]
    o = torch.empty(65536, dtype=torch.int64, device=dev)
    oh = torch.empty(65536, dtype=torch.int64).pin_memory()
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    for n in (1, 200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            for k in range(n):
                d[k & 1].copy_(h[k & 3], non_blocking=True)
            e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    with torch.cuda.stream(s):
        e0.record()
        for k in range(200):
            oh.copy_(o, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 200
    print(f"{tag}: H2D 8 MB {ms*1e3:.1f} us = {8.388608/ms:.1f} GB/s; D2H 0.5 MB {ms2*1e3:.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        node = sys.argv[1]
        if node != "none":
            cpus = cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
            os.sched_setaffinity(0, cpus)
        measure(f"bound to node {node}")
        sys.exit(0)
    print(subprocess.run("nvidia-smi topo -m; nproc; cat /sys/devices/system/node/node*/cpulist; "
                         "nvidia-smi --query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv",
                         shell=True, capture_output=True, text=True).stdout)
    bdf = subprocess.run("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0", shell=True,
                         capture_output=True, text=True).stdout.strip().lower()
    bdf = bdf[4:] if len(bdf) > 12 else bdf
    for p in glob.glob(f"/sys/bus/pci/devices/*{bdf[-10:]}/numa_node") + glob.glob(f"/sys/bus/pci/devices/{bdf}/numa_node"):
        print(p, open(p).read().strip())
    print("affinity now:", len(os.sched_getaffinity(0)), "cpus")
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    for n in ["none"] + [str(x) for x in nodes]:
        subprocess.run([sys.executable, __file__, n])
