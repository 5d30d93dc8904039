#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi topo -m
echo; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
echo; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo "$(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done
echo; python - <<'PY'
import os
print("affinity of this process:", len(os.sched_getaffinity(0)), "cpus")
PY
} > gpurun_out/topo.txt 2>&1
cat > /tmp/h2d.py <<'PY'
import os, sys, time, torch
r = int(os.environ.get("LOCAL_RANK", "0")); bind = os.environ.get("BIND", "0") == "1"
torch.cuda.set_device(r)
if bind:
    bdf = torch.cuda.get_device_properties(r).pci_bus_id if hasattr(torch.cuda.get_device_properties(r), "pci_bus_id") else None
    import subprocess
    bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(r)], capture_output=True, text=True).stdout.strip().lower()
    bdf = bdf[4:] if bdf.startswith("0000") and len(bdf) > 12 else bdf
    path = "/sys/bus/pci/devices/%s/local_cpulist" % bdf
    cpus = set()
    for part in open(path).read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    os.sched_setaffinity(0, cpus)
import torch.distributed as dist
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
h = torch.empty(400 << 20, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty_like(h, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dist.barrier(device_ids=[r]); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gb = 20 * h.numel() / dt / 1e9
t = torch.tensor([gb], device="cuda"); lst = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
dist.all_gather(lst, t)
if r == 0: print("bind=%s world=%d H2D GB/s per rank:" % (bind, dist.get_world_size()), [round(float(x), 1) for x in lst], "sum", round(sum(float(x) for x in lst), 1))
dist.destroy_process_group()
PY
for n in 2 4 8; do for b in 0 1; do
BIND=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n /tmp/h2d.py 2>/dev/null | grep "H2D" >> gpurun_out/topo.txt
done; done
cat gpurun_out/topo.txt
