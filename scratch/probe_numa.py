"""Probe: NUMA placement of pinned host memory vs D2H bandwidth (one process per visible GPU when launched under torchrun)."""
import glob, os, subprocess, sys, time
import torch

rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
    for n in sorted(glob.glob("/sys/devices/system/node/node*")):
        print(n, open(n + "/cpulist").read().strip())
    print("affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus")
torch.cuda.set_device(rank)
bus = torch.cuda.get_device_properties(rank)
pci = None
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(rank)
    pci = pynvml.nvmlDeviceGetPciInfo(h).busId
    if isinstance(pci, bytes): pci = pci.decode()
except Exception as e:
    print("pynvml", e)
node = None
if pci:
    p = "/sys/bus/pci/devices/" + pci.lower()[-12:] + "/numa_node"
    if os.path.exists(p):
        node = int(open(p).read())
print(f"rank {rank}: pci {pci} numa_node {node}", flush=True)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))

def bw(tag):
    n = 64 << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = time.perf_counter()
    for _ in range(20): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"rank {rank} {tag}: D2H {20 * n / dt / 1e9:.1f} GB/s", flush=True)
    if world > 1: dist.barrier()

bw("default placement")
if node is not None and node >= 0:
    cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    ids = set()
    for part in cpus.split(","):
        a, _, b = part.partition("-")
        ids.update(range(int(a), int(b or a) + 1))
    ids &= os.sched_getaffinity(0) if False else ids
    try:
        os.sched_setaffinity(0, ids)
        bw(f"affinity node {node}")
    except Exception as e:
        print("setaffinity failed", e)
