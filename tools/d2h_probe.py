"""Concurrent device->host bandwidth per rank (pinned memory), with and without binding the process to the GPU's
NUMA-local CPUs before the pinned allocation.  torchrun --nproc-per-node N tools/d2h_probe.py [bind]"""
import os, subprocess, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("LOCAL_RANK", "0"))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
torch.cuda.set_device(rank)
if bind:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(rank)
        n_cpu = os.cpu_count()
        words = (n_cpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1 and 64 * w + b < n_cpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # noqa: BLE001
        print("rank", rank, "bind failed:", e)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
nbytes = 1 << 30
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
h_ = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
for _ in range(2):
    h_.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(8):
    h_.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("rank %d bind=%s cpus=%d: D2H %.1f GB/s" % (rank, bind, len(os.sched_getaffinity(0)), 8 * nbytes / dt / 1e9), flush=True)
dist.barrier()
if False and rank == 0 and not bind:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
dist.destroy_process_group()
