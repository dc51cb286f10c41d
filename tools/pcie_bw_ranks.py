"""Host-link ceiling of the e2e path (qs_step_host) with every GPU of the box copying at once.  Launch under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bw_ranks.py
Each rank copies pinned host <-> device buffers of its own GPU (64 MB device-to-host, 16 MB host-to-device: the proportions of one
1,048,576-env step of qs_step_host), first alone (rank by rank), then all ranks together between barriers; rank 0 prints the
per-rank and the aggregate GB/s, plus the GPU / NUMA topology the driver reports."""
import os, subprocess, time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
D2H, H2D, K = 64 << 20, 16 << 20, 20
dd = torch.empty(D2H, dtype=torch.uint8, device="cuda"); hd = torch.empty(D2H, dtype=torch.uint8).pin_memory()
du = torch.empty(H2D, dtype=torch.uint8, device="cuda"); hu = torch.empty(H2D, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode):
    for _ in range(2 + K):
        if _ == 2:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s1): hd.copy_(dd, non_blocking=True)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s2): du.copy_(hu, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    return ((D2H if mode in ("d2h", "both") else 0) + (H2D if mode in ("h2d", "both") else 0)) / dt / 1e9


def gather(x):
    t = torch.zeros(world, dtype=torch.float64); t[rank] = x
    dist.all_reduce(t)
    return t.tolist()


for mode in ("d2h", "h2d", "both"):
    alone = 0.0
    for r in range(world):                       # one rank at a time
        dist.barrier()
        if r == rank:
            alone = run(mode)
    dist.barrier()
    a = gather(alone)
    dist.barrier()
    together = run(mode)                         # every rank at once
    b = gather(together)
    if rank == 0:
        print("%-4s alone    per rank GB/s: %s" % (mode, " ".join("%5.1f" % v for v in a)))
        print("%-4s together per rank GB/s: %s   aggregate %.1f GB/s" % (mode, " ".join("%5.1f" % v for v in b), sum(b)), flush=True)
if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            if cmd[0] == "lscpu":
                out = "\n".join(l for l in out.splitlines() if any(k in l for k in ("NUMA", "Socket", "Model name", "CPU(s):")))
            print(out, flush=True)
        except Exception as ex:
            print(cmd, ex)
dist.destroy_process_group()
