#!/usr/bin/env python3
"""Host-side ceiling of the end-to-end leg: pinned-host -> device cudaMemcpyAsync bandwidth per GPU and in aggregate.
  torchrun --nproc-per-node N scripts/probe_h2d.py [--mb 512] [--bind 0|1]
Every rank copies its own pinned buffer to its own GPU: first one rank at a time (the others idle), then all ranks at once.
--bind 1 pins the process to the CPUs NVML reports as local to its GPU BEFORE the pinned buffer is allocated, so the
buffer's pages land on that GPU's NUMA node.  Rank 0 prints one JSON line (plus `nvidia-smi topo -m` to stderr)."""
import argparse, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from trackingbench_slam_b200.hostbind import bind_to_gpu

ap = argparse.ArgumentParser()
ap.add_argument('--mb', type=int, default=512)
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--bind', type=int, default=1)
a = ap.parse_args()
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
bound = bind_to_gpu(local) if a.bind else None
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = a.mb << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device='cuda')
st = torch.cuda.Stream()

def run():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        d.copy_(h, non_blocking=True)
        st.synchronize()
        e0.record(st)
        for _ in range(a.reps):
            d.copy_(h, non_blocking=True)
        e1.record(st)
        st.synchronize()
    return n * a.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

solo = 0.0
for r in range(world):          # one rank at a time
    barrier()
    if r == rank:
        solo = run()
barrier()
t0 = time.perf_counter()
together = run()                # all ranks at once
barrier()
wall = time.perf_counter() - t0
vals = torch.tensor([solo, together], dtype=torch.float64, device='cuda')
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    solo_l = [float(v[0]) for v in allv]
    tog_l = [float(v[1]) for v in allv]
    print(json.dumps({'probe': 'pinned H2D cudaMemcpyAsync, %d MB x %d per rank' % (a.mb, a.reps), 'n_gpus': world, 'numa_bind': bool(a.bind),
                      'bound_cpus_rank0': bound, 'solo_GBps_per_gpu': [round(x, 2) for x in solo_l],
                      'concurrent_GBps_per_gpu': [round(x, 2) for x in tog_l], 'concurrent_aggregate_GBps': round(sum(tog_l), 2),
                      'host_cpus': os.cpu_count()}))
    try:
        sys.stderr.write(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout)
    except Exception:
        pass
if world > 1:
    dist.destroy_process_group()
