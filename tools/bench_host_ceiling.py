"""Host-memory ceiling of the host-buffer step (fp_step_host) when every GPU of the box runs it at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_host_ceiling.py

Every rank repeats one step's traffic -- 10.5 MB of pinned actions host -> device, 1.2 MB of reward + done device -> host
(131 072 envs) -- (a) alone, one rank after the other, and (b) all ranks at once; rank 0 prints the per-link and
aggregate GB/s.  If (b) per link is well below (a), the limiter is the host side (one NUMA domain's memory / root
complexes), not the GPU and not a collective (there is none on this path)."""
import json, os, subprocess, time
import torch, torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
E = 131072
h_act = torch.rand(E * 20, dtype=torch.float32).pin_memory(); d_act = torch.empty_like(h_act, device=dev)
d_out = torch.rand(E * 9 // 4 + 1, dtype=torch.float32, device=dev); h_out = torch.empty_like(d_out, device="cpu").pin_memory()
nbytes = h_act.numel() * 4 + d_out.numel() * 4


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(iters=200):
    for _ in range(10):
        d_act.copy_(h_act, non_blocking=True); h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        d_act.copy_(h_act, non_blocking=True); h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * iters / (time.perf_counter() - t0) / 1e9


alone = torch.zeros(world, dtype=torch.float64, device=dev)
for r in range(world):
    barrier()
    if r == rank:
        alone[r] = run()
barrier()
together = torch.zeros(world, dtype=torch.float64, device=dev)
together[rank] = run()
barrier()
if world > 1:
    dist.all_reduce(alone); dist.all_reduce(together)
if rank == 0:
    topo = ""
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        pass
    print(json.dumps({"ranks": world, "bytes_per_step": nbytes, "alone_GBps_per_link": [round(float(x), 1) for x in alone],
                      "together_GBps_per_link": [round(float(x), 1) for x in together],
                      "aggregate_together_GBps": round(float(together.sum()), 1), "aggregate_if_independent_GBps": round(float(alone.sum()), 1),
                      "host_cores": os.cpu_count()}))
    print(topo)
if world > 1:
    dist.destroy_process_group()
