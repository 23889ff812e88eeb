"""H2D bandwidth of a 10.5 MB pinned buffer (one step's actions at 131072 envs) split over k streams."""
import time, torch
dev = torch.device("cuda:0")
n = 131072 * 20
h = torch.rand(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)
for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    c = n // k
    def run():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    for _ in range(5): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): run(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(f"streams={k}: {dt*1e6:.1f} us per 10.5 MB = {n*4/dt/1e9:.1f} GB/s")
big = torch.rand(64 << 20, dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device=dev)
dbig.copy_(big, non_blocking=True); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"256 MB: {big.numel()*4/dt/1e9:.1f} GB/s")
