"""Config-2 style microbenchmark: batched power flow only (fp_power_flow) over N envs, CUDA events."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
flows = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
env = BatchedFlexProvisionEnv({"pf_max_iter": int(os.environ.get("PF_MAX_ITER", "32"))}, n_envs=8, device=dev, profiles=synthetic_profiles(network, 5, T=2000))
rng = np.random.default_rng(0)
lvl = rng.uniform(0.35, 1.0, (n, 1))
p = torch.from_numpy(network.base_p[None, 1:] * lvl * (1 + 0.1 * rng.standard_normal((n, 32)))).to(dev)
q = torch.from_numpy(network.base_q[None, 1:] * lvl * (1 + 0.1 * rng.standard_normal((n, 32)))).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    out = env.power_flow(p, q, want_flows=bool(flows))
ts = []
for _ in range(iters):
    if not os.environ.get("NO_FLUSH"):
        flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(int(os.environ.get("REPEAT", "1"))):
        out = env.power_flow(p, q, want_flows=bool(flows))
    b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
it = out["iters"].cpu().numpy()
print(json.dumps({"lib": os.environ.get("FLEXGPU_LIB", "default"), "n": n, "ms": ms, "solves_per_s": n / ms * 1e3,
                  "iters_hist": np.bincount(it).tolist(), "failed": int(out["failed"].sum())}))
