"""e2e microbenchmark of fp_step_host (pinned host buffers in/out) for a given FLEXGPU_HOST_CHUNKS."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
env = BatchedFlexProvisionEnv(None, n_envs=E, device=dev, profiles=synthetic_profiles(network, 5, T=105216), seed=5)
env.reset(return_obs=False)
h_act = torch.rand(4, E, 20, dtype=torch.float32).pin_memory()
d_act = torch.empty(E, 20, dtype=torch.float32, device=dev)
for k in range(5):
    env.step_host(h_act[k % 4].numpy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(60):
    env.step_host(h_act[k % 4].numpy())
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 60
# components
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
a.record(); d_act.copy_(h_act[0], non_blocking=True); b.record(); torch.cuda.synchronize()
h2d = a.elapsed_time(b)
print(json.dumps({"chunks": os.environ.get("FLEXGPU_HOST_CHUNKS", "default"), "envs": E, "us_per_step": dt * 1e6,
                  "env_steps_per_s": E / dt, "h2d_us_alone": h2d * 1e3}))
