"""Step-kernel time as a function of the number of sweep passes (pf_tol knob): separates the per-pass
cost of the fixed point from the fixed per-tile cost (loads, setpoints, final pass, epilogue)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
prof = synthetic_profiles(network, 5, T=105216)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for tol in (1e3, 1e-1, 1e-3, 1e-6, 1e-9):
    env = BatchedFlexProvisionEnv(dict(pf_tol=tol), n_envs=E, device=dev, profiles=prof, seed=5)
    env.reset(return_obs=False)
    acts = torch.rand(4, E, 5, 4, device=dev)
    ts = []
    for k in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[k % 4], want_info=True); b.record(); torch.cuda.synchronize()
        if k >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    it = env.pf_iterations.float()
    print(json.dumps({"pf_tol": tol, "passes_mean": float(it.mean()), "passes_max": int(it.max()), "us_median": float(np.median(ts))}))
    env.close()
