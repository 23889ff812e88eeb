"""Step-kernel time for combinations of solver knobs: python tools/bench_step_knobs.py ENVS key=v1,v2 [key=...]
e.g.  python tools/bench_step_knobs.py 131072 pf_f32_passes=0,4,5,6
Each combination: fresh env, Philox reset, 25 timed steps (L2 flushed before each, CUDA events), median."""
import itertools, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles


def parse(v):
    try:
        return int(v)
    except ValueError:
        try:
            return float(v)
        except ValueError:
            return v


E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
knobs = {k: [parse(x) for x in v.split(",")] for k, v in (a.split("=") for a in sys.argv[2:])}
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
prof = synthetic_profiles(network, 5, T=105216)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for combo in itertools.product(*knobs.values()):
    cfg = dict(zip(knobs.keys(), combo))
    envk = {k[4:]: str(v) for k, v in cfg.items() if k.startswith("ENV_")}
    os.environ.update(envk)
    env = BatchedFlexProvisionEnv({k: v for k, v in cfg.items() if not k.startswith("ENV_")}, n_envs=E, device=dev,
                                  profiles=prof, network=network.dict, seed=5)
    env.reset(return_obs=False)
    acts = torch.rand(4, E, 5, 4, device=dev)
    ts, to = [], []
    NOFLUSH = bool(os.environ.get("NOFLUSH"))              # NOFLUSH=1: no flush between steps -- the GPU then idles before each launch and the events also see the launch path (+11-13 us): not a kernel time
    for k in range(30):
        if not NOFLUSH:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[k % 4], want_info=True); b.record(); torch.cuda.synchronize()
        if k >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    for k in range(12):
        if not NOFLUSH:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[k % 4], want_info=True, return_obs="ring"); b.record(); torch.cuda.synchronize()
        if k >= 4:
            to.append(a.elapsed_time(b) * 1e3)
    it = env.pf_iterations.float()
    us = float(np.median(ts))
    print(json.dumps({**cfg, "envs": E, "passes_mean": round(float(it.mean()), 3), "passes_max": int(it.max()),
                      "us_median": round(us, 2), "env_steps_per_s": round(E / us * 1e6 / 1e9, 4),
                      "hbm_frac": round(E * 1256 / (us * 1e-6) / 6553.3e9, 4), "step_ring_us": round(float(np.median(to)), 2)}), flush=True)
    env.close()
