"""Host-side cost of the Python verbs (small batch so that the GPU is never the bottleneck)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
dev = torch.device("cuda:0")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = BatchedFlexProvisionEnv(None, n_envs=E, device=dev, profiles=synthetic_profiles(Network(create_network(DEFAULT_ENV_ARGS)), 5, T=105216), seed=5)
env.reset()
a = torch.rand(E, 5, 4, device=dev)
def t(f, n=1000):
    env.reset(return_obs=False)                      # stay inside the episode slice
    for _ in range(10): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if k % 80 == 79: env.reset(return_obs=False)
        f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("envs", E)
print("step(want_info=True)  us/call", round(t(lambda: env.step(a)), 1))
print("step(want_info=False) us/call", round(t(lambda: env.step(a, want_info=False)), 1))
print("get_obs()             us/call", round(t(lambda: env.get_obs()), 1))
print("get_state()           us/call", round(t(lambda: env.get_state()), 1))
print("step+get_obs          us/call", round(t(lambda: (env.step(a, want_info=False), env.get_obs())), 1))
print("step(return_obs=True)  us/call", round(t(lambda: env.step(a, want_info=False, return_obs=True)), 1))
