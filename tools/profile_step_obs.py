"""Driver for an ncu capture of the fused step+observation kernel: 12 L2-flushed calls of step(return_obs=True)
on 131 072 envs (ncu -k regex:k_env_t -s 9 -c 1 ... python tools/profile_step_obs.py)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
dev = torch.device("cuda:0")
net = Network(create_network(DEFAULT_ENV_ARGS))
E = 131072
env = BatchedFlexProvisionEnv(None, n_envs=E, device=dev, profiles=synthetic_profiles(net, 5, T=105216))
env.reset(return_obs=False)
a = torch.rand(E, 5, 4, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for k in range(12):
    flush.zero_()
    env.step(a, want_info=False, return_obs=True)
torch.cuda.synchronize()
