import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
dev = torch.device("cuda:0")
net = Network(create_network(DEFAULT_ENV_ARGS))
E = 131072
env = BatchedFlexProvisionEnv(None, n_envs=E, device=dev, profiles=synthetic_profiles(net, 5, T=105216))
env.reset(return_obs=False)
a = torch.rand(E, 5, 4, device=dev)
env.step(a, want_info=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for k in range(30):
    flush.zero_()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); st = env.get_state(); s1.record(); torch.cuda.synchronize(); ts.append(s0.elapsed_time(s1))
print("get_state us", float(np.median(ts)) * 1e3, tuple(st.shape), st.dtype)
ts = []
for k in range(30):
    flush.zero_()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); ob = env.get_obs(); s1.record(); torch.cuda.synchronize(); ts.append(s0.elapsed_time(s1))
print("get_obs (push + window view) us", float(np.median(ts)) * 1e3, tuple(ob.shape), ob.dtype)
