"""Device rollout with every Transition written (DeviceRollout, record_envs = all): a few steps for an ncu launch list,
or timed with CUDA events.  python tools/profile_rollout.py [ENVS] [STEPS] [time]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
from flexgpu.policy import DevicePolicy, DeviceRollout, TRANSITION_FIELDS
from flexgpu.predictor import DeviceReplayBuffer

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
timed = len(sys.argv) > 3
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
prof = synthetic_profiles(network, 5, T=105216)
env = BatchedFlexProvisionEnv({}, n_envs=E, device=dev, profiles=prof, network=network.dict, seed=5)
rng = np.random.default_rng(7)
sd = {"fc1.weight": rng.normal(0, 0.1, (64, 149)), "fc1.bias": rng.uniform(-0.08, 0.08, 64), "layernorm.weight": np.ones(64),
      "layernorm.bias": np.zeros(64), "rnn.weight_ih": rng.uniform(-0.125, 0.125, (192, 64)),
      "rnn.weight_hh": rng.uniform(-0.125, 0.125, (192, 64)), "rnn.bias_ih": rng.uniform(-0.125, 0.125, 192),
      "rnn.bias_hh": rng.uniform(-0.125, 0.125, 192), "fc2.weight": rng.normal(0, 0.1, (4, 64)), "fc2.bias": rng.uniform(-0.125, 0.125, 4)}
pol = DevicePolicy(sd, device=dev, std=1.0, seed=11)
buf = DeviceReplayBuffer(2 * E, TRANSITION_FIELDS, device=dev)
ro = DeviceRollout(env, pol, replay=buf, record_envs=E)
ro.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for k in range(steps + (3 if timed else 0)):
    if timed:
        flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ro.step(); b.record(); torch.cuda.synchronize()
    if k >= 3:
        ts.append(a.elapsed_time(b) * 1e3)
if timed:
    print(json.dumps({"envs": E, "rollout_with_writes_us": round(float(np.median(ts)), 1)}))
