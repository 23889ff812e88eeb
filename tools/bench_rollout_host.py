"""Host side of the device rollout: python tools/bench_rollout_host.py ENVS RECORDED_ENVS -- per step, the time the host
needs to enqueue DeviceRollout.step() (2 launches; 6 with transition writes) against the GPU time of the step, 80 steps
back to back without a synchronisation (no L2 flush).  B200: 36 us of host time per step (88 us with transition writes)
against 318 / 175 / 41 us of GPU time at 131072 / 65536 / 8192 envs: the loop is GPU-bound from ~8 k envs up."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
from flexgpu.policy import DevicePolicy, DeviceRollout, TRANSITION_FIELDS
from flexgpu.predictor import DeviceReplayBuffer
E = int(sys.argv[1]); rec = int(sys.argv[2]); critic = len(sys.argv) > 3
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
buf = DeviceReplayBuffer(max(2 * rec, 64), TRANSITION_FIELDS, device=dev) if rec else None
ro = DeviceRollout(env, pol, replay=buf, record_envs=rec) if rec else DeviceRollout(env, pol)
ro.reset()
for _ in range(10): ro.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
N = 80
for _ in range(N): ro.step()
b.record(); t1 = time.perf_counter()      # host enqueue time
torch.cuda.synchronize(); t2 = time.perf_counter()
print(json.dumps({"envs": E, "rec": rec, "host_enqueue_us_per_step": (t1 - t0) / N * 1e6, "wall_us_per_step": (t2 - t0) / N * 1e6, "gpu_us_per_step": a.elapsed_time(b) / N * 1e3}))
