"""Config-4 microbenchmark: predictor + safety penalty -> replay ring over N envs (CUDA events)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
from flexgpu.predictor import DeviceReplayBuffer, VoltagePredictor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
env = BatchedFlexProvisionEnv(None, n_envs=8, device=dev, profiles=synthetic_profiles(network, 5, T=2000))
g = np.load(os.path.join(ROOT, "tests", "golden", "predictor_golden.npz"))
pred = VoltagePredictor.from_linear_model(env, g["coef"], g["intercept"], g["x_scale"], g["x_min"], g["y_scale"], g["y_min"])
buf = DeviceReplayBuffer(2 * n, {"v_pred": 33, "safety_penalty": 1}, device=dev)
X = torch.rand(n, 66, device=dev) * 0.4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    pred.predict(X, want_vhat=False, want_penalty=False, sink=buf)
ts = []
for _ in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pred.predict(X, want_vhat=False, want_penalty=False, sink=buf); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
bytes_alg = n * (66 * 4 + 33 * 4 + 4)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flops = 2.0 * n * 66 * 33
print(json.dumps({"workload": f"predictor+penalty->replay, {n} envs (BASELINE config 4)", "ms": ms, "envs_per_s": n / ms * 1e3,
                  "hbm_gbs_achieved": bytes_alg / ms / 1e6, "hbm_frac": bytes_alg / ms / 1e6 / peaks["hbm_gbs"],
                  "useful_tflops": flops / ms / 1e9, "issued_tf32_tflops": 3 * 2.0 * n * 72 * 48 / ms / 1e9,
                  "bytes_per_env": bytes_alg / n}))
