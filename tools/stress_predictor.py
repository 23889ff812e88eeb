"""Stress check of the warp-specialised predictor pipeline: every row of many launches (sizes that end on
full, partial and odd tiles; concurrent memory traffic on a second stream to perturb the timing) against a
torch fp64 reference of the same affine map and penalty.  usage: python tools/stress_predictor.py [iters]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
from flexgpu.predictor import DeviceReplayBuffer, VoltagePredictor

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
network = Network(create_network(DEFAULT_ENV_ARGS))
env = BatchedFlexProvisionEnv(None, n_envs=8, device=dev, profiles=synthetic_profiles(network, 5, T=2000))
g = np.load(os.path.join(ROOT, "tests", "golden", "predictor_golden.npz"))
pred = VoltagePredictor.from_linear_model(env, g["coef"], g["intercept"], g["x_scale"], g["x_min"], g["y_scale"], g["y_min"])
A = torch.from_numpy(pred.A).to(dev); c = torch.from_numpy(pred.c).to(dev)
base = torch.from_numpy(np.stack([network.base_p, network.base_q], axis=1).reshape(-1)).to(dev)
sizes = [1, 127, 128, 129, 4097, 18944, 148 * 128 * 5, 148 * 128 * 5 + 77, 262144, 262144 + 33, 1 << 20]
noise = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream()
gen = torch.Generator(device=dev); gen.manual_seed(7)
worst_v, worst_p, launches = 0.0, 0.0, 0
for it in range(iters):
    n = sizes[it % len(sizes)]
    scale = 0.7 + 1.3 * torch.rand(n, 66, device=dev, dtype=torch.float64, generator=gen)   # up to 2x load: some rows violate
    X = (base[None, :] * scale).float().contiguous()
    buf = DeviceReplayBuffer(n + 5, {"v_pred": 33, "safety_penalty": 1}, device=dev)
    buf.reserve(n + 5)                                                                         # full ring: physical row = logical row
    pos = int(it * 3) % (n + 5)                                                                # wrapping ring segments too
    if it % 2:
        with torch.cuda.stream(side):
            noise.add_(1)                                                                       # HBM traffic racing the kernel
    vhat, pen = pred.predict(X, sink=buf, pos=pos)
    torch.cuda.synchronize()
    ref = X.double() @ A.T + c
    ev = float((vhat.double() - ref).abs().max())
    vd = vhat.double()
    pref = 1000.0 * (torch.clamp(pred.v_min - vd, min=0) + torch.clamp(vd - pred.v_max, min=0)).sum(dim=1)
    ep = float(((pen - pref).abs() / (1.0 + pref.abs())).max())
    idx = (pos + torch.arange(n, device=dev)) % (n + 5)
    ring_ok = torch.equal(buf.get_batch(n + 5, start=0)["v_pred"][idx], vhat)
    worst_v, worst_p, launches = max(worst_v, ev), max(worst_p, ep), launches + 1
    buf.close()
    assert ev < 2e-6 and ep < 1e-12 and ring_ok, (it, n, ev, ep, ring_ok)
print(json.dumps({"launches": launches, "max_abs_err_vs_fp64": worst_v, "max_rel_penalty_err": worst_p, "ring": "bit-identical"}))
