"""Debug helper: per-array max differences between the CUDA path and the C mirror."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]
from oracle import c_mirror, env_ref, ieee33
from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles

variant = sys.argv[1] if len(sys.argv) > 1 else "thread"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
network = Network(create_network(DEFAULT_ENV_ARGS))
prof = synthetic_profiles(network, 5, T=4000, seed=0)
fo = c_mirror.make_net(ieee33.tree_arrays(ieee33.create_network()), env_ref.DEFAULT_ARGS, env_ref.DEFAULT_ARGS["buildings"],
                       variant=c_mirror.VARIANT_THREAD if variant == "thread" else c_mirror.VARIANT_WARP)
env = BatchedFlexProvisionEnv({"kernel_variant": variant}, n_envs=n, device="cuda:0", profiles=prof)
mb = c_mirror.MirrorBatch(fo, prof.as_dict(), n, keep_flows=True)
env.keep_line_flows(True)
rng = np.random.default_rng(1)
start = rng.integers(0, env.max_start(), n).astype(np.int32)
e0 = rng.uniform(0.9 * 0.0125, 1.1 * 0.0125, (n, 5)); a0 = rng.uniform(0, 1, (n, 20))
env.reset(start, e0, a0, return_obs=False); mb.reset(start, e0, a0)


def report(tag):
    d = {
        "V": (env.voltages.cpu().numpy(), mb.V), "E_cur": (env.ess_energy.cpu().numpy(), mb.E_cur),
        "E_init": (env.initial_ess_energy.cpu().numpy(), mb.E_init), "setp": (env.setpoints.cpu().numpy(), mb.setp),
        "cum": (env.cumulative_reward.cpu().numpy(), mb.cum), "iters": (env.pf_iterations.cpu().numpy(), mb.iters),
        "flags": (env.flags.cpu().numpy(), mb.flags), "vcount": (env.violation_count.cpu().numpy(), mb.vcount),
        "P": (env.line_flows[0].cpu().numpy(), mb.pfl), "Q": (env.line_flows[1].cpu().numpy(), mb.qfl),
        "ell": (env.line_flows[2].cpu().numpy(), mb.isq),
    }
    for k, (a, b) in d.items():
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        bad = np.argwhere(a != b)
        print(tag, k, "max|diff| %.3e" % np.max(np.abs(a - b)), "n_diff", len(bad), "first", bad[:3].tolist())


report("reset")
for t in range(3):
    a = rng.normal(0.3, 0.5, (n, 20)).astype(np.float32)
    r, d, info = env.step(torch.from_numpy(a)); rr, dd, ii = mb.step(a)
    print("step", t, "reward diff %.3e" % np.max(np.abs(r.cpu().numpy() - rr)), "info diff %.3e" % np.max(np.abs(env._info.cpu().numpy() - ii)))
    report(f"step{t}")
