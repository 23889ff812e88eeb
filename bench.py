#!/usr/bin/env python
"""bench.py -- 33-bus env-steps/s (incl. power flow) of the fused flex_provision step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]

A "step" is one fused env step (action application, injections, DistFlow power flow,
constraint masks, reward/penalty, ESS update, bookkeeping: flexibility_provision_env.py:241-356)
over the E environments resident on each GPU.  Workload: E = 131 072 envs per GPU -- one GPU's
shard of BASELINE config 5 (2^20 envs over 8 B200), the configuration the 1/2/4/8-GPU metric is
quoted on; at N = 8 the run IS config 5.  Every rank holds its own E envs (weak scaling, no
collective on the step path; one NCCL all-reduce of the 16-double statistics vector after the
timed region).  Config 3 (65 536 envs on one B200) is measured in the same run at N = 1 and
reported as `config3`.  Episodes are 95 steps long (quirk Q1), so a Philox auto-reset launch runs
inside the timed region every 94 steps -- it is part of a rollout and is counted.

Timing: W untimed steps, then K steps, each bracketed by CUDA events on the launching stream
with an L2 flush (256 MiB memset, untimed) before it -- a good part of the per-GPU working set
would otherwise sit in the 126 MB L2.  value = total envs * K / max-over-ranks(sum of step times).
e2e: the same step through fp_step_host (pinned host actions in, reward+done out, copies and
the stream sync inside the timed region), wall-clocked.

--impl reference: the CPU arm.  The reference's own implementation (Pyomo + IPOPT) is not
installable here, so this times the oracle port: oracle/c/flex_oracle.c (OpenMP,
all host cores) on a bounded sample of the same workload, plus the single-core Python
restatement (oracle/env_ref.py) for scale.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "safe-marl_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "33-bus env-steps/sec (incl. power flow)"
UNIT = "env-steps/s"
B_ALG = 1256            # algorithmic bytes per env-step (DESIGN.md "Roofline")
EPISODE_STEPS = 94      # steps between auto-resets (an episode terminates at its 95th step)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_mirror(n, prof, seed):
    from oracle import c_mirror, env_ref, ieee33
    fonet = c_mirror.make_net(ieee33.tree_arrays(ieee33.create_network()), env_ref.DEFAULT_ARGS,
                              env_ref.DEFAULT_ARGS["buildings"])
    mb = c_mirror.MirrorBatch(fonet, prof.as_dict(), n)
    mb.reset_random(seed)
    return mb


def cpu_port_rate(prof, n_envs, budget_s, seed=5):
    """The C port (OpenMP, all cores) on a bounded sample: `n` envs stepped until budget_s."""
    import numpy as np
    from oracle import c_mirror
    cores = c_mirror.use_all_cores()
    n = min(n_envs, 65536)
    mb = make_mirror(n, prof, seed)
    rng = np.random.default_rng(0)
    acts = rng.uniform(0, 1, (4, n, 20)).astype(np.float32)
    mb.step(acts[0])                                           # warm-up (page faults, thread pool)
    t0 = time.perf_counter(); k = 0
    while True:
        mb.step(acts[k % 4]); k += 1
        el = time.perf_counter() - t0
        if el >= budget_s or k >= 90:
            break
    return n * k / el, cores, f"{n} envs x {k} steps of the C port (oracle/c/flex_oracle.c, OpenMP) in {el:.2f} s"


def python_restatement_rate(prof, budget_s=5.0):
    """Single-core Python restatement (env_ref + sweep power flow): the reference's structure
    (per-env Python objects) minus Pyomo/IPOPT, i.e. a generous stand-in for its per-env step."""
    import numpy as np
    from oracle import env_ref, ieee33
    env = env_ref.RefFlexEnv(dict(env_ref.DEFAULT_ARGS), ieee33.create_network(), prof.as_dict(),
                             rng=np.random.RandomState(0), pf_method='sweep')
    rng = np.random.RandomState(1)
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < budget_s:
        _, done, _ = env.step(rng.uniform(0, 1, 20))
        k += 1
        if done:
            env.reset()
    return k / (time.perf_counter() - t0)


def ncu_traffic(envs_per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_env_t<STEP> launch from the committed
    `ncu --set full` capture (profiles/r1_step_kernel_traffic.json), if it was taken at this launch size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_step_kernel_traffic.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch"]) if int(d["envs"]) == int(envs_per_gpu) else None
    except (OSError, ValueError, KeyError):
        return None


def workload_name(envs_per_gpu):
    return (f"fused_env_step_{envs_per_gpu}_envs_per_gpu (one GPU's shard of BASELINE config 5 = 2^20 envs over 8 GPUs; "
            f"N=8 is config 5)")


def run_reference(args):
    """--impl reference: the CPU arm (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from flexgpu import DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles
    network = Network(create_network(DEFAULT_ENV_ARGS))
    prof = synthetic_profiles(network, 5, T=args.rows, seed=0)
    import numpy as np
    from oracle import c_mirror
    cores = c_mirror.use_all_cores()                             # torchrun exports OMP_NUM_THREADS=1
    n = min(args.envs_per_gpu, 65536)
    mb = make_mirror(n, prof, 5)
    rng = np.random.default_rng(0)
    acts = rng.uniform(0, 1, (4, n, 20)).astype(np.float32)
    steps = max(1, min(args.steps, 60)); warm = max(1, min(args.warmup, 3))
    for w in range(warm):
        mb.step(acts[w % 4])
    t0 = time.perf_counter()
    for k in range(steps):
        mb.step(acts[k % 4])
    el = time.perf_counter() - t0
    value = n * steps / el
    py = python_restatement_rate(prof, 5.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs_per_gpu), "envs_per_gpu": args.envs_per_gpu,
                   "total_envs": args.envs_per_gpu * args.gpus, "profile_rows": args.rows,
                   "actions": "fp32 uniform(0,1), host memory",
                   "sample": f"each step = {n} envs of that workload (bounded CPU sample), rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} envs x {steps} steps, C port of the step (oracle/c/flex_oracle.c, OpenMP); "
                                   f"the reference itself (Pyomo+IPOPT per step) is not installable; "
                                   f"single-core Python restatement: {py:.0f} env-steps/s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--rows", type=int, default=105216, help="profile dataset rows (bundled-data shape)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="skip the extra 65 536-env (config 3) measurement at N=1")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush (diagnostics only)")
    ap.add_argument("--no-config4", action="store_true", help="skip the extra predictor (config 4) measurement at N=1")
    ap.add_argument("--no-obs", action="store_true", help="skip the extra step+get_obs measurement (rollout-loop cost)")
    ap.add_argument("--variant", default="thread", choices=["thread", "warp"], help="kernel variant (see include/flexgpu.h)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from flexgpu import BatchedFlexProvisionEnv, DEFAULT_ENV_ARGS, Network, create_network, synthetic_profiles

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3); K = args.steps; E = args.envs_per_gpu

    network = Network(create_network(DEFAULT_ENV_ARGS))
    prof = synthetic_profiles(network, 5, T=args.rows, seed=0)
    env = BatchedFlexProvisionEnv({"kernel_variant": args.variant}, n_envs=E, device=dev, profiles=prof, seed=5,
                                  env_offset=rank * E)
    env.reset(return_obs=False)
    n_act = 8
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    acts = torch.rand(n_act, E, 5, 4, device=dev, dtype=torch.float32, generator=g)     # resident in HBM
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"t": 0}

    def one_step(k, timed):
        if state["t"] == EPISODE_STEPS:
            env.reset(return_obs=False)                       # Philox auto-reset, one launch
            state["t"] = 0
        env.step(acts[k % n_act], want_info=True)
        state["t"] += 1

    for k in range(W):
        if not args.no_flush:
            flush.zero_()
        one_step(k, False)
    barrier()
    env.episode_stats(reduce=False, reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    barrier()
    launches0 = env.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for k in range(K):
        if not args.no_flush:
            flush.zero_()
        ev[k][0].record(stream)
        one_step(W + k, True)
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = env.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # kernel-only duration of the dominant kernel (k_env<STEP>): steps without a reset in them
    step_ms = sorted(a.elapsed_time(b) for a, b in ev)
    kern_ms = statistics.median(step_ms)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: through fp_step_host with pinned HOST buffers (copies + sync inside)
    h_act = torch.rand(4, E, 20, dtype=torch.float32).pin_memory()
    Ke = max(10, min(K, 100))
    env.reset(return_obs=False); state["t"] = 0
    for k in range(3):
        env.step_host(h_act[k % 4].numpy())
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        if 3 + k == EPISODE_STEPS:
            env.reset(return_obs=False)
        r_h, d_h, _ = env.step_host(h_act[k % 4].numpy())
    barrier()
    e2e_s = time.perf_counter() - t0

    obs_extra = None
    if not args.no_obs:
        env.reset(return_obs=False)
        for k in range(3):                                        # warm-up: allocates the observation window ring
            env.step(acts[k % n_act], want_info=False, return_obs=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record(stream)
        for k in range(60):
            env.step(acts[k % n_act], want_info=False, return_obs=True)
        e1.record(stream); barrier()
        obs_extra = E * world * 60 / (e0.elapsed_time(e1) * 1e-3)

    # ---- BASELINE config 3 (65 536 envs on one B200), same method, N = 1 only
    config3 = None
    if world == 1 and E != 65536 and not args.no_config3:
        E3 = 65536
        env3 = BatchedFlexProvisionEnv({"kernel_variant": args.variant}, n_envs=E3, device=dev, profiles=prof, seed=5)
        env3.reset(return_obs=False)
        acts3 = torch.rand(4, E3, 5, 4, device=dev, dtype=torch.float32, generator=g)
        K3 = 60
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K3)]
        for k in range(5):
            flush.zero_(); env3.step(acts3[k % 4], want_info=True)
        for k in range(K3):
            flush.zero_()
            ev3[k][0].record(stream); env3.step(acts3[k % 4], want_info=True); ev3[k][1].record(stream)
        torch.cuda.synchronize()
        ms3 = [a.elapsed_time(b) for a, b in ev3]
        config3 = {"workload": "fused_env_step_65536_envs (BASELINE config 3)", "value": E3 * K3 / (sum(ms3) * 1e-3),
                   "unit": UNIT, "steps": K3, "kernel_ms_median": statistics.median(ms3),
                   "roofline_frac": B_ALG * E3 / (statistics.median(ms3) * 1e-3) / 1e9 / peaks()[0]}
        env3.close()

    # ---- BASELINE config 4 (voltage predictor + safety penalty -> replay ring, 262 144 envs), N = 1 only
    config4 = None
    if world == 1 and not args.no_config4:
        import numpy as np
        from flexgpu import Network, create_network, DEFAULT_ENV_ARGS
        from flexgpu.predictor import DeviceReplayBuffer, VoltagePredictor
        E4 = 262144
        gp = np.load(os.path.join(ROOT, "tests", "golden", "predictor_golden.npz"))
        pred = VoltagePredictor.from_linear_model(env, gp["coef"], gp["intercept"], gp["x_scale"], gp["x_min"], gp["y_scale"], gp["y_min"])
        net4 = Network(create_network(DEFAULT_ENV_ARGS))
        base4 = torch.from_numpy(np.stack([net4.base_p, net4.base_q], axis=1).reshape(-1)).to(dev)
        # X = interleaved [p_n, q_n] of load scenarios: base loads x U(0.7, 1.3) (data_generation.py:27-36)
        X4 = (base4[None, :] * (0.7 + 0.6 * torch.rand(E4, 66, device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(4)))).float().contiguous()
        ring = DeviceReplayBuffer(2 * E4, {"v_pred": 33, "safety_penalty": 1}, device=dev)
        K4 = 40
        ev4 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K4)]
        for k in range(5):
            pred.predict(X4, want_vhat=False, want_penalty=False, sink=ring)
        for k in range(K4):
            flush.zero_()
            ev4[k][0].record(stream); pred.predict(X4, want_vhat=False, want_penalty=False, sink=ring); ev4[k][1].record(stream)
        torch.cuda.synchronize()
        ms4 = statistics.median(a.elapsed_time(b) for a, b in ev4)
        b4 = 66 * 4 + 33 * 4 + 4                                  # X row in, Vhat row + penalty into the ring
        config4 = {"workload": "predictor_penalty_to_replay_ring_262144_envs (BASELINE config 4)", "value": E4 / (ms4 * 1e-3),
                   "unit": "envs/s", "kernel": "k_predict (tcgen05 kind::tf32, 3xTF32, A operand in TMEM)", "kernel_ms_median": ms4,
                   "bytes_per_env": b4, "roofline_frac": b4 * E4 / (ms4 * 1e-3) / 1e9 / peaks()[0], "steps": K4}
        ring.close()

    t = torch.tensor([dev_ms, e2e_s, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, kern_ms = (float(x) for x in t)
    stats = env.episode_stats(reduce=True)                     # the one NCCL collective (16 doubles)
    total_envs = E * world
    value = total_envs * K / (dev_ms * 1e-3)
    e2e = total_envs * Ke / e2e_s

    if rank == 0:
        peak, peak_src = peaks()
        achieved = B_ALG * E / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(E),
                       "kernel_variant": args.variant, "envs_per_gpu": E, "total_envs": total_envs, "profile_rows": args.rows,
                       "actions": "fp32 uniform(0,1), resident in HBM", "auto_reset_every": EPISODE_STEPS,
                       "l2": "flushed before every timed step (256 MiB memset, untimed)" if not args.no_flush else "NOT flushed",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks",
                       "wall_s_incl_flush": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": E * 20 * 4, "d2h_bytes_per_step": E * 9,
                    "steps": Ke, "api": "BatchedFlexProvisionEnv.step_host -> fp_step_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(E), "kernel": {"thread": "k_env_t<STEP>", "warp": "k_env<STEP>"}[args.variant], "bytes_per_env_step": B_ALG,
                         "kernel_ms_median": kern_ms, "peak_source": peak_src,
                         "note": "not HBM-bound: 17 fp64 ops x 32 lines x ~6 one-pass sweeps (a warp runs the maximum of its 32 envs; mean 5.2) + the final pass + the setpoint "
                                 "arithmetic = ~4400 fp64 lane-ops per env-step, which cap the kernel at ~3.9e9 env-steps/s "
                                 "on the fp64 pipe (59 lane-ops/clk/SM measured), i.e. 75 % of this HBM roofline at best; "
                                 "see DESIGN.md section 3"},
            "stats": {k: float(v) for k, v in stats.items()},
        }
        if obs_extra is not None:
            line["step_plus_get_obs_env_steps_per_s"] = obs_extra
        if config3 is not None:
            line["config3"] = config3
        if config4 is not None:
            line["config4"] = config4
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_port_rate(prof, E, budget_s=12.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
