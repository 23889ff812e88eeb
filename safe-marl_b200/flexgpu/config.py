"""Environment configuration: the reference's env YAML keys -> the C-ABI FpConfig.

Key names and defaults follow madrl/args/env_args/flex_provision.yaml:3-33 of the
reference; `alg` (train_agent.py:67) selects the 'safemaddpg' pass-through branch of
step (flexibility_provision_env.py:268-274).
"""
from collections import namedtuple
from math import acos, tan

from . import _lib

DEFAULT_ENV_ARGS = dict(
    history=24, pv_scale=0.15, demand_scale=1.0, reactive_scale=1.0, v_max=1.1, v_min=0.9,
    data_path="./data", episode_limit=96, action_low=0, action_high=1.0, seed=0, e_min=0.0,
    e_max=0.025, pv_cost=0.05, ess_cost=0.03, discomfort_coeff=0.15, voltage_coeff=1.0,
    p_ch_max=0.005, p_dis_max=0.005, eta_ch=0.9, eta_dis=0.9, cos_phi_max=0.95,
    max_power_reduction=0.5, sample_interval="15min", buildings=[5, 10, 15, 20, 25],
    pv_nodes=[5, 10, 15, 20, 25], ess_nodes=[5, 10, 15, 20, 25], v_nom=12.66, s_nom=1000,
    pv_cap=0.15,
)

# Solver knobs that the reference leaves to IPOPT's defaults (utils/pf.py:101-102).
DEFAULT_SOLVER_ARGS = dict(
    kernel_variant="thread",  # "thread" (one thread per env, throughput) | "warp" (one warp per env)
    pf_tol=None,        # None -> DEFAULT_PF_TOL[variant]; see DESIGN.md "convergence"
    pf_max_iter=32,     # exceeding it counts as solver failure (:314-337)
    pf_f32_passes=None,  # None -> DEFAULT_PF_F32_PASSES[variant]: opening passes of a solve that run in fp32
    fail_penalty=200.0,  # :336
    e_next_lb=-1e-8,    # E_next >= 0 (pf.py:46) relaxed by IPOPT's bound_relax_factor
)


# thread variant: max |dl| (squared current) between sweeps; warp variant: max |dv| (squared voltage)
# thread: residual of the current row (pf.py:85-88) before the last update of the currents; that
# update contracts it by another ~0.05, so 1e-5 leaves V / P,Q / I within 1e-11 / 6e-9 / 6e-8 p.u. of the
# Newton solution (parity bar 1e-6) and saves the pass that 1e-6 would cost every warp (the slowest of its
# 32 envs decides): 7.0 -> 6.0 passes per tile on the bench workload.  warp: max |dv|.
DEFAULT_PF_TOL = {"thread": 1e-5, "warp": 1e-9}
# thread: the first four passes of the fixed point run in fp32 (error after four passes ~1e-5 relative either
# way, fp32 rounding noise ~2e-7), the fp64 passes that follow converge on the same tolerance -- same pass count
# and the same accuracy against Newton as an all-fp64 solve (tests hold 1e-8 on the reference's vectors)
DEFAULT_PF_F32_PASSES = {"thread": 5, "warp": 0}


def convert(dictionary):
    """Same helper as flexibility_provision_env.py:14-15."""
    return namedtuple('GenericDict', dictionary.keys())(**dictionary)


def normalize_args(kwargs):
    """Accept a dict or a namedtuple (reference :37-40) and fill in defaults."""
    if kwargs is None:
        kwargs = {}
    if not isinstance(kwargs, dict):
        kwargs = kwargs._asdict()
    a = dict(DEFAULT_ENV_ARGS)
    a.update(DEFAULT_SOLVER_ARGS)
    a.update(kwargs)
    return a


def load_env_yaml(path):
    """Read the reference's env YAML (the `env_args` block)."""
    import yaml
    with open(path, "r") as f:
        return yaml.safe_load(f)["env_args"]


def make_fp_config(args, network):
    """Build the FpConfig for `network` (see network.Network)."""
    if list(args["buildings"]) != list(args["pv_nodes"]) or list(args["buildings"]) != list(args["ess_nodes"]):
        # the reference indexes PV/ESS dicts with building ids (quirk Q8) -- it only works
        # when the three lists coincide, so anything else is rejected instead of guessed.
        raise ValueError("buildings, pv_nodes and ess_nodes must be identical lists")
    c = _lib.FpConfig()
    nb = network.n_bus
    na = len(args["buildings"])
    if nb > _lib.FP_MAX_BUS:
        raise ValueError(f"at most {_lib.FP_MAX_BUS} buses (one lane per line)")
    if na > _lib.FP_MAX_AGENTS:
        raise ValueError(f"at most {_lib.FP_MAX_AGENTS} buildings")
    c.n_bus, c.n_agents = nb, na
    c.history = int(args["history"])
    c.episode_limit = int(args["episode_limit"])
    c.raw_actions = 1 if args.get("alg", None) == "safemaddpg" else 0
    c.pf_max_iter = int(args["pf_max_iter"])
    variant = args.get("kernel_variant", "thread")
    if variant not in _lib.VARIANTS:
        raise ValueError("kernel_variant must be 'thread' or 'warp'")
    c.variant = _lib.VARIANTS[variant]
    c.pf_tol = float(DEFAULT_PF_TOL[variant] if args.get("pf_tol") is None else args["pf_tol"])
    c.pf_f32_passes = int(DEFAULT_PF_F32_PASSES[variant] if args.get("pf_f32_passes") is None else args["pf_f32_passes"])
    c.v_min, c.v_max = float(args["v_min"]), float(args["v_max"])
    c.e_min, c.e_max = float(args["e_min"]), float(args["e_max"])
    c.p_ch_max, c.p_dis_max = float(args["p_ch_max"]), float(args["p_dis_max"])
    c.eta_ch, c.eta_dis = float(args["eta_ch"]), float(args["eta_dis"])
    c.max_power_reduction = float(args["max_power_reduction"])
    c.kappa = tan(acos(args["cos_phi_max"]))      # same libm call as the reference (:623)
    c.pv_cost, c.ess_cost = float(args["pv_cost"]), float(args["ess_cost"])
    c.discomfort_coeff, c.voltage_coeff = float(args["discomfort_coeff"]), float(args["voltage_coeff"])
    c.delta_t = 24 / args["episode_limit"]        # utils/pf.py:23-24
    c.fail_penalty = float(args["fail_penalty"])
    c.e_next_lb = float(args["e_next_lb"])
    for i in range(nb):
        c.parent[i] = int(network.parent[i])
        c.r[i] = float(network.r[i])
        c.x[i] = float(network.x[i])
        c.imax[i] = float(network.imax[i])
    for i, b in enumerate(args["buildings"]):
        c.agent_bus[i] = network.position[b]
    c.action_low, c.action_high = float(args.get("action_low", 0.0)), float(args.get("action_high", 1.0))
    return c
