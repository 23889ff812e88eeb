"""GPU parity on a feeder that is NOT the IEEE 33-bus case: the run-time-table instantiation of the
thread kernels (any radial tree with <= 33 buses) and the warp kernels, against the C mirror built
from the same tree.  The feeder has two lines leaving the slack bus, two laterals on one bus and a
nested lateral -- every chain/slot case of the one-pass sweep that the IEEE 33-bus shape lacks."""
import numpy as np
import pytest
import torch

from oracle import c_mirror, ieee33, pf_ref

pytestmark = pytest.mark.gpu

#        1 ---- 2 ---- 3 ---- 4 ---- 5
#        |      |\            |
#        |      | 8 -- 9      6 -- 7
#        |      10             \
#        |                      13
#        11 -- 12 -- 14
LINES = [(1, 2, .30, .20), (2, 3, .45, .25), (3, 4, .40, .30), (4, 5, .80, .60), (4, 6, .50, .35), (6, 7, .90, .70),
         (2, 8, .35, .15), (8, 9, .60, .45), (2, 10, .55, .40), (1, 11, .25, .12), (11, 12, .70, .50), (6, 13, .65, .20),
         (12, 14, .95, .85)]
LOADS = {2: (120, 60), 3: (90, 40), 4: (200, 90), 5: (150, 70), 6: (60, 30), 7: (210, 100), 8: (100, 50), 9: (80, 30),
         10: (140, 60), 11: (60, 20), 12: (300, 150), 13: (75, 35), 14: (180, 80)}
BUILDINGS = [5, 9, 14]


def custom_network(args):
    zbase = args["v_nom"] ** 2 * 1000 / args["s_nom"]
    ibase = args["s_nom"] / args["v_nom"]
    nodes = list(range(1, 15))
    return {
        'bus_numbers': nodes,
        'line_connections': [(f, t) for (f, t, _, _) in LINES],
        'line_resistances': {(f, t): r / zbase for (f, t, r, _) in LINES},
        'line_reactances': {(f, t): x / zbase for (f, t, _, x) in LINES},
        'max_line_currents': {(f, t): 150.0 / ibase for (f, t, _, _) in LINES},
        'bus_types': {n: (1 if n == 1 else 0) for n in nodes},
        'active_power_demand': {n: LOADS.get(n, (0, 0))[0] / args["s_nom"] for n in nodes},
        'reactive_power_demand': {n: LOADS.get(n, (0, 0))[1] / args["s_nom"] for n in nodes},
        'buildings': BUILDINGS, 'PVs_at_buildings': BUILDINGS, 'ESSs_at_buildings': BUILDINGS,
    }


@pytest.fixture(scope="module", params=["thread", "warp"])
def setup(request, cuda, args):
    from flexgpu import BatchedFlexProvisionEnv
    from flexgpu.network import Network
    from flexgpu.profiles import synthetic_profiles
    a = dict(args); a.update(buildings=BUILDINGS, pv_nodes=BUILDINGS, ess_nodes=BUILDINGS, kernel_variant=request.param)
    net = custom_network(a)
    prof = synthetic_profiles(Network(net), len(BUILDINGS), T=1500, seed=3)
    tree = ieee33.tree_arrays(net)
    fonet = c_mirror.make_net(tree, a, BUILDINGS,
                              variant=c_mirror.VARIANT_WARP if request.param == "warp" else c_mirror.VARIANT_THREAD)
    n = 200
    env = BatchedFlexProvisionEnv(a, n_envs=n, device=cuda, network=net, profiles=prof)
    yield env, fonet, prof, tree, Network(net), n
    env.close()


def test_power_flow_bit_exact_vs_mirror_and_close_to_newton(setup):
    env, fonet, prof, tree, network, n = setup
    rng = np.random.default_rng(1)
    p = network.base_p[None, 1:] * rng.uniform(0.5, 1.6, (300, network.n_bus - 1))
    q = network.base_q[None, 1:] * rng.uniform(0.5, 1.6, (300, network.n_bus - 1))
    out = {k: (v.cpu().numpy() if v is not None else None) for k, v in env.power_flow(p, q).items()}
    ref = c_mirror.mirror_power_flow(fonet, p, q)
    for k in ("V", "P", "Q", "Isq"):
        assert np.array_equal(out[k], ref[k]), k
    assert np.array_equal(out["iters"], ref["iters"]) and not out["failed"].any()
    for i in (0, 17, 299):                                         # independent dense Newton on pf.py:65-98
        sol = pf_ref.solve_newton(tree, np.concatenate([[0.0], p[i]]), np.concatenate([[0.0], q[i]]))
        assert np.max(np.abs(out["V"][i] - np.sqrt(sol['v']))) < 1e-8
        assert np.max(np.abs(out["P"][i] - sol['P'][1:])) < 1e-7


def test_env_steps_bit_exact_vs_mirror(setup):
    env, fonet, prof, tree, network, n = setup
    mb = c_mirror.MirrorBatch(fonet, prof.as_dict(), n)
    rng = np.random.default_rng(2)
    na = len(BUILDINGS)
    start = rng.integers(0, env.max_start(), n).astype(np.int32)
    e0 = rng.uniform(0.9 * 0.0125, 1.1 * 0.0125, (n, na)); a0 = rng.uniform(0, 1, (n, 4 * na))
    env.reset(start, e0, a0, return_obs=False); mb.reset(start, e0, a0)
    for t in range(6):
        a = rng.normal(0.4, 0.5, (n, 4 * na))
        r, d, info = env.step(torch.from_numpy(a), want_info=True)
        rr, dd, ii = mb.step(a)
        assert np.array_equal(r.cpu().numpy(), rr) and np.array_equal(d.cpu().numpy(), dd)
        assert np.array_equal(info["voltage_penalty"].cpu().numpy(), ii[:, 5])
    assert np.array_equal(env.voltages.cpu().numpy(), mb.V)
    assert np.array_equal(env.ess_energy.cpu().numpy(), mb.E_cur)
    assert np.array_equal(env.violation_mask.cpu().numpy().view(np.uint64), mb.vmask)
    assert np.array_equal(env.pf_iterations.cpu().numpy(), mb.iters)
    obs = env.get_obs(); state = env.get_state()
    assert obs.shape == (n, na, 6 * env.history) and state.shape == (n, 3 * 14 + 2 * na + 1)
    # step(return_obs=True) on a run-time-table feeder (fewer agents: packed observation row, two-launch
    # fallback of fp_step_obs) == step + get_obs, and the window ring agrees with the fp64 history ring
    a = rng.normal(0.4, 0.5, (n, 4 * na))
    want = env.get_obs(push=False, dtype=torch.float64)            # what the next push must NOT yet contain
    r, d, info, o2 = env.step(torch.from_numpy(a), return_obs=True)
    rr, dd, ii = mb.step(a)
    assert np.array_equal(r.cpu().numpy(), rr)
    cur = mb.current_obs()                                         # [n, na, 6] after the step
    assert np.array_equal(o2[:, :, -6:].cpu().numpy(), cur.astype(np.float32))
    assert torch.equal(o2[:, :, :-6], obs[:, :, 6:])               # shifted by one entry
    assert torch.equal(env.get_obs(push=False, dtype=torch.float64)[:, :, :-6].float(), o2[:, :, 6:])
