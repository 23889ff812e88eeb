"""GPU: the batched safety projection (fp_safety_project) -- SAFEMADDPG.safety_layer_optimization (madrl/models/
safemaddpg.py:176-299) for every env at once -- against the oracle's restatement (oracle/safety_ref.py: parse_actions with
the env's own helper rules + the closed-form optimum, itself checked against the QP as the reference poses it by
tests/test_oracle_safety.py).  Parity with the reference's solver (Gurobi) is UNPINNED: it is not installable here.
Tolerance: the kernel and the oracle evaluate the same fp64 formulas; outputs are fp32 (the reference casts the solver's
result to float32, :115), so 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import env_ref, ieee33, safety_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_predictor.npz")


def _stub(args, env, e):
    """A RefFlexEnv-shaped object carrying GPU env e's current values (what parse_actions / the QP read)."""
    st = env.get_state(torch.float64)[e].cpu().numpy()             # [P(33), Q(33), Ppv(5), V(33), price, E(5)]
    ref = env_ref.RefFlexEnv.__new__(env_ref.RefFlexEnv)
    ref.args = env_ref._Args(dict(env_ref.DEFAULT_ARGS)); ref.args.update(args)
    ref.base_powergrid = ieee33.create_network()
    G = ref.base_powergrid
    ref.current_active_demand = {b: float(st[i]) for i, b in enumerate(G['bus_numbers'])}
    ref.current_reactive_demand = {b: float(st[33 + i]) for i, b in enumerate(G['bus_numbers'])}
    ref.current_pv_power = {b: float(st[66 + i]) for i, b in enumerate(G['PVs_at_buildings'])}
    ref.current_ess_energy = {b: float(st[66 + 5 + 33 + 1 + i]) for i, b in enumerate(G['ESSs_at_buildings'])}
    return ref


@pytest.mark.parametrize("model", ["reference_fit", "synthetic"])
def test_safety_projection_matches_oracle(cuda, profiles, args, model):
    from flexgpu import BatchedFlexProvisionEnv
    n = 3000
    env = BatchedFlexProvisionEnv(None, n_envs=n, device=cuda, profiles=profiles, seed=21)
    env.reset(return_obs=False)
    g = torch.Generator(device=cuda).manual_seed(3)
    for _ in range(3):
        env.step(torch.rand(n, 5, 4, device=cuda, generator=g), want_info=False)
    rng = np.random.default_rng(4)
    if model == "reference_fit":                 # the regressor the reference's own training script fits (scaled space, as it loads it)
        gp = np.load(GOLD)
        coef, icpt = gp["coef"].astype(np.float64), gp["intercept"].astype(np.float64)
    else:                                        # a mix of in-limit, recoverable and hopeless predictions
        coef = rng.normal(0, 0.02, (33, 66)) - 0.004
        icpt = 1.0 + rng.uniform(-0.12, 0.12, 33)
    env.load_safety_model(coef, icpt)
    actions = (torch.randn(n, 5, 4, device=cuda, generator=g) * 0.6 + 0.4).float()       # raw policy outputs, some outside [0, 1]
    adj, slack, moved = env.safety_project(actions, want_info=True)
    adj_agent = env.safety_project(actions, layout="agent")
    assert adj.shape == (n, 20) and adj_agent.shape == (n, 5, 4)
    assert torch.equal(adj.view(n, 4, 5).permute(0, 2, 1), adj_agent)                     # the two layouts hold the same numbers
    a_np, adj_np, slack_np, moved_np = actions.cpu().numpy(), adj.cpu().numpy(), slack.cpu().numpy(), moved.cpu().numpy()
    n_moved = 0
    for e in list(range(0, n, 37)) + [n - 1]:
        ref = _stub(args, env, e)
        want, wslack, wlam = safety_ref.safety_layer(ref, a_np[e].astype(np.float64), coef, icpt)
        assert np.max(np.abs(adj_np[e] - want) / np.maximum(1.0, np.abs(want))) < 1e-6, (e, adj_np[e], want)
        assert np.allclose(slack_np[e], wslack, rtol=1e-9, atol=1e-12)
        assert np.array_equal(moved_np[e], wlam > 0)
        n_moved += int((wlam > 0).sum())
    assert n_moved > 0
    if model == "synthetic":
        frac = float(moved.float().mean())
        assert 0.05 < frac < 0.95, frac                     # the case mix really exercises both branches
        assert bool((adj.view(n, 4, 5)[:, :3] >= 0).all())  # x, c, d stay non-negative (their Pyomo domains)
    env.close()
