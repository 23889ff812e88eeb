"""CPU: the closed-form safety projection (what the CUDA kernel implements) against the QP posed as the reference's
Pyomo model poses it (madrl/models/safemaddpg.py:176-299), solved by a generic solver (scipy trust-constr).  Gurobi is absent:
parity with the reference's solver output is unpinned (oracle/safety_ref.py says so)."""
import numpy as np
import pytest

from oracle import safety_ref


def _case(rng, kind):
    nbus, blds = 33, [4, 9, 14, 19, 24]                         # bus positions of buildings 5, 10, 15, 20, 25
    is_b = -np.ones(nbus, dtype=int)
    for k, n in enumerate(blds):
        is_b[n] = k
    Pd = rng.uniform(0.02, 0.4, nbus); Qd = rng.uniform(0.01, 0.2, nbus)
    sP = rng.uniform(-0.9, -0.05, nbus); sQ = rng.uniform(-0.6, -0.02, nbus)       # more load -> lower voltage
    y0 = np.stack([rng.uniform(0, 0.5, 5), rng.uniform(0, 0.005, 5), rng.uniform(0, 0.005, 5), rng.uniform(-0.02, 0.02, 5)], axis=1)
    y0[rng.random(5) < 0.5, 1] = 0.0; y0[y0[:, 1] > 0, 2] = 0.0              # never both charging and discharging
    P0, Q0 = Pd.copy(), Qd.copy()
    for k, n in enumerate(blds):
        P0[n] = Pd[n] * (1 - y0[k, 0]) + y0[k, 1] - y0[k, 2]; Q0[n] = Qd[n] + y0[k, 3]
    V_free = sP * P0 + sQ * Q0                                            # prediction at the proposed actions, before the intercept
    if kind == "inside":
        b = 1.0 - V_free
    elif kind == "low":
        b = 0.9 - V_free - rng.uniform(0.0, 0.02, nbus)                 # a little below v_min: recoverable
    elif kind == "very_low":
        b = 0.9 - V_free - rng.uniform(5.0, 400.0, nbus)                # hopeless: the multiplier saturates at 1000
    elif kind == "high":
        b = 1.1 - V_free + rng.uniform(0.0, 0.02, nbus)
    else:
        b = 1.0 - V_free + rng.uniform(-0.25, 0.25, nbus)
    return y0, Pd, Qd, is_b, sP, sQ, b, blds


@pytest.mark.parametrize("kind", ["inside", "low", "very_low", "high", "mixed"])
def test_closed_form_matches_the_posed_qp(kind):
    rng = np.random.default_rng({"inside": 1, "low": 2, "very_low": 3, "high": 4, "mixed": 5}[kind])
    for trial in range(1):                                       # the interior-point solve takes ~7 s per case
        y0, Pd, Qd, is_b, sP, sQ, b, blds = _case(rng, kind)
        want, res = safety_ref.solve_reference_qp(y0, Pd, Qd, is_b, sP, sQ, b, 0.9, 1.1)
        got = np.stack([safety_ref.project_closed_form(y0[k], Pd[n], Qd[n], sP[n], sQ[n], b[n], 0.9, 1.1)[0] for k, n in enumerate(blds)])

        def cost(y):
            c = float(((y - y0) ** 2).sum())
            for k, n in enumerate(blds):
                V = sP[n] * (Pd[n] * (1 - y[k, 0]) + y[k, 1] - y[k, 2]) + sQ[n] * (Qd[n] + y[k, 3]) + b[n]
                c += 1000.0 * (max(0.0, 0.9 - V) + max(0.0, V - 1.1))
            return c
        # the closed form is at least as good as the generic solver's point and (the QP is strictly convex in y) close to it
        assert cost(got) <= cost(want) * (1 + 1e-7) + 1e-9, (kind, trial, cost(got), cost(want))
        assert (got[:, :3] >= 0).all()
        if kind == "inside":
            assert np.array_equal(got, y0)
            continue
        scale = max(1.0, float(np.abs(want - y0).max()))
        assert np.max(np.abs(got - want)) < 1e-6 * scale, (kind, trial, np.max(np.abs(got - want)))


def test_closed_form_kkt_on_random_cases():
    """First-order optimality without any solver: no feasible perturbation lowers the exact-penalty cost."""
    rng = np.random.default_rng(9)
    for _ in range(200):
        y0 = np.array([rng.uniform(0, 0.5), rng.uniform(0, 0.005) * (rng.random() < 0.5), 0.0, rng.uniform(-0.02, 0.02)])
        if y0[1] == 0.0:
            y0[2] = rng.uniform(0, 0.005)
        Pd, Qd, sP, sQ = rng.uniform(0.02, 0.4), rng.uniform(0.01, 0.2), rng.uniform(-0.9, -0.05), rng.uniform(-0.6, -0.02)
        b = 1.0 - sP * Pd - sQ * Qd + rng.choice([-1, 1]) * rng.uniform(0.1, 0.4)
        y, slack, lam = safety_ref.project_closed_form(y0, Pd, Qd, sP, sQ, b, 0.9, 1.1)

        def cost(z):
            V = sP * (Pd * (1 - z[0]) + z[1] - z[2]) + sQ * (Qd + z[3]) + b
            return float(((z - y0) ** 2).sum()) + 1000.0 * (max(0.0, 0.9 - V) + max(0.0, V - 1.1))
        c0 = cost(y)
        for _ in range(40):
            z = y + rng.normal(0, 1e-4, 4) * rng.choice([1.0, 10.0, 100.0])
            z[:3] = np.maximum(z[:3], 0.0)
            assert cost(z) >= c0 - 1e-12 * max(1.0, c0)
        assert 0.0 <= lam <= 1000.0 and slack >= 0.0
