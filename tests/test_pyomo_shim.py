"""CPU: the Pyomo stand-in that lets the reference's own utils/pf.py run in the build container
(oracle/pyomo_shim.py, used by tests/golden/make_ref_golden.py).  The model below is written the way
utils/pf.py:13-98 writes its model (same component kinds, same rule shapes) on a 4-bus feeder with a
lateral; the shim must solve it, verify the rules, and reject a model whose rules cannot hold."""
import numpy as np
import pytest

from oracle import pyomo_shim as pyo
from oracle import pf_ref
from oracle.ieee33 import tree_arrays


def _model(extra_loss=0.0):
    buses = [1, 2, 3, 4]
    lines = [(1, 2), (2, 3), (2, 4)]
    R = {(1, 2): 0.01, (2, 3): 0.02, (2, 4): 0.015}
    X = {(1, 2): 0.02, (2, 3): 0.01, (2, 4): 0.03}
    Pn = {1: 0.0, 2: 0.3, 3: 0.5, 4: -0.1}
    Qn = {1: 0.0, 2: 0.1, 3: 0.2, 4: 0.05}
    m = pyo.ConcreteModel()
    m.N = pyo.Set(initialize=buses); m.L = pyo.Set(initialize=lines)
    m.R = pyo.Param(m.L, initialize=R); m.X = pyo.Param(m.L, initialize=X)
    m.Pnet = pyo.Param(m.N, initialize=Pn); m.Qnet = pyo.Param(m.N, initialize=Qn)
    m.Vsqr = pyo.Var(m.N, within=pyo.NonNegativeReals)
    m.Pl = pyo.Var(m.L, within=pyo.Reals); m.Ql = pyo.Var(m.L, within=pyo.Reals)
    m.Isqr = pyo.Var(m.L, within=pyo.NonNegativeReals)
    m.Ps = pyo.Var(m.N, within=pyo.Reals); m.Qs = pyo.Var(m.N, within=pyo.Reals)
    for n in m.N:
        if n == 1:
            m.Vsqr[n].fix(1)
        else:
            m.Ps[n].fix(0); m.Qs[n].fix(0)
    m.obj = pyo.Objective(rule=lambda mm: sum(mm.R[i, j] * mm.Isqr[i, j] for (i, j) in mm.L), sense=pyo.minimize)
    m.active_power_balance = pyo.Constraint(m.N, rule=lambda mm, n: (
        sum(mm.Pl[i, j] for (i, j) in mm.L if j == n)
        - sum(mm.Pl[i, j] + mm.R[i, j] * mm.Isqr[i, j] for (i, j) in mm.L if i == n) + mm.Ps[n] - mm.Pnet[n] == 0))
    m.reactive_power_balance = pyo.Constraint(m.N, rule=lambda mm, n: (
        sum(mm.Ql[i, j] for (i, j) in mm.L if j == n)
        - sum(mm.Ql[i, j] + mm.X[i, j] * mm.Isqr[i, j] for (i, j) in mm.L if i == n) + mm.Qs[n] - mm.Qnet[n] == 0))
    m.current = pyo.Constraint(m.L, rule=lambda mm, i, j: mm.Isqr[i, j] * mm.Vsqr[j] == (mm.Pl[i, j] ** 2 + mm.Ql[i, j] ** 2))
    m.voltage_drop = pyo.Constraint(m.L, rule=lambda mm, i, j: (
        mm.Vsqr[i] - 2 * (mm.R[i, j] * mm.Pl[i, j] + mm.X[i, j] * mm.Ql[i, j])
        - (mm.R[i, j] ** 2 + mm.X[i, j] ** 2) * mm.Isqr[i, j] + extra_loss == mm.Vsqr[j]))
    return m, buses, lines, R, X, Pn, Qn


def test_shim_solves_and_verifies_a_pf_model():
    m, buses, lines, R, X, Pn, Qn = _model()
    res = pyo.SolverFactory('ipopt').solve(m, tee=False)
    assert res.solver.status == pyo.SolverStatus.ok
    assert pyo._NewtonInPlaceOfIpopt.last_max_residual < 1e-12
    net = dict(bus_numbers=buses, line_connections=lines, line_resistances=R, line_reactances=X,
               max_line_currents={l: 0.0 for l in lines}, bus_types={n: int(n == 1) for n in buses})
    sol = pf_ref.solve_sweep(tree_arrays(net), np.array([Pn[n] for n in buses]), np.array([Qn[n] for n in buses]))
    assert np.max(np.abs(np.array([m.Vsqr[n].value for n in buses]) - sol['v'])) < 1e-12
    assert abs(m.Pl[(1, 2)].value - sol['P'][1]) < 1e-12 and abs(m.Isqr[(2, 4)].value - sol['ell'][3]) < 1e-12
    # the slack injection closes its own balance row
    assert abs(m.Ps[1].value - (m.Pl[(1, 2)].value + R[(1, 2)] * m.Isqr[(1, 2)].value)) < 1e-12


def test_shim_rejects_rules_the_solution_does_not_satisfy():
    """A model whose voltage-drop rule differs from the DistFlow equations the Newton solver solves
    must come back as a solver failure (that is what makes the generated fixtures a check of the
    reference's rules and not only of the oracle)."""
    m, *_ = _model(extra_loss=1e-6)
    res = pyo.SolverFactory('ipopt').solve(m)
    assert res.solver.status != pyo.SolverStatus.ok
    assert pyo._NewtonInPlaceOfIpopt.last_max_residual > 5e-7


def test_expression_arithmetic_and_equality_residuals():
    v = pyo.VarData(); v.value = 3.0
    e = 2 * v + 1 - v / 2 + v ** 2
    assert isinstance(e, pyo.Expr) and abs(e.v - (6 + 1 - 1.5 + 9)) < 1e-15
    r = (v * v == 9.5)
    assert isinstance(r, pyo.Residual) and abs(r.value + 0.5) < 1e-15
    assert sum(x for x in [v, v]).v == 6.0
    with pytest.raises(AssertionError):
        pyo.SolverFactory('gurobi')
