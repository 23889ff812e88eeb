"""fp64 CPU power flow: two independent solvers of the reference's DistFlow system.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY: pinned on the reference's own model code through
oracle/pyomo_shim.py (tests/golden/ref_pf.npz); IPOPT itself is absent.

The reference states the power flow as a square nonlinear system inside a Pyomo
model (utils/pf.py:65-98) and lets IPOPT find a root; the objective
(utils/pf.py:59-62) is inert because the feasible set is a set of isolated roots.
Here the same equations are solved two ways:

  * `solve_newton`  -- dense Newton on the 4*(nb-1) unknowns (P, Q, l, v) from a
                       flat start; residual <= 1e-13.  Independent of any sweep.
  * `solve_sweep`   -- textbook backward/forward sweep, sequential sums in bus
                       order (NOT the kernel's scan order).

`power_flow_solver` / `power_flow_solver_simplified` wrap them behind the
reference's signatures (utils/pf.py:10 and :115) and return the same dict layout
(utils/pf.py:108-113, :187-192).
"""
import numpy as np

from .ieee33 import tree_arrays

ETA_CH = 0.9       # flex_provision.yaml:24
ETA_DIS = 0.9      # flex_provision.yaml:25
EPISODE_LIMIT = 96  # flex_provision.yaml:11


class SolverFailure(Exception):
    """Mirrors `raise Exception('Solver failed to find a solution')` (utils/pf.py:104-105)."""


def residuals(tree, p, q, P, Q, ell, v):
    """Residuals of utils/pf.py:65-94 at (P, Q, l, v); arrays indexed by bus position.

    p, q are net consumptions per bus (slack entry ignored), P/Q/ell are indexed by
    the child bus of each line, v is squared voltage (v[slack] == 1).
    Returns (rP, rQ, rI, rV), each of length nb (slack entries are 0).
    """
    parent, R, X = tree['parent'], tree['R'], tree['X']
    nb = len(parent)
    rP = np.zeros(nb)
    rQ = np.zeros(nb)
    rI = np.zeros(nb)
    rV = np.zeros(nb)
    for j in range(nb):
        if parent[j] < 0:
            continue
        # pf.py:65-83 at bus j: inflow - sum(outflow + loss) - net consumption == 0
        rP[j] += P[j] - p[j]
        rQ[j] += Q[j] - q[j]
        i = parent[j]
        if parent[i] >= 0:
            rP[i] -= P[j] + R[j] * ell[j]
            rQ[i] -= Q[j] + X[j] * ell[j]
        rI[j] = ell[j] * v[j] - (P[j] ** 2 + Q[j] ** 2)                       # pf.py:85-88
        rV[j] = v[i] - 2 * (R[j] * P[j] + X[j] * Q[j]) - (R[j] ** 2 + X[j] ** 2) * ell[j] - v[j]  # pf.py:90-94
    return rP, rQ, rI, rV


def solve_newton(tree, p, q, tol=1e-13, max_iter=30):
    """Dense Newton on F(P,Q,l,v)=0 from a flat start.  Returns dict of arrays (bus-indexed)."""
    parent, R, X = tree['parent'], tree['R'], tree['X']
    nb = len(parent)
    ns = [j for j in range(nb) if parent[j] >= 0]
    m = len(ns)
    col = {j: k for k, j in enumerate(ns)}
    P = np.zeros(nb)
    Q = np.zeros(nb)
    ell = np.zeros(nb)
    v = np.ones(nb)
    # flat start: lossless flows
    for j in tree['order'][::-1]:
        if parent[j] < 0:
            continue
        P[j] += p[j]
        Q[j] += q[j]
        i = parent[j]
        if parent[i] >= 0:
            P[i] += P[j]
            Q[i] += Q[j]
    for it in range(max_iter):
        rP, rQ, rI, rV = residuals(tree, p, q, P, Q, ell, v)
        F = np.concatenate([rP[ns], rQ[ns], rI[ns], rV[ns]])
        if np.max(np.abs(F)) <= tol:
            break
        J = np.zeros((4 * m, 4 * m))
        oP, oQ, oL, oV = 0, m, 2 * m, 3 * m
        for j in ns:
            k = col[j]
            i = parent[j]
            # rP[j] += P[j]; rP[i] -= P[j] + R[j] ell[j]
            J[oP + k, oP + k] += 1.0
            J[oQ + k, oQ + k] += 1.0
            if parent[i] >= 0:
                ki = col[i]
                J[oP + ki, oP + k] -= 1.0
                J[oP + ki, oL + k] -= R[j]
                J[oQ + ki, oQ + k] -= 1.0
                J[oQ + ki, oL + k] -= X[j]
            # rI[j] = ell v - P^2 - Q^2
            J[oL + k, oL + k] = v[j]
            J[oL + k, oV + k] = ell[j]
            J[oL + k, oP + k] = -2 * P[j]
            J[oL + k, oQ + k] = -2 * Q[j]
            # rV[j] = v[i] - 2(RP+XQ) - z2 ell - v[j]
            J[oV + k, oP + k] = -2 * R[j]
            J[oV + k, oQ + k] = -2 * X[j]
            J[oV + k, oL + k] = -(R[j] ** 2 + X[j] ** 2)
            J[oV + k, oV + k] = -1.0
            if parent[i] >= 0:
                J[oV + k, oV + col[i]] = 1.0
        try:
            dx = np.linalg.solve(J, -F)
        except np.linalg.LinAlgError as e:
            raise SolverFailure(str(e))
        P[ns] += dx[oP:oP + m]
        Q[ns] += dx[oQ:oQ + m]
        ell[ns] += dx[oL:oL + m]
        v[ns] += dx[oV:oV + m]
        if not np.all(np.isfinite(v)) or np.any(v[ns] <= 0):
            raise SolverFailure('voltage collapsed')
    else:
        raise SolverFailure('Newton did not converge')
    return dict(P=P, Q=Q, ell=ell, v=v, iters=it)


def solve_sweep(tree, p, q, tol=1e-13, max_iter=100):
    """Plain backward/forward sweep, sequential sums (bus order).  Independent of the kernel."""
    parent, R, X = tree['parent'], tree['R'], tree['X']
    nb = len(parent)
    order = tree['order']
    ell = np.zeros(nb)
    v = np.ones(nb)
    for it in range(1, max_iter + 1):
        P = np.zeros(nb)
        Q = np.zeros(nb)
        for j in order[::-1]:            # children before parents
            if parent[j] < 0:
                continue
            P[j] += p[j]
            Q[j] += q[j]
            i = parent[j]
            if parent[i] >= 0:
                P[i] += P[j] + R[j] * ell[j]
                Q[i] += Q[j] + X[j] * ell[j]
        v_new = np.ones(nb)
        for j in order:                  # parents before children
            if parent[j] < 0:
                continue
            v_new[j] = v_new[parent[j]] - 2 * (R[j] * P[j] + X[j] * Q[j]) - (R[j] ** 2 + X[j] ** 2) * ell[j]
        if not np.all(np.isfinite(v_new)) or np.any(v_new <= 0):
            raise SolverFailure('voltage collapsed')
        ell_new = np.where(parent >= 0, (P ** 2 + Q ** 2) / v_new, 0.0)
        dv = np.max(np.abs(v_new - v))
        dl = np.max(np.abs(ell_new - ell))
        v, ell = v_new, ell_new
        if dv <= tol and dl <= tol:
            return dict(P=P, Q=Q, ell=ell, v=v, iters=it)
    raise SolverFailure('sweep did not converge')


def _pack(net, tree, sol, with_ess=None):
    buses = tree['buses']
    pos = tree['pos']
    voltages = {n: float(np.sqrt(sol['v'][pos[n]])) for n in buses}            # pf.py:108
    currents = {}
    flows = {}
    for (f, t) in net['line_connections']:
        # the line's child is whichever end has the other as parent
        j = pos[t] if tree['parent'][pos[t]] == pos[f] else pos[f]
        currents[(f, t)] = float(np.sqrt(sol['ell'][j]))                        # pf.py:109
        flows[(f, t)] = (float(sol['P'][j]), float(sol['Q'][j]))                # pf.py:110
    out = {'Voltages': voltages, 'Currents': currents, 'Power Flows': flows}
    if with_ess is not None:
        out['Next ESS Energy'] = with_ess
    return out


def net_injections(net, Pload, Qload, Pred, Ppv, Qpv, Pesc, Pesd):
    """Net consumption per bus implied by the balance rows utils/pf.py:65-83 (bus order)."""
    buses = net['bus_numbers']
    p = np.zeros(len(buses))
    q = np.zeros(len(buses))
    for i, n in enumerate(buses):
        pn = Pload[n]
        if n in net['buildings']:
            pn = pn - Pred[n]
        if n in net['PVs_at_buildings']:
            pn = pn - Ppv[n]
        if n in net['ESSs_at_buildings']:
            pn = pn + Pesc[n]
            pn = pn - Pesd[n]
        qn = Qload[n]
        if n in net['PVs_at_buildings']:
            qn = qn - Qpv[n]
        p[i] = pn
        q[i] = qn
    return p, q


def power_flow_solver(net, active_power_demand, reactive_power_demand, power_reduction,
                      pv_active_power, pv_reactive_power, ess_charging, ess_discharging,
                      initial_ess_energy, method='newton', tree=None):
    """Same signature and return layout as utils/pf.py:10-113."""
    tree = tree or tree_arrays(net)
    p, q = net_injections(net, active_power_demand, reactive_power_demand, power_reduction,
                          pv_active_power, pv_reactive_power, ess_charging, ess_discharging)
    sol = (solve_newton if method == 'newton' else solve_sweep)(tree, p, q)
    dt = 24 / EPISODE_LIMIT                                                      # pf.py:23-24
    e_next = {}
    for k in net['ESSs_at_buildings']:
        e = initial_ess_energy[k] + dt * (ETA_CH * ess_charging[k] - (1 / ETA_DIS) * ess_discharging[k])  # pf.py:96-98
        # E_next is a NonNegativeReals variable (pf.py:46); IPOPT relaxes bounds by 1e-8.
        if e < -1e-8:
            raise SolverFailure('E_next below its lower bound')
        e_next[k] = float(e)
    return _pack(net, tree, sol, with_ess=e_next)


def power_flow_solver_simplified(net, P_net, Q_net, method='newton', tree=None):
    """Same signature and return layout as utils/pf.py:115-192."""
    tree = tree or tree_arrays(net)
    p = np.array([P_net[n] for n in net['bus_numbers']], dtype=np.float64)
    q = np.array([Q_net[n] for n in net['bus_numbers']], dtype=np.float64)
    sol = (solve_newton if method == 'newton' else solve_sweep)(tree, p, q)
    return _pack(net, tree, sol)
