"""fp64 restatement of the safety layer of SAFEMADDPG (oracle side): madrl/models/safemaddpg.py:143-299.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference solves this QP with Gurobi
(gurobipy==11.0.1 through Pyomo, safemaddpg.py:280-281), which is not installable here, and ships no test or
golden vector for it.  What pins this restatement instead: `solve_reference_qp` poses the problem exactly as the
reference's Pyomo model does -- all 20 action variables, P_net / Q_net of all 33 buses, 66 slack variables, the row-sum
form of the voltage prediction (:266,272), penalty 1000 -- and hands it to a generic solver (scipy trust-constr, an interior-point method);
`project_closed_form`, the algorithm the CUDA kernel implements, must agree with it (tests/test_oracle_safety.py).

  parse_actions        safemaddpg.py:143-174 (scaling, clipping, ESS logic of the env, :621-677)
  solve_reference_qp   safemaddpg.py:176-299 as posed, generic solver
  project_closed_form  the same optimum in closed form

Structure that makes the closed form possible: V_pred of bus n depends only on that bus's own P_net / Q_net, and those
only on the bus's own four action variables, so the QP separates into one 4-variable problem per building (the other
28 buses only contribute a constant slack cost):
    min (x - x0)^2 + (c - c0)^2 + (d - d0)^2 + (g - g0)^2 + 1000 (s_lo + s_up)
    s.t. V = sP (Pd (1 - x) + c - d) + sQ (Qd + g) + b,  v_min - s_lo <= V <= v_max + s_up,  x, c, d, s >= 0.
With a = dV/d(x, c, d, g) = (-sP Pd, sP, -sP, sQ) and the violated side's deficit delta > 0, the KKT conditions give
y(lam) = max(y0 + (lam / 2) a, 0) on the three bounded variables, g(lam) = g0 + (lam / 2) a_g, with the constraint's
multiplier lam in [0, 1000] (1000 = the slack's price): the gain a . (y(lam) - y0) is concave piecewise linear in lam,
and lam* is the smallest lam that closes the deficit, or 1000 if none does (the remainder is slack).
"""
import numpy as np

from . import env_ref

W_SLACK = 1000.0


def parse_actions(env, actions):
    """safemaddpg.py:143-174 on one env: `actions` [5, 4] raw policy outputs -> dicts of applied setpoints."""
    a = np.asarray(actions, dtype=np.float64).reshape(len(env.base_powergrid['buildings']), 4)
    pct, ch, dis, qpv = {}, {}, {}, {}
    for i, b in enumerate(env.base_powergrid['buildings']):
        pct[b] = env.args['max_power_reduction'] * a[i, 0]
        ch[b] = env.args['p_ch_max'] * a[i, 1]
        dis[b] = env.args['p_dis_max'] * a[i, 2]
        qpv[b] = env._scale_and_clip_q_pv(a[i, 3], env.current_pv_power[b])
    pct = env.clip_percentage_reduction(pct)
    ch, dis = env.adjust_ess_actions(ch, dis)
    for k in ch:
        ch[k], dis[k] = env._clip_power_charging_discharging(ch[k], dis[k], env.current_ess_energy[k])
    return pct, ch, dis, qpv


def row_sums(coef, intercept):
    """safemaddpg.py:182-184: W_P = coef[:, :33], W_Q = coef[:, 33:]; :266,272 use their ROW SUMS."""
    coef = np.asarray(coef, dtype=np.float64)
    n = coef.shape[0]
    return coef[:, :n].sum(axis=1), coef[:, n:].sum(axis=1), np.asarray(intercept, dtype=np.float64)


def solve_reference_qp(y0, Pd, Qd, is_building, sP, sQ, b, v_min, v_max, w=W_SLACK):
    """The QP as the reference poses it.  y0 [nb_buildings, 4] proposed (x, c, d, g); Pd, Qd [33] current demands;
    is_building [33] index of the building at each bus or -1.  Returns the optimal [nb_buildings, 4]."""
    from scipy.optimize import minimize
    nbld, nbus = y0.shape[0], len(Pd)
    nv = 4 * nbld + 2 * nbus                                   # actions, slack_lower[33], slack_upper[33]
    # scale the variables so that SLSQP sees O(1) numbers: actions in units of 1e-3, slacks in units of 1e-3
    S = 1e-3

    def vpred(z):
        y = z[:4 * nbld].reshape(nbld, 4) * S
        P = Pd.copy(); Q = Qd.copy()
        for n in range(nbus):
            k = is_building[n]
            if k >= 0:
                P[n] = Pd[n] * (1.0 - y[k, 0]) + y[k, 1] - y[k, 2]
                Q[n] = Qd[n] + y[k, 3]
        return sP * P + sQ * Q + b

    # V_pred is affine in z: V = V(0) + J z  (exact: probe the columns)
    from scipy.optimize import Bounds, LinearConstraint
    v_at0 = vpred(np.zeros(nv))
    J = np.zeros((nbus, nv))
    for j in range(4 * nbld):
        e = np.zeros(nv); e[j] = 1.0
        J[:, j] = vpred(e) - v_at0
    A_lo = J.copy(); A_lo[:, 4 * nbld:4 * nbld + nbus] += S * np.eye(nbus)          # V + s_lo >= v_min
    A_up = -J.copy(); A_up[:, 4 * nbld + nbus:] += S * np.eye(nbus)                  # -V + s_up >= -v_max
    A = np.vstack([A_lo, A_up]) / S
    lb = np.concatenate([v_min - v_at0, v_at0 - v_max]) / S
    lo = np.full(nv, -np.inf); lo[4 * nbld:] = 0.0
    for k in range(nbld):
        lo[4 * k:4 * k + 3] = 0.0
    grad_lin = np.zeros(nv); grad_lin[4 * nbld:] = w * S / S ** 2
    y0s = y0.reshape(-1) / S

    def obj(z):
        d = z[:4 * nbld] - y0s
        return float(d @ d + grad_lin @ z)

    def jac(z):
        g = grad_lin.copy(); g[:4 * nbld] += 2.0 * (z[:4 * nbld] - y0s)
        return g

    def hess(z):
        H = np.zeros((nv, nv)); H[np.arange(4 * nbld), np.arange(4 * nbld)] = 2.0
        return H
    z0 = np.zeros(nv)
    z0[:4 * nbld] = y0s
    v0 = vpred(z0)
    z0[4 * nbld:4 * nbld + nbus] = np.maximum(0.0, v_min - v0) / S + 1e-3
    z0[4 * nbld + nbus:] = np.maximum(0.0, v0 - v_max) / S + 1e-3
    res = minimize(obj, z0, jac=jac, hess=hess, method="trust-constr", bounds=Bounds(lo, np.full(nv, np.inf)),
                   constraints=[LinearConstraint(A, lb, np.full(2 * nbus, np.inf))],
                   options={"maxiter": 3000, "gtol": 1e-10, "xtol": 1e-14, "barrier_tol": 1e-12})
    return res.x[:4 * nbld].reshape(nbld, 4) * S, res


def project_closed_form(y0, Pd, Qd, sP, sQ, b, v_min, v_max, w=W_SLACK):
    """One building: y0 = (x0, c0, d0, g0), scalars Pd, Qd, sP, sQ, b -> (y*, slack, lam*)."""
    y0 = np.asarray(y0, dtype=np.float64)
    a = np.array([-sP * Pd, sP, -sP, sQ])
    V0 = sP * (Pd * (1.0 - y0[0]) + y0[1] - y0[2]) + sQ * (Qd + y0[3]) + b
    if V0 < v_min:
        delta, sgn = v_min - V0, 1.0
    elif V0 > v_max:
        delta, sgn = V0 - v_max, -1.0
    else:
        return y0.copy(), 0.0, 0.0
    a = sgn * a                                                # direction that closes the deficit
    # breakpoints: a bounded variable moving down hits zero at lam_i = 2 y0_i / |a_i|
    brk = sorted(2.0 * y0[i] / -a[i] for i in range(3) if a[i] < 0.0)
    lam, gain = 0.0, 0.0

    def slope(l):                                              # d gain / d lam just above l
        s = a[3] * a[3] / 2.0
        for i in range(3):
            if a[i] > 0.0 or (a[i] < 0.0 and l < 2.0 * y0[i] / -a[i]):
                s += a[i] * a[i] / 2.0
        return s
    for nxt in brk + [w]:
        nxt = min(nxt, w)
        if nxt <= lam:
            continue
        s = slope(lam)
        if s > 0.0 and gain + s * (nxt - lam) >= delta:
            lam = lam + (delta - gain) / s
            gain = delta
            break
        gain += s * (nxt - lam)
        lam = nxt
        if lam >= w:
            break
    y = y0.copy()
    for i in range(3):
        y[i] = max(y0[i] + 0.5 * lam * a[i], 0.0)
    y[3] = y0[3] + 0.5 * lam * a[3]
    return y, max(0.0, delta - gain), lam


def safety_layer(env, actions, coef, intercept):
    """safety_layer_optimization (:176-299) for one env through the closed form: adjusted actions in the reference's
    return layout [x(5) | c(5) | d(5) | g(5)] (:290-296), per-building slack, per-building multiplier."""
    sP, sQ, b = row_sums(coef, intercept)
    pct, ch, dis, qpv = parse_actions(env, actions)
    bus = env.base_powergrid['bus_numbers']
    out = np.zeros((len(env.base_powergrid['buildings']), 4)); slack = np.zeros(len(out)); lam = np.zeros(len(out))
    for k, bld in enumerate(env.base_powergrid['buildings']):
        n = bus.index(bld)
        y, s, l = project_closed_form([pct[bld], ch[bld], dis[bld], qpv[bld]], env.current_active_demand[bld],
                                      env.current_reactive_demand[bld], sP[n], sQ[n], b[n], env.args['v_min'], env.args['v_max'])
        out[k], slack[k], lam[k] = y, s, l
    return out.T.reshape(-1), slack, lam
