"""CPU: pins the power-flow oracle (no GPU).  Known answers K1-K3 of BASELINE.md section 5,
Newton vs sweep vs the C mirror, and the committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import c_mirror, ieee33, pf_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def base_loads(net):
    p = np.array([net['active_power_demand'][n] for n in net['bus_numbers']])
    q = np.array([net['reactive_power_demand'][n] for n in net['bus_numbers']])
    return p, q


def test_per_unit_conversion(net):
    # utils/create_net.py:22-24 with v_nom = 12.66 kV, s_nom = 1000 kVA
    zbase = 12.66 ** 2 * 1000 / 1000
    assert abs(zbase - 160.2756) < 1e-10
    assert net['line_resistances'][(1, 2)] == 0.0922 / zbase
    assert net['line_reactances'][(32, 33)] == 0.5302 / zbase
    assert net['active_power_demand'][25] == 0.42 and net['reactive_power_demand'][30] == 0.6
    assert abs(sum(net['active_power_demand'].values()) - 3.715) < 1e-12      # 3715 kW
    assert abs(sum(net['reactive_power_demand'].values()) - 2.300) < 1e-12    # 2300 kvar
    assert len(net['line_connections']) == 32 and net['bus_types'][1] == 1


@pytest.mark.parametrize("solver", ["newton", "sweep"])
def test_K1_ieee33_base_case(net, tree, solver):
    """Literature: V_min = 0.9131 p.u. @ bus 18, losses 202.68 kW / 135.14 kvar (Baran & Wu)."""
    p, q = base_loads(net)
    sol = (pf_ref.solve_newton if solver == "newton" else pf_ref.solve_sweep)(tree, p, q)
    V = np.sqrt(sol['v'])
    assert int(np.argmin(V)) + 1 == 18
    assert abs(V.min() - 0.913090) < 5e-7
    assert abs((tree['R'] * sol['ell']).sum() * 1000 - 202.6771) < 1e-3
    assert abs((tree['X'] * sol['ell']).sum() * 1000 - 135.1410) < 1e-3


def test_K2_run_pf_operating_point(net):
    """run_pf.py:36-57: code-defined operating point through the reference-shaped API."""
    B = net['buildings']
    Pl = {n: (0 if net['bus_types'][n] == 1 else 0.1) for n in net['bus_numbers']}
    Ql = {n: (0 if net['bus_types'][n] == 1 else 0.005) for n in net['bus_numbers']}
    res = pf_ref.power_flow_solver(net, Pl, Ql, {n: 0.1 * 0.5 for n in B}, {n: 0.5 * 0.15 for n in B},
                                   {n: 0 for n in B}, {n: 0.005 for n in B}, {n: 0 for n in B},
                                   {n: 0.025 / 2 for n in B})
    assert abs(res['Voltages'][18] - 0.94137528) < 1e-8
    assert abs(res['Voltages'][33] - 0.95744575) < 1e-8
    loss = sum(net['line_resistances'][k] * res['Currents'][k] ** 2 for k in res['Currents'])
    assert abs(loss - 0.08366532) < 1e-8
    for k in B:
        assert abs(res['Next ESS Energy'][k] - 0.013625) < 1e-15            # 0.0125 + 0.25*0.9*0.005
    assert set(res) == {'Voltages', 'Currents', 'Power Flows', 'Next ESS Energy'}  # utils/pf.py:113
    assert set(res['Currents']) == set(net['line_connections'])


def test_K3_residuals_and_solver_agreement(net, tree):
    rng = np.random.RandomState(3)
    p0, q0 = base_loads(net)
    for _ in range(10):
        p = p0 * rng.uniform(0.2, 1.6, 33)
        q = q0 * rng.uniform(0.2, 1.6, 33)
        a = pf_ref.solve_newton(tree, p, q)
        b = pf_ref.solve_sweep(tree, p, q)
        for key in ('P', 'Q', 'v'):
            assert np.max(np.abs(a[key] - b[key])) < 1e-10
        assert np.max(np.abs(a['ell'] - b['ell']) / np.maximum(1.0, a['ell'])) < 1e-10
        for r in pf_ref.residuals(tree, p, q, a['P'], a['Q'], a['ell'], a['v']):
            assert np.max(np.abs(r)) < 1e-12                                  # utils/pf.py:65-94


def test_c_mirror_matches_newton_within_tolerance(fonet, net, tree):
    """The C mirror (kernel op order, pf_tol 1e-9) vs dense Newton: 1e-6 p.u. is the parity bar."""
    rng = np.random.RandomState(11)
    p0, q0 = base_loads(net)
    n = 200
    p = p0[None, 1:] * rng.uniform(0.3, 1.3, (n, 32))
    q = q0[None, 1:] * rng.uniform(0.3, 1.3, (n, 32))
    m = c_mirror.mirror_power_flow(fonet, p, q)
    assert not m['failed'].any() and m['iters'].max() <= 8
    for e in range(0, n, 10):
        a = pf_ref.solve_newton(tree, np.concatenate(([0.0], p[e])), np.concatenate(([0.0], q[e])))
        assert np.max(np.abs(np.sqrt(a['v']) - m['V'][e])) < 1e-9
        assert np.max(np.abs(a['P'][1:] - m['P'][e])) < 1e-8
        assert np.max(np.abs(a['Q'][1:] - m['Q'][e])) < 1e-8
        assert np.max(np.abs(np.sqrt(a['ell'][1:]) - np.sqrt(m['Isq'][e]))) < 1e-7      # currents: 10x inside the 1e-6 bar


def test_golden_pf_vectors(fonet, tree):
    g = np.load(os.path.join(GOLD, "pf_golden.npz"))
    m = c_mirror.mirror_power_flow(fonet, g['p'], g['q'])
    assert np.max(np.abs(m['V'] - g['V'])) < 1e-9
    assert np.max(np.abs(m['P'] - g['P'])) < 1e-8 and np.max(np.abs(m['Q'] - g['Q'])) < 1e-8
    for e in (0, 5, 40):                      # the plain sweep agrees with the stored Newton answers too
        b = pf_ref.solve_sweep(tree, np.concatenate(([0.0], g['p'][e])), np.concatenate(([0.0], g['q'][e])))
        assert np.max(np.abs(np.sqrt(b['v']) - g['V'][e])) < 1e-11
    # K1 is row 0 of the fixture
    assert abs(g['V'][0].min() - 0.913090) < 5e-7


def test_mirror_reports_failure_on_collapse(fonet, net):
    """Loads far beyond the feeder's capability -> no high-voltage root -> failure flag."""
    p0, q0 = base_loads(net)
    m = c_mirror.mirror_power_flow(fonet, 8.0 * p0[None, 1:], 8.0 * q0[None, 1:])
    assert m['failed'][0]
    with pytest.raises(pf_ref.SolverFailure):
        pf_ref.solve_sweep(ieee33.tree_arrays(net), 8.0 * p0, 8.0 * q0)


def test_mirror_zero_and_negative_load(fonet):
    m = c_mirror.mirror_power_flow(fonet, np.zeros((1, 32)), np.zeros((1, 32)))
    assert np.all(m['V'] == 1.0) and m['iters'][0] == 1 + fonet.pf_f32 and not m['failed'][0]    # opening fp32 passes + one fp64 pass
    # reverse power flow (generation) raises voltages above 1
    m = c_mirror.mirror_power_flow(fonet, -0.05 * np.ones((1, 32)), np.zeros((1, 32)))
    assert m['V'][0, 1:].min() > 1.0 and not m['failed'][0]


# ------------------------------------------------------------------------------------------------
# Fixtures produced by the REFERENCE'S OWN code (tests/golden/make_ref_golden.py: utils/create_net.py
# and utils/pf.py imported from the reference checkout; Pyomo/IPOPT replaced by oracle/pyomo_shim.py,
# which solves with Newton and then evaluates every constraint rule of the reference on the solution).
def test_reference_create_network_equals_oracle(net):
    g = np.load(os.path.join(GOLD, "ref_network.npz"))
    lines = [tuple(int(x) for x in l) for l in g['lines']]
    assert list(g['bus_numbers']) == list(net['bus_numbers']) and sorted(lines) == sorted(net['line_connections'])
    assert np.array_equal(g['R'], [net['line_resistances'][l] for l in lines])       # create_net.py:22
    assert np.array_equal(g['X'], [net['line_reactances'][l] for l in lines])
    assert np.array_equal(g['imax'], [net['max_line_currents'][l] for l in lines])   # create_net.py:24
    assert np.array_equal(g['p'], [net['active_power_demand'][n] for n in net['bus_numbers']])   # create_net.py:17
    assert np.array_equal(g['q'], [net['reactive_power_demand'][n] for n in net['bus_numbers']])
    assert np.array_equal(g['bus_types'], [net['bus_types'][n] for n in net['bus_numbers']])
    assert list(g['buildings']) == list(net['buildings'])


def test_reference_power_flow_solver_outputs(fonet, tree):
    """run_pf.py's operating point through the reference's power_flow_solver and 32 scenarios through
    power_flow_solver_simplified: the sweep oracle, the Newton oracle and the C mirror agree with what
    the reference's extraction code (pf.py:108-113) returned."""
    g = np.load(os.path.join(GOLD, "ref_pf.npz"))
    assert g['max_residual'].max() < 1e-12 and g['runpf_max_residual'][0] < 1e-12   # the reference's own rules hold
    # K2 (run_pf.py:36-57)
    assert abs(g['runpf_V'][17] - 0.94137528) < 1e-8 and abs(g['runpf_V'][32] - 0.95744575) < 1e-8
    assert np.max(np.abs(g['runpf_E_next'] - 0.013625)) < 1e-17                     # pf.py:96-98
    m = c_mirror.mirror_power_flow(fonet, g['p'][:, 1:], g['q'][:, 1:])
    assert not m['failed'].any()
    assert np.max(np.abs(m['V'] - g['V'])) < 1e-9
    assert np.max(np.abs(m['P'] - g['P'])) < 1e-8 and np.max(np.abs(m['Q'] - g['Q'])) < 1e-8
    assert np.max(np.abs(np.sqrt(m['Isq']) - g['I'])) < 1e-7                        # currents: 10x inside the 1e-6 bar
    for e in (0, 7, 31):
        b = pf_ref.solve_sweep(tree, g['p'][e], g['q'][e])
        assert np.max(np.abs(np.sqrt(b['v']) - g['V'][e])) < 1e-11
        assert np.max(np.abs(np.sqrt(b['ell'][1:]) - g['I'][e])) < 1e-10
