"""IEEE 33-bus feeder data + the reference's per-unit conversion (oracle side).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY: data unpinned -- the reference's
Lines_33.xlsx / Nodes_33.xlsx are Git-LFS pointers, so the feeder is restated
from the public Baran & Wu (1989) 33-bus case.  The `Imax` column is not part of
that case; `IMAX_A` below is a documented synthetic rating.

Follows utils/create_net.py:8-38 (dict keys and p.u. conversion at :17-24).
"""
import numpy as np

V_NOM_KV = 12.66      # madrl/args/env_args/flex_provision.yaml:31
S_NOM_KVA = 1000.0    # flex_provision.yaml:32

# (FROM, TO, R ohm, X ohm) -- Baran & Wu 1989, 32 branches, bus 1 = substation
LINES = [
    (1, 2, 0.0922, 0.0470), (2, 3, 0.4930, 0.2511), (3, 4, 0.3660, 0.1864),
    (4, 5, 0.3811, 0.1941), (5, 6, 0.8190, 0.7070), (6, 7, 0.1872, 0.6188),
    (7, 8, 0.7114, 0.2351), (8, 9, 1.0300, 0.7400), (9, 10, 1.0440, 0.7400),
    (10, 11, 0.1966, 0.0650), (11, 12, 0.3744, 0.1238), (12, 13, 1.4680, 1.1550),
    (13, 14, 0.5416, 0.7129), (14, 15, 0.5910, 0.5260), (15, 16, 0.7463, 0.5450),
    (16, 17, 1.2890, 1.7210), (17, 18, 0.7320, 0.5740), (2, 19, 0.1640, 0.1565),
    (19, 20, 1.5042, 1.3554), (20, 21, 0.4095, 0.4784), (21, 22, 0.7089, 0.9373),
    (3, 23, 0.4512, 0.3083), (23, 24, 0.8980, 0.7091), (24, 25, 0.8960, 0.7011),
    (6, 26, 0.2030, 0.1034), (26, 27, 0.2842, 0.1447), (27, 28, 1.0590, 0.9337),
    (28, 29, 0.8042, 0.7006), (29, 30, 0.5075, 0.2585), (30, 31, 0.9744, 0.9630),
    (31, 32, 0.3105, 0.3619), (32, 33, 0.3410, 0.5302),
]

# (kW, kvar) at buses 2..33
LOADS = [
    (100, 60), (90, 40), (120, 80), (60, 30), (60, 20), (200, 100), (200, 100),
    (60, 20), (60, 20), (45, 30), (60, 35), (60, 35), (120, 80), (60, 10),
    (60, 20), (60, 20), (90, 40), (90, 40), (90, 40), (90, 40), (90, 40),
    (90, 50), (420, 200), (420, 200), (60, 25), (60, 25), (60, 20), (120, 70),
    (200, 600), (150, 70), (210, 100), (60, 40),
]

# Synthetic thermal ratings in ampere (NOT from the reference; its Imax column is
# unavailable).  400 A on the trunk head, 200 A elsewhere.
IMAX_A = {(f, t): (400.0 if t <= 6 else 200.0) for (f, t, _, _) in LINES}

BUILDINGS = [5, 10, 15, 20, 25]   # flex_provision.yaml:28-30


def create_network(v_nom=V_NOM_KV, s_nom=S_NOM_KVA, buildings=BUILDINGS):
    """Restates utils/create_net.py:8-38 on the public IEEE-33 data."""
    zbase = v_nom ** 2 * 1000 / s_nom          # create_net.py:22
    ibase = s_nom / v_nom                      # create_net.py:24
    nodes = list(range(1, 34))
    types = {n: (1 if n == 1 else 0) for n in nodes}
    pd_ = {1: 0.0}
    qd_ = {1: 0.0}
    for n, (p, q) in zip(range(2, 34), LOADS):
        pd_[n] = p / s_nom                     # create_net.py:17
        qd_[n] = q / s_nom                     # create_net.py:18
    lines = [(f, t) for (f, t, _, _) in LINES]
    r = {(f, t): rr / zbase for (f, t, rr, _) in LINES}
    x = {(f, t): xx / zbase for (f, t, _, xx) in LINES}
    imax = {k: v / ibase for k, v in IMAX_A.items()}
    return {
        'bus_numbers': nodes,
        'line_connections': lines,
        'line_resistances': r,
        'line_reactances': x,
        'max_line_currents': imax,
        'bus_types': types,
        'active_power_demand': pd_,
        'reactive_power_demand': qd_,
        'buildings': list(buildings),
        'PVs_at_buildings': list(buildings),
        'ESSs_at_buildings': list(buildings),
    }


def tree_arrays(net):
    """Arrays in *bus order* for the numpy solvers.

    Returns dict with: buses (list), slack index, and per non-slack bus j (position
    in `buses`): parent position, R, X of the line feeding j, Imax.  A processing
    order (parents before children) is included.
    """
    buses = list(net['bus_numbers'])
    pos = {b: i for i, b in enumerate(buses)}
    nb = len(buses)
    slack = [pos[b] for b in buses if net['bus_types'][b] == 1]
    assert len(slack) == 1
    slack = slack[0]
    # orient every line away from the slack bus (radial feeder)
    adj = {i: [] for i in range(nb)}
    for (f, t) in net['line_connections']:
        adj[pos[f]].append((pos[t], (f, t)))
        adj[pos[t]].append((pos[f], (f, t)))
    parent = np.full(nb, -1, dtype=np.int64)
    R = np.zeros(nb)
    X = np.zeros(nb)
    imax = np.zeros(nb)
    order = []
    seen = {slack}
    stack = [slack]
    while stack:
        u = stack.pop()
        order.append(u)
        for (w, key) in sorted(adj[u], reverse=True):
            if w not in seen:
                seen.add(w)
                parent[w] = u
                R[w] = net['line_resistances'][key]
                X[w] = net['line_reactances'][key]
                imax[w] = net['max_line_currents'][key]
                stack.append(w)
    assert len(order) == nb, "network is not connected"
    assert len(net['line_connections']) == nb - 1, "network is not radial"
    return dict(buses=buses, pos=pos, slack=slack, parent=parent, R=R, X=X,
                imax=imax, order=np.array(order, dtype=np.int64))
