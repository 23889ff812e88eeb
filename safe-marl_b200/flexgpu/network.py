"""Feeder topology: the reference's `create_network()` dict -> position-indexed arrays.

`create_network` mirrors utils/create_net.py:8-38 (same keys, same per-unit conversion at
:17-24).  The reference reads Nodes_33.xlsx / Lines_33.xlsx, which are Git-LFS payloads that
do not ship with the repository; when they are not readable the public IEEE 33-bus case
(Baran & Wu 1989) is used, which is what those files describe.
"""
import os
import warnings

import numpy as np

from .profiles import _is_lfs_pointer

# (FROM, TO, R ohm, X ohm, Imax A) -- Baran & Wu 33-bus; Imax is a synthetic rating
# (400 A head of the trunk, 200 A elsewhere): the reference's column is not public.
_IEEE33_LINES = [
    (1, 2, .0922, .0470), (2, 3, .4930, .2511), (3, 4, .3660, .1864), (4, 5, .3811, .1941),
    (5, 6, .8190, .7070), (6, 7, .1872, .6188), (7, 8, .7114, .2351), (8, 9, 1.0300, .7400),
    (9, 10, 1.0440, .7400), (10, 11, .1966, .0650), (11, 12, .3744, .1238), (12, 13, 1.4680, 1.1550),
    (13, 14, .5416, .7129), (14, 15, .5910, .5260), (15, 16, .7463, .5450), (16, 17, 1.2890, 1.7210),
    (17, 18, .7320, .5740), (2, 19, .1640, .1565), (19, 20, 1.5042, 1.3554), (20, 21, .4095, .4784),
    (21, 22, .7089, .9373), (3, 23, .4512, .3083), (23, 24, .8980, .7091), (24, 25, .8960, .7011),
    (6, 26, .2030, .1034), (26, 27, .2842, .1447), (27, 28, 1.0590, .9337), (28, 29, .8042, .7006),
    (29, 30, .5075, .2585), (30, 31, .9744, .9630), (31, 32, .3105, .3619), (32, 33, .3410, .5302),
]
_IEEE33_LOADS_KW_KVAR = [
    (100, 60), (90, 40), (120, 80), (60, 30), (60, 20), (200, 100), (200, 100), (60, 20), (60, 20),
    (45, 30), (60, 35), (60, 35), (120, 80), (60, 10), (60, 20), (60, 20), (90, 40), (90, 40),
    (90, 40), (90, 40), (90, 40), (90, 50), (420, 200), (420, 200), (60, 25), (60, 25), (60, 20),
    (120, 70), (200, 600), (150, 70), (210, 100), (60, 40),
]


def _ieee33_tables():
    nodes = [dict(NODES=1, Tb=1, PDn=0.0, QDn=0.0)]
    for n, (p, q) in zip(range(2, 34), _IEEE33_LOADS_KW_KVAR):
        nodes.append(dict(NODES=n, Tb=0, PDn=float(p), QDn=float(q)))
    lines = [dict(FROM=f, TO=t, R=r, X=x, Imax=(400.0 if t <= 6 else 200.0)) for (f, t, r, x) in _IEEE33_LINES]
    return nodes, lines


def _read_xlsx_tables(data_path):
    import pandas as pd
    nodes = pd.read_excel(os.path.join(data_path, "Nodes_33.xlsx")).to_dict("records")
    lines = pd.read_excel(os.path.join(data_path, "Lines_33.xlsx")).to_dict("records")
    return nodes, lines


def create_network(env_args, data_path=None):
    """Same return dict as utils/create_net.py:27-38."""
    nodes = lines = None
    if data_path is not None:
        files = [os.path.join(data_path, n) for n in ("Nodes_33.xlsx", "Lines_33.xlsx")]
        if all(os.path.isfile(f) and not _is_lfs_pointer(f) for f in files):
            nodes, lines = _read_xlsx_tables(data_path)     # real files: a read error is the caller's to see
        else:
            # absent, or the Git-LFS pointers the repository ships: the public case the files describe,
            # with SYNTHETIC line ratings (the Imax column is not public)
            warnings.warn(f"Nodes_33.xlsx / Lines_33.xlsx not readable under {data_path!r}: using the public "
                          "IEEE 33-bus case (Baran & Wu 1989) with synthetic Imax ratings", stacklevel=2)
    if nodes is None:
        nodes, lines = _ieee33_tables()
    s_nom, v_nom = env_args["s_nom"], env_args["v_nom"]
    zbase = v_nom ** 2 * 1000 / s_nom               # create_net.py:22
    ibase = s_nom / v_nom                           # create_net.py:24
    return {
        'bus_numbers': [r['NODES'] for r in nodes],
        'line_connections': [(r['FROM'], r['TO']) for r in lines],
        'line_resistances': {(r['FROM'], r['TO']): r['R'] / zbase for r in lines},
        'line_reactances': {(r['FROM'], r['TO']): r['X'] / zbase for r in lines},
        'max_line_currents': {(r['FROM'], r['TO']): r['Imax'] / ibase for r in lines},
        'bus_types': {r['NODES']: r['Tb'] for r in nodes},
        'active_power_demand': {r['NODES']: r['PDn'] / s_nom for r in nodes},
        'reactive_power_demand': {r['NODES']: r['QDn'] / s_nom for r in nodes},
        'buildings': env_args['buildings'],
        'PVs_at_buildings': env_args['pv_nodes'],
        'ESSs_at_buildings': env_args['ess_nodes'],
    }


class Network:
    """Position-indexed view of the reference's network dict (slack must be position 0)."""

    def __init__(self, net):
        self.dict = net
        buses = list(net['bus_numbers'])
        self.buses = buses
        self.n_bus = len(buses)
        self.position = {b: i for i, b in enumerate(buses)}
        slack = [b for b in buses if net['bus_types'][b] == 1]
        if len(slack) != 1 or self.position[slack[0]] != 0:
            # the reference prepends the slack column (flexibility_provision_env.py:489-490),
            # i.e. it assumes the slack bus is the first entry of bus_numbers.
            raise ValueError("exactly one slack bus, listed first in bus_numbers, is required")
        if len(net['line_connections']) != self.n_bus - 1:
            raise ValueError("the feeder must be radial (n_lines == n_bus - 1)")
        adj = {i: [] for i in range(self.n_bus)}
        for key in net['line_connections']:
            f, t = key
            adj[self.position[f]].append((self.position[t], key))
            adj[self.position[t]].append((self.position[f], key))
        self.parent = np.full(self.n_bus, -1, dtype=np.int32)
        self.r = np.zeros(self.n_bus)
        self.x = np.zeros(self.n_bus)
        self.imax = np.zeros(self.n_bus)
        self.line_of_bus = {}
        seen, frontier = {0}, [0]
        while frontier:
            u = frontier.pop()
            for (w, key) in adj[u]:
                if w in seen:
                    continue
                seen.add(w)
                self.parent[w] = u
                self.r[w] = net['line_resistances'][key]
                self.x[w] = net['line_reactances'][key]
                self.imax[w] = net['max_line_currents'][key]
                self.line_of_bus[w] = key
                frontier.append(w)
        if len(seen) != self.n_bus:
            raise ValueError("the feeder is not connected")
        self.base_p = np.array([net['active_power_demand'][b] for b in buses], dtype=np.float64)
        self.base_q = np.array([net['reactive_power_demand'][b] for b in buses], dtype=np.float64)
