"""Single-env fp64 restatement of the reference's FlexibilityProvisionEnv (oracle side).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY: pinned on the
reference's own env code -- tests/golden/ref_env_*.npz are episodes produced by
the reference's FlexibilityProvisionEnv itself (tests/golden/make_ref_golden.py)
and this restatement reproduces them to 1e-16 (tests/test_oracle_env.py); the
power flow behind both is Newton (oracle/pf_ref.py), not IPOPT.

Behavioural restatement of
madrl/environments/flex_provision/flexibility_provision_env.py -- every method
cites the lines it follows.  It is deliberately NOT structured like the
reference (no pandas, no deepcopy of dict state, no logging); the profile data
are handed in as already-resampled arrays, which is what the reference holds
after `resample_data` (:467-471).  All arithmetic is fp64; actions are widened
to fp64 before any arithmetic (the reference pins numpy 1.26 where
python_float * np.float32 -> float64).

Quirks reproduced on purpose: Q1 (row indexing), Q2 (E_init lag),
Q3 (no dt in the clip), Q4 (signed PV cost), Q7 (get_obs side effect), Q9 (RNG
draw order in reset), Q9b (`solver_failed` key only on failure).
"""
from math import acos, tan

import numpy as np

from . import pf_ref
from .ieee33 import create_network, tree_arrays

DEFAULT_ARGS = dict(                       # madrl/args/env_args/flex_provision.yaml:3-33
    history=24, pv_scale=0.15, demand_scale=1.0, reactive_scale=1.0, v_max=1.1, v_min=0.9,
    episode_limit=96, action_low=0, action_high=1.0, seed=0, e_min=0.0, e_max=0.025,
    pv_cost=0.05, ess_cost=0.03, discomfort_coeff=0.15, voltage_coeff=1.0, p_ch_max=0.005,
    p_dis_max=0.005, eta_ch=0.9, eta_dis=0.9, cos_phi_max=0.95, max_power_reduction=0.5,
    sample_interval="15min", buildings=[5, 10, 15, 20, 25], pv_nodes=[5, 10, 15, 20, 25],
    ess_nodes=[5, 10, 15, 20, 25], v_nom=12.66, s_nom=1000, pv_cap=0.15,
)


class _Args(dict):
    __getattr__ = dict.__getitem__


class RefFlexEnv:
    """One environment.  `profiles` = dict(P[T,32], Q[T,32], PV[T,5], price[T]) at the
    sample interval; `rng` = a numpy RandomState standing in for the global
    `np.random` the reference seeds at :49."""

    def __init__(self, args=None, net=None, profiles=None, rng=None, pf_method='newton',
                 time_delta=15, force_fail=None):
        a = dict(DEFAULT_ARGS)
        a.update(args or {})
        self.args = _Args(a)
        self.model = a.get('alg', None)                                   # :43
        self.rng = rng if rng is not None else np.random.RandomState(self.args.seed)  # :49
        self.base_powergrid = net or create_network(self.args.v_nom, self.args.s_nom, self.args.buildings)
        self._tree = tree_arrays(self.base_powergrid)
        self._pf_method = pf_method
        self._force_fail = force_fail or (lambda env: False)              # fault-injection hook (tests)
        self.prof = profiles
        self.time_delta = time_delta                                      # :422
        self.episode_limit = self.args.episode_limit                      # :61
        self.history = self.args.history                                  # :65
        self.n_agents = len(self.base_powergrid['buildings'])             # :66
        self.n_actions = 4                                                # :67
        self.agent_ids = self.base_powergrid['buildings']                 # :68
        obs, state = self.reset()                                         # :69
        self.obs_size = obs[0].shape[0]                                   # :71
        self.state_size = state.shape[0]                                  # :72

    # ------------------------------------------------------------------ helpers
    def _kappa(self):
        return tan(acos(self.args.cos_phi_max))                           # :623

    def _scale_and_clip_q_pv(self, reactive_action, active_power):       # :621-626
        c = self._kappa() * active_power
        return float(np.clip(-c + reactive_action * (c - (-c)), -c, c))

    def clip_percentage_reduction(self, d):                               # :676-677
        return {k: float(np.clip(v, 0, self.args.max_power_reduction)) for k, v in d.items()}

    def adjust_ess_actions(self, ch, dis):                                # :663-674
        for k in ch:
            if ch[k] > 0 and dis[k] > 0:
                if ch[k] > dis[k]:
                    ch[k] -= dis[k]
                    dis[k] = 0
                else:
                    dis[k] -= ch[k]
                    ch[k] = 0
        return ch, dis

    def _clip_power_charging_discharging(self, charging, discharging, e_now):  # :628-661
        A = self.args
        charging = float(np.clip(charging, 0, A.p_ch_max))
        discharging = float(np.clip(discharging, 0, A.p_dis_max))
        e_next = e_now + A.eta_ch * charging - (1 / A.eta_dis) * discharging    # no dt (Q3)
        if e_next > A.e_max:
            excess = e_next - A.e_max
            if charging > excess / A.eta_ch:
                charging -= excess / A.eta_ch
            else:
                discharging += (excess - charging * A.eta_ch) * A.eta_dis
                charging = 0
        elif e_next < A.e_min:
            lack = A.e_min - e_next
            if discharging > lack * A.eta_dis:
                discharging -= lack * A.eta_dis
            else:
                charging += (lack - discharging / A.eta_dis) / A.eta_ch
                discharging = 0
        charging = float(np.clip(charging, 0, A.p_ch_max))
        discharging = float(np.clip(discharging, 0, A.p_dis_max))
        return charging, discharging

    # ---------------------------------------------------------- profile indexing
    def _n_days(self):
        # (index[-1] - index[0]).days with a uniform `time_delta`-minute index      :421
        T = self.prof['P'].shape[0]
        return ((T - 1) * self.time_delta) // (24 * 60)

    def _draw_start(self):
        per_hour = 60 // self.time_delta
        hour = self.rng.choice(24)                                                  # :85,:412
        episode_days = (self.episode_limit // (24 * per_hour)) + 1                  # :423
        day = self.rng.choice(self._n_days() - episode_days)                        # :86,:424
        interval = self.rng.choice(per_hour)                                        # :87,:416
        return int(day), int(hour), int(interval)

    def start_index(self, day, hour, interval):
        per_hour = 60 // self.time_delta
        return interval + hour * per_hour + day * 24 * per_hour                     # :477

    def _slice_episode(self, start):
        n = self.episode_limit + self.history + 1                                   # :478
        P = self.prof['P'][start:start + n]
        Q = self.prof['Q'][start:start + n]
        z = np.zeros((P.shape[0], 1))
        self.active_demand_history = np.hstack((z, P))                              # :489-492
        self.reactive_demand_history = np.hstack((z, Q))                            # :510-513
        self.pv_history = self.prof['PV'][start:start + n]                          # :523-530
        self.price_history = self.prof['price'][start:start + n].reshape(-1, 1)     # :540-547

    def _set_demand_pv_prices(self):                                                # :609-619
        t = self.steps
        G = self.base_powergrid
        pv = self.pv_history[t]
        self.current_pv_power = {g: float(pv[i]) for i, g in enumerate(G['PVs_at_buildings'])}
        self.current_active_demand = {b: float(self.active_demand_history[t, i]) for i, b in enumerate(G['bus_numbers'])}
        self.current_reactive_demand = {b: float(self.reactive_demand_history[t, i]) for i, b in enumerate(G['bus_numbers'])}
        self.current_price = self.price_history[t].copy()                           # shape (1,)

    # ------------------------------------------------------------------ actions
    def get_action(self):                                                           # :716-719
        return self.rng.uniform(low=self.args.action_low, high=self.args.action_high,
                                size=self.n_agents * self.n_actions)

    def _parse_scaled(self, actions):
        """:113-118 / :276-281 -- scale raw actions to setpoints."""
        A, G = self.args, self.base_powergrid
        pct, ch, dis, qpv = {}, {}, {}, {}
        for i in range(self.n_agents):
            b = G['buildings'][i]
            pct[b] = A.max_power_reduction * float(actions[i * 4])
            ch[b] = A.p_ch_max * float(actions[i * 4 + 1])
            dis[b] = A.p_dis_max * float(actions[i * 4 + 2])
            qpv[b] = self._scale_and_clip_q_pv(float(actions[i * 4 + 3]), self.current_pv_power[b])
        return pct, ch, dis, qpv

    def _solve(self, e_init):
        if self._force_fail(self):
            raise pf_ref.SolverFailure('injected')
        return pf_ref.power_flow_solver(
            self.base_powergrid, self.current_active_demand, self.current_reactive_demand,
            self.power_reduction, self.current_pv_power, self.q_pv, self.ess_charging,
            self.ess_discharging, e_init, method=self._pf_method, tree=self._tree)

    # -------------------------------------------------------------------- reset
    def _reset_common(self, draw):
        G = self.base_powergrid
        self.steps = 1                                                              # :76
        self.cumulative_reward = 0                                                  # :77
        self.obs_history = {i: [] for i in range(self.n_agents)}                    # :79-80
        while True:
            day, hour, interval = draw()
            self.start_day, self.start_hour, self.start_interval = day, hour, interval
            self._slice_episode(self.start_index(day, hour, interval))              # :92-95
            self._set_demand_pv_prices()                                            # :98
            lo, hi = 0.9 * (self.args.e_max / 2), 1.1 * (self.args.e_max / 2)
            self.initial_ess_energy = {k: float(self.rng.uniform(lo, hi)) for k in G['ESSs_at_buildings']}  # :100
            actions = self.get_action()                                             # :103
            self._last_reset_draw = dict(start=self.start_index(day, hour, interval),
                                         e0=np.array([self.initial_ess_energy[k] for k in G['ESSs_at_buildings']]),
                                         a0=np.array(actions, dtype=np.float64))
            pct, ch, dis, qpv = self._parse_scaled(actions)                         # :113-118
            self.percentage_reduction = self.clip_percentage_reduction(pct)         # :121
            self.power_reduction = {b: self.current_active_demand[b] * self.percentage_reduction[b] for b in G['buildings']}  # :124
            self.ess_charging, self.ess_discharging = self.adjust_ess_actions(ch, dis)  # :127
            self.q_pv = qpv
            for k in self.ess_charging:                                             # :129-130 (clip vs E0)
                self.ess_charging[k], self.ess_discharging[k] = self._clip_power_charging_discharging(
                    self.ess_charging[k], self.ess_discharging[k], self.initial_ess_energy[k])
            try:
                res = self._solve(self.initial_ess_energy)                          # :134-144
            except pf_ref.SolverFailure:
                continue                                                            # :150-153
            self.current_voltage = res['Voltages']                                  # :146
            self.current_ess_energy = res['Next ESS Energy']                        # :147
            self.last_result = res
            break
        return self.get_obs(), self.get_state()                                     # :155

    def reset(self):                                                                # :74-155
        return self._reset_common(self._draw_start)

    def manual_reset(self, day, hour, interval):                                    # :157-239
        return self._reset_common(lambda: (day, hour, interval))

    def reset_with(self, start, e0, a0):
        """reset() with the draws handed in (test hook): episode starting at dataset row `start`, initial ESS
        energies `e0`, initial actions `a0` -- what the reference draws at :85-87, :100, :103."""
        class _Given:
            def __init__(self, e0, a0):
                self.e0, self.a0 = list(np.asarray(e0, dtype=np.float64)), np.asarray(a0, dtype=np.float64)

            def uniform(self, low=0.0, high=1.0, size=None):
                return self.a0.copy() if size is not None else self.e0.pop(0)
        per_day = 24 * (60 // self.time_delta)
        day, rem = divmod(int(start), per_day)
        hour, interval = divmod(rem, 60 // self.time_delta)
        keep, self.rng = self.rng, _Given(e0, a0)
        try:
            return self._reset_common(lambda: (day, hour, interval))
        finally:
            self.rng = keep

    # --------------------------------------------------------------------- step
    def step(self, actions):                                                        # :241-356
        G, A = self.base_powergrid, self.args
        keep = (dict(self.current_voltage), dict(self.current_ess_energy), dict(self.power_reduction),
                dict(self.ess_charging), dict(self.ess_discharging), dict(self.q_pv))   # :246-257
        actions = np.asarray(actions).reshape(self.n_agents * self.n_actions).astype(np.float64)  # :260 (+Q6)
        if self.model == 'safemaddpg':                                              # :268-274
            pct = {G['buildings'][i]: float(actions[i * 4]) for i in range(self.n_agents)}
            ch = {G['buildings'][i]: float(actions[i * 4 + 1]) for i in range(self.n_agents)}
            dis = {G['buildings'][i]: float(actions[i * 4 + 2]) for i in range(self.n_agents)}
            qpv = {G['buildings'][i]: float(actions[i * 4 + 3]) for i in range(self.n_agents)}
        else:
            pct, ch, dis, qpv = self._parse_scaled(actions)                         # :276-281
        self.percentage_reduction = self.clip_percentage_reduction(pct)             # :284
        self.ess_charging, self.ess_discharging = self.adjust_ess_actions(ch, dis)  # :287
        self.q_pv = qpv
        for k in self.ess_charging:                                                 # :289-290 (clip vs E_cur)
            self.ess_charging[k], self.ess_discharging[k] = self._clip_power_charging_discharging(
                self.ess_charging[k], self.ess_discharging[k], self.current_ess_energy[k])
        self.power_reduction = {b: self.current_active_demand[b] * self.percentage_reduction[b] for b in G['buildings']}  # :293
        solvable = False
        try:
            if np.isnan(actions).any():                                             # np.clip keeps a NaN: the NLP gets a NaN injection
                raise pf_ref.SolverFailure("NaN action")                             # and the solve raises (:314)
            res = self._solve(self.initial_ess_energy)                              # :298-308 (Q2: E_init)
            self.current_voltage = res['Voltages']
            self.current_ess_energy = res['Next ESS Energy']
            self.last_result = res
            solvable = True
        except pf_ref.SolverFailure:                                                # :314-328
            (self.current_voltage, self.current_ess_energy, self.power_reduction,
             self.ess_charging, self.ess_discharging, self.q_pv) = keep
        reward, info = self.calculate_reward(self.power_reduction, self.ess_charging,
                                             self.ess_discharging, self.q_pv, self.current_voltage)  # :330-335
        if not solvable:
            reward -= 200                                                           # :336
            info["solver_failed"] = True                                            # :337
        self._set_demand_pv_prices()                                                # :340 (Q1: row = steps)
        self.steps += 1                                                             # :342
        self.cumulative_reward += reward                                            # :343
        terminated = bool(self.steps >= self.episode_limit or not solvable)         # :345-348
        self.initial_ess_energy = self.current_ess_energy                           # :354
        return reward, terminated, info

    def calculate_reward(self, power_reduction, ess_charging, ess_discharging, q_pv, voltages):  # :679-706
        A = self.args
        lam = self.current_price
        revenue = sum(lam * power_reduction[b] for b in power_reduction)
        der_cost = sum(A.pv_cost * q_pv[g] for g in q_pv)                            # signed (Q4)
        ess_cost = sum(A.ess_cost * (ess_charging[k] + ess_discharging[k]) for k in ess_charging)
        discomfort = sum(A.discomfort_coeff * power_reduction[b] ** 2 for b in power_reduction)
        vpen = sum(A.voltage_coeff * max(0, v - A.v_max, A.v_min - v) for v in voltages.values())
        reward = revenue - der_cost - ess_cost - discomfort - vpen
        f = lambda x: float(np.asarray(x).reshape(-1)[0])        # lambda_flex has shape (1,) (:689-694)
        info = {
            'reward': f(reward), 'revenue': f(revenue), 'der_cost': f(der_cost),
            'ess_cost': f(ess_cost), 'discomfort_penalty': f(discomfort),
            'voltage_penalty': f(vpen), 'cumulative_reward': self.cumulative_reward,
        }
        return f(reward), info

    # ------------------------------------------------------------- observations
    def get_state(self):                                                            # :358-368
        G = self.base_powergrid
        s = [self.current_active_demand[b] for b in G['bus_numbers']]
        s += [self.current_reactive_demand[b] for b in G['bus_numbers']]
        s += [self.current_pv_power[g] for g in G['PVs_at_buildings']]
        s += [self.current_voltage[b] for b in G['bus_numbers']]
        s += [self.current_price[0]]
        s += [self.current_ess_energy[k] for k in G['ESSs_at_buildings']]
        return np.array(s)

    def get_obs(self):                                                              # :370-403 (Q7)
        out = []
        for i in range(self.n_agents):
            b = self.agent_ids[i]
            cur = np.array([self.current_active_demand[b], self.current_reactive_demand[b],
                            self.current_pv_power[b], self.current_voltage[b],
                            self.current_price[0], self.current_ess_energy[b]])
            if self.history > 1:
                past = self.obs_history[i][-(self.history - 1):] if self.history > 1 else []
                pad = [np.zeros_like(cur)] * (self.history - 1 - len(past))
                out.append(np.concatenate(pad + past + [cur], axis=0))
                self.obs_history[i].append(cur.copy())
            else:
                out.append(cur)
        return out

    def get_obs_agent(self, agent_id):                                              # :405-408
        return self.get_obs()[agent_id]

    def get_obs_size(self):
        return self.obs_size

    def get_state_size(self):
        return self.state_size

    def get_avail_actions(self):                                                    # :721-726
        return np.expand_dims(np.array([self.get_avail_agent_actions(i) for i in range(self.n_agents)]), axis=0)

    def get_avail_agent_actions(self, agent_id):                                    # :728-730
        return [1] * self.n_actions

    def get_total_actions(self):
        return self.n_actions

    def get_num_of_agents(self):
        return self.n_agents

    def get_env_info(self):                                                         # multiagentenv.py:61-67
        return {"state_shape": self.get_state_size(), "obs_shape": self.get_obs_size(),
                "n_actions": self.get_total_actions(), "n_agents": self.n_agents,
                "episode_limit": self.episode_limit}

    # accessor set used by utils/tester.py:35-61                                     :740-778
    def _get_bus_v(self):
        return np.array([self.current_voltage[b] for b in self.base_powergrid['bus_numbers']])

    def _get_bus_active(self):
        return np.array([self.current_active_demand[b] for b in self.base_powergrid['bus_numbers']])

    def _get_bus_reactive(self):
        return np.array([self.current_reactive_demand[b] for b in self.base_powergrid['bus_numbers']])

    def _get_pv_active(self):
        return np.array([self.current_pv_power[g] for g in self.base_powergrid['PVs_at_buildings']])

    def _get_pv_reactive(self):
        return np.array([self.q_pv[g] for g in self.base_powergrid['PVs_at_buildings']])

    def _get_ess_energy(self):
        return np.array([self.current_ess_energy[k] for k in self.base_powergrid['ESSs_at_buildings']])

    def _get_power_reduction(self):
        return np.array([self.power_reduction[b] for b in self.base_powergrid['buildings']])

    def _get_ess_charging(self):
        return np.array([self.ess_charging[k] for k in self.base_powergrid['ESSs_at_buildings']])

    def _get_ess_discharging(self):
        return np.array([self.ess_discharging[k] for k in self.base_powergrid['ESSs_at_buildings']])

    def _get_price(self):
        return np.array([self.current_price])
