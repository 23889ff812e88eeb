"""FlexibilityProvisionEnv -- N=1 drop-in for the reference's environment class.

Same constructor, verbs, attribute names and Python return types as
madrl/environments/flex_provision/flexibility_provision_env.py, as exercised by its three
consumers (madrl/models/model.py:208-283, utils/tester.py:18-63, run_env.py:38-170).  All
arithmetic runs in the CUDA kernels through BatchedFlexProvisionEnv(n_envs=1); this class only
(a) draws what the reference draws from the *global* numpy RNG, in the same order (quirk Q9),
and (b) converts device state to the reference's dicts / lists / floats.
"""
import numpy as np
import torch

from . import _lib
from .env import BatchedFlexProvisionEnv


class ActionSpace(object):                 # flexibility_provision_env.py:17-20
    def __init__(self, low, high):
        self.low = low
        self.high = high


class FlexibilityProvisionEnv:
    def __init__(self, kwargs, device="cuda:0", network=None, profiles=None):
        self._b = BatchedFlexProvisionEnv(kwargs, n_envs=1, device=device, network=network, profiles=profiles)
        b = self._b
        self.args = b.args
        self.model = getattr(self.args, 'alg', None)                                  # :43
        self.data_path = self.args.data_path                                          # :46
        np.random.seed(self.args.seed)                                                # :49 (global RNG, as the reference)
        self.base_powergrid = b.base_powergrid                                        # :52
        self.episode_limit = b.episode_limit                                          # :61
        self.action_space = ActionSpace(low=self.args.action_low, high=self.args.action_high)  # :64
        self.history = b.history
        self.n_agents = b.n_agents
        self.n_actions = b.n_actions
        self.agent_ids = b.agent_ids
        self.time_delta = b.time_delta
        self._last_actions = np.zeros(self.n_agents * self.n_actions)
        self._last_scaled = True
        agents_obs, state = self.reset()                                              # :69
        self.obs_size = agents_obs[0].shape[0]                                        # :71
        self.state_size = state.shape[0]                                              # :72

    # ------------------------------------------------------------------ RNG draws (Q9)
    def _select_start_hour(self):                                                     # :410-412
        return np.random.choice(24)

    def _select_start_interval(self):                                                 # :414-416
        return np.random.choice(60 // self.time_delta)

    def _select_start_day(self):                                                      # :418-424
        pv_days = self._b.profiles.n_days()
        episode_days = (self.episode_limit // (24 * (60 // self.time_delta))) + 1
        return np.random.choice(pv_days - episode_days)

    def get_action(self):                                                             # :716-719
        return np.random.uniform(low=self.action_space.low, high=self.action_space.high,
                                 size=self.n_agents * self.n_actions)

    def _start_row(self):                                                             # :477
        per_hour = 60 // self.time_delta
        return self.start_interval + self.start_hour * per_hour + self.start_day * 24 * per_hour

    def _device_reset(self):
        """E0 and a0 draws (:100, :103) + the device reset.  Returns True if solvable."""
        lo, hi = 0.9 * (self.args.e_max / 2), 1.1 * (self.args.e_max / 2)
        e0 = np.array([np.random.uniform(lo, hi) for _ in self.base_powergrid['ESSs_at_buildings']])
        a0 = self.get_action()
        self._last_actions = np.asarray(a0, dtype=np.float64)
        self._last_scaled = True
        self._b.reset(start_index=np.array([self._start_row()], dtype=np.int32), e0=e0[None], a0=a0[None],
                      return_obs=False)
        return not bool(int(self._b.flags[0]) & _lib.FLAG_RESET_FAILED)

    def reset(self):                                                                  # :74-155
        while True:
            self.start_hour = int(self._select_start_hour())
            self.start_day = int(self._select_start_day())
            self.start_interval = int(self._select_start_interval())
            if self._device_reset():
                break
            print("The power flow for the current initialization cannot be solved.")    # :151
        return self.get_obs(), self.get_state()

    def manual_reset(self, day, hour, interval):                                      # :157-239
        self.start_day, self.start_hour, self.start_interval = day, hour, interval
        if not self._device_reset():
            # the reference would retry the identical problem forever (:215-237)
            raise RuntimeError("The power flow for the current initialization cannot be solved.")
        return self.get_obs(), self.get_state()

    # ------------------------------------------------------------------ step
    def step(self, actions):                                                          # :241-356
        actions = np.asarray(actions).reshape((self.n_agents * self.n_actions))       # :260
        if actions.dtype not in (np.float32, np.float64):
            actions = actions.astype(np.float64)
        self._last_actions = actions.astype(np.float64)
        self._last_scaled = self.model != 'safemaddpg'
        reward, done, info = self._b.step(torch.from_numpy(np.ascontiguousarray(actions)).view(1, -1))
        row = self._b._info[0].cpu().numpy()
        out = {k: float(row[i]) for i, k in enumerate(_lib.INFO_KEYS[:7])}
        failed = row[_lib.INFO_KEYS.index("solver_failed")] != 0.0
        if failed:
            print("The power flow for the current step cannot be solved.")             # :315
            out["solver_failed"] = True                                                # :337 (key only on failure, Q9b)
        terminated = bool(done[0].item())
        r = float(reward[0].item())
        if terminated:
            print(f"Episode terminated at time: {self.steps} with return: {self.cumulative_reward:2.4f}.")  # :351
        return r, terminated, out

    # ------------------------------------------------------------------ observations
    def get_state(self):                                                              # :358-368
        return self._b.get_state(dtype=torch.float64)[0].cpu().numpy().copy()

    def get_obs(self):                                                                # :370-403 (Q7: pushes history)
        o = self._b.get_obs(push=True, dtype=torch.float64)[0].cpu().numpy()
        return [o[i].copy() for i in range(self.n_agents)]

    def get_obs_agent(self, agent_id):                                                # :405-408
        return self.get_obs()[agent_id]

    def get_obs_size(self):
        return self.obs_size

    def get_state_size(self):
        return self.state_size

    def get_avail_actions(self):                                                      # :721-726
        return np.expand_dims(np.array([self.get_avail_agent_actions(i) for i in range(self.n_agents)]), axis=0)

    def get_avail_agent_actions(self, agent_id):                                      # :728-730
        return [1] * self.n_actions

    def get_total_actions(self):
        return self.n_actions

    def get_num_of_agents(self):
        return self.n_agents

    def get_env_info(self):                                                           # multiagentenv.py:61-67
        return {"state_shape": self.get_state_size(), "obs_shape": self.get_obs_size(),
                "n_actions": self.get_total_actions(), "n_agents": self.n_agents,
                "episode_limit": self.episode_limit}

    def close(self):
        self._b.close()

    # ------------------------------------------------------------------ attributes read by callers
    def _row(self):
        steps = self.steps
        return int(self._b.start_index[0]) + (steps - 1 if steps > 1 else 1)          # quirk Q1

    @property
    def steps(self):
        return int(self._b.steps[0])

    @property
    def cumulative_reward(self):
        return float(self._b.cumulative_reward[0])

    def _by(self, keys, values):
        return {k: float(v) for k, v in zip(keys, values)}

    @property
    def current_voltage(self):
        return self._by(self.base_powergrid['bus_numbers'], self._b.voltages[0].cpu().numpy())

    @property
    def current_ess_energy(self):
        return self._by(self.base_powergrid['ESSs_at_buildings'], self._b.ess_energy[0].cpu().numpy())

    @property
    def initial_ess_energy(self):
        return self._by(self.base_powergrid['ESSs_at_buildings'], self._b.initial_ess_energy[0].cpu().numpy())

    @property
    def current_active_demand(self):
        p = np.concatenate(([0.0], self._b.profiles.P[self._row()]))                   # :489-490
        return self._by(self.base_powergrid['bus_numbers'], p)

    @property
    def current_reactive_demand(self):
        q = np.concatenate(([0.0], self._b.profiles.Q[self._row()]))
        return self._by(self.base_powergrid['bus_numbers'], q)

    @property
    def current_pv_power(self):
        return self._by(self.base_powergrid['PVs_at_buildings'], self._b.profiles.PV[self._row()])

    @property
    def current_price(self):
        return np.array([self._b.profiles.price[self._row()]])                         # shape (1,) like :619

    def _setpoint(self, k, keys):
        return self._by(keys, self._b.setpoints[0, k].cpu().numpy())

    @property
    def power_reduction(self):
        return self._setpoint(0, self.base_powergrid['buildings'])

    @property
    def ess_charging(self):
        return self._setpoint(1, self.base_powergrid['ESSs_at_buildings'])

    @property
    def ess_discharging(self):
        return self._setpoint(2, self.base_powergrid['ESSs_at_buildings'])

    @property
    def q_pv(self):
        return self._setpoint(3, self.base_powergrid['PVs_at_buildings'])

    @property
    def percentage_reduction(self):
        # not rolled back on solver failure in the reference either (:318-328 omit it)
        a = self._last_actions[0::4]
        mpr = self.args.max_power_reduction
        pct = mpr * a if self._last_scaled else a
        return self._by(self.base_powergrid['buildings'], np.clip(pct, 0, mpr))

    # accessor set used by utils/tester.py:35-61                                      :740-778
    def _get_bus_v(self):
        return self._b.voltages[0].cpu().numpy().copy()

    def _get_bus_active(self):
        return np.array(list(self.current_active_demand.values()))

    def _get_bus_reactive(self):
        return np.array(list(self.current_reactive_demand.values()))

    def _get_pv_active(self):
        return np.array(list(self.current_pv_power.values()))

    def _get_pv_reactive(self):
        return self._b.setpoints[0, 3].cpu().numpy().copy()

    def _get_ess_energy(self):
        return self._b.ess_energy[0].cpu().numpy().copy()

    def _get_power_reduction(self):
        return self._b.setpoints[0, 0].cpu().numpy().copy()

    def _get_ess_charging(self):
        return self._b.setpoints[0, 1].cpu().numpy().copy()

    def _get_ess_discharging(self):
        return self._b.setpoints[0, 2].cpu().numpy().copy()

    def _get_price(self):
        return np.array([self.current_price])
