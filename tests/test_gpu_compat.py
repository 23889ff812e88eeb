"""GPU: the N=1 drop-in `FlexibilityProvisionEnv` against the oracle's single-env restatement,
driven the way the reference's three consumers drive it (run_env.py:56-92,
madrl/models/model.py:208-254, utils/tester.py:18-63): same global-RNG draw order (Q9), same
Python return types, same accessor values (BASELINE config 1)."""
import numpy as np
import pytest

from oracle import env_ref, ieee33

pytestmark = pytest.mark.gpu


def pair(cuda, profiles, seed, **extra):
    from flexgpu import FlexibilityProvisionEnv
    a = dict(env_ref.DEFAULT_ARGS); a.update(extra); a["seed"] = seed
    ref = env_ref.RefFlexEnv(a, ieee33.create_network(), profiles.as_dict(), rng=np.random.RandomState(seed))
    env = FlexibilityProvisionEnv(dict(a), device=cuda, profiles=profiles)       # seeds np.random itself (:49)
    return env, ref


def test_run_env_style_episode(cuda, profiles):
    """run_env.py: reset, then 96 x (get_obs, get_state, N(0, 0.5) actions, step), done ignored."""
    env, ref = pair(cuda, profiles, 0)
    obs, state = env.reset(); robs, rstate = ref.reset()
    assert isinstance(obs, list) and len(obs) == 5 and obs[0].shape == (144,) and state.shape == (110,)
    assert obs[0].dtype == np.float64
    assert (env.start_day, env.start_hour, env.start_interval) == (ref.start_day, ref.start_hour, ref.start_interval)
    rng = np.random.RandomState(42)
    for t in range(96):
        o, s = env.get_obs(), env.get_state(); ro, rs = ref.get_obs(), ref.get_state()
        assert np.max(np.abs(np.array(o) - np.array(ro))) < 1e-8 and np.max(np.abs(s - rs)) < 1e-8
        a = np.concatenate([rng.normal(0, 0.5, 4) for _ in range(5)])
        r, term, info = env.step(a); rr, rterm, rinfo = ref.step(a)
        assert type(r) is float and type(term) is bool and isinstance(info, dict)
        assert set(info) == set(rinfo) and term == rterm
        assert abs(r - rr) <= 1e-6 * abs(rr) + 1e-12
        for k in rinfo:
            assert abs(info[k] - rinfo[k]) <= 1e-6 * abs(rinfo[k]) + 1e-9, k
        for name in ("percentage_reduction", "ess_charging", "ess_discharging", "q_pv", "power_reduction",
                     "current_ess_energy", "current_voltage"):
            mine, theirs = getattr(env, name), getattr(ref, name)
            assert list(mine) == list(theirs), name
            assert max(abs(mine[k] - theirs[k]) for k in theirs) < 1e-8, name
    assert env.steps == ref.steps == 97 and abs(env.cumulative_reward - ref.cumulative_reward) < 1e-6
    env.close()


def test_tester_accessors_and_manual_reset(cuda, profiles):
    """utils/tester.py:18-63: manual_reset(day, hour, quarter) then the _get_* accessor set per step."""
    env, ref = pair(cuda, profiles, 3)
    o, s = env.manual_reset(10, 7, 2); ro, rs = ref.manual_reset(10, 7, 2)
    assert np.max(np.abs(np.array(o) - np.array(ro))) < 1e-8 and np.max(np.abs(s - rs)) < 1e-8
    rng = np.random.RandomState(1)
    for t in range(12):
        a = rng.uniform(0.5, 1.0, 20).astype(np.float32)                          # translate_action output (Q5, Q6)
        env.step(a); ref.step(a)
        for name in ("_get_bus_v", "_get_bus_active", "_get_bus_reactive", "_get_pv_active", "_get_pv_reactive",
                     "_get_ess_energy", "_get_power_reduction", "_get_ess_charging", "_get_ess_discharging",
                     "_get_price"):
            mine, theirs = getattr(env, name)(), getattr(ref, name)()
            assert np.shape(mine) == np.shape(theirs), name
            assert np.max(np.abs(np.asarray(mine) - np.asarray(theirs))) < 1e-8, name
    env.close()


def test_sizes_and_static_api(cuda, profiles):
    env, ref = pair(cuda, profiles, 1)
    assert env.get_obs_size() == 144 and env.get_state_size() == 110
    assert env.get_total_actions() == 4 and env.get_num_of_agents() == 5
    assert env.get_env_info() == ref.get_env_info()
    assert np.array_equal(env.get_avail_actions(), ref.get_avail_actions())
    assert env.get_avail_agent_actions(2) == [1, 1, 1, 1]
    assert env.agent_ids == [5, 10, 15, 20, 25] and env.episode_limit == 96
    assert env.get_action().shape == (20,)
    assert set(env.base_powergrid) == set(ref.base_powergrid)
    assert env.get_obs_agent(1).shape == (144,)
    env.close()


def test_train_process_style_rollout_terminates_at_95(cuda, profiles):
    """model.py:208-254: reset, then step/get_obs once per step until done."""
    env, ref = pair(cuda, profiles, 5)
    env.reset(); ref.reset()
    rng = np.random.RandomState(2)
    for t in range(240):
        a = (0.5 * (np.clip(rng.normal(0.5, 0.6, 20), 0, 1) + 1.0)).astype(np.float32)   # util.py:121-129
        r, d, _ = env.step(a); rr, rd, _ = ref.step(a)
        o = env.get_obs(); ro = ref.get_obs()
        assert d == rd and abs(r - rr) <= 1e-6 * abs(rr) + 1e-12
        assert np.max(np.abs(np.array(o) - np.array(ro))) < 1e-8
        if d:
            break
    assert t + 1 == 95
    env.close()
