"""Generates the committed golden fixtures from the INDEPENDENT oracle (dense Newton power
flow + the Python env restatement), never from the CUDA path or its C mirror.

The reference itself cannot be run here (pyomo / ipopt absent) and ships no vectors, so these
fixtures pin the *oracle's math* and guard against regressions; the literature known answers
(K1) and the run_pf.py operating point (K2) are asserted separately in tests/test_oracle_pf.py.

    PYTHONPATH=. python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "safe-marl_b200")]

from oracle import env_ref, ieee33, pf_ref  # noqa: E402
from flexgpu.config import DEFAULT_ENV_ARGS  # noqa: E402
from flexgpu.network import Network, create_network  # noqa: E402
from flexgpu.profiles import synthetic_profiles  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pf_golden():
    net = ieee33.create_network()
    tree = ieee33.tree_arrays(net)
    rng = np.random.RandomState(42)
    base_p = np.array([net['active_power_demand'][n] for n in net['bus_numbers']])
    base_q = np.array([net['reactive_power_demand'][n] for n in net['bus_numbers']])
    # +-30 % scenarios as safety_signal/data_generation.py:31-36, plus heavier/lighter loadings
    scen_p, scen_q = [base_p.copy()], [base_q.copy()]
    for k in range(47):
        s = 1.0 if k < 32 else (0.2 + 0.1 * (k - 32))
        scen_p.append(s * base_p * (1 + rng.uniform(-0.3, 0.3, 33)))
        scen_q.append(s * base_q * (1 + rng.uniform(-0.3, 0.3, 33)))
    out = dict(p=[], q=[], V=[], P=[], Q=[], ell=[])
    for p, q in zip(scen_p, scen_q):
        sol = pf_ref.solve_newton(tree, p, q)
        out['p'].append(p[1:]); out['q'].append(q[1:])
        out['V'].append(np.sqrt(sol['v'])); out['P'].append(sol['P'][1:]); out['Q'].append(sol['Q'][1:])
        out['ell'].append(sol['ell'][1:])
    np.savez_compressed(os.path.join(HERE, "pf_golden.npz"), **{k: np.array(v) for k, v in out.items()})


def env_golden():
    network = Network(create_network(DEFAULT_ENV_ARGS))
    prof = synthetic_profiles(network, 5, T=3000, seed=0)
    for tag, maker in (("normal", lambda r: r.normal(0, 0.5, 20)),                       # run_env.py:86
                       ("train", lambda r: (0.5 * (np.clip(r.normal(0.5, 0.6, 20), 0, 1) + 1.0)).astype(np.float32))):  # util.py:121-129
        env = env_ref.RefFlexEnv(dict(env_ref.DEFAULT_ARGS), ieee33.create_network(), prof.as_dict(),
                                 rng=np.random.RandomState(0))
        obs0, state0 = env.reset()
        d = env._last_reset_draw
        s0 = d['start']
        n_rows = env.episode_limit + env.history + 1
        rec = dict(P=prof.P[s0:s0 + n_rows], Q=prof.Q[s0:s0 + n_rows], PV=prof.PV[s0:s0 + n_rows],
                   price=prof.price[s0:s0 + n_rows], e0=d['e0'], a0=d['a0'],
                   obs0=np.array(obs0), state0=state0, V0=env._get_bus_v(), E0=env._get_ess_energy())
        r = np.random.RandomState(7)
        acts, rew, done, info, V, E, setp, states, obs = [], [], [], [], [], [], [], [], {}
        keys = ['reward', 'revenue', 'der_cost', 'ess_cost', 'discomfort_penalty', 'voltage_penalty', 'cumulative_reward']
        for t in range(95):
            a = maker(r)
            rr, dd, ii = env.step(a)
            o = env.get_obs()                                  # one call per step, as model.py:223
            acts.append(np.asarray(a, dtype=np.float64)); rew.append(rr); done.append(dd)
            info.append([ii[k] for k in keys]); V.append(env._get_bus_v()); E.append(env._get_ess_energy())
            setp.append([env._get_power_reduction(), env._get_ess_charging(), env._get_ess_discharging(),
                         env._get_pv_reactive()])
            states.append(env.get_state())
            if t in (0, 1, 2, 30, 94):
                obs[t] = np.array(o)
        rec.update(actions=np.array(acts), actions_f32=np.array([tag == "train"]), reward=np.array(rew),
                   done=np.array(done), info=np.array(info), V=np.array(V), E=np.array(E), setp=np.array(setp),
                   state=np.array(states), obs_steps=np.array(sorted(obs)), obs=np.array([obs[k] for k in sorted(obs)]))
        np.savez_compressed(os.path.join(HERE, f"env_golden_{tag}.npz"), **rec)


def predictor_golden():
    """scikit-learn's own fit + predict (the reference's pipeline, train_safety_signal_model.py:33-46,73)
    on 1000 scenarios generated as data_generation.py:27-58 does, oracle power flow in place of IPOPT."""
    from oracle import predictor_ref
    X, Y = predictor_ref.generate_scenarios(1000, seed=0)
    model, sx, sy = predictor_ref.fit_pipeline(X, Y)
    rng = np.random.RandomState(3)
    Xt = X[rng.choice(len(X), 48, replace=False)] * (1 + rng.uniform(-0.6, 0.6, (48, 66)))   # some far outside the fit range
    Xt[40:] *= 3.0                                                                         # heavy loading -> under-voltage
    Vt = predictor_ref.predict(model, sx, sy, Xt)
    coef = np.array([e.coef_ for e in model.estimators_]); icpt = np.array([e.intercept_ for e in model.estimators_])
    np.savez_compressed(os.path.join(HERE, "predictor_golden.npz"), coef=coef, intercept=icpt, x_scale=sx.scale_,
                        x_min=sx.min_, y_scale=sy.scale_, y_min=sy.min_, X=Xt, V=Vt,
                        penalty=predictor_ref.slack_penalty(Vt), V_rowsum=predictor_ref.rowsum_predict(coef, icpt, Xt))


if __name__ == "__main__":
    pf_golden()
    env_golden()
    predictor_golden()
    print("golden fixtures written to", HERE)
