"""fp64 restatement of the safety-signal voltage predictor and the safety penalty (oracle side).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY: pinned on the reference's own scripts --
tests/golden/ref_predictor.npz holds the scenarios, the fitted regressor, the (unsaved) scalers and the
predictions produced by safety_signal/data_generation.py and train_safety_signal_model.py executed
unchanged (tests/golden/make_ref_golden.py; file I/O intercepted, Newton behind the Pyomo stand-in), and
tests/test_oracle_predictor.py checks this restatement against them: scenarios bit for bit, voltages
1e-10, predictions 1e-9.  Not pinned: the reference's shipped training data (Git-LFS pointers) and
scikit-learn 1.5.2 vs the version in this image.

  generate_scenarios   safety_signal/data_generation.py:27-58  (+-30 % loads -> simplified PF -> V)
  fit_pipeline         safety_signal/train_safety_signal_model.py:33-46,73
  predict              model.predict on scaled inputs, un-scaled back to p.u.
  rowsum_predict       the form madrl/models/safemaddpg.py:182-184,266,272 actually evaluates
  slack_penalty        optimal slack cost at fixed actions, safemaddpg.py:205-229
"""
import numpy as np

from . import pf_ref
from .ieee33 import create_network, tree_arrays


def generate_scenarios(n=1000, variation=0.3, seed=0, net=None):
    net = net or create_network()
    tree = tree_arrays(net)
    rng = np.random.RandomState(seed)
    buses = net['bus_numbers']
    P0 = np.array([net['active_power_demand'][b] for b in buses])
    Q0 = np.array([net['reactive_power_demand'][b] for b in buses])
    X, Y = [], []
    for _ in range(n):
        # the reference draws P and Q of one bus in two separate comprehensions (:31-36)
        p = P0 * (1 + rng.uniform(-variation, variation, len(buses)))
        q = Q0 * (1 + rng.uniform(-variation, variation, len(buses)))
        sol = pf_ref.solve_sweep(tree, p, q)
        X.append(np.stack([p, q], axis=1).reshape(-1))          # interleaved [P1, Q1, P2, Q2, ...] (:48-50)
        Y.append(np.sqrt(sol['v']))
    return np.array(X), np.array(Y)


def fit_pipeline(X, Y, test_size=0.2, random_state=42):
    from sklearn.linear_model import LinearRegression
    from sklearn.model_selection import train_test_split
    from sklearn.multioutput import MultiOutputRegressor
    from sklearn.preprocessing import MinMaxScaler
    sx, sy = MinMaxScaler(), MinMaxScaler()
    Xs, Ys = sx.fit_transform(X), sy.fit_transform(Y)
    Xtr, Xte, Ytr, Yte = train_test_split(Xs, Ys, test_size=test_size, random_state=random_state)
    model = MultiOutputRegressor(LinearRegression()).fit(Xtr, Ytr)
    return model, sx, sy


def predict(model, sx, sy, X):
    """sklearn's own fp64 path: scale, predict, un-scale."""
    return sy.inverse_transform(model.predict(sx.transform(np.asarray(X, dtype=np.float64))))


def affine_predict(A, c, X):
    return np.asarray(X, dtype=np.float64) @ np.asarray(A).T + np.asarray(c)


def rowsum_predict(coef, intercept, X):
    """safemaddpg.py:182-184 splits coef as [:, :33] / [:, 33:] and :266,272 multiplies the ROW SUMS
    by the bus's own P_net / Q_net."""
    X = np.asarray(X, dtype=np.float64)
    n = coef.shape[0]
    return coef[:, :n].sum(axis=1) * X[:, 0::2] + coef[:, n:].sum(axis=1) * X[:, 1::2] + intercept


def slack_penalty(V, v_min=0.9, v_max=1.1, weight=1000.0):
    V = np.asarray(V, dtype=np.float64)
    return weight * (np.maximum(0.0, v_min - V) + np.maximum(0.0, V - v_max)).sum(axis=-1)
