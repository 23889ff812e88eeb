"""CPU: pins the predictor / replay oracle against scikit-learn's own pipeline (the golden fixture
was produced by sklearn fit + predict, tests/golden/make_golden.py) and the reference's
TransReplayBuffer semantics (utils/replay_buffer.py:3-30)."""
import os

import numpy as np
import pytest

from oracle import predictor_ref, replay_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "predictor_golden.npz")


def fold(g):
    """The host-side folding flexgpu.predictor.VoltagePredictor.from_linear_model performs."""
    W, b = g["coef"], g["intercept"]
    A = W * g["x_scale"][None, :] / g["y_scale"][:, None]
    c = (W @ g["x_min"] + b - g["y_min"]) / g["y_scale"]
    return A, c


def test_folded_affine_map_equals_sklearn_predict():
    g = np.load(GOLD)
    A, c = fold(g)
    V = predictor_ref.affine_predict(A, c, g["X"])
    assert np.max(np.abs(V - g["V"])) < 1e-12
    assert np.max(np.abs(predictor_ref.slack_penalty(V) - g["penalty"])) < 1e-8
    assert (g["penalty"] == 0.0).sum() >= 5 and g["penalty"].max() > 1000.0             # both regimes are in the fixture
    assert abs(g["V"][:, 0] - 1.0).max() < 1e-12                                     # slack bus: constant column


def test_pipeline_reproduces_itself():
    X, Y = predictor_ref.generate_scenarios(60, seed=0)
    m, sx, sy = predictor_ref.fit_pipeline(X, Y)
    P = predictor_ref.predict(m, sx, sy, X)
    assert X.shape == (60, 66) and Y.shape == (60, 33) and np.max(np.abs(P - Y)) < 5e-3
    assert np.all(X[:, 0] == 0) and np.all(X[:, 1] == 0)                             # slack bus has no load
    assert np.allclose(X[:, 2::2].mean(0) / X[:, 2::2].mean(0), 1.0)


def test_rowsum_quirk_form_differs_from_the_proper_map():
    g = np.load(GOLD)
    V = predictor_ref.rowsum_predict(g["coef"], g["intercept"], g["X"])
    assert np.array_equal(V, g["V_rowsum"]) and np.max(np.abs(V - g["V"])) > 1e-3      # it is a different model


def test_replay_fifo_and_contiguous_window():
    buf = replay_ref.RefTransReplayBuffer(5)
    for i in range(8):
        buf.add_experience({"x": np.array([i], dtype=np.float32)})
    assert len(buf) == 5 and [int(t["x"][0]) for t in buf.buffer] == [3, 4, 5, 6, 7]    # oldest popped (:23-27)
    rng = np.random.RandomState(0)
    for _ in range(20):
        batch, start = buf.get_batch(3, rng)
        assert 0 <= start <= 2 and [int(t["x"][0]) for t in batch] == [3 + start, 4 + start, 5 + start]   # consecutive (:17-20)
    buf.clear()
    assert len(buf) == 0
    with pytest.raises(ValueError):
        buf.get_batch(1, rng)
