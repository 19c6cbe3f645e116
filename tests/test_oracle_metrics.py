"""CPU: the metric oracle (restated torchmetrics 0.9.0 collection of utils/scripts_utils.py:80-91) against scikit-learn's
independent implementations of the same definitions, plus its degenerate cases."""
import numpy as np
import pytest

from oracle import metrics_oracle as mx


@pytest.mark.parametrize("seed,rate", [(0, 0.3), (1, 0.01), (2, 0.9)])
def test_oracle_matches_sklearn(seed, rate):
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(seed)
    y = (rng.random(20000) < rate).astype(np.int32)
    pred = np.clip(0.55 * y + rng.random(20000) * 0.6, 0, 0.999)
    tau = 0.65
    got = mx.values_from_counts(mx.confusion_counts(pred, y, tau))
    hard = (pred >= tau).astype(np.int32)
    want = {
        "JaccardIndex": sk.jaccard_score(y, hard, average="macro"),
        "Precision": sk.precision_score(y, hard),
        "Recall": sk.recall_score(y, hard),
        "F1Score": sk.f1_score(y, hard),
        "FBetaScore": sk.fbeta_score(y, hard, beta=0.5),
    }
    for k in mx.NAMES:
        assert abs(float(got[k]) - want[k]) < 2e-6, (k, got[k], want[k])


def test_oracle_degenerate_and_accumulation():
    # nothing predicted, nothing to find: every ratio has a zero denominator -> 0; class 0 has IoU 1 -> Jaccard 0.5
    v = mx.values_from_counts([0, 0, 100, 0])
    assert float(v["Precision"]) == 0 and float(v["Recall"]) == 0 and float(v["F1Score"]) == 0 and float(v["FBetaScore"]) == 0
    assert float(v["JaccardIndex"]) == 0.5
    # threshold is inclusive (preds >= tau)
    assert mx.confusion_counts(np.array([0.65, 0.6499]), np.array([1, 1]), 0.65).tolist() == [1, 0, 0, 1]
    # float32 predictions are compared with float32(tau)
    p32 = np.array([np.float32(0.65)], dtype=np.float32)
    assert mx.confusion_counts(p32, np.array([1]), 0.65).tolist() == [1, 0, 0, 0]
    # forward returns batch values, compute the accumulated ones
    rng = np.random.default_rng(3)
    coll = mx.MetricCollectionOracle(0.65)
    tot = np.zeros(4, dtype=np.int64)
    for _ in range(3):
        p, y = rng.random(1000), (rng.random(1000) < 0.2).astype(np.int32)
        b = coll(p, y)
        c = mx.confusion_counts(p, y, 0.65)
        tot += c
        assert b == mx.values_from_counts(c)
    assert coll.compute() == mx.values_from_counts(tot)
    coll.reset()
    assert coll.state.sum() == 0


def test_collection_value_formulas_on_cpu_tensors():
    """the float32 formulas of the drop-in collection (scene-net_b200/utils/scripts_utils.py::_values) against the
    oracle for random and degenerate count vectors — pure torch arithmetic on [tp, fp, tn, fn], no kernel involved"""
    import torch
    from scenenet_b200.utils.scripts_utils import _values, METRIC_NAMES
    rng = np.random.default_rng(7)
    cases = [rng.integers(0, 10 ** rng.integers(1, 8), 4) for _ in range(200)]
    cases += [np.array(c) for c in ([0, 0, 0, 0], [0, 0, 100, 0], [5, 0, 0, 0], [0, 7, 0, 0], [0, 0, 0, 9], [1, 1, 1, 1])]
    for c in cases:
        got = _values(torch.tensor(c, dtype=torch.int64))
        want = mx.values_from_counts(c)
        assert list(got.keys()) == list(METRIC_NAMES)
        for k in METRIC_NAMES:
            assert got[k].dtype == torch.float32
            assert abs(float(got[k]) - float(want[k])) <= 1.2e-7, (c, k, float(got[k]), float(want[k]))
