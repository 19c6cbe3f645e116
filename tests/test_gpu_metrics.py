"""GPU: confusion counts at a threshold (csrc/metrics.cu) and the drop-in metric collection (utils/scripts_utils.py
mirror of the reference's init_metrics) against the oracle restatement of torchmetrics 0.9.0 (oracle/metrics_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mx

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sb():
    import scenenet_b200 as sb
    return sb


@pytest.mark.parametrize("pdt", [torch.float64, torch.float32])
@pytest.mark.parametrize("ydt", [torch.int32, torch.float64, torch.float32, torch.uint8, torch.int64, torch.bool])
@pytest.mark.parametrize("n", [0, 1, 5, 4099, 1 << 20])
def test_counts_bit_exact(pdt, ydt, n):
    sb = _sb()
    g = torch.Generator().manual_seed(n + 7)
    y = (torch.rand(n, generator=g) < 0.1)
    pred = (0.5 * y + 0.6 * torch.rand(n, generator=g)).clamp(0, 0.999).to(pdt)
    if n > 4:
        pred[:4] = torch.tensor([0.65, 0.6499999, 0.6500001, 0.0]).to(pdt)  # the boundary itself
    yy = y.to(ydt)
    total = torch.tensor([5, 6, 7, 8], dtype=torch.int64, device=DEV)
    got = sb.ops.confusion_counts(pred.to(DEV), yy.to(DEV), 0.65, total=total)
    want = mx.confusion_counts(pred.numpy(), y.numpy(), 0.65)
    assert got.cpu().numpy().tolist() == want.tolist()
    assert total.cpu().numpy().tolist() == (want + np.array([5, 6, 7, 8])).tolist()


def test_collection_matches_oracle_over_an_epoch():
    sb = _sb()
    from scenenet_b200.utils.scripts_utils import init_metrics, METRIC_NAMES
    coll = init_metrics(tau=0.65).to(DEV)
    ora = mx.MetricCollectionOracle(0.65)
    g = torch.Generator().manual_seed(11)
    for step in range(4):
        y = (torch.rand((2, 1, 16, 16, 16), generator=g) < 0.05).double()
        pred = (0.6 * y + 0.5 * torch.rand(y.shape, generator=g)).clamp(0, 0.999)
        # the reference's call (lit_model_wrappers.py:170-171)
        out = coll(torch.flatten(pred.to(DEV)), torch.flatten(y.to(DEV)).to(torch.int))
        out.update()
        want = ora(pred.numpy(), y.numpy())
        assert list(out.keys()) == list(METRIC_NAMES)
        for k in METRIC_NAMES:
            assert out[k].dtype == torch.float32 and abs(float(out[k]) - float(want[k])) <= 1.2e-7, (step, k)
    res, want = coll.compute(), ora.compute()
    for k in METRIC_NAMES:
        assert abs(float(res[k]) - float(want[k])) <= 1.2e-7, k
    assert coll.confusion.cpu().numpy().tolist() == ora.state.tolist()
    coll.reset()
    assert int(coll.confusion.sum()) == 0
    empty = coll.compute()
    assert float(empty["Precision"]) == 0.0 and float(empty["JaccardIndex"]) == 0.0


def test_collection_on_model_output_full_size():
    """config-2 sized batch through the module, then the metric pass: counts equal torch's own thresholding"""
    sb = _sb()
    from scenenet_b200.utils.scripts_utils import init_metrics
    torch.manual_seed(0)
    model = sb.SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5)).to(DEV)
    g = torch.Generator().manual_seed(1234)
    x = (torch.rand((32, 1, 64, 64, 64), generator=g) < 0.016).double().to(DEV)
    y = (torch.rand((32, 1, 64, 64, 64), generator=g) < 3e-4).double().to(DEV)
    with torch.no_grad():
        pred = model(x)
    coll = init_metrics(0.65).to(DEV)
    coll(torch.flatten(pred), torch.flatten(y).to(torch.int))
    pos, tgt = pred.flatten() >= 0.65, y.flatten() != 0
    want = [int((pos & tgt).sum()), int((pos & ~tgt).sum()), int((~pos & ~tgt).sum()), int((~pos & tgt).sum())]
    assert coll.confusion.cpu().tolist() == want
    assert sum(want) == pred.numel()
