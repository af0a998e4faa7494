"""CPU tests of the training-step utilities (deepsense6g_tii_b200/train.py) against restatements of the reference's
FocalLoss / EMA (train2_seq.py:291-334) and the dataset value ranges (data2_seq.py)."""
import math

import torch
from torch import nn

from deepsense6g_tii_b200.train import EMA, FocalLoss, synthetic_batch


class _RefEMA:
    """The reference's per-parameter Python loop (train2_seq.py:303-334), restated."""

    def __init__(self, model, decay):
        self.model, self.decay, self.shadow, self.backup = model, decay, {}, {}

    def register(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                self.shadow[n] = p.data.clone()

    def update(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                self.shadow[n] = ((1.0 - self.decay) * p.data + self.decay * self.shadow[n]).clone()


def test_ema_multi_tensor_matches_reference_loop():
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(8, 16), nn.LayerNorm(16), nn.Linear(16, 4))
    m[1].bias.requires_grad_(False)  # frozen parameters are skipped, as in the reference
    a, b = EMA(m, 0.999), _RefEMA(m, 0.999)
    a.register(); b.register()
    assert set(a.shadow) == set(b.shadow) and "1.bias" not in a.shadow
    for step in range(5):
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.1 * torch.randn_like(p))
        a.update(); b.update()
    for n in a.shadow:
        assert torch.allclose(a.shadow[n], b.shadow[n], rtol=1e-6, atol=1e-7), n
    # apply_shadow / restore swap the tensors in and out (validate with EMA weights, train2_seq.py:159-160, 220-221)
    before = {n: p.data.clone() for n, p in m.named_parameters()}
    a.apply_shadow()
    assert torch.equal(m[0].weight.data, a.shadow["0.weight"])
    a.restore()
    for n, p in m.named_parameters():
        assert torch.equal(p.data, before[n])


def test_focal_loss_matches_formula_for_soft_and_index_targets():
    torch.manual_seed(1)
    x = torch.randn(6, 64)
    soft = torch.rand(6, 64)

    def formula(x, t, alpha=0.25, gamma=2.0):
        p = torch.sigmoid(x)
        ce = -(t * torch.log(p) + (1 - t) * torch.log(1 - p))
        pt = p * t + (1 - p) * (1 - t)
        return ((alpha * t + (1 - alpha) * (1 - t)) * ce * (1 - pt) ** gamma).mean()

    crit = FocalLoss()
    assert torch.allclose(crit(x, soft), formula(x, soft), rtol=1e-5, atol=1e-7)
    idx = torch.randint(0, 64, (6,))
    assert torch.allclose(crit(x, idx), formula(x, torch.nn.functional.one_hot(idx, 64).float()), rtol=1e-5, atol=1e-7)


def test_synthetic_batch_shapes_and_value_ranges():
    g = torch.Generator().manual_seed(2)
    imgs, lids, rads, gps, soft, beam = synthetic_batch(3, seq_len=5, size=32, generator=g)
    assert len(imgs) == len(lids) == len(rads) == 5
    assert imgs[0].shape == (3, 3, 32, 32) and lids[0].shape == (3, 1, 32, 32) and rads[0].shape == (3, 2, 32, 32)
    assert gps.shape == (3, 2, 2) and soft.shape == (3, 64) and beam.shape == (3,)
    assert imgs[0].min() >= 0 and imgs[0].max() <= 255 and torch.equal(imgs[0], imgs[0].round())
    vals = set((lids[0] * 5).round().unique().tolist())
    assert vals <= {0.0, 1.0, 2.0, 3.0, 4.0, 5.0} and float((lids[0] == 0).float().mean()) > 0.85
    assert rads[0].min() >= 0 and rads[0].max() <= 1
    assert torch.equal(gps[:, :, 0], gps[:, :, 1]) and gps.abs().max() <= math.pi / 2
    # soft target: peak 1.25 * N(0; 0.5) at the beam, zero beyond +-5 beams
    peak = 1.25 / (0.5 * math.sqrt(2 * math.pi))
    for b in range(3):
        assert abs(float(soft[b, beam[b]]) - peak) < 1e-5
        far = (torch.arange(64) - beam[b]).abs() > 5
        assert float(soft[b][far].abs().max()) == 0.0


def test_fused_optimizer_refuses_cpu_parameters():
    """No CPU fallback on the optimizer side either (SURVEY.md §8(f)1 kernel lives in libdsfuse.so)."""
    import pytest
    from deepsense6g_tii_b200.optim import FusedAdamWEMA
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdamWEMA([p]).step()
