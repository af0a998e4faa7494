"""Training-step utilities around the drop-in model: the pieces of ``Engine.train`` (train2_seq.py:94-136) a
caller needs to run ``BASELINE.json`` configs[2] ("full model2_seq training step bf16, DDP batch-sharded, focal loss +
EMA, synthetic data").

  * ``FocalLoss``   train2_seq.py:291-301  (torchvision ``sigmoid_focal_loss``, mean; stays stock PyTorch — north_star)
  * ``EMA``         train2_seq.py:303-334  same API (register / update / apply_shadow / restore) and the same arithmetic
                    ``shadow = (1 - decay) * param + decay * shadow``, but as ONE multi-tensor lerp per update instead of a
                    Python loop over ~300 parameter tensors
  * ``synthetic_batch``  value ranges of the dataset code (data2_seq.py:141,149,162-167,204-206,277-280)
  * ``train_step``  forward -> focal loss -> backward -> optimizer.step -> EMA update, the order of train2_seq.py:117-134

The trunks (torchvision ResNets), the GPS MLP, the loss and the optimizer are stock PyTorch; only the four fusion
stages inside ``TransFuser.encoder`` run on the dsfuse kernels.
"""
import math

import torch
from torch import nn


class FocalLoss(nn.Module):
    """``FocalLoss(gamma=2, alpha=0.25)`` of train2_seq.py:291-301: integer targets are one-hot encoded over 64 beams,
    soft targets are used as they are; sigmoid focal loss, mean reduction."""

    def __init__(self, gamma=2, alpha=0.25):
        super().__init__()
        self.gamma = gamma
        self.alpha = alpha

    def forward(self, input, target):
        from torchvision.ops import sigmoid_focal_loss
        if target.dim() == 1:
            target = torch.nn.functional.one_hot(target, num_classes=64)
        return sigmoid_focal_loss(input, target.float(), alpha=self.alpha, gamma=self.gamma, reduction="mean")


class EMA:
    """Exponential moving average of the trainable parameters (train2_seq.py:303-334), multi-tensor."""

    def __init__(self, model, decay):
        self.model = model
        self.decay = decay
        self.shadow = {}
        self.backup = {}

    def _named(self):
        return [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]

    def register(self):
        self.shadow = {n: p.data.clone() for n, p in self._named()}

    def update(self):
        if getattr(self, "_fused_updates", 0) > 0:   # optim.FusedAdamWEMA.step() has already folded this update into its launch
            self._fused_updates -= 1
            return
        named = self._named()
        if not named:
            return
        missing = [n for n, _ in named if n not in self.shadow]
        if missing:
            raise AssertionError("EMA.update before register(): %s" % missing[:3])
        # shadow <- decay * shadow + (1 - decay) * param  ==  lerp(shadow, param, 1 - decay)
        torch._foreach_lerp_([self.shadow[n] for n, _ in named], [p.data for _, p in named], 1.0 - self.decay)

    def apply_shadow(self):
        for n, p in self._named():
            if n not in self.shadow:
                raise AssertionError("EMA.apply_shadow before register(): %s" % n)
            self.backup[n] = p.data
            p.data = self.shadow[n]

    def restore(self):
        for n, p in self._named():
            if n not in self.backup:
                raise AssertionError("EMA.restore without apply_shadow(): %s" % n)
            p.data = self.backup[n]
        self.backup = {}


def synthetic_batch(batch, seq_len=5, size=256, generator=None, device="cpu", pin=False):
    """One synthetic training batch with the dataset's shapes and value ranges: camera uint8-valued floats in [0, 255]
    (data2_seq.py:141), LiDAR-BEV histogram in {0, .2, ..., 1} with ~95 % zeros (:204-206), radar maps uniform [0, 1]
    (:149), GPS = the same calibrated angle in both components of each of the 2 rows (:277-280), soft target
    1.25 * N(beam, 0.5) on +-5 beams (:162-167).  Returns (images, lidars, radars, gps, soft_target, beam_index)."""
    g = generator
    imgs = [torch.randint(0, 256, (batch, 3, size, size), generator=g).float() for _ in range(seq_len)]
    lids = []
    for _ in range(seq_len):
        occ = (torch.rand(batch, 1, size, size, generator=g) < 0.05).float()
        lids.append(occ * torch.randint(1, 6, (batch, 1, size, size), generator=g).float() / 5.0)
    rads = [torch.rand(batch, 2, size, size, generator=g) for _ in range(seq_len)]
    ang = (torch.rand(batch, 2, 1, generator=g) - 0.5) * math.pi
    gps = ang.expand(batch, 2, 2).contiguous()
    beam = torch.randint(0, 64, (batch,), generator=g)
    idx = torch.arange(64).view(1, 64).float()
    d = idx - beam.view(-1, 1).float()
    soft = 1.25 * torch.exp(-0.5 * (d / 0.5) ** 2) / (0.5 * math.sqrt(2 * math.pi))
    soft = soft * (d.abs() <= 5).float()
    out = [imgs, lids, rads, gps, soft, beam]

    def mv(t):
        if pin:
            t = t.pin_memory()
        return t.to(device, non_blocking=True) if str(device) != "cpu" else t

    return [mv(t) for t in imgs], [mv(t) for t in lids], [mv(t) for t in rads], mv(gps), mv(soft), mv(beam)


def train_step(model, batch, criterion, optimizer, ema=None, autocast_dtype=None, grad_sync=None):
    """One iteration of ``Engine.train`` (train2_seq.py:105-134): zero_grad(set_to_none) -> forward -> focal loss on the
    soft target -> backward -> optimizer.step -> ema.update.  ``autocast_dtype`` (e.g. torch.bfloat16) runs the stock
    trunks / MLPs under ``torch.autocast``; the fusion stages use ``config.fusion_dtype`` regardless.  ``grad_sync``: a
    ``dist.DataParallelGrads`` (one process per GPU): gradients are averaged over ranks before the optimizer step.  Returns
    the loss tensor (no host sync)."""
    imgs, lids, rads, gps, soft, _ = batch
    if grad_sync is not None:   # dist.DataParallelGrads: fixed gradient views of one flat bucket instead of fresh tensors
        grad_sync.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    if autocast_dtype is not None:
        with torch.autocast("cuda", dtype=autocast_dtype):
            pred = model(imgs, lids, rads, gps)
    else:
        pred = model(imgs, lids, rads, gps)
    loss = criterion(pred.float(), soft)
    loss.backward()
    if grad_sync is not None:
        grad_sync.sync()
    optimizer.step()
    if ema is not None:
        ema.update()
    return loss
