"""CPU/GPU restatement of ``Encoder.forward`` / ``TransFuser.forward`` (model2_seq.py:473-597, 880-894)
in plain PyTorch, written against the *attribute names* the reference modules use
(``image_encoder.features``, ``lidar_encoder._model``, ``radar_encoder._model``, ``vel_emb1..4``,
``transformer1..4``, ``join``), so the same function runs on the reference's own ``Encoder`` object and
on the drop-in ``deepsense6g_tii_b200.Encoder`` (whose GPT math it replaces by ``fusion_ref``).

TEST INFRASTRUCTURE ONLY.  Pinned in ``tests/test_oracle.py`` against the live reference ``Encoder``.
"""
import torch

from . import fusion_ref as R


def normalize_imagenet(x):
    """model2_seq.py:36-45."""
    x = x.clone()
    x[:, 0] = (x[:, 0] / 255.0 - 0.485) / 0.229
    x[:, 1] = (x[:, 1] / 255.0 - 0.456) / 0.224
    x[:, 2] = (x[:, 2] / 255.0 - 0.406) / 0.225
    return x


def _gpt_params(gpt):
    return {k: v for k, v in gpt.named_parameters()}


def encoder_forward(enc, image_list, lidar_list, radar_list, gps, stage_autocast=False):
    """Restates model2_seq.py:473-597 (four fusion stages interleaved with the three ResNet trunks).
    stage_autocast=True evaluates ONLY the fusion stages under torch.autocast(bf16) (trunks stay fp32): the stock-PyTorch
    calibration for the bf16 tensor-core mode."""
    cfg = enc.config
    S = cfg.seq_len
    image_list = [normalize_imagenet(t) for t in image_list]
    bz, _, h, w = lidar_list[0].shape
    V = len(image_list) // S
    img = torch.stack(image_list, dim=1).view(bz * V * S, image_list[0].shape[1], h, w)
    lid = torch.stack(lidar_list, dim=1).view(bz * S, lidar_list[0].shape[1], h, w)
    rad = torch.stack(radar_list, dim=1).view(bz * S, radar_list[0].shape[1], h, w)
    ie, le, re_ = enc.image_encoder.features, enc.lidar_encoder._model, enc.radar_encoder._model

    def stem(m, x):
        return m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(x)))))

    feats = [stem(ie, img), stem(le, lid), stem(re_, rad)]
    g = gps
    for k, (vel, gpt) in enumerate(((enc.vel_emb1, enc.transformer1), (enc.vel_emb2, enc.transformer2),
                                    (enc.vel_emb3, enc.transformer3), (enc.vel_emb4, enc.transformer4))):
        if k > 0:
            layer = "layer%d" % (k + 1)
            feats = [getattr(ie, layer)(feats[0]), getattr(le, layer)(feats[1]), getattr(re_, layer)(feats[2])]
        g = vel(g)
        with torch.autocast(feats[0].device.type, dtype=torch.bfloat16, enabled=stage_autocast):
            feats, g = R.fusion_stage(_gpt_params(gpt), feats, g, cfg.n_head, S, cfg.vert_anchors, cfg.horz_anchors, V)
        feats = [f.float() if stage_autocast else f for f in feats]
        g = g.float() if stage_autocast else g
    pooled = [torch.flatten(ie.avgpool(feats[0]), 1).view(bz, V * S, -1),
              torch.flatten(le.avgpool(feats[1]), 1).view(bz, S, -1),
              torch.flatten(re_.avgpool(feats[2]), 1).view(bz, S, -1)]
    return torch.cat(pooled + [g], dim=1).sum(dim=1)


def transfuser_forward(model, image_list, lidar_list, radar_list, gps, stage_autocast=False):
    """model2_seq.py:880-894."""
    return model.join(encoder_forward(model.encoder, image_list, lidar_list, radar_list, gps, stage_autocast))
