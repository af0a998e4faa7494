"""Drop-in modules for the reference's GPT fusion path (model2_seq.py).

Same constructors, ``forward`` signatures, attribute names and ``state_dict`` keys as the reference
classes, so reference checkpoints load with ``strict=True`` (including the ``module.`` DataParallel
prefix handled by the caller, my_test.py:11):

  * ``SelfAttention`` / ``Block``   model2_seq.py:74-134   (parameter containers; math is fused)
  * ``GPT``                         model2_seq.py:175-287
  * ``ImageCNN`` / ``LidarEncoder`` model2_seq.py:12-72    (stock torchvision ResNets, out of scope)
  * ``Encoder``                     model2_seq.py:406-597
  * ``TransFuser``                  model2_seq.py:850-894  (wired to the GPT ``Encoder``)

The fusion stage itself (pool -> tokens -> 8 blocks -> ln_f -> upsample -> residual add, forward and
backward) runs in ``functional.FusionStageFn`` on the hand-written sm_100a kernels.  There is no
PyTorch fallback: on a CPU tensor or without ``libdsfuse.so`` the modules raise.

Extra, optional config attributes (duck-typed ``config`` object, config_seq.py:3-45):
  ``fusion_dtype``            torch.bfloat16 (default, tensor-core mode) or torch.float32 (parity mode)
  ``modality_missing``        None | 'image' | 'lidar' | 'radar' | 'lidar_radar'  (mambafuser_seq.py:361-391)
  ``modality_missing_type``   'zerolike' | 'randlike'
  ``missing_fast_path``       bool (default True): eval-mode cache of the stem output of a zeroed branch
  ``pretrained``              bool; torchvision ImageNet weights for the trunks (needs a local cache)
"""
import torch
import torch.nn.functional as F
from torch import nn

from .functional import fusion_stage, param_names, pooled_tail, pooled_tail_supported, stem_pack, stem_pack_supported


class SelfAttention(nn.Module):
    """Parameter container with the reference layout (model2_seq.py:79-91)."""

    def __init__(self, n_embd, n_head, attn_pdrop, resid_pdrop):
        super().__init__()
        if n_embd % n_head != 0:
            raise AssertionError("n_embd must be divisible by n_head")
        self.key = nn.Linear(n_embd, n_embd)
        self.query = nn.Linear(n_embd, n_embd)
        self.value = nn.Linear(n_embd, n_embd)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.resid_drop = nn.Dropout(resid_pdrop)
        self.proj = nn.Linear(n_embd, n_embd)
        self.n_head = n_head


class Block(nn.Module):
    """Parameter container with the reference layout (model2_seq.py:116-126)."""

    def __init__(self, n_embd, n_head, block_exp, attn_pdrop, resid_pdrop):
        super().__init__()
        self.ln1 = nn.LayerNorm(n_embd)
        self.ln2 = nn.LayerNorm(n_embd)
        self.attn = SelfAttention(n_embd, n_head, attn_pdrop, resid_pdrop)
        self.mlp = nn.Sequential(
            nn.Linear(n_embd, block_exp * n_embd),
            nn.ReLU(True),
            nn.Linear(block_exp * n_embd, n_embd),
            nn.Dropout(resid_pdrop),
        )


class GPT(nn.Module):
    """B200-native replacement of ``model2_seq.GPT`` (same ctor / forward / state_dict)."""

    def __init__(self, n_embd, n_head, block_exp, n_layer, vert_anchors, horz_anchors, seq_len,
                 embd_pdrop, attn_pdrop, resid_pdrop, config):
        super().__init__()
        self.n_embd = n_embd
        self.n_head = n_head
        self.n_layer = n_layer
        self.seq_len = seq_len
        self.vert_anchors = vert_anchors
        self.horz_anchors = horz_anchors
        self.config = config
        n_tok = (config.n_views + 2) * seq_len * vert_anchors * horz_anchors + 2
        self.pos_emb = nn.Parameter(torch.zeros(1, n_tok, n_embd))
        self.drop = nn.Dropout(embd_pdrop)
        self.blocks = nn.Sequential(*[Block(n_embd, n_head, block_exp, attn_pdrop, resid_pdrop) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(n_embd)
        self.block_size = seq_len
        self._pdrop = (float(embd_pdrop), float(attn_pdrop), float(resid_pdrop))
        # graph-safe dropout: a device-resident call counter (not part of the state dict) that a captured step bumps at
        # every replay; its value is XOR-ed into the seed inside the kernels
        self.register_buffer("_drop_counter", torch.zeros(1, dtype=torch.int64), persistent=False)
        self._drop_base_seed = None
        self._drop_capture = None  # tests: a dict that receives seed/step and the attention keep-bitmaps of the last call
        self.apply(self._init_weights)
        self._names = param_names(n_layer)

    def get_block_size(self):
        return self.block_size

    def _init_weights(self, module):
        # reference init law, model2_seq.py:207-214
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def configure_optimizers(self):
        """Same contract as the reference method (model2_seq.py:216-246): two optimizer parameter groups — weights of
        Linear layers with weight decay 0.01; every bias, the LayerNorm weights and ``pos_emb`` with weight decay 0 —
        each group in sorted parameter-name order."""
        decay, no_decay = [], ["pos_emb"]
        for mod_name, mod in self.named_modules():
            for p_name, _ in mod.named_parameters(recurse=False):
                full = "%s.%s" % (mod_name, p_name) if mod_name else p_name
                if p_name == "bias" or (p_name == "weight" and isinstance(mod, (nn.LayerNorm, nn.BatchNorm2d))):
                    no_decay.append(full)
                elif p_name == "weight" and isinstance(mod, (nn.Linear, nn.Conv2d)):
                    decay.append(full)
        table = dict(self.named_parameters())
        return [{"params": [table[n] for n in sorted(decay)], "weight_decay": 0.01},
                {"params": [table[n] for n in sorted(set(no_decay))], "weight_decay": 0.0}]

    def _flat_params(self):
        table = dict(self.named_parameters())
        return [table[n] for n in self._names]

    def _stage_cfg(self, residual):
        dropout = None
        if self.training and any(p > 0.0 for p in self._pdrop):
            # nn.Dropout semantics (model2_seq.py:104,109,125,272): active in train() only.  A fresh seed per call is
            # drawn from torch's CPU generator (so torch.manual_seed governs it); the masks themselves are Philox
            # functions of (seed, site, element) computed inside the kernels.
            if self._drop_counter.is_cuda and torch.cuda.is_current_stream_capturing():
                # The host-side seed is frozen into the captured kernels' arguments, so freshness must come from the device:
                # bump the counter (captured -> once per replay) and hand this call its own snapshot of it, which stays
                # valid for the backward even if the module is called again before that backward runs.
                if self._drop_base_seed is None:
                    self._drop_base_seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
                seed = self._drop_base_seed
                self._drop_counter.add_(1)
                seed_dev = self._drop_counter.clone()
            else:
                seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
                seed_dev = None
            # data-parallel replicas share torch.manual_seed (train2_seq.py:430-434): fold the rank in so that each replica
            # draws its own masks, like independent per-device generators do under nn.DataParallel
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                seed ^= ((dist.get_rank() + 1) * 0x9E3779B97F4A7C15) & ((1 << 62) - 1)
            dropout = dict(embd=self._pdrop[0], attn=self._pdrop[1], resid=self._pdrop[2], seed=seed, step=0, seed_dev=seed_dev)
            if self._drop_capture is not None:
                self._drop_capture.clear()
                self._drop_capture.update(seed=seed, step=0)
                dropout["capture"] = self._drop_capture
        return dict(dropout=dropout, seq_len=self.seq_len, n_views=self.config.n_views, vert_anchors=self.vert_anchors,
                    horz_anchors=self.horz_anchors, n_head=self.n_head, n_layer=self.n_layer,
                    compute_dtype=getattr(self.config, "fusion_dtype", torch.bfloat16), residual=residual,
                    grad_hook=getattr(self, "_grad_reducer", None), shadows=self._shadow_cfg())

    # ---------------------------------------------------------------- persistent bf16 weight shadows (optim.FusedAdamWEMA)
    def enable_persistent_shadows(self):
        """Keep the bf16 weight shadows of every block (plain + transposed, q/k/v fused; see ``functional._forward_bf16``) in
        buffers owned by this module instead of re-packing them at the start of every forward.  The forward re-packs only when
        a parameter changed behind the shadows' back (``_shadow_key``); ``optim.FusedAdamWEMA`` writes them as part of its
        update and calls ``mark_shadows_fresh``."""
        if getattr(self, "_shadow_bufs", None) is None:
            self._shadow_bufs = [None] * self.n_layer
            self._shadow_fresh_key = None

    def shadow_views(self, i):
        C = self.n_embd
        F = self.blocks[i].mlp[0].weight.shape[0]
        dev = self.pos_emb.device
        b = self._shadow_bufs[i]
        if b is None or b["flat"].device != dev:
            sizes = [3 * C * C, 3 * C * C, C * C, C * C, F * C, F * C, C * F, C * F]
            flat = torch.empty(sum(sizes), device=dev, dtype=torch.bfloat16)
            v, off = [], 0
            for n in sizes:
                v.append(flat[off:off + n])
                off += n
            b = dict(flat=flat, wqkv=v[0].view(3 * C, C), wqkv_t=v[1].view(C, 3 * C), wp=v[2].view(C, C), wp_t=v[3].view(C, C),
                     w1=v[4].view(F, C), w1_t=v[5].view(C, F), w2=v[6].view(C, F), w2_t=v[7].view(F, C),
                     bqkv=torch.empty(3 * C, device=dev, dtype=torch.float32))
            self._shadow_bufs[i] = b
            self._shadow_fresh_key = None
        return {k: t for k, t in b.items() if k != "flat"}

    def _shadow_key(self):
        return tuple((p.data_ptr(), p._version) for p in self._flat_params())

    def mark_shadows_fresh(self):
        self._shadow_fresh_key = self._shadow_key()

    def _shadow_cfg(self):
        if getattr(self, "_shadow_bufs", None) is None or getattr(self.config, "fusion_dtype", torch.bfloat16) != torch.bfloat16:
            return None
        views = [self.shadow_views(i) for i in range(self.n_layer)]
        return dict(views=views, fresh=self._shadow_fresh_key is not None and self._shadow_fresh_key == self._shadow_key(), owner=self)

    def set_grad_reducer(self, reducer):
        """Data-parallel training: a ``dist.OverlappedGradReducer`` that averages this GPT's gradients block by block
        while the backward is still running (bf16 path).  With it set, ``.grad`` of the GPT parameters is already
        averaged over ranks after ``backward()`` — do not all-reduce them again (exclude them from DDP)."""
        self._grad_reducer = reducer

    def forward(self, image_tensor, lidar_tensor, radar_tensor, gps):
        """Reference semantics (model2_seq.py:248-287): inputs are already pooled to the anchor grid,
        outputs are the un-tokenised anchor maps + the two GPS tokens."""
        return fusion_stage(self._stage_cfg(False), image_tensor, lidar_tensor, radar_tensor, gps, self._flat_params())

    def fuse(self, image_features, lidar_features, radar_features, gps_embd):
        """Whole stage on full-resolution trunk features (model2_seq.py:515-526): returns
        (image', lidar', radar', gps_out) with feat' = feat + upsample(GPT(pool(feat)))."""
        return fusion_stage(self._stage_cfg(True), image_features, lidar_features, radar_features, gps_embd, self._flat_params())


# --------------------------------------------------------------------------------------------- trunks
_IMAGENET_STATS = {}


def normalize_imagenet(x):
    """ImageNet mean/std on 0-255 input (model2_seq.py:36-45).  The constants are cached per (device, dtype): building
    them from Python lists on every call is a host-to-device copy, which a CUDA-graph capture of the step forbids."""
    key = (x.device, x.dtype)
    if key not in _IMAGENET_STATS:
        _IMAGENET_STATS[key] = (torch.tensor([0.485, 0.456, 0.406], dtype=x.dtype).view(1, 3, 1, 1).to(x.device),
                                torch.tensor([0.229, 0.224, 0.225], dtype=x.dtype).view(1, 3, 1, 1).to(x.device))
    mean, std = _IMAGENET_STATS[key]
    return (x / 255.0 - mean) / std


def _resnet(kind, pretrained):
    from torchvision import models
    ctor = getattr(models, kind)
    if pretrained:
        return ctor(weights="DEFAULT")
    return ctor(weights=None)


class ImageCNN(nn.Module):
    """ResNet-34 image trunk, fc removed (model2_seq.py:12-34)."""

    def __init__(self, c_dim, normalize=True, pretrained=False):
        super().__init__()
        self.normalize = normalize
        self.features = _resnet("resnet34", pretrained)
        self.features.fc = nn.Sequential()

    def forward(self, inputs):
        c = 0
        for x in inputs:
            if self.normalize:
                x = normalize_imagenet(x)
            c = c + self.features(x)
        return c


class LidarEncoder(nn.Module):
    """ResNet-18 trunk with an ``in_channels`` stem, fc removed (model2_seq.py:48-72)."""

    def __init__(self, num_classes=512, in_channels=2, pretrained=False):
        super().__init__()
        self._model = _resnet("resnet18", pretrained)
        self._model.fc = nn.Sequential()
        old = self._model.conv1
        self._model.conv1 = nn.Conv2d(in_channels, out_channels=old.out_channels, kernel_size=old.kernel_size,
                                      stride=old.stride, padding=old.padding, bias=old.bias)

    def forward(self, inputs):
        feats = 0
        for x in inputs:
            feats = feats + self._model(x)
        return feats


def _missing(cfg, name):
    mm = getattr(cfg, "modality_missing", None)
    return mm == name or (mm == "lidar_radar" and name in ("lidar", "radar"))


class Encoder(nn.Module):
    """Multi-scale fusion encoder (model2_seq.py:406-597) with the four GPT stages on dsfuse kernels."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        pre = bool(getattr(config, "pretrained", False))
        self.avgpool = nn.AdaptiveAvgPool2d((config.vert_anchors, config.horz_anchors))  # kept for state/attr parity; fused
        self.image_encoder = ImageCNN(512, normalize=True, pretrained=pre)
        self.lidar_encoder = LidarEncoder(num_classes=512, in_channels=1, pretrained=pre)
        self.radar_encoder = LidarEncoder(num_classes=512, in_channels=2 if getattr(config, "add_velocity", 0) else 1,
                                          pretrained=pre)
        self.vel_emb1 = nn.Linear(2, 64)
        self.vel_emb2 = nn.Linear(64, 128)
        self.vel_emb3 = nn.Linear(128, 256)
        self.vel_emb4 = nn.Linear(256, 512)

        def gpt(c):
            return GPT(n_embd=c, n_head=config.n_head, block_exp=config.block_exp, n_layer=config.n_layer,
                       vert_anchors=config.vert_anchors, horz_anchors=config.horz_anchors, seq_len=config.seq_len,
                       embd_pdrop=config.embd_pdrop, attn_pdrop=config.attn_pdrop, resid_pdrop=config.resid_pdrop,
                       config=config)

        self._stem_cache = {}  # missing-modality fast path, see _stem()
        self.transformer1 = gpt(64)
        self.transformer2 = gpt(128)
        self.transformer3 = gpt(256)
        self.transformer4 = gpt(512)

    def _stem(self, name, m, x):
        """conv1 -> bn1 -> relu -> maxpool -> layer1 of one trunk (model2_seq.py:495-512).  Missing-modality fast path
        (SURVEY.md §8f item 3): in eval mode a branch whose input was replaced by zeros (mambafuser_seq.py:418-420) yields
        the same feature map for every frame, so it is computed once for ONE frame, cached until any parameter or
        buffer of that stem changes, and broadcast — two of the three stems drop out of the missing-modality sweep."""
        def run(t):
            return m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(t)))))

        cfg = self.config
        if (self.training or not _missing(cfg, name) or getattr(cfg, "modality_missing_type", "zerolike") != "zerolike"
                or not getattr(cfg, "missing_fast_path", True)):
            return run(x)
        parts = [m.conv1, m.bn1, m.layer1]
        version = sum(int(t._version) for mod in parts for t in list(mod.parameters()) + list(mod.buffers()))
        key = (name, tuple(x.shape[1:]), x.dtype, x.device, x.is_contiguous(memory_format=torch.channels_last), torch.is_autocast_enabled())
        hit = self._stem_cache.get(key)
        if hit is None or hit[0] != version:
            with torch.no_grad():
                hit = (version, run(x[:1]))
            self._stem_cache[key] = hit
        one = hit[1]
        nhwc = one.is_contiguous(memory_format=torch.channels_last) and not one.is_contiguous()
        return one.expand(x.shape[0], *one.shape[1:]).contiguous(memory_format=torch.channels_last if nhwc else torch.contiguous_format)

    def _stack(self, frames, conv1, normalize):
        """The stacked input of one trunk (model2_seq.py:481-482, 491-493): ``normalize_imagenet`` per frame, ``torch.stack(dim=1)``,
        ``.view(B*S, C, H, W)``.  For plain fp32 CUDA frames this is one ``dsf_stem_pack`` launch that also writes the dtype
        (bf16 under autocast) and storage order (channels_last when conv1's weight is) cuDNN is about to ask for; any other input
        takes the reference's op sequence."""
        if getattr(self.config, "fused_stem_tail", True) and stem_pack_supported(frames):
            dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
            if dtype in (torch.float32, torch.bfloat16):
                wt = conv1.weight
                nhwc = wt.dim() == 4 and wt.is_contiguous(memory_format=torch.channels_last) and not wt.is_contiguous()
                return stem_pack(list(frames), normalize, dtype, nhwc)
        if normalize:
            frames = [normalize_imagenet(t) for t in frames]
        f0 = frames[0]
        # (the reference uses .view here, :491-493; reshape also accepts channels_last frames)
        return torch.stack(list(frames), dim=1).reshape(f0.shape[0] * len(frames), f0.shape[1], f0.shape[2], f0.shape[3])

    def _apply_missing(self, name, t):
        # semantics of mambafuser_seq.py:361-391, 418-420: replace the stacked input ahead of conv1
        if not _missing(self.config, name):
            return t
        if getattr(self.config, "modality_missing_type", "zerolike") == "randlike":
            return torch.rand_like(t)
        return torch.zeros_like(t)

    def forward(self, image_list, lidar_list, radar_list, gps, velocity=None, rebuild_modality_feat_list=None):
        cfg = self.config
        bz, _, h, w = lidar_list[0].shape
        cfg.n_views = len(image_list) // cfg.seq_len  # the reference mutates the shared config too (:489)
        S, V = cfg.seq_len, cfg.n_views
        ie, le, re_ = self.image_encoder.features, self.lidar_encoder._model, self.radar_encoder._model
        img = self._stack(image_list, ie.conv1, self.image_encoder.normalize)
        lid = self._stack(lidar_list, le.conv1, False)
        rad = self._stack(radar_list, re_.conv1, False)
        img, lid, rad = self._apply_missing("image", img), self._apply_missing("lidar", lid), self._apply_missing("radar", rad)

        f_img, f_lid, f_rad = self._stem("image", ie, img), self._stem("lidar", le, lid), self._stem("radar", re_, rad)
        g = self.vel_emb1(gps)
        f_img, f_lid, f_rad, g = self.transformer1.fuse(f_img, f_lid, f_rad, g)
        f_img, f_lid, f_rad = ie.layer2(f_img), le.layer2(f_lid), re_.layer2(f_rad)
        g = self.vel_emb2(g)
        f_img, f_lid, f_rad, g = self.transformer2.fuse(f_img, f_lid, f_rad, g)
        f_img, f_lid, f_rad = ie.layer3(f_img), le.layer3(f_lid), re_.layer3(f_rad)
        g = self.vel_emb3(g)
        f_img, f_lid, f_rad, g = self.transformer3.fuse(f_img, f_lid, f_rad, g)
        f_img, f_lid, f_rad = ie.layer4(f_img), le.layer4(f_lid), re_.layer4(f_rad)
        g = self.vel_emb4(g)
        f_img, f_lid, f_rad, g = self.transformer4.fuse(f_img, f_lid, f_rad, g)

        if getattr(cfg, "fused_stem_tail", True) and pooled_tail_supported(f_img, f_lid, f_rad, g):
            return pooled_tail(f_img, f_lid, f_rad, g, bz)  # avgpool + flatten + cat + sum (:581-595) in one launch
        p_img = torch.flatten(ie.avgpool(f_img), 1).view(bz, V * S, -1)
        p_lid = torch.flatten(le.avgpool(f_lid), 1).view(bz, S, -1)
        p_rad = torch.flatten(re_.avgpool(f_rad), 1).view(bz, S, -1)
        fused = torch.cat([p_img, p_lid, p_rad, g.to(p_img.dtype)], dim=1)  # (B, (V+2)S + 2, 512)
        return torch.sum(fused, dim=1)


class TransFuser(nn.Module):
    """``model2_seq.TransFuser`` (model2_seq.py:850-894) with the GPT ``Encoder`` wired in."""

    def __init__(self, config, device, pretrain_weight=False):
        super().__init__()
        self.device = device
        self.config = config
        self.pred_len = config.pred_len
        self.encoder = Encoder(config).to(self.device)
        self.join = nn.Sequential(
            nn.Linear(512, 256),
            nn.ReLU(inplace=True),
            nn.Linear(256, 128),
            nn.ReLU(inplace=True),
            nn.Linear(128, 64),
        ).to(self.device)
        if pretrain_weight:
            self.load_pretrained_weight()

    def load_pretrained_weight(self, path="mamba_fusion.pth"):
        self.load_state_dict(torch.load(path, map_location=self.device))

    def forward(self, image_list, lidar_list, radar_list, gps, rebuild_modality_feat_list=None, velocity=None):
        fused = self.encoder(image_list, lidar_list, radar_list, gps)
        return self.join(fused)
