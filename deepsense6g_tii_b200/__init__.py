"""deepsense6g_tii_b200 — B200-native (sm_100a) GPT fusion stage of szy4017/DeepSense6G_TII.

Only the hot path named by BASELINE.json is here: drop-in ``GPT`` / ``Encoder`` / ``TransFuser``
modules (reference API, model2_seq.py) over the hand-written CUDA kernels in ``csrc/`` reached
through the C ABI in ``include/dsfuse.h``.
"""
from .modules import GPT, Block, Encoder, ImageCNN, LidarEncoder, SelfAttention, TransFuser, normalize_imagenet  # noqa: F401
from .functional import fusion_stage, param_names  # noqa: F401

__all__ = ["GPT", "Block", "SelfAttention", "Encoder", "ImageCNN", "LidarEncoder", "TransFuser",
           "normalize_imagenet", "fusion_stage", "param_names"]
