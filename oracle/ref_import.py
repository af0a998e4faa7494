"""Import the *unmodified* reference (``/root/reference``) with the two shims SURVEY.md §8c lists.

TEST INFRASTRUCTURE ONLY.  This module works only in the authoring container (the GPU box has no
``/root/reference``).  It is used to (1) validate the plain-torch restatement in
``oracle/fusion_ref.py`` and (2) generate the committed golden vectors under ``tests/golden/``
(``oracle/make_golden.py``).

Shims (documentation of how the reference is imported, not product code):
  * ``model2_seq.py:9`` does ``from mamba_ssm import Mamba`` at import time; ``mamba_ssm`` is not
    installed and the GPT path never touches it -> a stub module is registered first.
  * ``model2_seq.py:23,59`` build ``models.resnet34(weights=True)`` which would download ImageNet
    weights -> torchvision constructors are wrapped to build random-init nets (BASELINE configs say
    "random init").
  * ``model2_seq.py:861`` hard-wires the Mamba encoder into ``TransFuser`` and ``:889`` passes five
    arguments; ``Encoder.forward`` (``:473``) takes four -> a subclass drops the fifth.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DSF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model2_seq.py"))


_cached = None


def load_reference():
    """Returns (model2_seq module, GlobalConfig class)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torchvision

    if "mamba_ssm" not in sys.modules:
        stub = types.ModuleType("mamba_ssm")
        stub.Mamba = type("Mamba", (), {})
        sys.modules["mamba_ssm"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    r34, r18 = torchvision.models.resnet34, torchvision.models.resnet18
    if not getattr(r34, "_dsf_wrapped", False):
        def _r34(weights=None, **k):
            return r34(weights=None)

        def _r18(weights=None, **k):
            return r18(weights=None)

        _r34._dsf_wrapped = True
        _r18._dsf_wrapped = True
        torchvision.models.resnet34 = _r34
        torchvision.models.resnet18 = _r18
    import model2_seq as M
    from config_seq import GlobalConfig

    class EncoderGPT(M.Encoder):
        def forward(self, a, b, c, d, rebuild_modality_feat_list=None):
            return super().forward(a, b, c, d)

    M.EncoderGPT = EncoderGPT
    M.EncoderWithMamba = EncoderGPT
    _cached = (M, GlobalConfig)
    return _cached


def make_config(**kw):
    _, GlobalConfig = load_reference()
    base = dict(add_velocity=1, embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    base.update(kw)
    return GlobalConfig(**base)
