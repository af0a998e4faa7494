"""Generate ``tests/golden/*.npz`` by executing the UNMODIFIED reference classes on CPU fp32.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs ``/root/reference``):

    python -m oracle.make_golden

Every fixture stores: the reference ``GPT.state_dict()`` (perturbed after the reference's own init so
biases / LayerNorm affine / pos_emb are non-trivial), the inputs, the reference outputs, and the
gradients of ``loss = sum_i <out_i, probe_i>`` w.r.t. inputs and parameters (probes stored too).
The reference calls exercised: ``model2_seq.GPT`` (model2_seq.py:175-287),
``nn.AdaptiveAvgPool2d`` (:414), ``F.interpolate(..., mode='bilinear')`` + add (:521-526).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _perturb(gpt, gen):
    with torch.no_grad():
        for name, prm in gpt.named_parameters():
            if name == "pos_emb" or name.endswith("bias"):
                prm.add_(torch.randn(prm.shape, generator=gen) * 0.05)
            elif "ln" in name and name.endswith("weight"):
                prm.add_(torch.randn(prm.shape, generator=gen) * 0.1)


def gpt_case(name, C, n_head, L, A, S, B, seed, scale=None):
    """GPT-level (scale None) or stage-level (pool -> GPT -> upsample -> add) fixture."""
    M, _ = ref_import.load_reference()
    cfg = ref_import.make_config(seq_len=S, vert_anchors=A, horz_anchors=A, n_head=n_head, n_layer=L, n_views=1)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    gpt = M.GPT(C, n_head, cfg.block_exp, L, A, A, S, 0.0, 0.0, 0.0, cfg)
    _perturb(gpt, gen)
    gpt.train()
    H = A * (scale or 1)
    feats = [torch.randn(B * S, C, H, H, generator=gen).requires_grad_(True) for _ in range(3)]
    gps = torch.randn(B, 2, C, generator=gen).requires_grad_(True)
    if scale is None:
        outs = gpt(feats[0], feats[1], feats[2], gps)
    else:
        pool = torch.nn.AdaptiveAvgPool2d((A, A))
        io, lo, ro, go = gpt(pool(feats[0]), pool(feats[1]), pool(feats[2]), gps)
        if scale > 1:
            io, lo, ro = [F.interpolate(t, scale_factor=scale, mode="bilinear") for t in (io, lo, ro)]
        outs = (feats[0] + io, feats[1] + lo, feats[2] + ro, go)
    probes = [torch.randn(o.shape, generator=gen) for o in outs]
    loss = sum((o * pr).sum() for o, pr in zip(outs, probes))
    loss.backward()
    d = {"meta": np.array([C, n_head, L, A, S, B, scale or 0], dtype=np.int64)}
    for k, v in gpt.state_dict().items():
        d["param/" + k] = v.detach().numpy()
    for k, prm in gpt.named_parameters():
        d["gparam/" + k] = prm.grad.numpy()
    for nm, t in zip(("img", "lidar", "radar"), feats):
        d["in/" + nm] = t.detach().numpy()
        d["gin/" + nm] = t.grad.numpy()
    d["in/gps"] = gps.detach().numpy()
    d["gin/gps"] = gps.grad.numpy()
    for nm, o, pr in zip(("img", "lidar", "radar", "gps"), outs, probes):
        d["out/" + nm] = o.detach().numpy()
        d["probe/" + nm] = pr.numpy()
    d["loss"] = np.array(loss.item(), dtype=np.float64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def dropout_case(name, C, n_head, L, A, S, B, seed, p_drop=0.1):
    """GPT-level fixture with the reference in train() mode and all four nn.Dropout sites at ``p_drop``
    (config_seq.py:39-41; model2_seq.py:104,109,125,272).  The Bernoulli draws of every nn.Dropout module are
    recovered with forward hooks (mask = (out != 0) / (1 - p); where the input is exactly 0 the mask is
    irrelevant) and stored next to the outputs / gradients, so the restatement can be checked on the same draws."""
    M, _ = ref_import.load_reference()
    cfg = ref_import.make_config(seq_len=S, vert_anchors=A, horz_anchors=A, n_head=n_head, n_layer=L, n_views=1)
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    gpt = M.GPT(C, n_head, cfg.block_exp, L, A, A, S, p_drop, p_drop, p_drop, cfg)
    _perturb(gpt, gen)
    gpt.train()
    masks = {}

    def hook(key):
        def fn(mod, inp, out):
            masks[key] = (out != 0).to(torch.float32) / (1.0 - mod.p)
        return fn

    gpt.drop.register_forward_hook(hook("embd"))
    for i, blk in enumerate(gpt.blocks):
        blk.attn.attn_drop.register_forward_hook(hook("attn.%d" % i))
        blk.attn.resid_drop.register_forward_hook(hook("proj.%d" % i))
        blk.mlp[3].register_forward_hook(hook("mlp.%d" % i))
    feats = [torch.randn(B * S, C, A, A, generator=gen).requires_grad_(True) for _ in range(3)]
    gps = torch.randn(B, 2, C, generator=gen).requires_grad_(True)
    outs = gpt(feats[0], feats[1], feats[2], gps)
    probes = [torch.randn(o.shape, generator=gen) for o in outs]
    loss = sum((o * pr).sum() for o, pr in zip(outs, probes))
    loss.backward()
    d = {"meta": np.array([C, n_head, L, A, S, B, 0], dtype=np.int64)}
    for k, v in gpt.state_dict().items():
        d["param/" + k] = v.detach().numpy()
    for k, prm in gpt.named_parameters():
        d["gparam/" + k] = prm.grad.numpy()
    for nm, t in zip(("img", "lidar", "radar"), feats):
        d["in/" + nm] = t.detach().numpy()
        d["gin/" + nm] = t.grad.numpy()
    d["in/gps"] = gps.detach().numpy()
    d["gin/gps"] = gps.grad.numpy()
    for nm, o, pr in zip(("img", "lidar", "radar", "gps"), outs, probes):
        d["out/" + nm] = o.detach().numpy()
        d["probe/" + nm] = pr.numpy()
    for k, m in masks.items():
        d["mask/" + k] = m.numpy()
    d["loss"] = np.array(loss.item(), dtype=np.float64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "masks:", sorted(masks))


def op_case():
    """Operator-level fixtures: AdaptiveAvgPool2d((A,A)) and bilinear interpolate at scales 2/4/8."""
    gen = torch.Generator().manual_seed(7)
    d = {}
    for scale in (1, 2, 4, 8):
        x = torch.randn(3, 5, 4 * scale, 4 * scale, generator=gen)
        d["pool_in/%d" % scale] = x.numpy()
        d["pool_out/%d" % scale] = torch.nn.AdaptiveAvgPool2d((4, 4))(x).numpy()
        if scale > 1:
            y = torch.randn(2, 3, 4, 4, generator=gen)
            d["up_in/%d" % scale] = y.numpy()
            d["up_out/%d" % scale] = F.interpolate(y, scale_factor=scale, mode="bilinear").numpy()
    # non-divisible adaptive pool (not used by the model, pins the window rule)
    x = torch.randn(2, 3, 10, 7, generator=gen)
    d["pool_in/ragged"] = x.numpy()
    d["pool_out/ragged"] = torch.nn.AdaptiveAvgPool2d((4, 4))(x).numpy()
    path = os.path.join(OUT, "ops.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--dropout-only" in sys.argv:  # added later; leaves the earlier fixtures byte-identical
        dropout_case("gpt_tiny_dropout", C=32, n_head=4, L=2, A=2, S=2, B=2, seed=21)
        return
    op_case()
    gpt_case("gpt_tiny", C=32, n_head=4, L=2, A=2, S=2, B=2, seed=11)
    gpt_case("gpt_c64_t962", C=64, n_head=4, L=2, A=8, S=5, B=1, seed=12)
    gpt_case("stage_tiny_s4", C=32, n_head=4, L=2, A=2, S=2, B=2, seed=13, scale=4)
    gpt_case("stage_tiny_s8", C=16, n_head=4, L=1, A=2, S=2, B=1, seed=14, scale=8)
    gpt_case("stage_tiny_s2", C=32, n_head=2, L=1, A=4, S=1, B=2, seed=15, scale=2)
    gpt_case("stage_tiny_s1", C=32, n_head=4, L=1, A=2, S=2, B=2, seed=16, scale=1)
    dropout_case("gpt_tiny_dropout", C=32, n_head=4, L=2, A=2, S=2, B=2, seed=21)


if __name__ == "__main__":
    main()
