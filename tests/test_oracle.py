"""CPU tests that pin the oracle restatement (oracle/fusion_ref.py).

1. against the committed golden vectors generated from the reference classes
   (oracle/make_golden.py; reference model2_seq.py:175-287, :414, :521-526);
2. against the reference itself when /root/reference is present (authoring container only).
"""
import numpy as np
import pytest
import torch

from conftest import GPT_CASES, STAGE_CASES, load_golden, rel_err, assert_close, GOLDEN
from oracle import fusion_ref as R
from oracle import ref_import

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order


def _run_oracle(g):
    cfg = g["cfg"]
    p = {k: v.clone().requires_grad_(True) for k, v in g["param"].items()}
    ins = {k: v.clone().requires_grad_(True) for k, v in g["in"].items()}
    if cfg["scale"] == 0:
        outs = R.gpt_forward(p, ins["img"], ins["lidar"], ins["radar"], ins["gps"], cfg["n_head"], cfg["S"])
    else:
        (a, b, c), gps = R.fusion_stage(p, (ins["img"], ins["lidar"], ins["radar"]), ins["gps"],
                                        cfg["n_head"], cfg["S"], cfg["A"], cfg["A"])
        outs = (a, b, c, gps)
    names = ("img", "lidar", "radar", "gps")
    loss = sum((o * g["probe"][n]).sum() for o, n in zip(outs, names))
    loss.backward()
    return dict(zip(names, outs)), p, ins, loss


@pytest.mark.parametrize("case", GPT_CASES + STAGE_CASES)
def test_oracle_matches_golden(case):
    g = load_golden(case)
    outs, p, ins, loss = _run_oracle(g)
    for n, o in outs.items():
        assert o.shape == g["out"][n].shape
        assert rel_err(o, g["out"][n]) < TOL, n
    for n, t in ins.items():
        assert rel_err(t.grad, g["gin"][n]) < TOL, n
    for n, t in p.items():
        assert_close(t.grad, g["gparam"][n], 5e-5, 1e-7, n)
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))


def test_oracle_dropout_matches_golden_on_the_reference_draws():
    """Reference GPT in train() mode, p = 0.1 at all four nn.Dropout sites: given the Bernoulli draws recorded from
    the reference run, the restatement reproduces outputs and every gradient."""
    g = load_golden("gpt_tiny_dropout")
    cfg = g["cfg"]
    assert set(g["mask"]) == {"embd"} | {"%s.%d" % (k, i) for k in ("attn", "proj", "mlp") for i in range(cfg["L"])}
    for m in g["mask"].values():  # 0 or 1/(1-p), roughly 10 % dropped
        assert set(torch.unique(m).tolist()) <= {0.0, float(torch.tensor(1.0) / torch.tensor(0.9))}
    p = {k: v.clone().requires_grad_(True) for k, v in g["param"].items()}
    ins = {k: v.clone().requires_grad_(True) for k, v in g["in"].items()}
    outs = R.gpt_forward(p, ins["img"], ins["lidar"], ins["radar"], ins["gps"], cfg["n_head"], cfg["S"], masks=g["mask"])
    names = ("img", "lidar", "radar", "gps")
    sum((o * g["probe"][n]).sum() for o, n in zip(outs, names)).backward()
    for o, n in zip(outs, names):
        assert rel_err(o, g["out"][n]) < TOL, n
    for n, t in ins.items():
        assert rel_err(t.grad, g["gin"][n]) < TOL, n
    for n, t in p.items():
        assert_close(t.grad, g["gparam"][n], 5e-5, 1e-7, n)
    # and the masks matter: without them the result is far away
    with torch.no_grad():
        plain = R.gpt_forward(g["param"], *[g["in"][n] for n in names], cfg["n_head"], cfg["S"])
    assert rel_err(plain[0], g["out"]["img"]) > 1e-2


def test_ops_golden():
    z = np.load(GOLDEN + "/ops.npz")
    for s in (1, 2, 4, 8):
        x = torch.from_numpy(z["pool_in/%d" % s])
        assert rel_err(R.anchor_pool(x, 4, 4), torch.from_numpy(z["pool_out/%d" % s])) < 1e-6
        if s > 1:
            y = torch.from_numpy(z["up_in/%d" % s])
            assert rel_err(R.bilinear_upsample(y, s), torch.from_numpy(z["up_out/%d" % s])) < 1e-6
    x = torch.from_numpy(z["pool_in/ragged"])
    assert rel_err(R.anchor_pool(x, 4, 4), torch.from_numpy(z["pool_out/ragged"])) < 1e-6


def test_token_order_and_count():
    # T = (V+2)*S*A*A + 2 (model2_seq.py:189); GPS tokens are the last two (:270)
    B, S, A, C = 2, 5, 8, 8
    img = torch.zeros(B * S, C, A, A); lid = torch.ones(B * S, C, A, A); rad = 2 * torch.ones(B * S, C, A, A)
    gps = 3 * torch.ones(B, 2, C)
    x = R.build_tokens(img, lid, rad, gps, torch.zeros(1, 962, C), S, 1)
    assert x.shape == (B, 962, C)
    assert (x[:, :320] == 0).all() and (x[:, 320:640] == 1).all() and (x[:, 640:960] == 2).all() and (x[:, 960:] == 3).all()
    # channels-last inside a token; (y, x) raster inside a frame
    img = torch.arange(B * S * C * A * A, dtype=torch.float32).reshape(B * S, C, A, A)
    x = R.build_tokens(img, lid, rad, gps, torch.zeros(1, 962, C), S, 1)
    assert x[1, 2 * 64 + 3 * 8 + 5, 4] == img[1 * S + 2, 4, 3, 5]


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree absent (GPU box)")
def test_oracle_matches_reference_live():
    """Full-size stage-4-like GPT (C=128 to keep it quick, T=962, L=8) against the live reference."""
    M, _ = ref_import.load_reference()
    cfg = ref_import.make_config()
    torch.manual_seed(3)
    gpt = M.GPT(128, 4, 4, 8, 8, 8, 5, 0., 0., 0., cfg)
    with torch.no_grad():
        gpt.pos_emb.normal_(0, 0.02)
    B = 1
    ins = [torch.randn(B * 5, 128, 8, 8) for _ in range(3)] + [torch.randn(B, 2, 128)]
    with torch.no_grad():
        ref = gpt(*ins)
        p = {k: v for k, v in gpt.state_dict().items()}
        got = R.gpt_forward(p, *ins, n_head=4, seq_len=5)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and rel_err(a, b) < TOL


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree absent (GPU box)")
def test_init_law_matches_reference():
    M, _ = ref_import.load_reference()
    cfg = ref_import.make_config()
    gpt = M.GPT(64, 4, 4, 2, 8, 8, 5, 0., 0., 0., cfg)
    p = R.init_gpt_params(64, 4, 4, 2, 962)
    sd = gpt.state_dict()
    assert set(sd.keys()) == set(p.keys())
    for k in sd:
        assert sd[k].shape == p[k].shape, k


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree absent (GPU box)")
def test_encoder_restatement_matches_reference_live():
    """oracle/model_ref.encoder_forward against the reference's own Encoder.forward (CPU fp32, B=1)."""
    from oracle import model_ref
    M, _ = ref_import.load_reference()
    cfg = ref_import.make_config(n_layer=2)
    torch.manual_seed(4)
    enc = M.Encoder(cfg).eval()
    with torch.no_grad():
        for k in (1, 2, 3, 4):
            getattr(enc, "transformer%d" % k).pos_emb.normal_(0, 0.02)
    gen = torch.Generator().manual_seed(5)
    imgs = [torch.rand(1, 3, 256, 256, generator=gen) * 255 for _ in range(5)]
    lids = [torch.rand(1, 1, 256, 256, generator=gen) for _ in range(5)]
    rads = [torch.rand(1, 2, 256, 256, generator=gen) for _ in range(5)]
    gps = torch.rand(1, 2, 2, generator=gen)
    with torch.no_grad():
        ref = enc(imgs, lids, rads, gps)
        got = model_ref.encoder_forward(enc, imgs, lids, rads, gps)
    assert got.shape == ref.shape == (1, 512)
    assert rel_err(got, ref) < 1e-4


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree absent (GPU box)")
def test_configure_optimizers_groups_match_reference():
    """GPT.configure_optimizers (model2_seq.py:216-246, dead code in the reference but part of the class API): the drop-in
    returns the same two parameter groups (names, order, weight decay)."""
    from deepsense6g_tii_b200.modules import GPT
    M, _ = ref_import.load_reference()
    cfg = ref_import.make_config()
    ref = M.GPT(64, 4, 4, 2, 8, 8, 5, 0., 0., 0., cfg)
    ours = GPT(64, 4, 4, 2, 8, 8, 5, 0., 0., 0., cfg)

    def named_groups(m):
        inv = {id(p): n for n, p in m.named_parameters()}
        return [([inv[id(p)] for p in g["params"]], g["weight_decay"]) for g in m.configure_optimizers()]

    assert named_groups(ours) == named_groups(ref)
