"""Shared helpers of the GPU parity tests and of tests/tools/diag_bf16_error.py: one fusion-stage problem evaluated by
the dsfuse kernels (through the C ABI) and by the oracle restatement (oracle/fusion_ref.py) in float64 on the same
device, with per-tensor relative L2 errors.

ReLU decisions.  The gradient of the stage is discontinuous where an mlp.0 pre-activation crosses 0 (nn.ReLU,
model2_seq.py:123).  Any evaluation that rounds an operand of mlp.0 lands a few pre-activations with |z| below its rounding
error on the other side: in bf16 about 0.2-0.3 % of the active units flip, each flipped unit is a 100 % error of dL/dz
there, and the relative error of the mlp.0 / ln2 gradients is sqrt(flip fraction) = 4-5e-2 for ANY bf16 evaluation (stock
torch.autocast included; tests/tools/bf16_error_model.py reproduces it on the CPU with nothing but round-to-bf16 inserted
into float64 math).  The oracle can therefore be evaluated a second time WITH the decisions the kernels took
(``relu.{i}`` masks of oracle.fusion_ref.block): against that reference only rounding proper is left, and every gradient
tensor has to meet north_star's 2e-2.
"""
import torch

from oracle import fusion_ref as R

S, NH = 5, 4


def make_problem(B, C, H, L, A, dev, seed=None, wscale=0.01):
    T = 3 * S * A * A + 2
    gen = torch.Generator().manual_seed(100 + C if seed is None else seed)
    p0 = R.init_gpt_params(C, NH, 4, L, T, generator=gen, pos_std=0.02)
    p0 = {k: (v + wscale * torch.randn(v.shape, generator=gen)).to(dev) for k, v in p0.items()}
    feats = [torch.randn(B * S, C, H, H, generator=gen).to(dev) for _ in range(3)]
    gps = torch.randn(B, 2, C, generator=gen).to(dev)
    probes = [torch.randn(f.shape, generator=gen).to(dev) for f in feats] + [torch.randn(B, 2, C, generator=gen).to(dev)]
    return dict(B=B, C=C, H=H, L=L, A=A, T=T, p0=p0, feats=feats, gps=gps, probes=probes)


def leafs(pb, dt=torch.float32):
    return ({k: v.to(dt).clone().requires_grad_(True) for k, v in pb["p0"].items()},
            [f.to(dt).clone().requires_grad_(True) for f in pb["feats"]] + [pb["gps"].to(dt).clone().requires_grad_(True)])


def run_oracle(pb, dt=torch.float64, autocast=False, relu_masks=None):
    """The oracle restatement; ``relu_masks``: {"relu.i": (B*T, 4C) tensor} whose sign pattern replaces ReLU's own decisions."""
    p, i = leafs(pb, dt)
    masks = None
    if relu_masks is not None:
        masks = {k: (v > 0).to(dt).view(pb["B"], pb["T"], -1) for k, v in relu_masks.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        (a, b, c), g = R.fusion_stage(p, i[:3], i[3], NH, S, pb["A"], pb["A"], masks=masks)
    outs = (a, b, c, g)
    sum((o.to(dt) * pr.to(dt)).sum() for o, pr in zip(outs, pb["probes"])).backward()
    return outs, p, i


def run_dsfuse(pb, mode, capture=None, **cfg_extra):
    from deepsense6g_tii_b200.functional import fusion_stage, param_names
    p, i = leafs(pb)
    cfg = dict(seq_len=S, n_views=1, vert_anchors=pb["A"], horz_anchors=pb["A"], n_head=NH, n_layer=pb["L"], compute_dtype=mode)
    if capture is not None:
        cfg["capture"] = capture
    cfg.update(cfg_extra)
    outs = fusion_stage(cfg, i[0], i[1], i[2], i[3], [p[n] for n in param_names(pb["L"])])
    sum((o.float() * pr).sum() for o, pr in zip(outs, pb["probes"])).backward()
    torch.cuda.synchronize()
    return outs, p, i


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def errors(res, ref, L):
    """{tensor name: relative L2 error}: out0-3, gin0-3, g/<param>; attn.key.bias (mathematically zero gradient) is reported
    relative to the query-bias gradient of the same block."""
    from deepsense6g_tii_b200.functional import param_names
    d = {}
    for j, (x, y) in enumerate(zip(res[0], ref[0])):
        d["out%d" % j] = rel(x, y)
    for j, (x, y) in enumerate(zip(res[2], ref[2])):
        d["gin%d" % j] = rel(x.grad, y.grad)
    for n in param_names(L):
        if n.endswith("attn.key.bias"):
            qn = float(ref[1][n.replace("key", "query")].grad.double().norm())
            d["g/" + n] = float((res[1][n].grad.double() - ref[1][n].grad.double()).norm()) / (qn + 1e-30)
        else:
            d["g/" + n] = rel(res[1][n].grad, ref[1][n].grad)
    return d


def worst(d, prefix=""):
    k = max((k for k in d if k.startswith(prefix)), key=lambda k: d[k])
    return d[k], k


def flip_fraction(capture, ref_capture_or_masks):
    """Worst-block fraction of the active mlp.0 units whose ReLU decision differs between two evaluations."""
    w = 0.0
    for k, a in capture.items():
        m, r = a > 0, ref_capture_or_masks[k].reshape(a.shape) > 0
        w = max(w, float((m != r).sum()) / max(1.0, float(r.sum())))
    return w


def oracle_relu_decisions(pb, dt=torch.float64):
    """The mlp.0 outputs of the float64 oracle, block by block (no autograd), for ``flip_fraction``."""
    p = {k: v.to(dt) for k, v in pb["p0"].items()}
    with torch.no_grad():
        pooled = [R.anchor_pool(f.to(dt), pb["A"], pb["A"]) for f in pb["feats"]]
        x = R.build_tokens(pooled[0], pooled[1], pooled[2], pb["gps"].to(dt), p["pos_emb"], S, 1)
        out = {}
        for i in range(pb["L"]):
            pre = "blocks.%d." % i
            xm = x + R.self_attention(R.layer_norm(x, p[pre + "ln1.weight"], p[pre + "ln1.bias"]), p, pre + "attn.", NH)
            h = R.layer_norm(xm, p[pre + "ln2.weight"], p[pre + "ln2.bias"])
            out["relu.%d" % i] = R.linear(h, p[pre + "mlp.0.weight"], p[pre + "mlp.0.bias"]).reshape(-1, 4 * pb["C"])
            x = R.block(x, p, i, NH)
    return out
