"""Diagnostic (GPU): one fusion stage on REALISTIC trunk features (ResNet stem output) with a structured
upstream gradient, mine (fp32 mode) vs oracle fp32 vs oracle fp64."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import fusion_ref as R
from deepsense6g_tii_b200 import TransFuser
from deepsense6g_tii_b200.functional import fusion_stage, param_names

dev = torch.device("cuda")
L = 2
cfg = types.SimpleNamespace(seq_len=5, pred_len=4, n_views=1, vert_anchors=8, horz_anchors=8, n_embd=512, block_exp=4, n_layer=L, n_head=4,
                            embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0, add_velocity=1, fusion_dtype=torch.float32)
torch.manual_seed(100)
m = TransFuser(cfg, dev).train()
B = 2
g = torch.Generator().manual_seed(0)
imgs = torch.cat([(torch.rand(B, 3, 256, 256, generator=g) * 255) for _ in range(5)]).to(dev)
lids = torch.cat([(torch.rand(B, 1, 256, 256, generator=g) < 0.05).float() for _ in range(5)]).to(dev)
rads = torch.cat([torch.rand(B, 2, 256, 256, generator=g) for _ in range(5)]).to(dev)
enc = m.encoder
with torch.no_grad():
    from deepsense6g_tii_b200.modules import normalize_imagenet
    def stem(mm, x): return mm.layer1(mm.maxpool(mm.relu(mm.bn1(mm.conv1(x)))))
    f = [stem(enc.image_encoder.features, normalize_imagenet(imgs)), stem(enc.lidar_encoder._model, lids), stem(enc.radar_encoder._model, rads)]
    gps = enc.vel_emb1(torch.rand(B, 2, 2, generator=g).to(dev))
print("feature stats:", [(float(t.mean()), float(t.std()), float(t.max())) for t in f])
gpt = enc.transformer1
with torch.no_grad():
    gpt.pos_emb.normal_(0, 0.02)
names = param_names(L)
p0 = {k: v.detach().clone() for k, v in gpt.named_parameters()}
# structured upstream gradients: spatially smooth + per-channel constant parts
up = [torch.randn(t.shape[0], t.shape[1], 1, 1, generator=g).to(dev) * 0.1 + torch.randn(t.shape, generator=g).to(dev) * 0.01 for t in f]
upg = torch.randn(B, 2, 64, generator=g).to(dev)

def run(kind):
    dt = torch.float64 if kind == "o64" else torch.float32
    p = {k: v.to(dt).clone().requires_grad_(True) for k, v in p0.items()}
    ins = [t.to(dt).clone().requires_grad_(True) for t in f] + [gps.to(dt).clone().requires_grad_(True)]
    if kind == "mine":
        c = dict(seq_len=5, n_views=1, vert_anchors=8, horz_anchors=8, n_head=4, n_layer=L, compute_dtype=torch.float32)
        outs = fusion_stage(c, ins[0], ins[1], ins[2], ins[3], [p[n] for n in names])
    else:
        (a, b, c_), go = R.fusion_stage(p, ins[:3], ins[3], 4, 5, 8, 8)
        outs = (a, b, c_, go)
    loss = sum((o * u.to(dt)).sum() for o, u in zip(outs, up + [upg]))
    loss.backward()
    return outs, p, ins

ref = run("o64")
for kind in ("mine", "o32"):
    res = run(kind)
    def rel(a, b): return float((a.double() - b).norm() / (b.norm() + 1e-300))
    print(kind, "out", ["%.2e" % rel(x, y) for x, y in zip(res[0], ref[0])])
    print(kind, "gin", ["%.2e" % rel(x.grad, y.grad) for x, y in zip(res[2], ref[2])])
    # split dfeat error: remove the passthrough part
    print(kind, "gin minus passthrough", ["%.2e" % rel(x.grad.double() - u.double(), y.grad - u.double()) for x, y, u in zip(res[2][:3], ref[2][:3], up)])
    errs = sorted([(rel(res[1][n].grad, ref[1][n].grad), n) for n in names if "key.bias" not in n], reverse=True)
    print(kind, "worst params", [("%.2e" % e, n) for e, n in errs[:4]])
