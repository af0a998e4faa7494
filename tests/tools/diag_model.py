"""Diagnostic (GPU): per-parameter gradient errors of the drop-in TransFuser vs the oracle forward on the same module."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import model_ref
from deepsense6g_tii_b200 import TransFuser

mode = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else torch.bfloat16
n_layer = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
if os.environ.get("NO_TF32", "1") == "1":
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
print("cudnn.allow_tf32 =", torch.backends.cudnn.allow_tf32)
cfg = types.SimpleNamespace(seq_len=5, pred_len=4, n_views=1, vert_anchors=8, horz_anchors=8, n_embd=512, block_exp=4, n_layer=n_layer, n_head=4,
                            embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0, add_velocity=1, fusion_dtype=mode)
torch.manual_seed(100)
m = TransFuser(cfg, dev).train()
with torch.no_grad():
    for k in (1, 2, 3, 4):
        getattr(m.encoder, "transformer%d" % k).pos_emb.normal_(0, 0.02)
B = 2
g = torch.Generator().manual_seed(0)
imgs = [(torch.rand(B, 3, 256, 256, generator=g) * 255).to(dev) for _ in range(5)]
lids = [(torch.rand(B, 1, 256, 256, generator=g) < 0.05).float().to(dev) for _ in range(5)]
rads = [torch.rand(B, 2, 256, 256, generator=g).to(dev) for _ in range(5)]
gps = torch.rand(B, 2, 2, generator=g).to(dev)
probe = torch.randn(B, 64, generator=g).to(dev)
out = m(imgs, lids, rads, gps)
(out * probe).sum().backward()
got = {n: p.grad.clone() for n, p in m.named_parameters()}
m.zero_grad(set_to_none=True)
ref = model_ref.transfuser_forward(m, imgs, lids, rads, gps)
(ref * probe).sum().backward()
ref_g = {n: p.grad.clone() for n, p in m.named_parameters()}
m.zero_grad(set_to_none=True)
ref2 = model_ref.transfuser_forward(m, imgs, lids, rads, gps)
(ref2 * probe).sum().backward()
print("oracle-vs-oracle logits %.3e  worst grad %.3e" % (float((ref2 - ref).norm() / ref.norm()),
      max(float((ref_g[n] - p.grad).norm() / (p.grad.norm() + 1e-30)) for n, p in m.named_parameters() if "key.bias" not in n)))
print("logits rel err %.3e" % float((out.float() - ref).norm() / ref.norm()))
rows = []
for n, p in m.named_parameters():
    e = float((got[n] - p.grad).norm() / (p.grad.norm() + 1e-30))
    rows.append((e, n, float(p.grad.norm())))
rows.sort(reverse=True)
for e, n, nr in rows[:25]:
    print("%.3e  |g|=%.3e  %s" % (e, nr, n))
import collections
by = collections.defaultdict(list)
for e, n, nr in rows:
    key = n.split(".")[1] if n.startswith("encoder.") else n.split(".")[0]
    by[key].append(e)
for k, v in by.items():
    print("%-20s max %.3e  median %.3e  (n=%d)" % (k, max(v), sorted(v)[len(v) // 2], len(v)))
