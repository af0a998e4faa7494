"""Diagnostic (GPU): per-tensor relative errors of the bf16 fused stage vs the fp32 oracle, next to the
errors of stock PyTorch bf16 autocast running the oracle code — calibrates what 'bf16 accuracy' means
for each gradient tensor.  Usage: python tests/tools/diag_bf16_error.py [B C H L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import fusion_ref as R
from deepsense6g_tii_b200.functional import fusion_stage, param_names

B, C, H, L = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (2, 64, 64, 8))]
S, A, nh = 5, 8, 4
T = 3 * S * A * A + 2
dev = torch.device("cuda")
gen = torch.Generator().manual_seed(100 + C)
p0 = R.init_gpt_params(C, nh, 4, L, T, generator=gen, pos_std=0.02)
p0 = {k: (v + 0.01 * torch.randn(v.shape, generator=gen)).to(dev) for k, v in p0.items()}
feats = [torch.randn(B * S, C, H, H, generator=gen).to(dev) for _ in range(3)]
gps = torch.randn(B, 2, C, generator=gen).to(dev)
probes = [torch.randn(f.shape, generator=gen).to(dev) for f in feats] + [torch.randn(B, 2, C, generator=gen).to(dev)]
names = param_names(L)

def leafs():
    return ({k: v.clone().requires_grad_(True) for k, v in p0.items()}, [f.clone().requires_grad_(True) for f in feats] + [gps.clone().requires_grad_(True)])

def run_oracle(autocast):
    p, i = leafs()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        (a, b, c), g = R.fusion_stage(p, i[:3], i[3], nh, S, A, A)
    outs = (a, b, c, g)
    sum((o.float() * pr).sum() for o, pr in zip(outs, probes)).backward()
    return outs, p, i

def run_mine(mode):
    p, i = leafs()
    cfg = dict(seq_len=S, n_views=1, vert_anchors=A, horz_anchors=A, n_head=nh, n_layer=L, compute_dtype=mode)
    outs = fusion_stage(cfg, i[0], i[1], i[2], i[3], [p[n] for n in names])
    sum((o.float() * pr).sum() for o, pr in zip(outs, probes)).backward()
    return outs, p, i

def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))

def run_oracle64():
    p = {k: v.double().clone().requires_grad_(True) for k, v in p0.items()}
    i = [f.double().clone().requires_grad_(True) for f in feats] + [gps.double().clone().requires_grad_(True)]
    (a, b, c), g = R.fusion_stage(p, i[:3], i[3], nh, S, A, A)
    outs = (a, b, c, g)
    sum((o * pr.double()).sum() for o, pr in zip(outs, probes)).backward()
    return outs, p, i

print("allow_tf32 matmul/cudnn:", torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32,
      "float32_matmul_precision:", torch.get_float32_matmul_precision(), "env:",
      {k: v for k, v in os.environ.items() if "TF32" in k})
ref = run_oracle64()
rows = {}
for tag, res in (("mine_f32", run_mine(torch.float32)), ("mine_bf16", run_oracle(False)), ("torch_autocast", run_oracle(True))):
    d = {}
    for j, (x, y) in enumerate(zip(res[0], ref[0])):
        d["out%d" % j] = rel(x, y)
    for j, (x, y) in enumerate(zip(res[2], ref[2])):
        d["gin%d" % j] = rel(x.grad, y.grad)
    for n in names:
        d["g/" + n] = rel(res[1][n].grad, ref[1][n].grad)
    rows[tag] = d
keys = list(rows["mine_bf16"].keys())
print("B=%d C=%d H=%d L=%d" % (B, C, H, L))
print("%-34s %10s %10s %10s" % ("tensor", "mine_f32", "mine_bf16", "autocast"))
worst = {t: (0, "") for t in rows}
for k in keys:
    if ".key.bias" in k:
        continue
    print("%-34s %10.2e %10.2e %10.2e" % (k, rows["mine_f32"][k], rows["mine_bf16"][k], rows["torch_autocast"][k]))
    for t in rows:
        if rows[t][k] > worst[t][0]:
            worst[t] = (rows[t][k], k)
print("WORST", {t: "%.2e %s" % v for t, v in worst.items()})
