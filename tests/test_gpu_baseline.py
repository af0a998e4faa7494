"""The practical GPU baseline: the reference's fusion-stage math in STOCK PyTorch eager on
the same B200 (cuBLAS / ATen kernels, fp32 and bf16 autocast), timed next to the dsfuse path on the bench workload
(BASELINE.json configs[1]: C = 512, 8 layers, 4 heads, T = 962, batch 12, fwd + bwd).  The oracle is only the thing being
compared against here; the numbers are written to gpurun_out/torch_gpu_baseline.json when that directory exists."""
import json
import os
import types

import pytest
import torch

from conftest import ROOT
from oracle import fusion_ref as R

pytestmark = pytest.mark.gpu

C, NH, L, A, S, V, B = 512, 4, 8, 8, 5, 1, 12
T = (V + 2) * S * A * A + 2


def _time(step, warm=3, iters=8):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_stage_faster_than_stock_torch_on_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from deepsense6g_tii_b200 import GPT
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(0)
    feats = [torch.randn(B * S, C, A, A, generator=gen).abs_().to(dev).requires_grad_(True) for _ in range(3)]
    gps = torch.randn(B, 2, C, generator=gen).to(dev).requires_grad_(True)
    probes = [torch.randn(B * S, C, A, A, generator=gen).to(dev) * 1e-3 for _ in range(3)] + [torch.randn(B, 2, C, generator=gen).to(dev) * 1e-3]

    cfg = types.SimpleNamespace(n_views=V, fusion_dtype=torch.bfloat16)
    torch.manual_seed(100)
    gpt = GPT(C, NH, 4, L, A, A, S, 0.0, 0.0, 0.0, cfg).to(dev)
    with torch.no_grad():
        gpt.pos_emb.normal_(0, 0.02)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in gpt.state_dict().items()}

    def clear():
        for t in list(p.values()) + list(gpt.parameters()) + feats + [gps]:
            t.grad = None

    def ours():
        clear()
        outs = gpt.fuse(feats[0], feats[1], feats[2], gps)
        sum((o.float() * pr).sum() for o, pr in zip(outs, probes)).backward()

    def stock(autocast):
        def step():
            clear()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                (a, b, c), g = R.fusion_stage(p, feats, gps, NH, S, A, A)
            sum((o.float() * pr).sum() for o, pr in zip((a, b, c, g), probes)).backward()
        return step

    t_ours = _time(ours)
    t_f32 = _time(stock(False))
    t_bf16 = _time(stock(True))
    res = {"workload": "gpt_fusion_stage C=512 L=8 nh=4 T=%d batch=%d fwd+bwd, eager launches (no CUDA graph) on both sides" % (T, B),
           "dsfuse_bf16_ms": t_ours, "stock_torch_fp32_ms": t_f32, "stock_torch_bf16_autocast_ms": t_bf16,
           "dsfuse_samples_per_s": B / t_ours * 1e3, "stock_torch_fp32_samples_per_s": B / t_f32 * 1e3,
           "stock_torch_bf16_autocast_samples_per_s": B / t_bf16 * 1e3,
           "speedup_vs_stock_bf16_autocast": t_bf16 / t_ours, "speedup_vs_stock_fp32": t_f32 / t_ours}
    print(json.dumps(res))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "torch_gpu_baseline.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert t_ours < t_bf16 and t_ours < t_f32, res
