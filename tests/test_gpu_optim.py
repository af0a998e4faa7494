"""optim.FusedAdamWEMA (dsf_adamw_ema_pack) against torch.optim.AdamW + the reference's EMA loop (train2_seq.py:131-134, 315-320, 539),
and the hand-over of the bf16 weight shadows to the fusion-stage forward."""
import types

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _model(dev, n_layer=2, C=64):
    from deepsense6g_tii_b200.modules import GPT
    cfg = types.SimpleNamespace(n_views=1, fusion_dtype=torch.bfloat16)
    torch.manual_seed(3)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gpt = GPT(C, 4, 4, n_layer, 8, 8, 5, 0.0, 0.0, 0.0, cfg)
            self.conv = torch.nn.Conv2d(3, 8, 3)          # a non-GPT parameter, stored channels_last (dense, permuted strides)
            self.scalar = torch.nn.Parameter(torch.randn(5))
    net = Net().to(dev).to(memory_format=torch.channels_last)
    with torch.no_grad():
        net.gpt.pos_emb.normal_(0, 0.02)
    return net


def _reference_ema_update(shadow, model, decay):
    """train2_seq.py:315-320, verbatim arithmetic."""
    for name, param in model.named_parameters():
        if param.requires_grad:
            shadow[name] = ((1.0 - decay) * param.data + decay * shadow[name]).clone()


def test_fused_adamw_ema_matches_torch_adamw_and_reference_ema(cuda_dev):
    import copy
    from deepsense6g_tii_b200.optim import FusedAdamWEMA
    from deepsense6g_tii_b200.train import EMA
    a = _model(cuda_dev)
    b = copy.deepcopy(a)
    ema = EMA(a, 0.999)
    ema.register()
    opt_a = FusedAdamWEMA(a.parameters(), lr=1e-3, weight_decay=0.01, ema=ema, gpts=[a.gpt])
    opt_b = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=0.01)
    shadow_b = {n: p.data.clone() for n, p in b.named_parameters()}
    gen = torch.Generator(device=cuda_dev).manual_seed(0)
    for step in range(6):
        for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            g = torch.randn(pa.shape, device=cuda_dev, generator=gen) * (0.1 + step)
            if pa.dim() == 4:
                g = g.contiguous(memory_format=torch.channels_last)
            pa.grad, pb.grad = g.clone(memory_format=torch.preserve_format), g.clone(memory_format=torch.preserve_format)
        opt_a.step()
        ema.update()                      # folded into opt_a.step(): must be a no-op now
        opt_b.step()
        _reference_ema_update(shadow_b, b, 0.999)
    torch.cuda.synchronize()
    assert opt_a.steps_done == 6
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        # fp32 arithmetic in a different association order (one fused pass vs torch's foreach kernels): a few ulp per step
        assert rel_err(pa, pb) < 2e-5, n
        assert rel_err(opt_a.state[pa]["exp_avg"], opt_b.state[pb]["exp_avg"]) < 2e-6, n
        assert rel_err(opt_a.state[pa]["exp_avg_sq"], opt_b.state[pb]["exp_avg_sq"]) < 2e-6, n
        assert rel_err(ema.shadow[n], shadow_b[n]) < 2e-5, n
    # the bf16 shadows the forward will read are exactly the repack of the updated fp32 weights
    for i, blk in enumerate(a.gpt.blocks):
        v = a.gpt.shadow_views(i)
        wqkv = torch.cat([blk.attn.query.weight, blk.attn.key.weight, blk.attn.value.weight], 0).detach()
        assert torch.equal(v["wqkv"], wqkv.bfloat16()) and torch.equal(v["wqkv_t"], wqkv.bfloat16().t())
        assert torch.equal(v["wp"], blk.attn.proj.weight.detach().bfloat16()) and torch.equal(v["wp_t"], blk.attn.proj.weight.detach().bfloat16().t())
        assert torch.equal(v["w1"], blk.mlp[0].weight.detach().bfloat16()) and torch.equal(v["w1_t"], blk.mlp[0].weight.detach().bfloat16().t())
        assert torch.equal(v["w2"], blk.mlp[2].weight.detach().bfloat16()) and torch.equal(v["w2_t"], blk.mlp[2].weight.detach().bfloat16().t())
        assert torch.equal(v["bqkv"], torch.cat([blk.attn.query.bias, blk.attn.key.bias, blk.attn.value.bias]).detach())


def test_forward_after_the_fused_step_launches_no_pack_kernels_and_matches_a_fresh_pack(cuda_dev):
    from deepsense6g_tii_b200 import _capi
    from deepsense6g_tii_b200.optim import FusedAdamWEMA
    net = _model(cuda_dev, n_layer=3)
    gpt = net.gpt
    ins = [torch.randn(10, 64, 8, 8, device=cuda_dev) for _ in range(3)] + [torch.randn(2, 2, 64, device=cuda_dev)]

    def fwd_bwd():
        n0 = _capi.launch_count()
        out = gpt(*ins)
        n1 = _capi.launch_count()
        sum(o.float().square().sum() for o in out).backward()
        return [o.detach().clone() for o in out], n1 - n0

    opt = FusedAdamWEMA(gpt.parameters(), lr=1e-3, gpts=[gpt])
    _, n_first = fwd_bwd()                # shadows not written yet: the forward packs them (3 launches)
    opt.step()
    opt.zero_grad(set_to_none=True)
    out_fused, n_fused = fwd_bwd()        # shadows written by the optimizer kernel
    assert n_first - n_fused == 3, (n_first, n_fused)
    gpt._shadow_fresh_key = None          # force a re-pack from the same fp32 weights
    out_packed, n_packed = fwd_bwd()
    assert n_packed == n_first
    for x, y in zip(out_fused, out_packed):
        assert torch.equal(x, y)
    # a parameter changed behind the optimizer's back (load_state_dict, manual edit) invalidates the shadows
    with torch.no_grad():
        gpt.blocks[0].mlp[0].weight.mul_(1.5)
    out_edit, n_edit = fwd_bwd()
    assert n_edit == n_first and not torch.equal(out_edit[0], out_packed[0])


def test_fused_training_step_replays_as_a_cuda_graph(cuda_dev):
    """fwd + bwd + fused optimizer captured once: every replay advances the device-resident step counter and equals the eager
    sequence of steps (same gradients up to atomics order -> parameters agree closely)."""
    import copy
    from deepsense6g_tii_b200.optim import FusedAdamWEMA
    from deepsense6g_tii_b200.train import EMA
    net = _model(cuda_dev)
    ref = copy.deepcopy(net)
    ins = [torch.randn(10, 64, 8, 8, device=cuda_dev) for _ in range(3)] + [torch.randn(2, 2, 64, device=cuda_dev)]

    def make(m):
        ema = EMA(m.gpt, 0.99)
        ema.register()
        opt = FusedAdamWEMA(m.gpt.parameters(), lr=1e-3, ema=ema, gpts=[m.gpt])

        def step():
            opt.zero_grad(set_to_none=True)
            out = m.gpt(*ins)
            sum(o.float().square().mean() for o in out).backward()
            opt.step()
            ema.update()
        return step, opt, ema

    step_g, opt_g, ema_g = make(net)
    step_e, opt_e, ema_e = make(ref)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step_g()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step_g()                          # step 3
    for _ in range(3):
        g.replay()                        # capture does not execute: steps 3, 4, 5
    torch.cuda.synchronize()
    for _ in range(5):
        step_e()
    torch.cuda.synchronize()
    assert opt_g.steps_done == 5 and opt_e.steps_done == 5
    for (n, pa), (_, pb) in zip(net.gpt.named_parameters(), ref.gpt.named_parameters()):
        if n.endswith("attn.key.bias"):   # mathematically zero gradient: Adam normalises pure atomics-order noise
            continue
        # Adam divides by sqrt(v): where a gradient is small, the atomics-order noise of the weight-gradient reductions is amplified
        assert rel_err(pa, pb) < 2e-2, n
        assert rel_err(ema_g.shadow[n], ema_e.shadow[n]) < 2e-2, n
