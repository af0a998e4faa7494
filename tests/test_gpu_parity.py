"""GPU parity at north_star's tolerance, tensor by tensor, including the configuration bench.py times.

Reference for every comparison: the oracle restatement of model2_seq.py:94-134, 248-287, 515-526 evaluated in float64 on the
same device (tests/parity_util.py).  bf16 mode: every output and — against the oracle evaluated with the kernels' own ReLU
decisions — every gradient tensor within 2e-2; the decision-free comparison is bounded by the flip model (relative error of
the mlp.0 / ln2 gradients = sqrt(fraction of flipped ReLU decisions), tests/tools/bf16_error_model.py), with the measured
fraction asserted to be what bf16 rounding of the mlp.0 operands produces and no more.
"""
import os
import sys

import pytest
import torch

import parity_util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SHAPES = [  # (B, C, H, L, A): the four stages of the 256x256 model at B = 2, the benched stage-4 configuration at B = 12,
    (2, 64, 64, 8, 8), (2, 128, 32, 8, 8), (2, 256, 16, 8, 8), (2, 512, 8, 8, 8),
    (12, 512, 8, 8, 8), (12, 64, 64, 8, 8),
]


def _check(got, cap, pb, ref, what):
    L = pb["L"]
    plain = U.errors(got, ref, L)
    matched = U.errors(got, U.run_oracle(pb, relu_masks=cap), L)
    flip = U.flip_fraction(cap, U.oracle_relu_decisions(pb))
    # attn.key.bias: mathematically zero gradient (softmax shift invariance); errors() reports its rounding noise relative to the
    # query-bias gradient of the same block
    bad = ["%s %s: %.3e > 2e-2" % (what, k, v) for k, v in matched.items() if v > (5e-2 if k.endswith("key.bias") else 2e-2)]
    assert not bad, bad
    w_out = U.worst(plain, "out")
    assert w_out[0] <= 1e-2, (what, w_out)
    # decision-free comparison: bf16 rounding of the mlp.0 operands flips 1-4e-3 of the active units (CPU model: 1.6e-3 at
    # C = 128); a flipped unit is a 100 % error of dL/dz, so no bf16 evaluation gets below sqrt(flip) on mlp.0 / ln2
    assert flip <= 5e-3, (what, "flipped ReLU decisions", flip)
    w_plain = U.worst({k: v for k, v in plain.items() if k[0] == "g" and not k.endswith("key.bias")})
    assert w_plain[0] <= 2e-2 + 1.25 * flip ** 0.5, (what, w_plain, flip)
    return plain, matched, flip


@pytest.mark.parametrize("B,C,H,L,A", [(2, 64, 64, 8, 8), (2, 128, 32, 8, 8), (12, 64, 64, 8, 8)])
def test_stage_bf16_chain_kernels_every_tensor_within_2e2(cuda_dev, monkeypatch, B, C, H, L, A):
    """The same bar with the row-local chain kernels of csrc/chain.cu serving the forward and the backward of the narrow stages."""
    monkeypatch.setenv("DSF_CHAIN", "2")
    monkeypatch.setenv("DSF_CHAIN_BWD", "2")
    pb = U.make_problem(B, C, H, L, A, cuda_dev)
    cap = {}
    got = U.run_dsfuse(pb, torch.bfloat16, capture=cap)
    _check(got, cap, pb, U.run_oracle(pb), "chain kernels")


@pytest.mark.parametrize("B,C,H,L,A", SHAPES)
def test_stage_bf16_every_tensor_within_2e2(cuda_dev, B, C, H, L, A):
    pb = U.make_problem(B, C, H, L, A, cuda_dev)
    cap = {}
    got = U.run_dsfuse(pb, torch.bfloat16, capture=cap)
    ref = U.run_oracle(pb)
    plain, matched, flip = _check(got, cap, pb, ref, "eager")
    print("B=%d C=%d: worst matched %.2e %s | worst plain %.2e %s | flips %.2e" % ((B, C) + U.worst(matched) + U.worst(
        {k: v for k, v in plain.items() if not k.endswith("key.bias")}) + (flip,)))


@pytest.mark.parametrize("B,C,H,L,A", [(12, 512, 8, 8, 8), (12, 128, 32, 8, 8), (12, 512, 16, 2, 16)],
                         ids=["benched-stage4", "stage2", "scaled-16x16-anchors"])
def test_graph_captured_step_as_bench_times_it_vs_oracle(cuda_dev, B, C, H, L, A):
    """The schedule bench.py times — the whole fwd+bwd captured in one CUDA graph by ``bench.capture_step``, weight-gradient
    GEMMs / bias sums / weight packs / the dQ kernel on side streams — replayed, against the float64 oracle; and against the
    eager launch sequence of the same inputs (identical forward, gradients equal up to atomics order)."""
    sys.path.insert(0, ROOT)
    import bench
    from deepsense6g_tii_b200.functional import fusion_stage, param_names
    pb = U.make_problem(B, C, H, L, A, cuda_dev)
    ref = U.run_oracle(pb)
    eager = U.run_dsfuse(pb, torch.bfloat16)
    p, i = U.leafs(pb)
    names = param_names(L)
    cap = {}
    cfg = dict(seq_len=U.S, n_views=1, vert_anchors=A, horz_anchors=A, n_head=U.NH, n_layer=L, compute_dtype=torch.bfloat16, capture=cap)

    def step():
        for t in list(p.values()) + i:
            t.grad = None
        outs = fusion_stage(cfg, i[0], i[1], i[2], i[3], [p[n] for n in names])
        sum((o.float() * pr).sum() for o, pr in zip(outs, pb["probes"])).backward()
        return outs

    graph, outs, n_launch = bench.capture_step(step)
    assert n_launch > 10 * L
    for t in list(p.values()) + i:      # the replay must produce everything itself
        t.grad.zero_()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    got = (outs, p, i)
    _check(got, cap, pb, ref, "graph replay")
    for a, b in zip(outs, eager[0]):
        assert torch.equal(a, b), "forward of the replay differs from the eager launch sequence"
    for n in names:
        if not n.endswith("key.bias"):
            assert U.rel(p[n].grad, eager[1][n].grad) <= 1e-3, n


def test_fp32_mode_every_tensor_within_1e3_given_the_relu_decisions(cuda_dev):
    """fp32 parity mode at the real stage-4 shape: 1e-3 on every tensor once the handful of |z| < 1e-7 ReLU decisions that any two
    fp32 evaluations take differently are matched (the decision-free test with its 2-3e-3 parameter bounds stays in test_gpu_stage)."""
    pb = U.make_problem(2, 512, 8, 8, 8, cuda_dev)
    cap = {}
    got = U.run_dsfuse(pb, torch.float32, capture=cap)
    matched = U.errors(got, U.run_oracle(pb, relu_masks=cap), pb["L"])
    bad = ["%s: %.3e" % (k, v) for k, v in matched.items() if v > 1e-3 and not k.endswith("key.bias")]
    assert not bad, bad
