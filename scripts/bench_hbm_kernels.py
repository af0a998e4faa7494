"""HBM-bound kernels at the four stage shapes (GPU): tokens fwd / bwd and upsample-add fwd / bwd, CUDA-event timed over rotating
buffer sets larger than L2 (so every launch reads its feature maps from HBM, as in the step).
Usage: python scripts/bench_hbm_kernels.py [once]   ("once": a single launch of every kernel, for ncu)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsense6g_tii_b200 import _capi as K  # noqa: E402

dev = torch.device("cuda")
K.check_device()
once = "once" in sys.argv[1:]
B, S, A = 12, 5, 8


def timeit(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


for (C, H) in [(64, 64), (128, 32), (256, 16), (512, 8)]:
    g = K.make_geom(B, S, 1, A, A, C, H, H, K.DSF_F32)
    T = 3 * S * A * A + 2
    e_f = 3 * B * S * C * H * H
    nset = 1 if once else max(2, int(400e6 // (e_f * 4)) + 1)
    sets = []
    for _ in range(nset):
        feats = [torch.randn(B * S, C, H, H, device=dev) for _ in range(3)]
        outs = [torch.empty_like(f) for f in feats]
        sets.append((feats, outs))
    gps = torch.randn(B, 2, C, device=dev)
    pos = torch.randn(1, T, C, device=dev)
    x = torch.empty(B * T, C, device=dev)
    y = torch.randn(B * T, C, device=dev)
    dy = torch.empty(B * T, C, device=dev)
    dgps = torch.empty(B, 2, C, device=dev)
    dpos = torch.zeros(1, T, C, device=dev)
    it = [0]

    def nxt():
        it[0] += 1
        return sets[it[0] % nset]

    runs = {
        "tokens_fwd": lambda: (lambda s: K.tokens_fwd(g, s[0][0], s[0][1], s[0][2], gps, pos, x))(nxt()),
        "tokens_bwd": lambda: (lambda s: K.tokens_bwd(g, y, s[0], s[1], dgps, dpos))(nxt()),
        "upsample_add_fwd": lambda: (lambda s: K.upsample_add_fwd(g, y, s[0], s[1]))(nxt()),
        "upsample_add_bwd": lambda: (lambda s: K.upsample_add_bwd(g, s[0], gps, dy))(nxt()),
    }
    e_t = B * T * C
    nbytes = {"tokens_fwd": 4 * (e_f + e_t), "tokens_bwd": 4 * (e_t + 2 * e_f), "upsample_add_fwd": 4 * (e_t + 2 * e_f), "upsample_add_bwd": 4 * (e_f + e_t)}
    line = "C=%d H=%d (%d MB of feature maps):" % (C, H, e_f * 4 // 1000000)
    for name, fn in runs.items():
        if once:
            fn()
            torch.cuda.synchronize()
            continue
        t = timeit(fn, 20, 5)
        line += "  %s %.1f us (%.0f GB/s)" % (name, t, nbytes[name] / t / 1e3)
    print(line, flush=True)
