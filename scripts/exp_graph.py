"""Experiment: eager launch vs CUDA-graph replay of one fwd+bwd fusion-stage step (bench shape)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
gpt = bench.build_gpt(dev)
gen = torch.Generator().manual_seed(0)
feats_h, gps_h, probes_h = bench.synth_inputs(gen, bench.BATCH)
feats = [f.to(dev).requires_grad_(True) for f in feats_h]
gps = gps_h.to(dev).requires_grad_(True)
probes = [p.to(dev) for p in probes_h]


def timeit(fn, n=20, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("eager   %.3f ms/step" % timeit(lambda: bench.one_step(gpt, feats, gps, probes)))
with torch.no_grad():
    print("fwd only (no_grad) %.3f ms" % timeit(lambda: gpt.fuse(feats[0], feats[1], feats[2], gps)))
# graph capture (side stream warm-up as torch requires)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        bench.one_step(gpt, feats, gps, probes)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
for p in gpt.parameters():
    p.grad = None
with torch.cuda.graph(g):
    loss = bench.one_step(gpt, feats, gps, probes)
torch.cuda.synchronize()
print("graph   %.3f ms/step" % timeit(g.replay))
ref = [p.grad.clone() for p in gpt.parameters()]
g.replay()
torch.cuda.synchronize()
print("replay grads equal:", all(torch.equal(a, p.grad) for a, p in zip(ref, gpt.parameters())), "loss", float(loss))
