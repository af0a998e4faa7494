"""Per-stage timing of the four fusion stages of the 256x256 model (C = 64/128/256/512, H = 64/32/16/8), batch 12:
graph-replayed fwd+bwd time and the per-family breakdown of an instrumented step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for (c, scale) in [(64, 8), (128, 4), (256, 2), (512, 1)]:
    bench.C, bench.SCALE = c, scale
    gpt = bench.build_gpt(dev)
    gen = torch.Generator().manual_seed(0)
    feats_h, gps_h, probes_h = bench.synth_inputs(gen, bench.BATCH)
    feats = [f.to(dev).requires_grad_(True) for f in feats_h]
    gps = gps_h.to(dev).requires_grad_(True)
    probes = [p.to(dev) for p in probes_h]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            bench.one_step(gpt, feats, gps, probes)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        bench.one_step(gpt, feats, gps, probes)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    fam = bench.profile_families(gpt, feats, gps, probes)
    print("C=%d H=%d: %.3f ms/step (graph)  families: %s" % (c, 8 * scale, e0.elapsed_time(e1) / 10,
          " ".join("%s=%.3f" % (k, v["ms"]) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]))))
    del g, gpt, feats
    torch.cuda.empty_cache()
