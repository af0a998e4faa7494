"""BASELINE.json configs[3]: missing-modality inference throughput sweep — eval mode, LiDAR and radar inputs replaced by
zeros ahead of conv1 (mambafuser_seq.py:361-391, 418-420), batch 1 ... 1024, bf16 fusion + bf16 autocast trunks.
Batches <= 16 replay the forward as one CUDA graph (they are launch-bound otherwise).  Prints samples/s with the zeroed-stem cache (deepsense6g_tii_b200.modules.Encoder._stem) on and off."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsense6g_tii_b200 import TransFuser  # noqa: E402
from deepsense6g_tii_b200.train import synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
cfg = types.SimpleNamespace(seq_len=5, pred_len=4, n_views=1, vert_anchors=8, horz_anchors=8, n_embd=512, block_exp=4, n_layer=8, n_head=4,
                            embd_pdrop=0.1, attn_pdrop=0.1, resid_pdrop=0.1, add_velocity=1, fusion_dtype=torch.bfloat16,
                            modality_missing="lidar_radar", modality_missing_type="zerolike")
torch.manual_seed(100)
model = TransFuser(cfg, dev).to(memory_format=torch.channels_last).eval()
batches = [int(b) for b in sys.argv[1:]] or [1, 4, 16, 64, 128, 256, 512, 1024]
for B in batches:
    imgs, lids, rads, gps, _, _ = synthetic_batch(B, 5, 256, generator=torch.Generator().manual_seed(B), device=dev)
    imgs = [t.contiguous(memory_format=torch.channels_last) for t in imgs]
    line = "batch %4d:" % B
    for fast in (True, False):
        cfg.missing_fast_path = fast

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return model(imgs, lids, rads, gps)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        if B <= 16:  # launch-bound at small batch: replay the forward as one CUDA graph
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run()
            run = g.replay
            run()
            torch.cuda.synchronize()
        n = max(3, min(20, 512 // B))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        line += "  %s %8.2f ms  %8.1f samples/s" % ("stem cache on " if fast else "stem cache off", ms, B / ms * 1e3)
    print(line, flush=True)
    del imgs, lids, rads
    torch.cuda.empty_cache()
