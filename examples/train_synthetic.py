#!/usr/bin/env python
"""End-to-end example: the reference's training loop (train2_seq.py:94-136, 430-434, 531-539) on the drop-in model.

    python examples/train_synthetic.py --steps 20                     # 1 GPU
    torchrun --nproc-per-node 8 examples/train_synthetic.py           # 8 GPUs, batch-sharded (replaces nn.DataParallel)

What changes for a user of szy4017/DeepSense6G_TII:
  * `from deepsense6g_tii_b200 import TransFuser` instead of `from model2_seq import TransFuser` (same constructor, same
    forward(fronts, lidars, radars, gps), same state-dict names — reference checkpoints load with strict=True);
  * one process per GPU (torchrun + DistributedDataParallel) instead of nn.DataParallel (train2_seq.py:538);
  * FocalLoss / EMA come from deepsense6g_tii_b200.train (same semantics as train2_seq.py:291-334).
Synthetic data stands in for the DeepSense 6G scenarios (the dataset is not shipped); value ranges follow data2_seq.py.
"""
import argparse
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from deepsense6g_tii_b200 import TransFuser  # noqa: E402
from deepsense6g_tii_b200.train import EMA, FocalLoss, synthetic_batch, train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch_size", type=int, default=12, help="per GPU (the reference's --batch_size is global)")
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--ema", type=int, default=1)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(100)  # train2_seq.py:430-434
    torch.backends.cudnn.benchmark = True
    # config_seq.GlobalConfig values (config_seq.py:3-45) + the runtime attributes train2_seq.py:414-425 copies on
    cfg = types.SimpleNamespace(seq_len=5, pred_len=4, n_views=1, vert_anchors=8, horz_anchors=8, n_embd=512, block_exp=4, n_layer=8,
                                n_head=4, embd_pdrop=0.1, attn_pdrop=0.1, resid_pdrop=0.1, add_velocity=1, fusion_dtype=torch.bfloat16)
    model = TransFuser(cfg, dev).to(memory_format=torch.channels_last).train()
    net = model
    enc = model.encoder
    gpts = [enc.transformer1, enc.transformer2, enc.transformer3, enc.transformer4]
    grad_sync = None
    if world > 1:   # one process per GPU replaces nn.DataParallel (train2_seq.py:538): same start, gradients averaged every step
        from deepsense6g_tii_b200 import dist as D
        D.broadcast_params(model.parameters())
        grad_sync = D.DataParallelGrads(model, gpts)
    criterion = FocalLoss()
    ema = EMA(model, 0.999) if args.ema else None
    if ema:
        ema.register()
    # optim.AdamW(model.parameters(), lr) + ema.update() (train2_seq.py:131-134, 539) as one dsfuse launch that also keeps the GPTs' bf16
    # weight shadows up to date
    from deepsense6g_tii_b200.optim import FusedAdamWEMA
    optimizer = FusedAdamWEMA(model.parameters(), lr=args.lr, ema=ema, gpts=gpts)
    gen = torch.Generator().manual_seed(rank)
    for step in range(args.steps):
        batch = synthetic_batch(args.batch_size, 5, 256, generator=gen, device=dev)
        loss = train_step(net, batch, criterion, optimizer, ema, autocast_dtype=torch.bfloat16, grad_sync=grad_sync)
        if rank == 0 and (step % 5 == 0 or step == args.steps - 1):
            print("step %3d  loss %.5f" % (step, float(loss.detach())))
    if ema:  # validate with the shadow weights as Engine.validate does (train2_seq.py:159-160, 220-221)
        ema.apply_shadow()
        model.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            imgs, lids, rads, gps, _, beam = synthetic_batch(args.batch_size, 5, 256, generator=gen, device=dev)
            pred = model(imgs, lids, rads, gps)
        if rank == 0:
            print("eval with EMA weights: logits", tuple(pred.shape), "top-1 beams", pred.argmax(-1)[:6].tolist())
        ema.restore()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
