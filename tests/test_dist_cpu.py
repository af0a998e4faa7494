"""world_size-2 gloo test (CPU) of the N>1 plumbing used by bench.py --gpus N: parameter broadcast from
rank 0 and the flat-bucket gradient all-reduce (mean) that replaces nn.DataParallel's reduce
(train2_seq.py:538)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepsense6g_tii_b200 import dist as D
    torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
    m = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    D.broadcast_params(m.parameters())
    w0 = torch.cat([p.data.reshape(-1) for p in m.parameters()]).clone()
    x = torch.full((3, 8), float(rank + 1))  # rank-dependent shard of the batch
    m(x).sum().backward()
    local = [p.grad.clone() for p in m.parameters()]
    buf = D.allreduce_grads(m.parameters())
    buf2 = D.allreduce_grads(m.parameters(), buf)  # bucket reuse (values already equal -> unchanged)
    q.put((rank, w0.numpy(), [g.numpy() for g in local], [p.grad.numpy().copy() for p in m.parameters()], buf2.data_ptr() == buf.data_ptr()))
    dist.destroy_process_group()


def test_broadcast_and_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, w_a, loc_a, red_a, reuse_a), (_, w_b, loc_b, red_b, reuse_b) = res
    import numpy as np
    assert np.array_equal(w_a, w_b)                   # same weights after the broadcast
    assert reuse_a and reuse_b
    for ga, gb, ra, rb in zip(loc_a, loc_b, red_a, red_b):
        assert np.allclose(ra, (ga + gb) / 2, atol=1e-6) and np.array_equal(ra, rb)


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepsense6g_tii_b200 import dist as D
    red = D.OverlappedGradReducer()
    # the fused backward hands over one flat bucket per transformer block (last block first), then the leftovers
    buckets = [torch.full((16,), float(10 * i + rank + 1)) for i in range(3)]
    tail = [torch.full((4,), float(100 + rank)), torch.full((2,), float(200 + rank))]
    for i in reversed(range(3)):
        red.block_ready(i, buckets[i])
    red.finish(tail)
    # deferred mode (one reducer shared by several stages): finish() only starts the collectives, wait_all() completes them
    red2 = D.OverlappedGradReducer(defer=True)
    stage_a, stage_b = torch.full((8,), float(rank + 1)), torch.full((8,), float(10 * (rank + 1)))
    red2.block_ready(0, stage_a)
    red2.finish([])
    red2.block_ready(0, stage_b)
    red2.finish([])
    red2.wait_all()
    q.put((rank, [b.numpy().copy() for b in buckets], [t.numpy().copy() for t in tail], red.active, stage_a.numpy().copy(), stage_b.numpy().copy()))
    dist.destroy_process_group()


def test_overlapped_grad_reducer_averages_block_buckets_world2():
    """dist.OverlappedGradReducer (the per-block all-reduce bench.py --gpus N runs inside the backward): every bucket and
    the leftover tensors end up as the mean over ranks on every rank."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, buckets, tail, active, stage_a, stage_b in res:
        assert active
        assert np.allclose(stage_a, 1.5) and np.allclose(stage_b, 15.0)
        for i, b in enumerate(buckets):
            assert np.allclose(b, 10 * i + 1.5)          # mean of (10i + 1) and (10i + 2)
        assert np.allclose(tail[0], 100.5) and np.allclose(tail[1], 200.5)


def _dpgrads_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepsense6g_tii_b200 import dist as D
    torch.manual_seed(7)
    m = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3), torch.nn.Flatten(), torch.nn.Linear(4 * 4 * 4, 3)).to(memory_format=torch.channels_last)
    dp = D.DataParallelGrads(m)           # no GPT stages: everything rides in the flat bucket
    views = [p.grad for p in m.parameters()]
    out = []
    for step in range(2):                 # the gradient views are reused: zero_grad() must clear the previous step
        dp.zero_grad()
        x = torch.full((2, 2, 6, 6), float(rank + 1 + step))
        m(x).sum().backward()
        local = [p.grad.clone() for p in m.parameters()]
        dp.sync()
        out.append(([g.numpy() for g in local], [p.grad.numpy().copy() for p in m.parameters()]))
    same_views = all(p.grad is v for p, v in zip(m.parameters(), views))
    strides_ok = all(p.grad.stride() == p.stride() for p in m.parameters())
    q.put((rank, out, same_views, strides_ok))
    dist.destroy_process_group()


def test_data_parallel_grads_flat_bucket_world2():
    """dist.DataParallelGrads (bench.py --workload model at N > 1): parameters keep fixed gradient views of one flat buffer with the
    parameter's own strides, and sync() leaves the mean over ranks in them."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dpgrads_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, out_a, views_a, strides_a), (_, out_b, views_b, strides_b) = res
    assert views_a and views_b and strides_a and strides_b
    for step in range(2):
        for la, lb, ra, rb in zip(out_a[step][0], out_b[step][0], out_a[step][1], out_b[step][1]):
            assert np.allclose(ra, (la + lb) / 2, atol=1e-5) and np.array_equal(ra, rb)
