"""Data-parallel plumbing for the fusion stage (one process per GPU; replaces nn.DataParallel,
train2_seq.py:538).  The path shards by batch only: the single exchange step per iteration is the
gradient all-reduce (mean) of the GPT parameters; rank 0's weights are broadcast once at start-up.
Backend: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def flat_size(params):
    return sum(p.numel() for p in params)


def broadcast_params(params, src=0):
    """Make every rank start from rank `src`'s weights (one flat broadcast)."""
    params = list(params)
    if not dist.is_initialized() or dist.get_world_size() == 1 or not params:
        return
    flat = torch.cat([p.data.reshape(-1) for p in params])
    dist.broadcast(flat, src)
    off = 0
    for p in params:
        n = p.numel()
        p.data.copy_(flat[off:off + n].view_as(p))
        off += n


def allreduce_grads(params, buf=None):
    """Average gradients over ranks through one flat bucket; returns the bucket (reusable)."""
    params = [p for p in params if p.grad is not None]
    if not dist.is_initialized() or dist.get_world_size() == 1 or not params:
        return buf
    n_total = flat_size(params)
    if buf is None or buf.numel() != n_total or buf.device != params[0].grad.device:
        buf = torch.empty(n_total, device=params[0].grad.device, dtype=params[0].grad.dtype)
    off = 0
    for p in params:
        n = p.numel()
        buf[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    dist.all_reduce(buf)
    buf.div_(dist.get_world_size())
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(buf[off:off + n].view_as(p.grad))
        off += n
    return buf


class OverlappedGradReducer:
    """Bucketed gradient all-reduce overlapped with the fusion stage's own backward.

    ``FusionStageFn`` keeps all gradients of one transformer block in ONE flat fp32 buffer and calls
    ``block_ready(i, flat)`` as soon as the kernels producing block i's gradients are enqueued; the reducer
    starts an asynchronous NCCL all-reduce (average) of that bucket, which runs while the backward of blocks
    i-1 ... 0 is still computing.  ``finish(tensors)`` reduces the few remaining tensors (pos_emb) and
    makes the compute stream wait for every outstanding collective, so the gradients autograd receives are
    already averaged.  With world_size 1 (or no process group) it is a no-op.

    ``defer=True``: ``finish`` only STARTS the remaining collectives; the caller calls ``wait_all()`` once, before the
    gradients are consumed (optimizer step / end of the training step).  One reducer shared by the four fusion stages of
    a model then lets the tail all-reduce of a stage run behind the kernels that follow it instead of stalling the stream
    at the end of every stage's backward.
    """

    def __init__(self, defer=False):
        self.active = dist.is_initialized() and dist.get_world_size() > 1
        self.defer = defer
        self._works = []

    def _start(self, t):
        if dist.get_backend() == "nccl":
            self._works.append((dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True), None))
        else:  # gloo (CPU tests) has no AVG: sum now, divide once the collective has completed
            self._works.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True), t))

    def block_ready(self, index, flat):
        if self.active:
            self._start(flat)

    def finish(self, tensors=()):
        if not self.active:
            return
        for t in tensors:
            self._start(t)
        if not self.defer:
            self.wait_all()

    def wait_all(self):
        """The current stream waits for every collective started so far (gloo: and the sums are turned into means)."""
        if not self.active:
            return
        world = dist.get_world_size()
        for w, t in self._works:
            w.wait()
            if t is not None:
                t.div_(world)
        self._works = []


class DataParallelGrads:
    """Gradient exchange of a whole drop-in model under one-process-per-GPU data parallelism, capturable in a CUDA graph
    (replaces nn.DataParallel's gather / reduce, train2_seq.py:538, and stock DistributedDataParallel, whose reducer hooks
    force eager launches):

      * the GPT fusion stages average their own gradients bucket by bucket inside their backward (``OverlappedGradReducer``,
        one shared instance, waits deferred to ``sync()``);
      * every other trainable parameter (ResNet trunks, GPS MLP, join head) owns a fixed view of ONE flat fp32 buffer as its
        ``.grad`` (autograd accumulates into it in place), which ``sync()`` averages with a single all-reduce.

    Usage per step: ``zero_grad()`` -> forward / loss / backward -> ``sync()`` -> optimizer step.  BatchNorm statistics stay
    per replica, as under nn.DataParallel."""

    def __init__(self, model, gpts=()):
        self.gpts = list(gpts)
        self.reducer = OverlappedGradReducer(defer=True)
        for g in self.gpts:
            g.set_grad_reducer(self.reducer)
        own = set(id(p) for g in self.gpts for p in g.parameters())
        self.params = [p for p in model.parameters() if p.requires_grad and id(p) not in own]
        self.flat = None
        if self.params:
            dev = self.params[0].device
            self.flat = torch.zeros(flat_size(self.params), device=dev, dtype=torch.float32)
            off = 0
            for p in self.params:   # same strides as the parameter (channels_last conv weights are dense but permuted)
                p.grad = torch.as_strided(self.flat, p.shape, p.stride(), off)
                off += p.numel()

    def zero_grad(self):
        if self.flat is not None:
            self.flat.zero_()
        for g in self.gpts:
            for p in g.parameters():
                p.grad = None

    def sync(self):
        self.reducer.wait_all()
        if self.flat is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat)
            self.flat.div_(dist.get_world_size())
