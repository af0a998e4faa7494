#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native GPT fusion stage.

Metric (BASELINE.json): train samples/sec (fwd+bwd).  Workload at every N: BASELINE.json configs[1],
"single GPT fusion stage microbench (n_embd 512, 8 layers, 4 heads, 8x8 anchors, ~960 tokens) fwd+bwd,
bf16" = the stage-4 fusion stage of model2_seq.py (Encoder.forward:571-579 + GPT:175-287) at per-GPU
batch 12, seq_len 5 -> T = 962 tokens, synthetic (60, 512, 8, 8) feature maps per modality.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU): every rank runs its own batch-12 shard (weak scaling)
and the GPT gradients are all-reduced over NCCL every step (DDP semantics of train2_seq.py:538 replaced
by one-process-per-GPU).  Rank 0 prints ONE JSON line.

--impl reference times the reference's own CPU implementation of the same path (the oracle port of
model2_seq.py in oracle/fusion_ref.py — /root/reference does not exist on the GPU box) on the host
cores, on a bounded sample (batch 2) of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, L, NH, A, S, V, BATCH, SCALE = 512, 8, 4, 8, 5, 1, 12, 1
T = (V + 2) * S * A * A + 2
WORKLOAD = "gpt_fusion_stage n_embd=512 n_layer=8 n_head=4 anchors=8x8 seq_len=5 T=962 batch=12/GPU fwd+bwd"
FWD_FLOPS_PER_SAMPLE = L * (24.0 * T * C * C + 4.0 * T * T * C)  # 63.58 GF
CPU_SAMPLE_BATCH = 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_inputs(gen, batch, device="cpu", pin=False):
    """Synthetic stage-4 inputs: post-ReLU-like trunk features (non-negative, ~unit scale) and GPS embeddings."""
    feats = [torch.randn(batch * S, C, A * SCALE, A * SCALE, generator=gen).abs_() for _ in range(3)]
    gps = torch.randn(batch, 2, C, generator=gen)
    probes = [torch.randn(f.shape, generator=gen) * 1e-3 for f in feats] + [torch.randn(batch, 2, C, generator=gen) * 1e-3]
    ts = feats + [gps] + probes
    if pin:
        ts = [t.pin_memory() for t in ts]
    if device != "cpu":
        ts = [t.to(device) for t in ts]
    return ts[:3], ts[3], ts[4:]


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn():
    from oracle import fusion_ref as R
    gen = torch.Generator().manual_seed(0)
    p = R.init_gpt_params(C, NH, 4, L, T, generator=gen, pos_std=0.02)
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    feats, gps, probes = synth_inputs(gen, CPU_SAMPLE_BATCH)
    feats = [f.requires_grad_(True) for f in feats]
    gps.requires_grad_(True)

    def step():
        for t in list(p.values()) + feats + [gps]:
            t.grad = None
        (a, b, c), g = R.fusion_stage(p, feats, gps, NH, S, A, A)
        loss = sum((o * pr).sum() for o, pr in zip((a, b, c, g), probes))
        loss.backward()
        return float(loss.detach())
    return step


def time_cpu(steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return CPU_SAMPLE_BATCH * steps / dt, dt / steps * 1e3


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    v, ms = time_cpu(steps, warmup)
    cores = torch.get_num_threads()
    sample = "oracle port (oracle/fusion_ref.py) of the same stage, fp32, batch %d per step, %d threads, %s" % (CPU_SAMPLE_BATCH, cores, cpu_model())
    print(json.dumps({
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample_batch": CPU_SAMPLE_BATCH},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------- GPU arm
PDROP = 0.0  # --dropout: embd/attn/resid probability of the stage workload (the reference trains with 0.1)


def build_gpt(device):
    import types
    from deepsense6g_tii_b200 import GPT
    cfg = types.SimpleNamespace(n_views=V, fusion_dtype=torch.bfloat16)
    torch.manual_seed(100)  # the reference's seed (train2_seq.py:430-434)
    m = GPT(C, NH, 4, L, A, A, S, PDROP, PDROP, PDROP, cfg)
    with torch.no_grad():
        m.pos_emb.normal_(0, 0.02)
    return m.to(device)


def profile_families_graph(gpt, feats, gps, probes):
    """Per-call device times INSIDE a CUDA graph: the instrumented step is captured with `external` CUDA events (event-record
    nodes) around every C-ABI call and replayed; the events then bracket the kernels exactly as the graph-launched timed
    region runs them (no host launch latency between a record and its kernel).  The side stream is switched off for this
    capture so that every family is timed alone on the GPU.  Returns None when external events are unavailable."""
    from deepsense6g_tii_b200 import _capi
    import deepsense6g_tii_b200.functional as Fn
    try:
        torch.cuda.Event(enable_timing=True, external=True)
    except TypeError:
        return None
    names = ["tokens_fwd", "tokens_bwd", "layernorm_fwd", "layernorm_bwd", "gemm_bf16_nt", "gemm_bf16_tn", "colsum", "attn_fwd", "attn_bwd",
             "upsample_add_fwd", "upsample_add_bwd", "pack_block_weights", "dropout_inplace"]
    rec, orig = [], {}
    for n in names:
        f = getattr(_capi, n)
        orig[n] = f

        def wrap(*a, _f=f, _n=n, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
            e0.record()
            _f(*a, **k)
            e1.record()
            shape = None
            if _n == "gemm_bf16_nt":
                shape = (a[0].shape[0], a[1].shape[0], a[0].shape[1])
            elif _n == "gemm_bf16_tn":
                shape = (a[0].shape[0], a[0].shape[1], a[1].shape[1])
            rec.append((_n, e0, e1, shape))
        setattr(_capi, n, wrap)
    old_env = os.environ.get("DSF_WGRAD_STREAM")
    os.environ["DSF_WGRAD_STREAM"] = "0"
    runs = []
    try:
        Fn.K = _capi
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_step(gpt, feats, gps, probes)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in gpt.parameters():
            p.grad = None
        del rec[:]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            one_step(gpt, feats, gps, probes)
        for _ in range(4):
            g.replay()
            torch.cuda.synchronize()
            runs.append([(n, e0.elapsed_time(e1), shape) for n, e0, e1, shape in rec])
        runs = runs[1:]
    except Exception as ex:  # keep the bench line alive: fall back to the eager instrumentation
        sys.stderr.write("bench.py: graph-instrumented profile failed (%s); using eager events\n" % ex)
        return None
    finally:
        for n, f in orig.items():
            setattr(_capi, n, f)
        if old_env is None:
            os.environ.pop("DSF_WGRAD_STREAM", None)
        else:
            os.environ["DSF_WGRAD_STREAM"] = old_env
    fam = {}
    for i, (n, _, shape) in enumerate(runs[0]):
        d = fam.setdefault(n, {"launch_calls": 0, "ms": 0.0, "flops": 0.0})
        d["launch_calls"] += 1
        d["ms"] += statistics.median(r[i][1] for r in runs)
        if shape is not None:
            d["flops"] += 2.0 * shape[0] * shape[1] * shape[2]
    b = feats[0].shape[0] // S
    if "attn_fwd" in fam:
        fam["attn_fwd"]["flops"] = L * 4.0 * T * T * C * b
    if "attn_bwd" in fam:
        fam["attn_bwd"]["flops"] = L * 2 * 4.0 * T * T * C * b  # 2x forward (recompute not counted)
    return fam


def profile_families(gpt, feats, gps, probes):
    """One instrumented fwd+bwd: CUDA events around every C-ABI call, summed per kernel family."""
    from deepsense6g_tii_b200 import _capi
    names = ["tokens_fwd", "tokens_bwd", "layernorm_fwd", "layernorm_bwd", "gemm_bf16_nt", "gemm_bf16_tn", "colsum", "relu_bwd",
             "attn_fwd", "attn_bwd", "upsample_add_fwd", "upsample_add_bwd", "cast_f32_bf16", "relu_bwd_colsum",
             "pack_block_weights"]
    rec, orig = [], {}
    for n in names:
        f = getattr(_capi, n)
        orig[n] = f

        def wrap(*a, _f=f, _n=n, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _f(*a, **k)
            e1.record()
            shape = None
            if _n == "gemm_bf16_nt":
                shape = (a[0].shape[0], a[1].shape[0], a[0].shape[1])
            elif _n == "gemm_bf16_tn":
                shape = (a[0].shape[0], a[0].shape[1], a[1].shape[1])
            rec.append((_n, e0, e1, shape))
        setattr(_capi, n, wrap)
    REPS = 3
    runs = []
    try:
        import deepsense6g_tii_b200.functional as Fn
        Fn.K = _capi
        for _ in range(REPS):
            del rec[:]
            torch.cuda.synchronize()
            # keep the GPU busy (~15 ms spin) while the host enqueues the whole instrumented step, so that the events
            # bracket back-to-back kernel execution and not host launch latency
            torch.cuda._sleep(30_000_000)
            one_step(gpt, feats, gps, probes)
            torch.cuda.synchronize()
            runs.append([(n, e0.elapsed_time(e1), shape) for n, e0, e1, shape in rec])
    finally:
        for n, f in orig.items():
            setattr(_capi, n, f)
    fam = {}
    for i, (n, _, shape) in enumerate(runs[0]):  # per-call median over the repetitions
        d = fam.setdefault(n, {"launch_calls": 0, "ms": 0.0, "flops": 0.0})
        d["launch_calls"] += 1
        d["ms"] += statistics.median(r[i][1] for r in runs)
        if shape is not None:
            d["flops"] += 2.0 * shape[0] * shape[1] * shape[2]
    b = feats[0].shape[0] // S
    if "attn_fwd" in fam:
        fam["attn_fwd"]["flops"] = L * 4.0 * T * T * C * b
    if "attn_bwd" in fam:
        fam["attn_bwd"]["flops"] = L * 2 * 4.0 * T * T * C * b  # 2x forward (recompute not counted)
    return fam


def hbm_family_bytes(batch):
    """Algorithmic bytes per step of the HBM-bound kernel families (elements x dtype size, what each launch must
    read and write once): M = batch*T token rows, E_f = feature-map elements of the three branches."""
    M = batch * T
    e_t = M * C
    e_f = 3 * batch * S * C * (A * SCALE) * (A * SCALE)
    F = 4 * C
    per_block_w = 4 * C * C + 2 * F * C  # q, k, v, proj + mlp.0 + mlp.2 weights
    return {
        "tokens_fwd": e_f * 4 + e_t * 4 + T * C * 4,
        "tokens_bwd": e_t * 4 + 2 * e_f * 4 + T * C * 4,                     # dx in; d(out) in for the residual branch, d(feat) out; dpos out
        "layernorm_fwd": (2 * L) * e_t * (4 + 2) + e_t * (4 + 4),           # 16 x (fp32 in, bf16 out) + ln_f (fp32 out)
        "layernorm_bwd": (2 * L) * e_t * (2 + 4 + 4 + 4 + 2) + e_t * (4 + 4 + 4 + 2),  # dy bf16, x, dx_add in; dx fp32 + bf16 copy out
        "colsum": L * M * (F + 3 * C) * 2,                                   # bf16 dL/d(mlp.0 out) and dqkv
        "pack_block_weights": L * per_block_w * (4 + 2 * 2),                 # fp32 in, plain + transposed bf16 out
        "upsample_add_fwd": e_t * 4 + 2 * e_f * 4,
        "upsample_add_bwd": e_f * 4 + e_t * 4,
    }


def one_step(gpt, feats, gps, probes):
    for p in gpt.parameters():
        p.grad = None
    for t in feats:
        t.grad = None
    gps.grad = None
    outs = gpt.fuse(feats[0], feats[1], feats[2], gps)
    loss = sum((o.float() * pr).sum() for o, pr in zip(outs, probes))
    loss.backward()
    return loss


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from deepsense6g_tii_b200 import _capi
    from deepsense6g_tii_b200 import dist as D
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _capi.check_device()
    _capi.set_pdl(args.pdl)
    sm_margin = int(os.environ.get("DSF_SM_MARGIN", "0")) if world > 1 else 0
    _capi.set_sm_margin(sm_margin)
    gpt = build_gpt(dev)
    model = gpt
    if world > 1:
        # same initial weights on every rank; gradients all-reduced (mean) over NCCL every step
        D.broadcast_params(gpt.parameters())
    gen = torch.Generator().manual_seed(rank)  # data generator seed 0 + rank
    feats_h, gps_h, probes_h = synth_inputs(gen, BATCH, pin=True)
    feats = [f.to(dev).requires_grad_(True) for f in feats_h]
    gps = gps_h.to(dev).requires_grad_(True)
    probes = [p.to(dev) for p in probes_h]
    if world > 1:
        # gradients are averaged bucket-by-bucket (one bucket per transformer block) while backward is still running
        gpt.set_grad_reducer(D.OverlappedGradReducer())

    def allreduce_grads():
        pass  # done inside backward by the overlapped reducer

    def step_eager():
        loss = one_step(model, feats, gps, probes)
        allreduce_grads()
        return loss

    loss_h = torch.empty((), pin_memory=True)

    # The whole fwd+bwd step (~190 kernel launches through the C ABI) is captured ONCE into a CUDA graph and replayed:
    # the launch sequence is static (fixed shapes, torch's caching allocator keeps the captured addresses alive), so
    # replay removes the per-launch host cost and the launch gaps between dependent kernels.
    graph, graph_launches, graph_note = None, 0, "eager launches"
    pdl_note = ", programmatic dependent launch on" if args.pdl else ", programmatic dependent launch off"
    # (N > 1: the bucketed NCCL all-reduces issued from inside the backward are captured too; DSF_GRAPH_DDP=0 opts out)
    use_graph = args.graph and (world == 1 or os.environ.get("DSF_GRAPH_DDP", "1") == "1")
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in gpt.parameters():
            p.grad = None
        graph = torch.cuda.CUDAGraph()
        n0 = _capi.launch_count()
        try:
            with torch.cuda.graph(graph):
                graph_loss = step_eager()
            graph_launches = _capi.launch_count() - n0
            graph_note = "whole step captured in one CUDA graph (%d dsfuse kernels per replay)" % graph_launches
        except Exception as ex:  # e.g. a collective that cannot be captured: keep the bench line alive with eager launches
            sys.stderr.write("bench.py: CUDA-graph capture failed on rank %d (%s); falling back to eager launches\n" % (rank, ex))
            graph, graph_note = None, "eager launches (graph capture failed: %s)" % type(ex).__name__
        torch.cuda.synchronize()
        if world > 1:  # all ranks must agree: one eager rank next to replaying ranks would dead-lock the collectives
            ok = torch.tensor([1 if graph is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                graph, graph_note = None, "eager launches (graph capture failed on some rank)"

    def step_resident():
        if graph is not None:
            graph.replay()
        else:
            step_eager()

    # e2e pipeline: every step copies ONE batch of inputs pinned-host -> device (23.6 MB) and reads the loss back.  The
    # copy of step k+1's batch runs on a copy stream into a double-buffered staging area while step k computes (what a
    # training data loader with pinned-memory prefetch does); the compute stream then moves the staged batch into the
    # step's input tensors (device-to-device, ~10 us) and runs the step.
    copy_stream = torch.cuda.Stream()
    host_in = feats_h + [gps_h]
    stage = [[torch.empty_like(t, device=dev) for t in host_in] for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]  # batch landed in stage[i]
    ev_free = [torch.cuda.Event() for _ in range(2)]   # the compute stream has consumed stage[i]
    e2e_k = [0, False]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[i])
            for d, h in zip(stage[i], host_in):
                d.copy_(h, non_blocking=True)
            ev_ready[i].record(copy_stream)

    def step_e2e():
        i = e2e_k[0] & 1
        if not e2e_k[1]:  # prime the pipeline (first warm-up step only)
            prefetch(i)
            e2e_k[1] = True
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ready[i])
        with torch.no_grad():
            for d, st_ in zip(feats + [gps], stage[i]):
                d.copy_(st_, non_blocking=True)
        ev_free[i].record(cur)
        prefetch(i ^ 1)  # next step's batch: overlaps this step's compute
        e2e_k[0] += 1
        if graph is not None:
            graph.replay()
            loss_h.copy_(graph_loss.detach(), non_blocking=True)
        else:
            loss_h.copy_(step_eager().detach(), non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, finalize=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _capi.launch_count()
        ncu_range = sampler is not None and os.environ.get("DSF_NCU_RANGE") == "1"
        if ncu_range:  # `ncu --profile-from-start off` then captures exactly the timed region
            torch.cuda.profiler.start()
        e0.record()
        for _ in range(steps):
            fn()
        if finalize is not None:
            finalize()
        e1.record()
        if ncu_range:
            torch.cuda.profiler.stop()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        n1 = _capi.launch_count() + (graph_launches * steps if graph is not None else 0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n1 - n0, clocks

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    ms, launches, clocks = timed(step_resident, steps, warmup, ClockSampler(local_rank) if rank == 0 else None)
    value = BATCH * world * steps / (ms * 1e-3)
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms / steps, "value": value, "gpu_launches": launches}))
        return
    # the K-th prefetch issued inside the region must also complete inside it: K steps <-> K host-to-device batch copies
    ms_e2e, _, _ = timed(step_e2e, steps, 2, finalize=lambda: torch.cuda.current_stream().wait_event(ev_ready[e2e_k[0] & 1]))
    e2e_value = BATCH * world * steps / (ms_e2e * 1e-3)

    synced = None
    if world > 1:  # the replayed / eager steps really exchanged gradients: every rank holds the same averaged pos_emb gradient
        g = gpt.pos_emb.grad.detach().reshape(-1)[:4096].contiguous()
        gs = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
        synced = all(torch.equal(gs[0], t) for t in gs[1:]) and bool(torch.isfinite(g).all()) and float(g.abs().sum()) > 0
    if rank != 0:
        return
    pk = peaks()
    gpt.set_grad_reducer(None)  # the instrumented step runs on rank 0 only: no collectives in it
    fam = profile_families_graph(gpt, feats, gps, probes) if graph is not None else None
    how = "external CUDA events captured around every kernel of a graph-replayed step (median of 3 replays, side stream off)"
    if fam is None:
        fam = profile_families(gpt, feats, gps, probes)
        how = "CUDA events around every kernel of an eager step (median of 3; adds ~3 us per call)"
    total_ms = sum(d["ms"] for d in fam.values())
    tc = {k: d for k, d in fam.items() if d["flops"] > 0}
    dom = max(tc, key=lambda k: tc[k]["ms"])
    achieved = tc[dom]["flops"] / (tc[dom]["ms"] * 1e-3) / 1e12
    traffic, traffic_src = None, None
    import glob
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))  # newest committed capture (tags sort by round)
    tp = tps[-1] if tps else ""
    if dom == "gemm_bf16_nt" and os.path.isfile(tp):  # dram bytes per launch of the dominant kernel, from the committed ncu capture
        tj = json.load(open(tp))
        traffic, traffic_src = tj["avg_dram_bytes_per_launch"], tj["source"]
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tc_sust"], "unit": "TFLOP/s",
                "frac": achieved / pk["tc_sust"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["src"] + " bf16_tflops_sustained",
                "launches_per_step": tc[dom]["launch_calls"], "avg_launch_ms": tc[dom]["ms"] / tc[dom]["launch_calls"],
                "share_of_step": tc[dom]["ms"] / total_ms, "timing": how,
                "families": {k: {"ms": round(d["ms"], 4), "calls": d["launch_calls"],
                                 "tflops": (d["flops"] / (d["ms"] * 1e-3) / 1e12) if d["flops"] else None} for k, d in fam.items()}}
    try:  # HBM-bound families: algorithmic GB/s against the measured copy bandwidth (launch-bound at 10-25 us per launch)
        hb = hbm_family_bytes(feats[0].shape[0] // S)
        for k, d in roofline["families"].items():
            if k in hb and d["ms"] > 0:
                d["gbps"] = round(hb[k] / (d["ms"] * 1e-3) / 1e9, 1)
                d["frac_of_hbm_peak"] = round(d["gbps"] / pk["hbm"], 3)
        roofline["hbm_peak_gbps"] = pk["hbm"]
    except Exception as ex:  # never lose the bench line over a diagnostic
        sys.stderr.write("bench.py: HBM family roofline skipped (%s)\n" % ex)
    cpu_v, cpu_ms = time_cpu(2, 1)
    cores = torch.get_num_threads()
    h2d = sum(t.numel() * t.element_size() for t in feats_h) + gps_h.numel() * gps_h.element_size()
    step_flops = 3.0 * FWD_FLOPS_PER_SAMPLE * BATCH
    out = {
        "metric": "train samples/sec (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "seq_len": S, "parallelism": "dp%d" % world,
                   "launch": graph_note + pdl_note, "dropout": PDROP,
                   "l2": "per-step working set ~1.5 GB of saved activations > 126 MB L2; no explicit flush",
                   "grad_allreduce": "NCCL all-reduce (avg) of 25.7 M fp32 grads per step, one bucket per block, overlapped with backward" if world > 1 else "none (1 GPU)",
                   "grads_identical_across_ranks": synced},
        "stage_tflops": step_flops * world / (ms / steps * 1e-3) / 1e12,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / steps,
                "pipeline": "one pinned-host -> device batch copy per step on a copy stream (double-buffered), overlapped with the previous step's compute; loss read back every step"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "cpu_baseline": {"value": cpu_v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "oracle port fp32, batch %d, 1 warm-up + 2 timed steps, %s" % (CPU_SAMPLE_BATCH, cpu_model())},
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------- full-model workload (configs[2])
def run_model(args, rank, world, local_rank):
    """BASELINE.json configs[2]: full model2_seq training step — drop-in TransFuser (ResNet trunks + 4 fusion stages on
    the dsfuse kernels + join MLP), bf16, batch 12 per GPU, seq_len 5, 256x256 inputs, dropout 0.1 (config_seq.py:39-41),
    focal loss + fused AdamW + EMA, synthetic data; N > 1: DistributedDataParallel over NCCL (batch-sharded)."""
    import types
    import torch.distributed as dist
    from deepsense6g_tii_b200 import TransFuser, _capi
    from deepsense6g_tii_b200.train import EMA, FocalLoss, synthetic_batch, train_step
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _capi.check_device()
    _capi.set_pdl(args.pdl)
    torch.backends.cudnn.benchmark = True  # as the reference (train2_seq.py:16)
    cfg = types.SimpleNamespace(seq_len=S, pred_len=4, n_views=1, vert_anchors=A, horz_anchors=A, n_embd=512, block_exp=4, n_layer=L,
                                n_head=NH, embd_pdrop=0.1, attn_pdrop=0.1, resid_pdrop=0.1, add_velocity=1, fusion_dtype=torch.bfloat16)
    torch.manual_seed(100)
    model = TransFuser(cfg, dev).to(memory_format=torch.channels_last).train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True)
    crit = FocalLoss()
    use_graph = args.graph and world == 1
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=use_graph)
    ema = EMA(model, 0.999)
    ema.register()
    gen = torch.Generator().manual_seed(rank)
    host = synthetic_batch(BATCH, S, 256, generator=gen, pin=True)

    def to_dev(b, stream=None):
        imgs, lids, rads, gps, soft, beam = b
        mv = lambda t: t.to(dev, non_blocking=True)
        return ([mv(t).contiguous(memory_format=torch.channels_last) for t in imgs], [mv(t) for t in lids], [mv(t) for t in rads],
                mv(gps), mv(soft), mv(beam))

    resident = to_dev(host)
    loss_h = torch.empty((), pin_memory=True)

    def step_eager(b=None):
        return train_step(net, resident if b is None else b, crit, opt, ema, autocast_dtype=torch.bfloat16)

    # N = 1: the whole training step (trunks, 4 fusion stages with graph-safe dropout, loss, backward, capturable fused
    # AdamW, multi-tensor EMA) is captured in one CUDA graph on static input tensors and replayed.
    graph, graph_loss, launch_note, graph_launches = None, None, "eager launches", 0
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _capi.launch_count()
        try:
            with torch.cuda.graph(graph):
                graph_loss = step_eager()
            launch_note = "whole training step captured in one CUDA graph (%d dsfuse kernels per replay)" % (_capi.launch_count() - n0)
            graph_launches = _capi.launch_count() - n0
        except Exception as ex:
            sys.stderr.write("bench.py: CUDA-graph capture of the model step failed (%s); eager launches\n" % ex)
            graph = None
        torch.cuda.synchronize()

    def step_resident():
        if graph is not None:
            graph.replay()
            return graph_loss
        return step_eager()

    # e2e: one pinned-host -> device batch copy (94 MB) per step into preallocated, double-buffered staging tensors on a
    # copy stream, overlapped with the previous step's compute; the step then reads the staged batch (graph mode: after a
    # device-to-device move into the graph's static input tensors)
    copy_stream = torch.cuda.Stream()
    flat = lambda b: list(b[0]) + list(b[1]) + list(b[2]) + list(b[3:5])
    host_flat = flat(host)
    stage = [[torch.empty_like(t, device=dev) for t in host_flat] for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    e2e_k = [0, False]

    def unflat(ts):
        return ts[0:S], ts[S:2 * S], ts[2 * S:3 * S], ts[3 * S], ts[3 * S + 1], None

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[i])
            for d, h in zip(stage[i], host_flat):
                d.copy_(h, non_blocking=True)
            ev_ready[i].record(copy_stream)

    def step_e2e():
        i = e2e_k[0] & 1
        if not e2e_k[1]:
            prefetch(i)
            e2e_k[1] = True
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ready[i])
        if graph is not None:
            with torch.no_grad():
                for dst, src in zip(flat(resident), stage[i]):
                    dst.copy_(src, non_blocking=True)
            ev_free[i].record(cur)
            prefetch(i ^ 1)
            graph.replay()
            loss_h.copy_(graph_loss.detach(), non_blocking=True)
        else:
            prefetch(i ^ 1)
            loss_h.copy_(step_eager(unflat(stage[i])).detach(), non_blocking=True)
            ev_free[i].record(cur)
        e2e_k[0] += 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, finalize=None):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _capi.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        if finalize is not None:
            finalize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (_capi.launch_count() - n0) + (graph_launches * steps if graph is not None else 0)

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(step_resident, steps, warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, steps, 2, finalize=lambda: torch.cuda.current_stream().wait_event(ev_ready[e2e_k[0] & 1]))
    loss_val = float(loss_h)
    if rank != 0:
        return
    h2d = sum(t.numel() * t.element_size() for t in host[0] + host[1] + host[2]) + sum(t.numel() * t.element_size() for t in host[3:])
    n_params = sum(p.numel() for p in model.parameters())
    print(json.dumps({
        "metric": "train samples/sec (fwd+bwd)", "value": BATCH * world * steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "full model2_seq TransFuser training step: fwd + focal loss + bwd + fused AdamW + EMA, batch 12/GPU, seq_len 5, "
                               "256x256 inputs, dropout 0.1, %d parameters, ResNet trunks in stock PyTorch (channels_last, bf16 autocast), "
                               "4 fusion stages on dsfuse kernels" % n_params,
                   "global_batch": BATCH * world, "seq_len": S, "parallelism": "dp%d" % world, "launch": launch_note + ", PDL " + ("on" if args.pdl else "off"),
                   "l2": "per-step working set (activations of 3 ResNets + 4 fusion stages) >> 126 MB L2; no explicit flush",
                   "grad_allreduce": "torch DistributedDataParallel (NCCL, bucketed, overlapped)" if world > 1 else "none (1 GPU)"},
        "e2e": {"value": BATCH * world * steps / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / steps, "pipeline": "next batch copied pinned-host -> device on a copy stream during the current step"},
        "gpu_launches": launches, "clocks": clocks, "final_loss": loss_val,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="stage", choices=["stage", "model"],
                    help="stage = BASELINE.json configs[1] (default, the driver's line); model = configs[2], the full training step")
    ap.add_argument("--quick", action="store_true", help="profiling runs: skip the e2e leg, the instrumented step and the CPU baseline")
    ap.add_argument("--no-pdl", dest="pdl", action="store_false", help="disable programmatic dependent launch of the hot kernels")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="stage workload: embd/attn/resid dropout probability (default 0 = the parity configuration; the reference trains with 0.1)")
    ap.add_argument("--anchors", type=int, default=8, choices=[8, 16],
                    help="16 = BASELINE.json configs[4], the scaled fusion stage: 16x16 anchors -> T = 3842 tokens (not the driver's line)")
    args = ap.parse_args()
    if args.dropout > 0:
        global PDROP
        PDROP = args.dropout
    if args.anchors != 8:
        global A, T, WORKLOAD, FWD_FLOPS_PER_SAMPLE
        A = args.anchors
        T = (V + 2) * S * A * A + 2
        WORKLOAD = "gpt_fusion_stage n_embd=512 n_layer=8 n_head=4 anchors=%dx%d seq_len=5 T=%d batch=12/GPU fwd+bwd (scaled config)" % (A, A, T)
        FWD_FLOPS_PER_SAMPLE = L * (24.0 * T * C * C + 4.0 * T * T * C)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fusion stage has no CPU path (use --impl reference for the CPU baseline)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "model":
            run_model(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
