"""Optimizer side of the training step on the dsfuse library.

``FusedAdamWEMA`` replaces the three per-iteration passes of ``Engine.train`` (train2_seq.py:131-134) —
``optimizer.step()`` of ``optim.AdamW(model.parameters(), lr)`` (:539), ``ema.update()`` (:315-320) and, on this path, the
fp32 -> bf16 repack of the GPT weights that otherwise opens every fusion-stage forward — by ONE multi-tensor kernel
(``dsf_adamw_ema_pack``): each parameter element is read and written once.  The registered ``GPT`` modules then find their
bf16 weight shadows already up to date and launch no pack kernels in the forward.

Same arithmetic as ``torch.optim.AdamW`` (decoupled weight decay, bias-corrected moments, eps outside the square root) and as
the reference EMA (``shadow = decay * shadow + (1 - decay) * param`` on the updated parameter).  The step counter lives on
the device, so a captured training step (CUDA graph) replays correctly.
"""
import ctypes

import torch

from . import _capi as K


class FusedAdamWEMA(torch.optim.Optimizer):
    """``FusedAdamWEMA(params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, ema=None, gpts=())``.

    ``ema``: a ``train.EMA`` whose ``register()`` has run; its shadows are updated inside the same launch and the
    ``ema.update()`` call that follows ``optimizer.step()`` in the reference loop becomes a no-op for that step.
    ``gpts``: ``modules.GPT`` instances whose bf16 weight shadows this optimizer maintains."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, ema=None, gpts=()):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.ema = ema
        self.gpts = list(gpts)
        for g in self.gpts:
            g.enable_persistent_shadows()
        self._step_dev = None
        self._sig = None
        self._tables = []   # per param group: dict(tab, tab_h (pinned), tab_d, t0_d, n, tiles), see _build
        self._keep = None

    # ------------------------------------------------------------------ table
    def _ema_of(self):
        if self.ema is None:
            return {}
        return {id(p): self.ema.shadow[n] for n, p in self.ema.model.named_parameters() if p.requires_grad and n in self.ema.shadow}

    def _shadow_targets(self):
        """{id(param): (shadow, shadow_t, copy_f32, rows, cols, row_off, ld_t)} for the GPT weights with bf16 shadows."""
        out = {}
        for g in self.gpts:
            for i, blk in enumerate(g.blocks):
                sh = g.shadow_views(i)
                C, F = g.n_embd, blk.mlp[0].weight.shape[0]
                a = blk.attn
                out[id(a.query.weight)] = (sh["wqkv"], sh["wqkv_t"], None, C, C, 0, 3 * C)
                out[id(a.key.weight)] = (sh["wqkv"], sh["wqkv_t"], None, C, C, C, 3 * C)
                out[id(a.value.weight)] = (sh["wqkv"], sh["wqkv_t"], None, C, C, 2 * C, 3 * C)
                out[id(a.proj.weight)] = (sh["wp"], sh["wp_t"], None, C, C, 0, C)
                out[id(blk.mlp[0].weight)] = (sh["w1"], sh["w1_t"], None, F, C, 0, F)
                out[id(blk.mlp[2].weight)] = (sh["w2"], sh["w2_t"], None, C, F, 0, C)
                out[id(a.query.bias)] = (None, None, sh["bqkv"][:C], 1, C, 0, 0)
                out[id(a.key.bias)] = (None, None, sh["bqkv"][C:2 * C], 1, C, 0, 0)
                out[id(a.value.bias)] = (None, None, sh["bqkv"][2 * C:], 1, C, 0, 0)
        return out

    def _build(self, work):
        """work: per group list of (p, g, m, v).  (Re)builds the tables: a ctypes array per parameter group, mirrored in a PINNED host
        buffer (so that the upload is a capturable memcpy node) and a device buffer the kernel reads."""
        ema, targets = self._ema_of(), self._shadow_targets()
        self._tables = []
        for group, items in zip(self.param_groups, work):
            if not items:
                self._tables.append(None)
                continue
            tab = (K.OptTensor * len(items))()
            tile0, tiles = [], 0
            ptr = lambda t: None if t is None else t.data_ptr()
            for k, (p, g, m, v) in enumerate(items):
                sh, sh_t, cp, rows, cols, row_off, ld_t = targets.get(id(p), (None, None, None, 1, p.numel(), 0, 0))
                tab[k] = K.OptTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), ptr(ema.get(id(p))), ptr(sh), ptr(sh_t), ptr(cp),
                                     rows, cols, row_off, ld_t, float(group["weight_decay"]), 0)
                tile0.append(tiles)
                tiles += K.opt_tiles(rows, cols, sh_t is not None)
            dev = items[0][0].device
            # two pinned host mirrors, used alternately: the one being rewritten was last read by the upload of two steps ago
            tab_h = [torch.empty((ctypes.sizeof(tab) + 15) // 16 * 16, dtype=torch.uint8).pin_memory() for _ in range(2)]
            t0_d = torch.tensor(tile0, dtype=torch.int32).to(dev)
            self._tables.append(dict(tab=tab, tab_h=tab_h, tab_d=torch.empty(tab_h[0].numel(), dtype=torch.uint8, device=dev), t0_d=t0_d,
                                     n=len(items), tiles=tiles, uploaded=[None, None], cur=0, dirty=True))

    def _refresh_grad_pointers(self, work):
        """Eager training allocates new gradient tensors every step: only the ``g`` column of the tables changes."""
        for tb, items in zip(self._tables, work):
            if tb is None:
                continue
            tab = tb["tab"]
            for k, (_, g, _, _) in enumerate(items):
                gp = g.data_ptr()
                if tab[k].g != gp:
                    tab[k].g = gp
                    tb["dirty"] = True

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        work, sig = [], []
        for group in self.param_groups:
            items = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdamWEMA: parameters must be float32 CUDA tensors (no CPU path)")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad
                if g.dtype != torch.float32 or g.stride() != p.stride():   # the kernel walks p, g, m, v as one dense blob each
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                items.append((p, g, st["exp_avg"], st["exp_avg_sq"]))
                sig.append((p.data_ptr(), g.data_ptr()))
            work.append(items)
        if not sig:
            return loss
        dev = next(it[0][0] for it in work if it).device
        with torch.cuda.device(dev):
            if self._step_dev is None:
                self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            ema_sig = None if self.ema is None else tuple(t.data_ptr() for t in self.ema.shadow.values())
            sig = (tuple(x[0] for x in sig), ema_sig, tuple(tuple(v.data_ptr() for v in g.shadow_views(0).values()) for g in self.gpts))
            if sig != self._sig:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FusedAdamWEMA: the tensor table must be built before a CUDA-graph capture — run one eager "
                                       "training step first (it allocates the moments and pins the host tables)")
                self._build(work)
                self._sig = sig
            else:
                self._refresh_grad_pointers(work)
            self._keep = work
            self._step_dev.add_(1)
            for group, tb in zip(self.param_groups, self._tables):
                if tb is None:
                    continue
                capturing = torch.cuda.is_current_stream_capturing()
                if tb["dirty"]:
                    tb["cur"] ^= 1
                    ev = tb["uploaded"][tb["cur"]]
                    if ev is not None and not capturing:   # (host synchronisation is illegal while a stream is being captured)
                        ev.synchronize()                   # never rewrite a host table an upload may still be reading
                    ctypes.memmove(tb["tab_h"][tb["cur"]].data_ptr(), tb["tab"], ctypes.sizeof(tb["tab"]))
                    tb["dirty"] = False
                # 30-60 KB read straight from pinned host memory by a small kernel (a kernel node under graph capture: replays re-read the
                # unchanged host table).  A host-to-device memcpy would queue behind the data loader's input prefetch on the copy engine.
                K.opt_upload_table(tb["tab_d"], tb["tab_h"][tb["cur"]])
                if not capturing:
                    tb["uploaded"][tb["cur"]] = torch.cuda.Event()
                    tb["uploaded"][tb["cur"]].record()
                b1, b2 = group["betas"]
                K.adamw_ema_pack(tb["tab_d"], tb["t0_d"], tb["n"], tb["tiles"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                 float(self.ema.decay) if self.ema is not None else 0.0, self._step_dev)
        for items in work:      # raw-pointer writes do not bump autograd's version counters
            for p, _, _, _ in items:
                torch.autograd.graph.increment_version(p)
        for g in self.gpts:
            g.mark_shadows_fresh()
        if self.ema is not None:
            self.ema._fused_updates = getattr(self.ema, "_fused_updates", 0) + 1
        return loss

    @property
    def steps_done(self):
        return 0 if self._step_dev is None else int(self._step_dev.item())
