"""Teacher-forced training step of the reference (train.py:702-797) as one device-side pipeline:
forward -> fused class-weighted cross-entropy -> backward -> Adam, with no host synchronisation
inside the step, and data-parallel over N GPUs (one process per GPU).

Data parallelism: every rank holds a replica; the batch is split across ranks.
  * The loss normaliser sum_i C[y_i] (train.py:736) is batch-GLOBAL, so the 16 doubles the loss
    kernel produces are all-reduced before the loss backward runs: gradients equal those of the
    single-process step on the concatenated batch.
  * Gradients live in one flat fp32 arena laid out in backward-completion order; as soon as a
    group of layers has finished its backward, its contiguous slice is all-reduced (NCCL over
    NVLink/NVSwitch) on a side stream while earlier layers' backward kernels keep running.
  * Adam (train.py:264 defaults) is one fused launch over the parameter arena and refreshes the
    bf16 weight shadows in the same pass.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _capi as K
from . import ops
from .loss import CATEGORIES, CONTROL_NAMES, loss_tables
from .model import GradArena, PackedBatch, ScoreTransformer, _Run


_SEED_OWNER = [None]        # id() of the TrainEngine whose step counter the library's seed pointer refers to


class ParamArena:
    """Re-homes the module's parameters into one flat fp32 buffer with GradArena's layout (the
    Parameters stay the same objects, so state_dict / checkpoints are unchanged) and keeps flat
    Adam moments and a flat bf16 shadow beside it."""

    def __init__(self, model: ScoreTransformer):
        self.layout = GradArena(model)
        lay = self.layout
        dev = lay.flat.device
        self.flat = torch.zeros(lay.total, dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.shadow = torch.zeros(lay.total, dtype=torch.bfloat16, device=dev) if model.compute_dtype == torch.bfloat16 else None
        params = dict(model.named_parameters())
        ext: Dict[str, torch.Tensor] = {}
        vpad = model.vpad
        for n in lay.order:
            p = params[n]
            o, numel = lay.offsets[n]
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            if self.shadow is not None and p.dim() == 2 and n != "embedding.weight":
                rows = vpad if n == "fc.weight" else p.shape[0]
                ext[n] = self.shadow[o:o + rows * p.shape[1]].view(rows, p.shape[1])
        if model.compute_dtype == torch.bfloat16:
            o, numel = lay.offsets["fc.bias"]
            ext["fc.bias"] = self.flat[o:o + vpad]
        model._w.external = ext
        self.refresh_shadow()
        # model.load_state_dict() copies into the arena views in place: the bf16 shadows must follow
        model.register_load_state_dict_post_hook(lambda module, incompatible: self.refresh_shadow())

    def refresh_shadow(self):
        if self.shadow is not None:
            ops.cast_f32_to_bf16(self.flat, self.shadow)


class GradBuckets:
    """Bucketed gradient all-reduce over the flat arena.  The arena is laid out in
    backward-completion order (fc, decoder layers top-down, encoder layers top-down, embedding),
    so each bucket is ONE contiguous slice that becomes final at a known point of the backward
    pass; `ready(prefix)` is called from _Run.backward and launches the all-reduce of every
    bucket whose groups are all final -- on `comm_stream` (NCCL) so it overlaps the remaining
    backward kernels, or inline (gloo, CPU tests)."""

    def __init__(self, model: ScoreTransformer, arena: GradArena, n_buckets: int, process_group, comm_stream=None):
        self.arena, self.pg, self.comm_stream = arena, process_group, comm_stream
        nd, ne = len(model.transformer.decoder.layers), len(model.transformer.encoder.layers)
        groups = ["fc.", "transformer.decoder.norm."]          # == GradArena.order == signalling order
        groups += [f"transformer.decoder.layers.{i}." for i in reversed(range(nd))]
        groups += ["transformer.encoder.norm."]
        groups += [f"transformer.encoder.layers.{i}." for i in reversed(range(ne))]
        groups += ["embedding."]
        self.groups = groups
        spans = {g: arena.span(g) for g in groups}
        if n_buckets and n_buckets > 0:
            per = max(1, math.ceil(len(groups) / n_buckets))
            chunks = [groups[i:i + per] for i in range(0, len(groups), per)]
        elif n_buckets == -1:
            # one bucket per layer, the embedding table alone in the last one (the smallest tail, the most collectives)
            chunks, pending = [], []
            for g in groups[:-1]:
                pending.append(g)
                if "layers." in g:
                    chunks.append(pending)
                    pending = []
            if pending:
                chunks.append(pending)
            chunks.append([groups[-1]])
        else:
            # default: THREE buckets -- (1) fc + the decoder stack, all-reduced while the encoder's backward runs; (2) the
            # encoder layers but the first, behind the first layer's backward; (3) the first encoder layer + the embedding
            # table, the exposed tail.  Measured on 8 x B200 (profiles/r02_f_dp8_timeline.txt): a 12.6-16.8 MB all-reduce
            # costs 90-100 us alone against 364 us for all 119 MB at once, and every collective kernel holds its SMs while
            # it waits for the slowest rank -- ten per-layer buckets kept NCCL kernels resident for 1.8 ms of a 12.9 ms
            # step (the exposed time beyond the 0.12 ms tail was compute slowed down under them), three need ~0.5 ms.
            first_enc = f"transformer.encoder.layers.0."
            i_enc = groups.index("transformer.encoder.norm.")
            dec, enc, tail = groups[:i_enc], [g for g in groups[i_enc:-1] if g != first_enc], [first_enc, groups[-1]]
            if ne < 2:
                dec, enc = dec + enc, []
            chunks = [c for c in (dec, enc, tail) if c]
        self.buckets = [(c, min(spans[g][0] for g in c), max(spans[g][1] for g in c)) for c in chunks]
        self.reset()

    def reset(self):
        self._done, self._launched, self.launch_order = set(), set(), []

    def ready(self, prefix: str):
        import torch.distributed as dist
        self._done.add(prefix)
        for idx, (chunk, b, e) in enumerate(self.buckets):
            if idx in self._launched or not all(g in self._done for g in chunk):
                continue
            self._launched.add(idx)
            self.launch_order.append(idx)
            view = self.arena.flat[b:e]
            if self.comm_stream is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg)
            else:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg)

    def finish(self):
        assert len(self._launched) == len(self.buckets), "a gradient bucket was never signalled"
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


class TrainEngine:
    def __init__(self, model: ScoreTransformer, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 eos_weight: float = 0.8, control_list: Sequence[str] = CONTROL_NAMES, process_group=None,
                 n_buckets: int = 0):
        K.require_cuda_device()
        self.model = model
        self.dev = model.embedding.weight.device
        self.lr, self.betas, self.eps = lr, betas, eps
        self.arena = ParamArena(model)
        self.grads = self.arena.layout                      # GradArena: flat fp32 grads + views
        model._grad_arena = self.grads
        self.grads.vpad = model.vpad
        W, C, cat = loss_tables(model.vocab_size, eos_weight, control_list)
        self.W, self.C, self.cat = W.to(self.dev), C.to(self.dev), cat.to(self.dev)
        self.ncat = len(CATEGORIES)
        self.sums = torch.zeros(K.XENT_MAX_SUMS, dtype=torch.float64, device=self.dev)
        self.denom = torch.zeros(K.XENT_MAX_SUMS, dtype=torch.float64, device=self.dev)   # [1] = batch-global sum C[y]
        self.step_count = 0
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.comm_stream = torch.cuda.Stream(device=self.dev, priority=-1) if self.world > 1 else None
        self.buckets = GradBuckets(model, self.grads, n_buckets, process_group, self.comm_stream) if self.world > 1 else None

    # ---- one step ---------------------------------------------------------------------
    def step(self, src, tgt_in, tgt_out, src_pad=None, tgt_pad=None, update: bool = True):
        """All arguments are device tensors.  Returns the device tensor `sums` (fp64[16]):
        sums[0]/sums[1] is the loss, sums[2+k]/sums[1] the k-th category term."""
        if not self.model.training:
            raise RuntimeError("TrainEngine.step needs model.train()")
        self.step_count += 1
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self.step_count * 0xD1342543DE82EF95
                + (0 if self.pg is None else 7919 * torch.distributed.get_rank(self.pg))) & 0xFFFFFFFFFFFFFFFF
        self._step_impl(src, tgt_in, tgt_out, src_pad, tgt_pad, seed, update, None)
        return self.sums

    def _step_impl(self, src, tgt_in, tgt_out, src_pad, tgt_pad, seed, update, step_dev, packed: Optional[PackedBatch] = None):
        m = self.model
        if packed is not None:
            pad_s = pad_t = None
            rows_t = packed.rows_t
            tgt_out = packed.tgt_out
        else:
            B, S = src.shape
            rows_t = B * tgt_in.shape[1]
            pad_s = None if src_pad is None else src_pad.to(torch.uint8)
            pad_t = None if tgt_pad is None else tgt_pad.to(torch.uint8)
        V, vp = m.vocab_size, m.vpad
        tg = tgt_out.reshape(-1)
        dp = self.world > 1
        if dp:
            # The normaliser sum_i C[y_i] (train.py:736) needs only the targets: compute it now and all-reduce it on
            # the communication stream while the forward pass runs, so the loss backward never waits for a collective
            import torch.distributed as dist
            ops.xent_denominator(tg, self.C, self.denom, V)
            self._on_comm(lambda: dist.all_reduce(self.denom, op=dist.ReduceOp.SUM, group=self.pg))
        run = _Run(m, src, tgt_in, pad_s, pad_t, pad_s, True, None, True, seed, False, packed=packed)
        logits = run.forward(save=True)                              # (rows, vpad) fp32
        lse = torch.empty(rows_t, dtype=torch.float32, device=self.dev)
        ops.xent_fwd(logits, tg, self.W, self.C, self.cat, self.ncat, lse, self.sums, V)
        dl = torch.empty(rows_t, vp, dtype=m.compute_dtype, device=self.dev)
        if dp:
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_event(self._denom_ready)
            ops.xent_bwd(logits, tg, self.W, lse, self.denom, dl, V, 1.0)
            # the numerators (loss value and the per-category terms that get logged) are reduced off the critical path
            self._on_comm(lambda: dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.pg), record=False)
        else:
            ops.xent_bwd(logits, tg, self.W, lse, self.sums, dl, V, 1.0)
        self.grads.flat.zero_()
        if self.buckets is not None:
            self.buckets.reset()
            m.grad_hook = self.buckets.ready
        try:
            run.backward(dl, self.grads.views)
        finally:
            m.grad_hook = None
        if self.buckets is not None:
            self.buckets.finish()
        if update:
            a = self.arena
            ops.adam_step(a.flat, self.grads.flat, a.m, a.v, a.shadow, self.step_count, self.lr, self.betas[0],
                          self.betas[1], self.eps, 1.0, step_dev=step_dev)

    # ---- padding-free batches ---------------------------------------------
    def step_packed(self, packed: PackedBatch, update: bool = True):
        """One training step on a batch without padding rows (model.PackedBatch).  Same arithmetic on the same tokens as
        step() on the padded batch: pad positions never reach the loss (ignore_index 0) nor any attention (masked keys)."""
        if not self.model.training:
            raise RuntimeError("TrainEngine.step needs model.train()")
        self.step_count += 1
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self.step_count * 0xD1342543DE82EF95
                + (0 if self.pg is None else 7919 * torch.distributed.get_rank(self.pg))) & 0xFFFFFFFFFFFFFFFF
        self._step_impl(None, None, None, None, None, seed, update, None, packed=packed)
        return self.sums

    def capture_packed(self, B: int, rows_s: int, rows_t: int, max_s: int, max_t: int):
        """CUDA graph of one packed step for fixed row counts: a replay serves every batch of B sequences whose packed
        rows fit (rows are rounded up; attention grids are sized for sequences up to max_s / max_t)."""
        dev = self.dev
        pad_src = torch.zeros(B, max_s, dtype=torch.int64, device=dev)
        pad_tgt = torch.zeros(B, max_t, dtype=torch.int64, device=dev)
        pad_src[:, :1] = 3
        pad_tgt[:, :1] = 3
        pk = PackedBatch.pack(pad_src, pad_tgt, pad_tgt, [1] * B, [1] * B, rows_s=rows_s, rows_t=rows_t)
        pk.max_s, pk.max_t = max_s, max_t
        self._g_pk = pk
        self._g_pad = dict(src=torch.zeros(B, max_s, dtype=torch.int64, device=dev),
                           tgt_in=torch.zeros(B, max_t, dtype=torch.int64, device=dev),
                           tgt_out=torch.zeros(B, max_t, dtype=torch.int64, device=dev))
        self._ctr = torch.full((1,), self.step_count, dtype=torch.int64, device=dev)
        K.check(K.lib().smer_set_seed_device_ptr(self._ctr.data_ptr()), "set_seed_device_ptr")
        _SEED_OWNER[0] = id(self)
        self._seed_base = (torch.initial_seed() * 0x9E3779B97F4A7C15
                           + (0 if self.pg is None else 7919 * torch.distributed.get_rank(self.pg))) & 0xFFFFFFFFFFFFFFFF
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step_impl(None, None, None, None, None, self._seed_base, False, self._ctr, packed=pk)
            scratch = [torch.zeros(1024, dtype=torch.float32, device=dev) for _ in range(4)]
            sh = torch.zeros(1024, dtype=torch.bfloat16, device=dev) if self.arena.shadow is not None else None
            one = torch.ones(1, dtype=torch.int64, device=dev)
            ops.adam_step(scratch[0], scratch[1], scratch[2], scratch[3], sh, 1, self.lr, self.betas[0], self.betas[1],
                          self.eps, 1.0, step_dev=one)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._ctr.add_(1)
            # the on-GPU collate (padded ids -> packed rows) is part of the captured step
            ops.pack_rows(self._g_pad["src"], pk.cu_s, pk.rows_s, pk.src_ids, pk.pos_s)
            ops.pack_rows(self._g_pad["tgt_in"], pk.cu_t, pk.rows_t, pk.tgt_in, pk.pos_t)
            ops.pack_rows(self._g_pad["tgt_out"], pk.cu_t, pk.rows_t, pk.tgt_out, pk.pos_t)
            self._step_impl(None, None, None, None, None, self._seed_base, True, self._ctr, packed=pk)
        return self

    def step_graph_packed(self, src, tgt_in, tgt_out, src_lens, tgt_lens):
        """Padded batch (host pinned or device tensors, as the reference's DataLoader yields them) + the HOST lengths of its
        sequences -> copies them in, packs them on the device and replays the captured packed step."""
        pk, gp = self._g_pk, self._g_pad
        B = src.shape[0]
        ls = torch.as_tensor(src_lens, dtype=torch.int32)
        lt = torch.as_tensor(tgt_lens, dtype=torch.int32)
        n_s, n_t = int(ls.sum()), int(lt.sum())
        if n_s > pk.rows_s or n_t > pk.rows_t or int(ls.max()) > pk.max_s or int(lt.max()) > pk.max_t or B != pk.B:
            raise RuntimeError("step_graph_packed: the batch does not fit the captured shape")
        cu = torch.zeros(2, B + 1, dtype=torch.int32)
        cu[0, 1:] = ls.cumsum(0)
        cu[1, 1:] = lt.cumsum(0)
        pk._cu2.copy_(cu.pin_memory() if src.is_pinned() else cu, non_blocking=True)
        gp["src"][:, : src.shape[1]].copy_(src, non_blocking=True)
        gp["tgt_in"][:, : tgt_in.shape[1]].copy_(tgt_in, non_blocking=True)
        gp["tgt_out"][:, : tgt_out.shape[1]].copy_(tgt_out, non_blocking=True)
        pk.n_s, pk.n_t = n_s, n_t
        self.step_count += 1
        self._graph.replay()
        return self.sums

    def _on_comm(self, fn, record=True):
        """Runs a collective on the communication stream after everything enqueued so far on the current stream
        (inline when there is no side stream: gloo / CPU tests).  record: remember its completion in `_denom_ready`."""
        if self.comm_stream is None:
            fn()
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            fn()
            if record:
                self._denom_ready = torch.cuda.Event()
                self._denom_ready.record(self.comm_stream)

    # ---- the same step captured once into a CUDA graph ---------------------------------
    def capture(self, B: int, S: int, T: int):
        """Captures one full step (forward, loss, backward, Adam) for fixed shapes.  The dropout
        seed and Adam's step number come from a device counter that the graph itself increments,
        so every replay is a new training step.  Under data parallelism the NCCL all-reduces (loss
        sums, gradient buckets on the side stream) are captured into the same graph."""
        dev = self.dev
        self._g_in = dict(src=torch.zeros(B, S, dtype=torch.int64, device=dev),
                          tgt_in=torch.zeros(B, T, dtype=torch.int64, device=dev),
                          tgt_out=torch.zeros(B, T, dtype=torch.int64, device=dev),
                          src_pad=torch.zeros(B, S, dtype=torch.bool, device=dev),
                          tgt_pad=torch.zeros(B, T, dtype=torch.bool, device=dev))
        gi = self._g_in
        gi["src"].fill_(3); gi["tgt_in"].fill_(3); gi["tgt_out"].fill_(3)
        self._ctr = torch.full((1,), self.step_count, dtype=torch.int64, device=dev)
        K.check(K.lib().smer_set_seed_device_ptr(self._ctr.data_ptr()), "set_seed_device_ptr")
        _SEED_OWNER[0] = id(self)
        self._seed_base = (torch.initial_seed() * 0x9E3779B97F4A7C15
                           + (0 if self.pg is None else 7919 * torch.distributed.get_rank(self.pg))) & 0xFFFFFFFFFFFFFFFF
        # warm-up outside capture (lazy kernel loads, cudaFuncSetAttribute), on a side stream
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            # forward/loss/backward on the dummy batch WITHOUT the update, and the Adam kernel on scratch buffers:
            # capture() must leave parameters, moments, shadows and the step counter exactly as it found them
            self._step_impl(gi["src"], gi["tgt_in"], gi["tgt_out"], gi["src_pad"], gi["tgt_pad"], self._seed_base, False,
                            self._ctr)
            scratch = [torch.zeros(1024, dtype=torch.float32, device=dev) for _ in range(4)]
            sh = torch.zeros(1024, dtype=torch.bfloat16, device=dev) if self.arena.shadow is not None else None
            one = torch.ones(1, dtype=torch.int64, device=dev)
            ops.adam_step(scratch[0], scratch[1], scratch[2], scratch[3], sh, 1, self.lr, self.betas[0], self.betas[1],
                          self.eps, 1.0, step_dev=one)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._ctr.add_(1)
            self._step_impl(gi["src"], gi["tgt_in"], gi["tgt_out"], gi["src_pad"], gi["tgt_pad"], self._seed_base, True,
                            self._ctr)
        return self

    def step_graph(self, src, tgt_in, tgt_out, src_pad, tgt_pad):
        """Copies one batch (host pinned or device tensors) into the captured step's input buffers
        and replays it.  Returns the device `sums` tensor."""
        gi = self._g_in
        gi["src"].copy_(src, non_blocking=True)
        gi["tgt_in"].copy_(tgt_in, non_blocking=True)
        gi["tgt_out"].copy_(tgt_out, non_blocking=True)
        gi["src_pad"].copy_(src_pad, non_blocking=True)
        gi["tgt_pad"].copy_(tgt_pad, non_blocking=True)
        self.step_count += 1
        self._graph.replay()
        return self.sums

    def release_graph(self):
        self._graph = None
        if _SEED_OWNER[0] == id(self):
            K.lib().smer_set_seed_device_ptr(None)
            _SEED_OWNER[0] = None

    def __del__(self):
        # the library keeps the address of `_ctr` (smer_set_seed_device_ptr): never leave it dangling
        try:
            if _SEED_OWNER[0] == id(self):
                K.lib().smer_set_seed_device_ptr(None)
                _SEED_OWNER[0] = None
        except Exception:
            pass

    # ---- optimizer state in torch.optim.Adam's layout (train.py:967-973 checkpoints) -------
    def _param_order(self):
        return [(n, p) for n, p in self.model.named_parameters()]

    def optimizer_state_dict(self) -> dict:
        """The flat-arena Adam state as `torch.optim.Adam(model.parameters(), lr).state_dict()` would hold it:
        loads into torch.optim.Adam / FusedAdam and is what train.py:970-973 saves as `optimizer_state_dict`."""
        opt = torch.optim.Adam([p for _, p in self._param_order()], lr=self.lr, betas=self.betas, eps=self.eps)
        lay = self.arena.layout
        for n, p in self._param_order():
            o, _ = lay.offsets[n]
            k = p.numel()
            opt.state[p] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.arena.m[o:o + k].view(p.shape).clone(),
                            "exp_avg_sq": self.arena.v[o:o + k].view(p.shape).clone()}
        return opt.state_dict()

    def load_optimizer_state_dict(self, sd: dict) -> None:
        """Inverse of optimizer_state_dict(); also accepts the `optimizer_state_dict` of a reference checkpoint."""
        params = self._param_order()
        opt = torch.optim.Adam([p for _, p in params], lr=self.lr, betas=self.betas, eps=self.eps)
        opt.load_state_dict(sd)
        lay = self.arena.layout
        steps = set()
        self.arena.m.zero_()
        self.arena.v.zero_()
        for n, p in params:
            st = opt.state.get(p)
            if not st:
                continue
            o, _ = lay.offsets[n]
            k = p.numel()
            self.arena.m[o:o + k].copy_(st["exp_avg"].reshape(-1))
            self.arena.v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise RuntimeError(f"TrainEngine keeps ONE Adam step number; the state dict holds {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        g = opt.param_groups[0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]
        if getattr(self, "_ctr", None) is not None:
            self._ctr.fill_(self.step_count)

    def loss_value(self) -> float:
        s = self.sums.tolist()
        return s[0] / s[1]


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr) semantics (train.py:264: default betas/eps, no weight decay,
    no amsgrad) with the update done by csrc/elementwise.cu:adam_kernel, one launch per tensor.
    For the module-API path (`loss.backward(); optim.step()`); TrainEngine uses the flat arena.

    The per-parameter state has torch.optim.Adam's layout (`step` 0-dim float tensor, `exp_avg` /
    `exp_avg_sq` shaped like the parameter), so the `optimizer_state_dict` of a reference checkpoint
    (train.py:967-973) loads into it and its own state_dict loads into torch.optim.Adam (train.py:291-295)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for g in self.param_groups:
            rows, total, max_n, step, touched = [], 0, 0, None, []
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if not torch.is_tensor(st["step"]):
                    st["step"] = torch.tensor(float(st["step"]))
                st["step"] += 1
                this_step = int(st["step"].item())             # host tensor (torch's default): no device sync
                m, v = st["exp_avg"], st["exp_avg_sq"]
                ok = (p.is_contiguous() and p.grad.is_contiguous() and m.is_contiguous() and v.is_contiguous()
                      and p.dtype == p.grad.dtype == m.dtype == v.dtype == torch.float32)
                if not ok:
                    raise RuntimeError("FusedAdam: parameters, gradients and state must be contiguous fp32 tensors")
                if step is not None and this_step != step:     # one table shares the bias correction: flush per step value
                    self._launch(rows, total, max_n, step, g)
                    rows, total, max_n = [], 0, 0
                step = this_step
                rows.append((p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()))
                touched.append(p)
                total += p.numel()
                max_n = max(max_n, p.numel())
            if rows:
                self._launch(rows, total, max_n, step, g)
            if touched:
                # The kernel wrote through raw pointers: tell autograd (and everything keyed on `_version`: the bf16
                # weight shadows of model._Weights, the decode cache stamp) that the parameters changed in place.
                torch.autograd.graph.increment_version(touched)
        return loss

    def _launch(self, rows, total, max_n, step, g):
        """All tensors of a group in one kernel launch: the (p, g, m, v, n) table goes through a pinned staging
        buffer to the device (gradients are fresh tensors every step, so the table is rebuilt every step)."""
        n = len(rows)
        dev = g["params"][0].device
        if getattr(self, "_tab_n", 0) < n:
            self._tab_host = torch.empty(n, 5, dtype=torch.int64).pin_memory()
            self._tab_dev = torch.empty(n, 5, dtype=torch.int64, device=dev)
            self._tab_n = n
            self._tab_rows = None
        if rows != getattr(self, "_tab_rows", None):       # the caching allocator usually hands the gradients the same
            ev = getattr(self, "_tab_ev", None)             # addresses every step: then the device table is still valid
            if ev is not None:
                ev.synchronize()               # the previous table has left the pinned buffer (long ago)
            self._tab_host[:n].copy_(torch.tensor(rows, dtype=torch.int64))
            self._tab_dev[:n].copy_(self._tab_host[:n], non_blocking=True)
            self._tab_ev = torch.cuda.Event()
            self._tab_ev.record()
            self._tab_rows = rows
        ops.adam_multi(self._tab_dev, n, max_n, total, step, g["lr"], g["betas"][0], g["betas"][1], g["eps"])
