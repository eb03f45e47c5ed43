"""KV-cached infilling decode (K12).

The reference has no cache: generation.model_generate (generation.py:209-225) re-runs the
whole encoder and decoder for every token.  Two entry points replace that arithmetic:

* `cached_forward` -- sits *inside* ScoreTransformer.forward for the call shape model_generate
  makes (batch 1, no padding masks, nopeek mask, eval mode).  The module remembers the piece
  (encoder memory, per-layer cross K/V) and the decoder prefix (per-layer self K/V, logits
  rows); a call whose `tgt` extends the cached prefix computes only the new positions, a
  diverging `tgt` (span regeneration, evaluation.py:1303-1335) truncates the cache first.
  generation.py therefore runs unchanged and gets the same (T,V) logits.
* `InfillDecoder` -- batched generation of many independent pieces entirely on the device:
  per step gather -> embed -> decoder layers (KV append + attention over the caches) -> fc ->
  grammar-masked sampling + span bookkeeping (csrc/sample.cu), no host round trip per token.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _capi as K
from . import ops
from . import spans

TARGET_CODES = {"r": 0, "d": 1, "o": 2, "p": 3, "t": 4}


# ----------------------------------------------------------------------------------------
# shared: encoder + per-layer cross K/V
# ----------------------------------------------------------------------------------------
def _encode(model, src, src_pad):
    """Runs the encoder stack; returns (mem (B*S,d), run)."""
    from .model import _Run
    B, S = src.shape
    dummy_tgt = src[:, :1]
    run = _Run(model, src, dummy_tgt, src_pad, None, src_pad, False, None, False, 0, False)
    mem = run.encode(save=False)
    return mem, run


def _cross_kv(model, run, mem, out=None):
    """Per decoder layer the K|V projection of the encoder memory; `out`: persistent buffers to fill."""
    d = model.d_model
    res = []
    for i, layer in enumerate(model.transformer.decoder.layers):
        ap = model._attn_p(layer.multihead_attn, f"transformer.decoder.layers.{i}.multihead_attn.")
        kv = out[i] if out is not None else torch.empty(mem.shape[0], 2 * d, dtype=model.compute_dtype, device=mem.device)
        ops.gemm_nt(mem, ap.w[d:3 * d], kv, bias=ap.b[d:])
        res.append(kv)
    return res


# ----------------------------------------------------------------------------------------
# drop-in incremental forward (batch 1)
# ----------------------------------------------------------------------------------------
class _PieceCache:
    def __init__(self, model, src):
        self.src_cpu = src.detach().cpu()
        self.S = src.shape[1]
        dev = src.device
        dt = model.compute_dtype
        d = model.d_model
        mem, run = _encode(model, src, None)
        self.cross = _cross_kv(model, run, mem)
        self.cap = 0
        self.self_kv: List[torch.Tensor] = []
        self.logits = None
        self.weights = None
        self.tokens: List[int] = []
        self.dev, self.dt, self.d = dev, dt, d
        self.nl = len(model.transformer.decoder.layers)
        self.V = model.vocab_size
        self.vpad = model.vpad

    def reserve(self, T: int, want_w: bool):
        if T > self.cap:
            cap = max(256, 1 << (T - 1).bit_length())
            old_kv, old_lg, old_w, n = self.self_kv, self.logits, self.weights, len(self.tokens)
            self.self_kv = [torch.empty(cap, 2 * self.d, dtype=self.dt, device=self.dev) for _ in range(self.nl)]
            self.logits = torch.empty(cap, self.vpad, dtype=torch.float32, device=self.dev)
            self.weights = torch.zeros(self.nl, cap, self.S, dtype=torch.float32, device=self.dev) if want_w else None
            if n:
                for a, b in zip(self.self_kv, old_kv):
                    a[:n].copy_(b[:n])
                self.logits[:n].copy_(old_lg[:n])
                if want_w and old_w is not None:
                    self.weights[:, :n].copy_(old_w[:, :n])
            self.cap = cap
        if want_w and self.weights is None:
            self.weights = torch.zeros(self.nl, self.cap, self.S, dtype=torch.float32, device=self.dev)
            self.tokens = []                      # rows for the cached prefix were never computed


def cached_forward(model, src, tgt, want_w: bool):
    """ScoreTransformer.forward for (1,S)/(1,T) inputs with the nopeek mask (generation.py:217)."""
    dev = src.device
    pc: Optional[_PieceCache] = model._decode_cache
    src_cpu = src.detach().cpu()
    stamp = sum(p._version for p in model.parameters())        # weights changed in place -> cache is stale
    if (pc is None or pc.dt != model.compute_dtype or pc.S != src.shape[1] or pc.stamp != stamp
            or not torch.equal(pc.src_cpu, src_cpu)):
        pc = _PieceCache(model, src)
        pc.stamp = stamp
        model._decode_cache = pc
    toks = tgt[0].tolist()
    T = len(toks)
    pc.reserve(T, want_w)
    P = 0
    lim = min(len(pc.tokens), T)
    while P < lim and pc.tokens[P] == toks[P]:
        P += 1
    if P < T:
        _extend(model, pc, tgt, P, T, want_w)
    pc.tokens = toks
    V = model.vocab_size
    logits = pc.logits[:T, :V].unsqueeze(0)
    nl = pc.nl
    if want_w:
        weights = pc.weights[:, :T].unsqueeze(0)                       # (1,Ld,T,S)
    else:
        weights = torch.zeros((), device=dev).expand(1, nl, T, pc.S)
    return logits, weights


def _extend(model, pc: _PieceCache, tgt, P: int, T: int, want_w: bool):
    """Computes decoder positions P..T-1 against the cached keys 0..P-1."""
    d, H, ff = model.d_model, model.nhead, model.dim_feedforward
    dh = d // H
    dt, dev = pc.dt, pc.dev
    k = T - P

    def new(r, c, dtype=None):
        return torch.empty(r, c, dtype=dtype or dt, device=dev)

    def ln(branch, resid, w, b):
        y = new(branch.shape[0], d)
        ops.layernorm_fwd(branch, resid, w, b, None, y, None, None)
        return y

    pe = model.pos_enc.pe.view(-1, d)
    y = new(k, d)
    ops.embed_pe(tgt[:, P:T].contiguous(), model.embedding.weight.detach(), pe, y, math.sqrt(d), P)
    for i, layer in enumerate(model.transformer.decoder.layers):
        lp = model._layer_p(layer, f"transformer.decoder.layers.{i}.")
        sa, ca = lp.sa, lp.ca
        q = new(k, d)
        ops.gemm_nt(y, sa.w[:d], q, bias=sa.b[:d])
        kv = pc.self_kv[i]
        ops.gemm_nt(y, sa.w[d:3 * d], kv[P:T], bias=sa.b[d:])          # KV-cache append
        o = new(k, d)
        a = ops.attn_args(q, kv[:T, :d], kv[:T, d:], o, 1, H, k, T, dh, causal=True, q_pos0=P)
        ops.attn_fwd(a)
        proj = new(k, d)
        ops.gemm_nt(o, sa.wo, proj, bias=sa.bo)
        y1 = ln(proj, y, *lp.ln[0])
        q2 = new(k, d)
        ops.gemm_nt(y1, ca.w[:d], q2, bias=ca.b[:d])
        ckv = pc.cross[i]
        o2 = new(k, d)
        lse = torch.empty(1, H, k, dtype=torch.float32, device=dev) if want_w else None
        a2 = ops.attn_args(q2, ckv[:, :d], ckv[:, d:], o2, 1, H, k, pc.S, dh, lse=lse)
        ops.attn_fwd(a2)
        if want_w:
            ops.attn_weights(a2, pc.weights[i, P:T])
        proj2 = new(k, d)
        ops.gemm_nt(o2, ca.wo, proj2, bias=ca.bo)
        y2 = ln(proj2, y1, *lp.ln[1])
        h = new(k, ff)
        ops.gemm_nt(y2, lp.w1, h, bias=lp.b1, flags=K.EPI_RELU)
        f = new(k, d)
        ops.gemm_nt(h, lp.w2, f, bias=lp.b2)
        y = ln(f, y2, *lp.ln[2])
    dn = model.transformer.decoder.norm
    yo = ln(y, None, dn.weight.detach(), dn.bias.detach())
    wfc, bfc = model._fc_p()
    ops.gemm_nt(yo, wfc, pc.logits[P:T], bias=bfc)


# ----------------------------------------------------------------------------------------
# batched on-device infilling
# ----------------------------------------------------------------------------------------
class InfillDecoder:
    """Generates the infill spans of many independent pieces (BASELINE config 4).

    pieces: list of 1-D int arrays -- the masked source sequences (output of
            generation.mask_bar_and_track), one `m_0` per span.
    targets: list of span-type strings per piece ('r','d','o','p','t'; generation.py:485-492).
    The per-token arithmetic is that of generation_all's loop (generation.py:528-687) with the
    K/V of earlier positions cached instead of recomputed.
    """

    def __init__(self, model, *, mode: str = "greedy", temperature: float = 1.0, top_p: float = 0.9, top_k: int = 0,
                 seed: int = 0, max_len: int = 1024, max_span: int = 100,
                 all_controls: Sequence[int] = tuple(range(242, 308)), splits: int = 1, use_graph: bool = True):
        K.require_cuda_device()
        if int(max_len) > model.pos_enc.pe.shape[0]:
            raise ValueError(f"max_len {max_len} exceeds the model's positional table ({model.pos_enc.pe.shape[0]} rows)")
        if int(max_len) < 2:
            raise ValueError("max_len must be at least 2")
        self.m = model
        self.mode = {"greedy": K.SAMPLE_GREEDY, "multinomial": K.SAMPLE_MULTINOMIAL, "top_p": K.SAMPLE_TOP_P,
                     "top_k": K.SAMPLE_TOP_K}[mode]
        self.temperature, self.top_p, self.top_k = float(temperature), float(top_p), int(top_k)
        self.seed, self.max_len, self.max_span = int(seed), int(max_len), int(max_span)
        self.splits = int(splits)
        self.use_graph = use_graph
        bm = np.zeros(10, dtype=np.uint32)
        for c in all_controls:
            bm[c >> 5] |= np.uint32(1 << (c & 31))
        self.control_bitmap_host = bm
        self.trace_distributions = False    # parity instrumentation: keep every step's masked softmax (n, max_len, V) fp64
        self.trace_masked = None
        self.steps_run = 0
        self.kernel_launches = 0
        self.profile = None                 # list of (kind, start, end) CUDA events when profiling one eager step

    # -- state ------------------------------------------------------------------------
    def _weight_ptrs(self):
        """Addresses of every weight tensor the captured step reads (bf16 shadows are re-made when weights change)."""
        ptrs = []
        for lp in self.layer_p:
            for ap in (lp.sa, lp.ca):
                ptrs += [ap.w.data_ptr(), ap.b.data_ptr(), ap.wo.data_ptr(), ap.bo.data_ptr()]
            ptrs += [lp.w1.data_ptr(), lp.b1.data_ptr(), lp.w2.data_ptr(), lp.b2.data_ptr()]
            ptrs += [t.data_ptr() for pair in lp.ln for t in pair]
        ptrs += [t.data_ptr() for t in self.fc_p]
        return tuple(ptrs)

    def _alloc(self, n, S, max_spans, dev):
        """Device buffers of one problem shape.  They persist across generate() calls of the same shape so that the
        captured CUDA graph (which bakes their addresses) can be replayed without being re-captured."""
        m = self.m
        d, H = m.d_model, m.nhead
        nl = len(m.transformer.decoder.layers)
        dt = m.compute_dtype
        L = self.max_len
        self.n, self.S, self.dev, self.max_spans = n, S, dev, max_spans
        self.src_dev = torch.zeros(n, S, dtype=torch.int64, device=dev)
        self.pad_dev = torch.zeros(n, S, dtype=torch.uint8, device=dev)
        self.src_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.cross = [torch.empty(n * S, 2 * d, dtype=dt, device=dev) for _ in range(nl)]
        self.self_kv = [torch.zeros(n, L, 2 * d, dtype=dt, device=dev) for _ in range(nl)]
        self.targets = torch.zeros(n, max_spans, dtype=torch.int8, device=dev)
        self.n_spans = torch.zeros(n, dtype=torch.int32, device=dev)
        self.nwd = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.tok_buf = torch.zeros(n, L, dtype=torch.int64, device=dev)
        self.cur_len = torch.ones(n, dtype=torch.int32, device=dev)
        self.fed_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.span_start = torch.zeros(n, dtype=torch.int32, device=dev)
        self.span_idx = torch.zeros(n, dtype=torch.int32, device=dev)
        self.done = torch.zeros(n, dtype=torch.int32, device=dev)
        self.gen_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self.state = torch.zeros(n, dtype=torch.int32, device=dev)
        self.bitmap = torch.from_numpy(self.control_bitmap_host.view(np.int32).copy()).to(dev)
        self.ids = torch.zeros(n, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(n, dtype=torch.int32, device=dev)
        ws = K.lib().smer_decode_attn_workspace_bytes(n, H, d // H, self.splits)
        self.ws = torch.empty(max(ws, 4) // 4, dtype=torch.float32, device=dev)
        # static activation buffers (graph-capturable step)
        ff = m.dim_feedforward
        mk = lambda c, t=None: torch.empty(n, c, dtype=t or dt, device=dev)
        self.buf = dict(x=mk(d), qkv=mk(3 * d), o=mk(d), proj=mk(d), y1=mk(d), z=mk(d), q2=mk(d), o2=mk(d), y2=mk(d),
                        h=mk(ff), f=mk(d), y3=mk(d), yo=mk(d), logits=mk(m.vpad, torch.float32), xres=mk(d), z3=mk(d))
        # small batches take the small-M projections with fused LayerNorm prologues (SMER_DECODE_SMALL=0/1 overrides)
        import os
        env = os.environ.get("SMER_DECODE_SMALL")
        self.small = dt == torch.bfloat16 and d % 128 == 0 and ff % 128 == 0 and d <= 1024 and (n <= 256 if env is None else env == "1")
        self.graph = None

    def _setup(self, pieces, targets, nwd, seq_base):
        m = self.m
        dev = m.embedding.weight.device
        n = len(pieces)
        lens = np.fromiter((len(p) for p in pieces), dtype=np.int64, count=n)
        S = (int(lens.max()) + 7) // 8 * 8
        src_np = np.zeros((n, S), dtype=np.int64)
        flat = np.concatenate([np.asarray(p, dtype=np.int64) for p in pieces])
        cols = np.arange(S)[None, :] < lens[:, None]               # valid positions, row-major == concat order
        src_np[cols] = flat
        nsp = np.fromiter((len(t) for t in targets), dtype=np.int64, count=n)
        max_spans = max(1, int(nsp.max()))
        tg_np = np.zeros((n, max_spans), dtype=np.int8)
        codes = np.fromiter((TARGET_CODES[c] for t in targets for c in t), dtype=np.int8, count=int(nsp.sum()))
        tg_np[np.arange(max_spans)[None, :] < nsp[:, None]] = codes
        self.layer_p = [m._layer_p(l, f"transformer.decoder.layers.{i}.") for i, l in enumerate(m.transformer.decoder.layers)]
        self.fc_p = m._fc_p()
        key = (n, S, max_spans, self.max_len, seq_base, str(dev), self.mode, self.temperature, self.top_p, self.top_k,
               self.seed, self.splits, self._weight_ptrs())
        if getattr(self, "_key", None) != key:
            self._alloc(n, S, max_spans, dev)
            self._key = key
        self.seq_base = seq_base
        src = torch.from_numpy(src_np).pin_memory()
        pad = torch.from_numpy((~cols).astype(np.uint8)).pin_memory()
        self.h2d_bytes = src.numel() * 8 + pad.numel()
        self.src_dev.copy_(src, non_blocking=True)
        self.pad_dev.copy_(pad, non_blocking=True)
        self.src_len.copy_(torch.from_numpy(lens.astype(np.int32)))
        self._ev0 = torch.cuda.Event(enable_timing=True)
        self._ev0.record()                                        # device work starts here (encoder)
        mem, run = _encode(m, self.src_dev, self.pad_dev)
        _cross_kv(m, run, mem, out=self.cross)
        del mem, run
        for kv in self.self_kv:
            kv.zero_()
        self.targets.copy_(torch.from_numpy(tg_np))
        self.n_spans.copy_(torch.from_numpy(nsp.astype(np.int32)))
        self.nwd.copy_(torch.tensor([1 if x else 0 for x in nwd], dtype=torch.uint8))
        self.tok_buf.zero_()
        self.tok_buf[:, 0] = 2                                     # every stream opens with m_0
        self.cur_len.fill_(1)
        for t in (self.fed_len, self.span_start, self.span_idx, self.gen_count, self.state, self.ids, self.pos):
            t.zero_()
        self.done.copy_((self.n_spans == 0).to(torch.int32))
        if self.trace_distributions:
            self.trace_masked = torch.zeros(n, self.max_len, m.vocab_size, dtype=torch.float64, device=dev)
            self.trace_span = torch.full((n, self.max_len), -1, dtype=torch.int32, device=dev)
            self.graph = None                                 # the trace buffer's address is baked into a captured step
        elif self.trace_masked is not None:
            self.trace_masked = None
            self.graph = None

    def _decode_attn(self, q, new_k, new_v, kc, vc, out, kv_len, key_pad, ld_cache, cache_stride, cache_len, ld_pad):
        m = self.m
        prof = self.profile
        if prof is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream())
        a = K.DecodeAttnArgs()
        a.q, a.new_k, a.new_v = q.data_ptr(), ops._p(new_k), ops._p(new_v)
        a.k_cache, a.v_cache, a.out = kc.data_ptr(), vc.data_ptr(), out.data_ptr()
        a.kv_len, a.key_pad, a.workspace = kv_len.data_ptr(), ops._p(key_pad), self.ws.data_ptr()
        a.ldq, a.ld_new, a.ldo = q.stride(0), (new_k.stride(0) if new_k is not None else 0), out.stride(0)
        a.ld_cache, a.cache_stride, a.ld_pad = ld_cache, cache_stride, ld_pad
        a.n_seq, a.H, a.dh, a.cache_len, a.splits = self.n, m.nhead, m.d_model // m.nhead, cache_len, self.splits
        a.dtype = K.dt(q)
        a.scale = 1.0 / math.sqrt(m.d_model // m.nhead)
        a.done = self.done.data_ptr()                 # finished pieces stop streaming their K/V
        K.check(K.lib().smer_decode_attn(C.byref(a), K.stream()), "decode_attn")
        if prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream())
            prof.append(("cross" if new_k is None else "self", e0, e1))

    def _step_small(self):
        """The decode step for small batches (n <= 256 pieces per GPU, bf16): every projection through the small-M kernel
        (a 64-row x 8-column slab per CTA, so that even a [n, 512] x [512, 512] product covers the GPU) with the LayerNorms
        fused into the prologue of the product that consumes them -- 35 launches per token instead of 48, each a programmatic
        dependent launch (csrc/common.cuh smer_launch_pdl) so that a kernel's weight loads start while its predecessor drains."""
        lib = K.lib()
        lib.smer_set_pdl(1)                # every launch below is a decode-chain kernel: programmatic dependent launches
        try:
            self._step_small_launches()
        finally:
            lib.smer_set_pdl(0)

    def _step_small_launches(self):
        m, b = self.m, self.buf
        d = m.d_model
        n, L, S = self.n, self.max_len, self.S
        lib = K.lib()
        K.check(lib.smer_decode_embed(self.tok_buf.data_ptr(), self.cur_len.data_ptr(), self.fed_len.data_ptr(),
                                      self.done.data_ptr(), self.pos.data_ptr(), m.embedding.weight.data_ptr(),
                                      m.pos_enc.pe.data_ptr(), b["x"].data_ptr(), K.dt(b["x"]), n, L, d, m.vocab_size,
                                      math.sqrt(d), K.stream()), "decode_embed")
        lin = ops.decode_linear
        x_in, prev_ln = b["x"], None                  # layer input: rows as they are, or pre-LN sums + that LayerNorm
        launches = 1
        for i, lp in enumerate(self.layer_p):
            sa, ca = lp.sa, lp.ca
            if prev_ln is None:
                lin(x_in, sa.w, b["qkv"], bias=sa.b)
                resid = x_in
            else:
                lin(x_in, sa.w, b["qkv"], bias=sa.b, ln=prev_ln, ln_out=b["xres"])
                resid = b["xres"]
            qkv, kv = b["qkv"], self.self_kv[i]
            self._decode_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kv[:, :, :d], kv[:, :, d:], b["o"], self.pos,
                              None, 2 * d, L * 2 * d, L, 0)
            lin(b["o"], sa.wo, b["z"], bias=sa.bo, resid=resid)                       # z1 = x + SelfAttn(x)
            lin(b["z"], ca.w[:d], b["q2"], bias=ca.b[:d], ln=lp.ln[0], ln_out=b["y1"])  # y1 = LN1(z1); q = y1 Wq
            ckv = self.cross[i]
            self._decode_attn(b["q2"], None, None, ckv[:, :d], ckv[:, d:], b["o2"], self.src_len, None, 2 * d,
                              S * 2 * d, S, 0)
            lin(b["o2"], ca.wo, b["z"], bias=ca.bo, resid=b["y1"])                    # z2 = y1 + CrossAttn
            lin(b["z"], lp.w1, b["h"], bias=lp.b1, relu=True, ln=lp.ln[1], ln_out=b["y2"])   # y2 = LN2(z2); h = relu(y2 W1)
            lin(b["h"], lp.w2, b["z3"], bias=lp.b2, resid=b["y2"])                    # z3 = y2 + FFN
            x_in, prev_ln = b["z3"], lp.ln[2]
            launches += 8
        dn = m.transformer.decoder.norm
        # LN3 of the last layer, then the decoder's final norm, both in the prologue of the vocabulary projection
        lin(x_in, self.fc_p[0], b["logits"], bias=self.fc_p[1], ln=prev_ln, ln2=(dn.weight.detach(), dn.bias.detach()))
        self._sample()
        self.launches_per_step = launches + 2

    def _sample(self):
        m, b = self.m, self.buf
        n, L = self.n, self.max_len
        a = K.SampleArgs()
        a.logits, a.ld, a.n_seq, a.V = b["logits"].data_ptr(), b["logits"].stride(0), n, m.vocab_size
        a.mode, a.temperature, a.top_p, a.top_k = self.mode, self.temperature, self.top_p, self.top_k
        a.seed, a.seq_base, a.step_base = self.seed, self.seq_base, 0
        a.state, a.targets, a.nwd, a.max_spans = self.state.data_ptr(), self.targets.data_ptr(), self.nwd.data_ptr(), self.max_spans
        a.raw_flags = a.raw_only_lo = a.raw_only_hi = None
        a.tok_buf, a.cur_len, a.span_start = self.tok_buf.data_ptr(), self.cur_len.data_ptr(), self.span_start.data_ptr()
        a.span_idx, a.fed_len, a.n_spans = self.span_idx.data_ptr(), self.fed_len.data_ptr(), self.n_spans.data_ptr()
        a.done, a.gen_count, a.control_bitmap = self.done.data_ptr(), self.gen_count.data_ptr(), self.bitmap.data_ptr()
        a.max_len, a.max_span = L, self.max_span
        a.out_token = a.out_probs = None
        a.trace_masked = self.trace_masked.data_ptr() if self.trace_masked is not None else None
        a.trace_span = self.trace_span.data_ptr() if self.trace_masked is not None else None
        K.check(K.lib().smer_sample_masked(C.byref(a), K.stream()), "sample_masked")

    def _step(self, step_idx_base: int):
        """One token for every unfinished piece.  All launches on the current stream; no sync."""
        if self.small:
            return self._step_small()
        m, b = self.m, self.buf
        d = m.d_model
        n, L, S = self.n, self.max_len, self.S
        lib = K.lib()
        K.check(lib.smer_decode_embed(self.tok_buf.data_ptr(), self.cur_len.data_ptr(), self.fed_len.data_ptr(),
                                      self.done.data_ptr(), self.pos.data_ptr(), m.embedding.weight.data_ptr(),
                                      m.pos_enc.pe.data_ptr(), b["x"].data_ptr(), K.dt(b["x"]), n, L, d, m.vocab_size,
                                      math.sqrt(d), K.stream()), "decode_embed")
        x = b["x"]
        launches = 1
        for i, lp in enumerate(self.layer_p):
            sa, ca = lp.sa, lp.ca
            qkv = b["qkv"]
            ops.gemm_nt(x, sa.w, qkv, bias=sa.b)
            kv = self.self_kv[i]
            self._decode_attn(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kv[:, :, :d], kv[:, :, d:], b["o"], self.pos,
                              None, 2 * d, L * 2 * d, L, 0)
            ops.gemm_nt(b["o"], sa.wo, b["proj"], bias=sa.bo)
            ops.layernorm_fwd(b["proj"], x, lp.ln[0][0], lp.ln[0][1], b["z"], b["y1"], None, None)
            ops.gemm_nt(b["y1"], ca.w[:d], b["q2"], bias=ca.b[:d])
            ckv = self.cross[i]
            self._decode_attn(b["q2"], None, None, ckv[:, :d], ckv[:, d:], b["o2"], self.src_len, None, 2 * d,
                              S * 2 * d, S, 0)
            ops.gemm_nt(b["o2"], ca.wo, b["proj"], bias=ca.bo)
            ops.layernorm_fwd(b["proj"], b["y1"], lp.ln[1][0], lp.ln[1][1], b["z"], b["y2"], None, None)
            ops.gemm_nt(b["y2"], lp.w1, b["h"], bias=lp.b1, flags=K.EPI_RELU)
            ops.gemm_nt(b["h"], lp.w2, b["f"], bias=lp.b2)
            ops.layernorm_fwd(b["f"], b["y2"], lp.ln[2][0], lp.ln[2][1], b["z"], b["y3"], None, None)
            x = b["y3"]
            launches += 11 + (2 if self.splits > 1 else 0)
        dn = m.transformer.decoder.norm
        ops.layernorm_fwd(x, None, dn.weight.detach(), dn.bias.detach(), None, b["yo"], None, None)
        ops.gemm_nt(b["yo"], self.fc_p[0], b["logits"], bias=self.fc_p[1])
        self._sample()
        self.launches_per_step = launches + 3

    def profile_step(self):
        """Runs ONE extra eager decode step (state is advanced by it) with CUDA events around the
        attention launches; returns algorithmic HBM bytes and milliseconds per kind."""
        self.profile = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pos = self.fed_len.clone()
        e0.record()
        self._step(0)
        e1.record()
        torch.cuda.synchronize()
        d, esz = self.m.d_model, (2 if self.m.compute_dtype == torch.bfloat16 else 4)
        nl = len(self.layer_p)
        cross_bytes = float(self.src_len.sum().item()) * 2 * d * esz          # per layer: K and V rows of every piece
        self_bytes = float((pos + 1).clamp(max=self.max_len).sum().item()) * 2 * d * esz
        out = {"step_ms": e0.elapsed_time(e1)}
        for kind, nbytes in (("cross", cross_bytes), ("self", self_bytes)):
            ms = sum(a.elapsed_time(b) for k, a, b in self.profile if k == kind)
            out[kind] = {"ms": ms, "bytes": nbytes * nl, "launches": sum(1 for k, _, _ in self.profile if k == kind)}
        self.profile = None
        return out

    # -- public -----------------------------------------------------------------------
    def infill(self, pieces, tracks_to_generate: Sequence[int], bars_to_generate: Sequence[int], **kw) -> Dict[str, object]:
        """Whole pieces in, completed pieces out -- the id-level equivalent of generation.generation_all
        (generation.py:468-702) for a batch: mask the selected (bar, track) spans (spans.mask_bar_and_track),
        decode all pieces together on the device, put the generated spans back (spans.restore_marked_input).
        Returns generate()'s dict plus `restored` (list of int64 arrays) and `src` (the masked inputs)."""
        srcs, targets, nwd = [], [], []
        for ids in pieces:
            src, _, _ = spans.mask_bar_and_track(ids, tracks_to_generate, bars_to_generate)
            srcs.append(src.tolist())
            targets.append(spans.mask_targets(ids, tracks_to_generate, bars_to_generate))
            nwd.append(spans.no_whole_duration(ids))
        kw.setdefault("nwd", nwd)
        res = self.generate(srcs, targets, **kw)
        res["src"] = srcs
        res["restored"] = [spans.restore_marked_input(s, g) for s, g in zip(srcs, res["streams"])]
        return res

    @torch.no_grad()
    def generate(self, pieces, targets, nwd: Optional[Sequence[bool]] = None, seq_base: int = 0, max_steps: int = 0,
                 check_every: int = 16) -> Dict[str, object]:
        if self.m.training:
            raise RuntimeError("InfillDecoder needs model.eval()")
        n = len(pieces)
        nwd = list(nwd) if nwd is not None else [False] * n
        self._setup(pieces, targets, nwd, seq_base)
        max_steps = max_steps or self.max_len
        steps = 0
        if not self.use_graph:
            self.graph = None
        elif self.graph is None or getattr(self, "_graph_steps", 0) != check_every:
            # (the graph is kept across generate() calls: same buffers, same weights -> same graph)
            # warm up once on a side stream (lazy module loads are not capturable), restore the state
            snap = [t.clone() for t in (self.tok_buf, self.cur_len, self.fed_len, self.span_start, self.span_idx,
                                        self.done, self.gen_count, self.state)]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._step(0)
            torch.cuda.current_stream().wait_stream(s)
            for t, c in zip((self.tok_buf, self.cur_len, self.fed_len, self.span_start, self.span_idx, self.done,
                             self.gen_count, self.state), snap):
                t.copy_(c)
            for kv in self.self_kv:
                pass                                   # row 0 is rewritten by the first real step
            # one graph = `check_every` decode steps (the steps between two done-checks): one launch per check
            # interval keeps the device fed even when the host is slow
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                for _ in range(check_every):
                    self._step(0)
            self._graph_steps = check_every
            for t, c in zip((self.tok_buf, self.cur_len, self.fed_len, self.span_start, self.span_idx, self.done,
                             self.gen_count, self.state), snap):
                t.copy_(c)
        trace = [] if getattr(self, "trace_intervals", False) else None     # diagnostics: device time per check interval
        live_hist, live = [], n - int(self.done.sum().item())
        self.check_every_used = check_every
        while steps < max_steps and live > 0:
            if trace is not None:
                trace.append(torch.cuda.Event(enable_timing=True))
                trace[-1].record()
            live_hist.append(live)
            if self.graph is not None:
                self.graph.replay()
                steps += check_every
            else:
                for _ in range(check_every):
                    self._step(steps)
                    steps += 1
            live = n - int(self.done.sum().item())     # one small D2H every `check_every` tokens
        self.interval_live = live_hist
        self.steps_run = steps
        self.kernel_launches = steps * self.launches_per_step
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        ev1.synchronize()
        self.device_ms = self._ev0.elapsed_time(ev1)              # encoder + cross K/V + decode loop on the device
        if trace is not None:
            trace.append(ev1)
            self.interval_ms = [a.elapsed_time(b) for a, b in zip(trace[:-1], trace[1:])]
        tok = self.tok_buf.cpu()
        lens = self.cur_len.cpu()
        gen = self.gen_count.cpu()
        self.d2h_bytes = tok.numel() * 8 + lens.numel() * 4 + gen.numel() * 4
        streams = [tok[i, : int(lens[i])].tolist() for i in range(n)]
        return {"streams": streams, "generated": gen.tolist(), "steps": steps, "done": self.done.cpu().tolist(),
                "device_ms": self.device_ms}
