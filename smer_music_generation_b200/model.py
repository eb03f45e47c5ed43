"""Host-side mirror of the reference's module API (model.py:59-106, transformer.py:16-470).

`ScoreTransformer` keeps the reference's constructor, `forward` signature, return tuple and
`state_dict` layout (128 entries, same names and shapes), so `train.py` / `generation.py`
run unchanged against it.  Nothing here computes with torch ops: the forward and backward
passes are sequences of calls into libsmer_b200.so (include/smer_b200.h) on the current CUDA
stream; torch only owns the buffers.  There is no CPU path -- calling the module without a
B200 raises.

Layout: activations are token-major ("batch-first", row = b*L + l), which only permutes the
reference's sequence-first storage.  Two arithmetic modes:
  * "bf16" (default): bf16 activations and weight shadows, fp32 accumulation, tcgen05 GEMMs.
  * "fp32": everything fp32 on CUDA cores -- the 1e-4 parity / greedy-decode path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _capi as K
from . import ops

VPAD_ALIGN = 64          # fc rows are padded to a multiple of this in the bf16 shadow (309 -> 320)

# dropout site ids (Philox counter high word): site = layer_site(kind, layer)
_SITE_EMB_SRC, _SITE_EMB_TGT = 1, 2
_KIND_ENC, _KIND_DEC = 1, 2
_SUB_ATTN_P, _SUB_DROP1, _SUB_FFN, _SUB_DROP2, _SUB_XATTN_P, _SUB_DROP3 = 0, 1, 2, 3, 4, 5


def _site(kind: int, layer: int, sub: int) -> int:
    return 16 + (kind * 64 + layer) * 8 + sub


# ----------------------------------------------------------------------------------------
# parameter containers: they only hold tensors under the reference's names
# ----------------------------------------------------------------------------------------
class _Linear(nn.Module):
    def __init__(self, fin: int, fout: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        self.bias = nn.Parameter(torch.empty(fout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(fin)
        nn.init.uniform_(self.bias, -bound, bound)


class _MultiheadAttention(nn.Module):
    """Parameter layout of nn.MultiheadAttention (packed Q|K|V rows), transformer.py:360,423."""

    def __init__(self, d: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _LayerNorm(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _EncoderLayer(nn.Module):
    def __init__(self, d: int, ff: int):
        super().__init__()
        self.self_attn = _MultiheadAttention(d)
        self.linear1 = _Linear(d, ff)
        self.linear2 = _Linear(ff, d)
        self.norm1 = _LayerNorm(d)
        self.norm2 = _LayerNorm(d)


class _DecoderLayer(nn.Module):
    def __init__(self, d: int, ff: int):
        super().__init__()
        self.self_attn = _MultiheadAttention(d)
        self.multihead_attn = _MultiheadAttention(d)
        self.linear1 = _Linear(d, ff)
        self.linear2 = _Linear(ff, d)
        self.norm1 = _LayerNorm(d)
        self.norm2 = _LayerNorm(d)
        self.norm3 = _LayerNorm(d)


class _Stack(nn.Module):
    def __init__(self, layers: List[nn.Module], d: int):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.norm = _LayerNorm(d)


class _Transformer(nn.Module):
    def __init__(self, d: int, ff: int, le: int, ld: int):
        super().__init__()
        self.encoder = _Stack([_EncoderLayer(d, ff) for _ in range(le)], d)
        self.decoder = _Stack([_DecoderLayer(d, ff) for _ in range(ld)], d)
        for p in self.parameters():            # transformer.py:137-142
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class _Embedding(nn.Module):
    def __init__(self, v: int, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(v, d))


class _PositionalEncoding(nn.Module):
    """model.py:110-125; the table is a persistent buffer `pe` of shape (max_len, 1, d)."""

    def __init__(self, d: int, max_len: int):
        super().__init__()
        pe = torch.zeros(max_len, d)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0).transpose(0, 1).contiguous())


# ----------------------------------------------------------------------------------------
# weight views in the compute dtype
# ----------------------------------------------------------------------------------------
class _Weights:
    """Name -> tensor the kernels read.  fp32 mode: the parameters themselves.  bf16 mode:
    bf16 shadows of the matrices (refreshed when a parameter's version changes); vectors
    (biases, LayerNorm affine) stay fp32.  `fc` is padded to `vpad` rows in bf16 mode so the
    vocabulary GEMM is 16-byte aligned (never visible in state_dict)."""

    def __init__(self, model: "ScoreTransformer"):
        self.model = model
        self._shadow: Dict[str, torch.Tensor] = {}
        self._key: Dict[str, Tuple[int, int]] = {}
        self.external: Dict[str, torch.Tensor] = {}      # shadows owned by trainer.ParamArena

    def mat(self, name: str, p: torch.Tensor, rows_pad: int = 0) -> torch.Tensor:
        if self.model.compute_dtype == torch.float32:
            return p.detach()
        if name in self.external:
            return self.external[name]
        key = (p.data_ptr(), p._version)
        sh = self._shadow.get(name)
        if sh is None or self._key.get(name) != key or sh.device != p.device:
            rows = max(p.shape[0], rows_pad)
            if sh is None or sh.shape[0] != rows or sh.device != p.device:
                sh = torch.zeros(rows, p.shape[1], dtype=torch.bfloat16, device=p.device)
                self._shadow[name] = sh
            ops.cast2d(p.detach(), sh[: p.shape[0]])
            self._key[name] = key
        return sh

    def vec(self, name: str, p: torch.Tensor, pad: int = 0) -> torch.Tensor:
        if pad <= p.shape[0]:
            return p.detach()
        if name in self.external:
            return self.external[name]
        key = (p.data_ptr(), p._version)
        sh = self._shadow.get(name)
        if sh is None or self._key.get(name) != key or sh.device != p.device:
            sh = torch.zeros(pad, dtype=torch.float32, device=p.device)
            sh[: p.shape[0]].copy_(p.detach())
            self._shadow[name] = sh
            self._key[name] = key
        return sh

    def invalidate(self):
        self._key.clear()


class _AttnP:
    __slots__ = ("w", "b", "wo", "bo", "name")


class _LayerP:
    __slots__ = ("sa", "ca", "w1", "b1", "w2", "b2", "ln", "name")


# ----------------------------------------------------------------------------------------
# the module
# ----------------------------------------------------------------------------------------
class ScoreTransformer(nn.Module):
    """Drop-in for the reference's model.ScoreTransformer (model.py:59-106).

    Extra keyword-only arguments (additions; the nine positional ones are the reference's):
      compute_dtype     "bf16" | "fp32"
      attention_weights "auto" | True | False -- the reference returns the head-averaged
          cross-attention probabilities (B, Ld, T, S) as a second value that no caller reads
          (SURVEY.md 0.3).  True computes them (a (B,Ld,T,S) fp32 tensor); False returns a
          zero-stride placeholder of that shape; "auto" = True in eval mode, False in training.
    `tgt_mask` may also be the string "causal": the nopeek mask (generation.py:193-206) is then
    never materialised; a tensor mask is inspected on the device and mapped to the causal flag
    when it is exactly that mask (any other float mask takes the additive-mask kernels).
    """

    def __init__(self, vocab_size, d_model, nhead, num_encoder_layers, num_decoder_layers, dim_feedforward,
                 max_seq_length, pos_dropout, trans_dropout, *, compute_dtype: str = "bf16",
                 attention_weights="auto"):
        super().__init__()
        if d_model % nhead:
            raise ValueError("d_model must be divisible by nhead")
        self.d_model = d_model
        self.nhead = nhead
        self.vocab_size = vocab_size
        self.dim_feedforward = dim_feedforward
        self.pos_dropout = float(pos_dropout)
        self.trans_dropout = float(trans_dropout)
        self.attention_weights = attention_weights
        self.compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[compute_dtype]
        self.embedding = _Embedding(vocab_size, d_model)
        self.pos_enc = _PositionalEncoding(d_model, max_seq_length)
        self.transformer = _Transformer(d_model, dim_feedforward, num_encoder_layers, num_decoder_layers)
        self.fc = _Linear(d_model, vocab_size)
        self._w = _Weights(self)
        self._calls = 0
        self._mask_cache: Optional[Tuple] = None
        self._decode_cache = None
        self.grad_hook = None            # set by parallel.DataParallelTrainer: called per finished bucket

    # -- helpers ----------------------------------------------------------------------
    @property
    def vpad(self) -> int:
        if self.compute_dtype == torch.float32:
            return self.vocab_size
        return (self.vocab_size + VPAD_ALIGN - 1) // VPAD_ALIGN * VPAD_ALIGN

    def set_compute_dtype(self, name: str) -> "ScoreTransformer":
        self.compute_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[name]
        self._w = _Weights(self)
        self._decode_cache = None
        return self

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._w = _Weights(self)
        self._decode_cache = None
        return r

    def _attn_p(self, m: _MultiheadAttention, name: str) -> _AttnP:
        a = _AttnP()
        a.name = name
        a.w = self._w.mat(name + "in_proj_weight", m.in_proj_weight)
        a.b = m.in_proj_bias.detach()
        a.wo = self._w.mat(name + "out_proj.weight", m.out_proj.weight)
        a.bo = m.out_proj.bias.detach()
        return a

    def _layer_p(self, layer, name: str) -> _LayerP:
        p = _LayerP()
        p.name = name
        p.sa = self._attn_p(layer.self_attn, name + "self_attn.")
        p.ca = self._attn_p(layer.multihead_attn, name + "multihead_attn.") if hasattr(layer, "multihead_attn") else None
        p.w1 = self._w.mat(name + "linear1.weight", layer.linear1.weight)
        p.b1 = layer.linear1.bias.detach()
        p.w2 = self._w.mat(name + "linear2.weight", layer.linear2.weight)
        p.b2 = layer.linear2.bias.detach()
        norms = [layer.norm1, layer.norm2] + ([layer.norm3] if hasattr(layer, "norm3") else [])
        p.ln = [(n.weight.detach(), n.bias.detach()) for n in norms]
        return p

    def _fc_p(self):
        return (self._w.mat("fc.weight", self.fc.weight, self.vpad), self._w.vec("fc.bias", self.fc.bias, self.vpad))

    def _classify_tgt_mask(self, tgt_mask, T: int, device):
        """-> (causal: bool, add_mask: Optional[(T,T) fp32 tensor])."""
        if tgt_mask is None:
            raise TypeError("tgt_mask is required (the reference indexes tgt_mask[0], model.py:95); "
                            "pass the nopeek mask or the string 'causal'")
        if isinstance(tgt_mask, str):
            if tgt_mask != "causal":
                raise ValueError("tgt_mask string must be 'causal'")
            return True, None
        m = tgt_mask[0] if tgt_mask.dim() == 3 else tgt_mask
        if m.shape != (T, T):
            raise RuntimeError(f"tgt_mask[0] must be ({T},{T}), got {tuple(m.shape)}")
        if m.dtype == torch.bool:
            m = torch.zeros(T, T, device=m.device).masked_fill_(m, float("-inf"))
        m = m.to(device=device, dtype=torch.float32)
        if m.stride(1) != 1:
            m = m.contiguous()
        key = (m.data_ptr(), m._version, T, m.stride(0))
        if self._mask_cache is not None and self._mask_cache[0] == key:
            kind = self._mask_cache[1]
        else:
            flags = torch.empty(3, dtype=torch.int32, device=device)
            ops.classify_mask(m, flags)
            bad, not_inf_above, nonzero_above = flags.tolist()       # one small D2H per new mask tensor
            kind = 2 if bad else (0 if not nonzero_above else (1 if not not_inf_above else 2))
            if T == 1:
                kind = 0 if not bad else 2
            self._mask_cache = (key, kind)
        if kind == 1:
            return True, None
        if kind == 0:
            return False, None
        return False, m

    @staticmethod
    def _pad_u8(mask, B, L, device):
        if mask is None:
            return None
        if mask.shape != (B, L):
            raise RuntimeError(f"key padding mask must be ({B},{L}), got {tuple(mask.shape)}")
        return mask.to(device=device).to(torch.uint8).contiguous()

    # -- forward ----------------------------------------------------------------------
    def forward(self, src, tgt, src_key_padding_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, tgt_mask=None):
        K.require_cuda_device()
        dev = self.embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("smer_b200.ScoreTransformer runs only on a CUDA (B200) device: call .to('cuda')")
        if src.dim() != 2 or tgt.dim() != 2 or src.shape[0] != tgt.shape[0]:
            raise RuntimeError("the batch number of src and tgt must be equal")      # transformer.py:117-118
        src = src.to(device=dev, dtype=torch.int64).contiguous()
        tgt = tgt.to(device=dev, dtype=torch.int64).contiguous()
        B, S = src.shape
        T = tgt.shape[1]
        if max(S, T) > self.pos_enc.pe.shape[0]:
            raise RuntimeError(f"sequence length {max(S, T)} exceeds max_seq_length {self.pos_enc.pe.shape[0]}")
        causal, add_mask = self._classify_tgt_mask(tgt_mask, T, dev)
        src_pad = self._pad_u8(src_key_padding_mask, B, S, dev)
        tgt_pad = self._pad_u8(tgt_key_padding_mask, B, T, dev)
        mem_pad = self._pad_u8(memory_key_padding_mask, B, S, dev)
        want_w = self.attention_weights
        if want_w == "auto":
            want_w = not self.training
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

        # incremental decode: batch-1, no pad masks, eval, causal (generation.py:209-219)
        if (not grad and not self.training and B == 1 and src_pad is None and tgt_pad is None and mem_pad is None
                and causal and add_mask is None and self.decode_cache_enabled):
            from .decode import cached_forward
            return cached_forward(self, src, tgt, bool(want_w))

        self._calls += 1
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._calls * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF
        run = _Run(self, src, tgt, src_pad, tgt_pad, mem_pad, causal, add_mask, self.training, seed, bool(want_w))
        if grad:
            params = [p for p in self.parameters()]
            logits = _StackFn.apply(run, *params)
        else:
            logits = run.forward(save=False)
        V = self.vocab_size
        out = logits.view(B, T, logits.shape[1])[:, :, :V]
        if want_w:
            weights = run.weights.permute(1, 0, 2, 3)           # (Ld,B,T,S) -> (B,Ld,T,S), model.py:101
        else:
            weights = torch.zeros((), device=dev).expand(B, len(self.transformer.decoder.layers), T, S)
        return out, weights

    decode_cache_enabled = True


# ----------------------------------------------------------------------------------------
# one forward(/backward) execution
# ----------------------------------------------------------------------------------------
class PackedBatch:
    """A batch without padding rows: the tokens of all sequences back to back, token-major, with int32
    prefix sums `cu_*` (B+1 entries, device) marking each sequence's rows.  The reference's collate functions pad every
    sequence of a bucket to the longest one (dataset.py:802-925); `pack()` removes those pads on the device from the padded
    batch and the per-sequence lengths, which the collate step knows on the host.  Row counts are rounded up to `align`
    (the extra "ghost" rows belong to no sequence, carry token id 0 and contribute nothing)."""

    def __init__(self, src_ids, tgt_in, tgt_out, pos_s, pos_t, cu_s, cu_t, n_s, n_t, max_s, max_t, B):
        self.src_ids, self.tgt_in, self.tgt_out = src_ids, tgt_in, tgt_out
        self.pos_s, self.pos_t, self.cu_s, self.cu_t = pos_s, pos_t, cu_s, cu_t
        self.n_s, self.n_t, self.max_s, self.max_t, self.B = n_s, n_t, max_s, max_t, B
        self.rows_s, self.rows_t = src_ids.numel(), tgt_in.numel()

    @staticmethod
    def pack(src, tgt_in, tgt_out, src_lens, tgt_lens, align: int = 128, rows_s: int = 0, rows_t: int = 0, out=None):
        """src / tgt_in / tgt_out: padded (B, S) / (B, T) int64 DEVICE tensors; src_lens / tgt_lens: HOST sequences of
        the un-padded lengths.  rows_s / rows_t force the packed row counts (fixed shapes for a captured step);
        out: a PackedBatch whose buffers are refilled in place."""
        dev = src.device
        B = src.shape[0]
        ls = [int(x) for x in src_lens]
        lt = [int(x) for x in tgt_lens]
        n_s, n_t = sum(ls), sum(lt)
        up = lambda n: (n + align - 1) // align * align
        rows_s, rows_t = max(rows_s, up(n_s)), max(rows_t, up(n_t))
        cu = torch.zeros(2, B + 1, dtype=torch.int32)
        cu[0, 1:] = torch.tensor(ls, dtype=torch.int32).cumsum(0)
        cu[1, 1:] = torch.tensor(lt, dtype=torch.int32).cumsum(0)
        if out is None:
            cu_d = cu.to(dev)
            mk = lambda n, dt: torch.empty(n, dtype=dt, device=dev)
            out = PackedBatch(mk(rows_s, torch.int64), mk(rows_t, torch.int64), mk(rows_t, torch.int64), mk(rows_s, torch.int32),
                              mk(rows_t, torch.int32), cu_d[0], cu_d[1], n_s, n_t, max(ls), max(lt), B)
            out._cu2 = cu_d
        else:
            if out.rows_s < up(n_s) or out.rows_t < up(n_t) or out.B != B:
                raise RuntimeError("PackedBatch.pack: the batch does not fit the preallocated rows")
            out._cu2.copy_(cu, non_blocking=True)
            out.n_s, out.n_t, out.max_s, out.max_t = n_s, n_t, max(ls), max(lt)
        ops.pack_rows(src, out.cu_s, out.rows_s, out.src_ids, out.pos_s)
        ops.pack_rows(tgt_in, out.cu_t, out.rows_t, out.tgt_in, out.pos_t)
        ops.pack_rows(tgt_out, out.cu_t, out.rows_t, out.tgt_out, out.pos_t)
        return out


class _Run:
    def __init__(self, model: ScoreTransformer, src, tgt, src_pad, tgt_pad, mem_pad, causal, add_mask, training,
                 seed, want_w, packed: Optional[PackedBatch] = None):
        self.m = model
        self.pk = packed
        if packed is not None:
            # packed rows: (B, S, T) = (sequences, longest source, longest target); pads do not exist
            src, tgt = packed.src_ids.view(1, -1), packed.tgt_in.view(1, -1)
            src_pad = tgt_pad = mem_pad = None
        self.src, self.tgt = src, tgt
        self.src_pad, self.tgt_pad, self.mem_pad = src_pad, tgt_pad, mem_pad
        self.causal, self.add_mask = causal, add_mask
        self.training, self.seed, self.want_w = training, seed, want_w
        self.dt = model.compute_dtype
        self.B, self.S = src.shape
        self.T = tgt.shape[1]
        self.rows_s, self.rows_t = self.B * self.S, self.B * self.T
        if packed is not None:
            if model.compute_dtype != torch.bfloat16 or want_w or not causal or add_mask is not None:
                raise RuntimeError("packed batches run on the bf16 tcgen05 path with the nopeek mask and without attention weights")
            self.B, self.S, self.T = packed.B, packed.max_s, packed.max_t
        self.d, self.H, self.ff = model.d_model, model.nhead, model.dim_feedforward
        self.dh = self.d // self.H
        self.pd = model.pos_dropout if training else 0.0
        self.td = model.trans_dropout if training else 0.0
        self.dev = src.device
        self.weights = None
        self.tape: Dict[str, object] = {}
        self.src_len = self._kv_len(src_pad)
        self.tgt_len = self._kv_len(tgt_pad)
        self.mem_len = self._kv_len(mem_pad)

    def _kv_len(self, pad):
        if pad is None:
            return None
        out = torch.empty(pad.shape[0], dtype=torch.int32, device=self.dev)
        ops.kv_len_from_pad(pad, out)
        return out

    def new(self, rows, cols, dtype=None):
        return torch.empty(rows, cols, dtype=dtype or self.dt, device=self.dev)

    # ---- building blocks ------------------------------------------------------------
    def _ln_fwd(self, branch, resid, ln, site, save_key, save):
        rows = branch.shape[0]
        y = self.new(rows, self.d)
        z = self.new(rows, self.d) if (resid is not None and save) else None
        mean = torch.empty(rows, dtype=torch.float32, device=self.dev) if save else None
        rstd = torch.empty(rows, dtype=torch.float32, device=self.dev) if save else None
        p = self.td if resid is not None else 0.0
        ops.layernorm_fwd(branch, resid, ln[0], ln[1], z, y, mean, rstd, dropout_p=p, seed=self.seed, site=site)
        if save:
            self.tape[save_key] = (z if z is not None else branch, mean, rstd)
        return y

    def _ln_bwd(self, dy, ln, site, save_key, grads, gname, has_resid=True, bias_name=None):
        z, mean, rstd = self.tape.pop(save_key)
        rows = dy.shape[0]
        dz = self.new(rows, self.d)
        p = self.td if has_resid else 0.0
        dbr = self.new(rows, self.d) if p > 0.0 else None
        ops.layernorm_bwd(dy, z, mean, rstd, ln[0], dz, dbr, grads[gname + "weight"], grads[gname + "bias"],
                          dropout_p=p, seed=self.seed, site=site,
                          dbias=grads[bias_name] if bias_name is not None else None)   # bias grad of the branch's Linear
        return dz, (dbr if dbr is not None else dz)

    def _attn_fwd(self, ap: _AttnP, xq, xkv, Lq, Lk, causal, key_pad, kv_len, add_mask, site_p, key, save, weights_out=None,
                  side="s"):
        B, d, H, dh = self.B, self.d, self.H, self.dh
        self_attn = xkv is None
        rq = xq.shape[0]
        rk = rq if self_attn else xkv.shape[0]
        pk = self.pk
        cu_q = cu_k = None
        if pk is not None:
            cu_q = pk.cu_s if side == "s" else pk.cu_t
            cu_k = cu_q if self_attn else pk.cu_s
        if self_attn:
            qkv = self.new(rq, 3 * d)
            ops.gemm_nt(xq, ap.w, qkv, bias=ap.b)
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
            kvbuf = None
        else:
            qkv = self.new(rq, d)
            ops.gemm_nt(xq, ap.w[:d], qkv, bias=ap.b[:d])
            kvbuf = self.new(rk, 2 * d)
            ops.gemm_nt(xkv, ap.w[d:3 * d], kvbuf, bias=ap.b[d:])
            q, k, v = qkv, kvbuf[:, :d], kvbuf[:, d:]
        o = self.new(rq, d)
        lse = torch.empty(H * rq if pk is not None else B * H * Lq, dtype=torch.float32, device=self.dev)
        a = ops.attn_args(q, k, v, o, B, H, Lq, Lk, dh, lse=lse, causal=causal, key_pad=key_pad, kv_len=kv_len,
                          add_mask=add_mask, dropout_p=self.td, seed=self.seed, site=site_p, cu_q=cu_q, cu_k=cu_k)
        ops.attn_fwd(a)
        if weights_out is not None:
            ops.attn_weights(a, weights_out)
        proj = self.new(rq, d)
        ops.gemm_nt(o, ap.wo, proj, bias=ap.bo)
        if save:
            self.tape[key] = (qkv, kvbuf, o, lse)
        return proj

    def _attn_bwd(self, ap: _AttnP, dproj, xq, xkv, Lq, Lk, causal, key_pad, kv_len, add_mask, site_p, key, grads,
                  resid_q, dmem=None, side="s"):
        """Returns grad wrt xq (residual `resid_q` folded in).  For cross-attention the K/V-side
        input gradient is accumulated into `dmem`."""
        B, d, H, dh = self.B, self.d, self.H, self.dh
        qkv, kvbuf, o, lse = self.tape.pop(key)
        n = ap.name
        rq = dproj.shape[0]
        pk = self.pk
        cu_q = cu_k = None
        ops.gemm_dw(dproj, o, grads[n + "out_proj.weight"])         # out_proj.bias: summed inside layernorm_bwd
        do = self.new(rq, d)
        ops.gemm_dx(dproj, ap.wo, do)
        dsum = torch.empty(H * rq if pk is not None else B * H * Lq, dtype=torch.float32, device=self.dev)
        self_attn = kvbuf is None
        if pk is not None:
            cu_q = pk.cu_s if side == "s" else pk.cu_t
            cu_k = cu_q if self_attn else pk.cu_s
        if self_attn:
            dqkv = self.new(rq, 3 * d)
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
            dq, dk, dv = dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:]
        else:
            dqkv = self.new(rq, d)
            dkv = self.new(kvbuf.shape[0], 2 * d)
            q, k, v = qkv, kvbuf[:, :d], kvbuf[:, d:]
            dq, dk, dv = dqkv, dkv[:, :d], dkv[:, d:]
        gw, gb = grads[n + "in_proj_weight"], grads[n + "in_proj_bias"]
        # in_proj_bias gradient = column sums of dq | dk | dv, accumulated by the attention backward kernels
        a = ops.attn_args(q, k, v, o, B, H, Lq, Lk, dh, lse=lse, causal=causal, key_pad=key_pad, kv_len=kv_len,
                          add_mask=add_mask, dropout_p=self.td, seed=self.seed, site=site_p, dout=do, dq=dq, dk=dk,
                          dv=dv, dsum=dsum, dbq=gb[:d], dbk=gb[d:2 * d], dbv=gb[2 * d:], cu_q=cu_q, cu_k=cu_k)
        ops.attn_bwd(a)
        dx = self.new(rq, d)
        if self_attn:
            ops.gemm_dw(dqkv, xq, gw)
            ops.gemm_dx(dqkv, ap.w, dx, resid=resid_q)
        else:
            ops.gemm_dw(dqkv, xq, gw[:d])
            ops.gemm_dx(dqkv, ap.w[:d], dx, resid=resid_q)
            ops.gemm_dw(dkv, xkv, gw[d:])
            ops.gemm_dx(dkv, ap.w[d:3 * d], dmem, resid=dmem)      # dmem += dkv . W_kv  (in place)
        return dx

    def _ffn_fwd(self, lp: _LayerP, x, site, key, save):
        rows = x.shape[0]
        h = self.new(rows, self.ff)
        ops.gemm_nt(x, lp.w1, h, bias=lp.b1, flags=K.EPI_RELU, dropout_p=self.td, seed=self.seed, site=site)
        f = self.new(rows, self.d)
        ops.gemm_nt(h, lp.w2, f, bias=lp.b2)
        if save:
            self.tape[key] = h
        return f

    def _ffn_bwd(self, lp: _LayerP, df, x, key, grads, resid):
        h = self.tape.pop(key)
        n = lp.name
        rows = df.shape[0]
        ops.gemm_dw(df, h, grads[n + "linear2.weight"])             # linear2.bias: summed inside layernorm_bwd
        dh = self.new(rows, self.ff)
        ops.gemm_dx(df, lp.w2, dh, resid=h, flags=K.EPI_GATE, dropout_p=self.td,
                    colsum_out=grads[n + "linear1.bias"])            # linear1.bias: summed in the gate epilogue
        ops.gemm_dw(dh, x, grads[n + "linear1.weight"])
        dx = self.new(rows, self.d)
        ops.gemm_dx(dh, lp.w1, dx, resid=resid)
        return dx

    # ---- whole passes ---------------------------------------------------------------
    def forward(self, save: bool) -> torch.Tensor:
        return self.decode(self.encode(save), save)

    def encode(self, save: bool) -> torch.Tensor:
        """Encoder stack (transformer.py:258-277) -> memory (B*S, d)."""
        m = self.m
        B, S, d = self.B, self.S, self.d
        pe = m.pos_enc.pe.view(-1, d)
        emb = m.embedding.weight.detach()
        scale = math.sqrt(d)
        x = self.new(self.rows_s, d)
        if self.pk is not None:
            ops.embed_pe_packed(self.pk.src_ids, self.pk.pos_s, emb, pe, x, scale, self.pd, self.seed, _SITE_EMB_SRC)
        else:
            ops.embed_pe(self.src, emb, pe, x, scale, 0, self.pd, self.seed, _SITE_EMB_SRC)
        enc = m.transformer.encoder
        self.enc_p = [m._layer_p(l, f"transformer.encoder.layers.{i}.") for i, l in enumerate(enc.layers)]
        for i, lp in enumerate(self.enc_p):
            if save:
                self.tape[f"e{i}.x"] = x
            a = self._attn_fwd(lp.sa, x, None, S, S, False, self.src_pad, self.src_len, None,
                               _site(_KIND_ENC, i, _SUB_ATTN_P), f"e{i}.sa", save)
            x1 = self._ln_fwd(a, x, lp.ln[0], _site(_KIND_ENC, i, _SUB_DROP1), f"e{i}.ln1", save)
            if save:
                self.tape[f"e{i}.x1"] = x1
            f = self._ffn_fwd(lp, x1, _site(_KIND_ENC, i, _SUB_FFN), f"e{i}.ffn", save)
            x = self._ln_fwd(f, x1, lp.ln[1], _site(_KIND_ENC, i, _SUB_DROP2), f"e{i}.ln2", save)
        self.enc_norm = (enc.norm.weight.detach(), enc.norm.bias.detach())
        mem = self._ln_fwd(x, None, self.enc_norm, 0, "enc.norm", save)
        if save:
            self.tape["mem"] = mem
        return mem

    def decode(self, mem: torch.Tensor, save: bool) -> torch.Tensor:
        """Decoder stack + fc (transformer.py:303-335, model.py:106) -> logits (B*T, vpad) fp32."""
        m = self.m
        B, S, T, d = self.B, self.S, self.T, self.d
        pe = m.pos_enc.pe.view(-1, d)
        emb = m.embedding.weight.detach()
        scale = math.sqrt(d)
        y = self.new(self.rows_t, d)
        if self.pk is not None:
            ops.embed_pe_packed(self.pk.tgt_in, self.pk.pos_t, emb, pe, y, scale, self.pd, self.seed, _SITE_EMB_TGT)
        else:
            ops.embed_pe(self.tgt, emb, pe, y, scale, 0, self.pd, self.seed, _SITE_EMB_TGT)
        dec = m.transformer.decoder
        self.dec_p = [m._layer_p(l, f"transformer.decoder.layers.{i}.") for i, l in enumerate(dec.layers)]
        if self.want_w:
            self.weights = torch.empty(len(self.dec_p), B, T, S, dtype=torch.float32, device=self.dev)
        for i, lp in enumerate(self.dec_p):
            if save:
                self.tape[f"d{i}.y"] = y
            a = self._attn_fwd(lp.sa, y, None, T, T, self.causal, self.tgt_pad, self.tgt_len, self.add_mask,
                               _site(_KIND_DEC, i, _SUB_ATTN_P), f"d{i}.sa", save, side="t")
            y1 = self._ln_fwd(a, y, lp.ln[0], _site(_KIND_DEC, i, _SUB_DROP1), f"d{i}.ln1", save)
            if save:
                self.tape[f"d{i}.y1"] = y1
            c = self._attn_fwd(lp.ca, y1, mem, T, S, False, self.mem_pad, self.mem_len, None,
                               _site(_KIND_DEC, i, _SUB_XATTN_P), f"d{i}.ca", save,
                               weights_out=self.weights[i] if self.want_w else None, side="t")
            y2 = self._ln_fwd(c, y1, lp.ln[1], _site(_KIND_DEC, i, _SUB_DROP2), f"d{i}.ln2", save)
            if save:
                self.tape[f"d{i}.y2"] = y2
            f = self._ffn_fwd(lp, y2, _site(_KIND_DEC, i, _SUB_FFN), f"d{i}.ffn", save)
            y = self._ln_fwd(f, y2, lp.ln[2], _site(_KIND_DEC, i, _SUB_DROP3), f"d{i}.ln3", save)
        self.dec_norm = (dec.norm.weight.detach(), dec.norm.bias.detach())
        yo = self._ln_fwd(y, None, self.dec_norm, 0, "dec.norm", save)
        if save:
            self.tape["yo"] = yo
        wfc, bfc = m._fc_p()
        self.fc_p = (wfc, bfc)
        logits = torch.empty(self.rows_t, wfc.shape[0], dtype=torch.float32, device=self.dev)
        ops.gemm_nt(yo, wfc, logits, bias=bfc)
        return logits

    def backward(self, dlogits: torch.Tensor, grads: Dict[str, torch.Tensor]) -> None:
        """dlogits: (B*T, vpad) in the compute dtype.  Accumulates into `grads` (fp32, zeroed)."""
        m = self.m
        B, S, T, d = self.B, self.S, self.T, self.d
        hook = m.grad_hook
        wfc, _ = self.fc_p
        yo = self.tape.pop("yo")
        ops.colsum(dlogits, grads["fc.bias"])
        ops.gemm_dw(dlogits, yo, grads["fc.weight"])
        dyo = self.new(self.rows_t, d)
        ops.gemm_dx(dlogits, wfc, dyo)
        if hook:
            hook("fc.")
        dy, _ = self._ln_bwd(dyo, self.dec_norm, 0, "dec.norm", grads, "transformer.decoder.norm.", has_resid=False)
        if hook:
            hook("transformer.decoder.norm.")
        mem = self.tape.pop("mem")
        dmem = torch.zeros(self.rows_s, d, dtype=self.dt, device=self.dev)
        for i in reversed(range(len(self.dec_p))):
            lp = self.dec_p[i]
            n = lp.name
            dz3, df = self._ln_bwd(dy, lp.ln[2], _site(_KIND_DEC, i, _SUB_DROP3), f"d{i}.ln3", grads, n + "norm3.",
                                   bias_name=n + "linear2.bias")
            dy2 = self._ffn_bwd(lp, df, self.tape.pop(f"d{i}.y2"), f"d{i}.ffn", grads, dz3)
            dz2, dc = self._ln_bwd(dy2, lp.ln[1], _site(_KIND_DEC, i, _SUB_DROP2), f"d{i}.ln2", grads, n + "norm2.",
                                   bias_name=n + "multihead_attn.out_proj.bias")
            dy1 = self._attn_bwd(lp.ca, dc, self.tape.pop(f"d{i}.y1"), mem, T, S, False, self.mem_pad, self.mem_len,
                                 None, _site(_KIND_DEC, i, _SUB_XATTN_P), f"d{i}.ca", grads, dz2, dmem, side="t")
            dz1, da = self._ln_bwd(dy1, lp.ln[0], _site(_KIND_DEC, i, _SUB_DROP1), f"d{i}.ln1", grads, n + "norm1.",
                                   bias_name=n + "self_attn.out_proj.bias")
            dy = self._attn_bwd(lp.sa, da, self.tape.pop(f"d{i}.y"), None, T, T, self.causal, self.tgt_pad,
                                self.tgt_len, self.add_mask, _site(_KIND_DEC, i, _SUB_ATTN_P), f"d{i}.sa", grads, dz1, side="t")
            if hook:
                hook(n)
        emb_scale = math.sqrt(d)
        ops.embed_bwd(self.tgt, dy, grads["embedding.weight"], emb_scale, self.pd, self.seed, _SITE_EMB_TGT)
        dx, _ = self._ln_bwd(dmem, self.enc_norm, 0, "enc.norm", grads, "transformer.encoder.norm.", has_resid=False)
        if hook:
            hook("transformer.encoder.norm.")
        for i in reversed(range(len(self.enc_p))):
            lp = self.enc_p[i]
            n = lp.name
            dz2, df = self._ln_bwd(dx, lp.ln[1], _site(_KIND_ENC, i, _SUB_DROP2), f"e{i}.ln2", grads, n + "norm2.",
                                   bias_name=n + "linear2.bias")
            dx1 = self._ffn_bwd(lp, df, self.tape.pop(f"e{i}.x1"), f"e{i}.ffn", grads, dz2)
            dz1, da = self._ln_bwd(dx1, lp.ln[0], _site(_KIND_ENC, i, _SUB_DROP1), f"e{i}.ln1", grads, n + "norm1.",
                                   bias_name=n + "self_attn.out_proj.bias")
            dx = self._attn_bwd(lp.sa, da, self.tape.pop(f"e{i}.x"), None, S, S, False, self.src_pad, self.src_len,
                                None, _site(_KIND_ENC, i, _SUB_ATTN_P), f"e{i}.sa", grads, dz1)
            if hook:
                hook(n)
        ops.embed_bwd(self.src, dx, grads["embedding.weight"], emb_scale, self.pd, self.seed, _SITE_EMB_SRC)
        if hook:
            hook("embedding.")
        self.tape.clear()


class GradArena:
    """Flat fp32 gradient buffer with one 256-byte aligned slot per parameter, in reverse
    execution order (fc, decoder layers top-down, encoder layers top-down, embedding) so that
    contiguous ranges finish together during backward (the NCCL buckets of parallel.py)."""

    def __init__(self, model: ScoreTransformer):
        names = [n for n, _ in model.named_parameters()]
        params = dict(model.named_parameters())
        order: List[str] = []
        order += [n for n in names if n.startswith("fc.")]
        order += [n for n in names if n.startswith("transformer.decoder.norm.")]
        nd = len(model.transformer.decoder.layers)
        for i in reversed(range(nd)):
            order += [n for n in names if n.startswith(f"transformer.decoder.layers.{i}.")]
        order += [n for n in names if n.startswith("transformer.encoder.norm.")]
        ne = len(model.transformer.encoder.layers)
        for i in reversed(range(ne)):
            order += [n for n in names if n.startswith(f"transformer.encoder.layers.{i}.")]
        order += [n for n in names if n.startswith("embedding.")]
        assert sorted(order) == sorted(names)
        self.order = order
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        vpad = model.vpad
        for n in order:
            p = params[n]
            numel = p.numel()
            if n == "fc.weight":
                numel = vpad * p.shape[1]
            elif n == "fc.bias":
                numel = vpad
            self.offsets[n] = (off, numel)
            off += (numel + 63) // 64 * 64
        self.total = off
        dev = model.embedding.weight.device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.views: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}           # exact-shape views handed to autograd
        for n in order:
            p = params[n]
            o, numel = self.offsets[n]
            if n == "fc.weight":
                full = self.flat[o:o + numel].view(vpad, p.shape[1])
                self.views[n] = full
                self.grads[n] = full[: p.shape[0]]
            elif n == "fc.bias":
                full = self.flat[o:o + numel]
                self.views[n] = full
                self.grads[n] = full[: p.shape[0]]
            else:
                v = self.flat[o:o + numel].view(p.shape)
                self.views[n] = v
                self.grads[n] = v

    def span(self, prefix: str) -> Tuple[int, int]:
        """[begin, end) element range of the parameters whose name starts with `prefix`."""
        sel = [self.offsets[n] for n in self.order if n.startswith(prefix)]
        b = min(o for o, _ in sel)
        e = max((o + (k + 63) // 64 * 64) for o, k in sel)
        return b, e


class _StackFn(torch.autograd.Function):
    """Autograd boundary: logits = f(parameters); backward runs the hand-written backward pass
    and hands each parameter its gradient (ordinary `.grad` tensors, as wandb.watch / Adam in
    train.py:264,661 expect)."""

    @staticmethod
    def forward(ctx, run: _Run, *params):
        ctx.run = run
        logits = run.forward(save=True)
        ctx.names = [n for n, _ in run.m.named_parameters()]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        run: _Run = ctx.run
        m = run.m
        arena = getattr(m, "_grad_arena", None)
        if arena is None or arena.flat.device != run.dev or arena.vpad != m.vpad:
            arena = GradArena(m)
            arena.vpad = m.vpad
            m._grad_arena = arena
        arena.flat.zero_()
        vp = m.vpad
        if dlogits.dtype == run.dt and dlogits.is_contiguous() and dlogits.shape[1] == vp:
            dl = dlogits
        else:
            dl = torch.empty(dlogits.shape[0], vp, dtype=run.dt, device=run.dev)
            src = dlogits if dlogits.stride(1) == 1 else dlogits.contiguous()
            ops.cast2d(src, dl, cols=min(src.shape[1], m.vocab_size))
        run.backward(dl, arena.views)
        out = [arena.grads[n] for n in ctx.names]
        ctx.run = None
        return (None, *out)
