"""Tensor-level wrappers over the C ABI (one function per kernel family).  All tensors are CUDA
tensors owned by PyTorch's caching allocator; kernels are enqueued on the current stream."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch

from . import _capi as K

_TC_GEMM = os.environ.get("SMER_GEMM", "tc")     # "simt" forces the CUDA-core GEMM (debug aid)
_TC_ATTN = os.environ.get("SMER_ATTN", "tc")
_NUM_SMS = None

# ---- optional per-kernel-family timing (bench.py): events on the launching stream ------------
PROFILE = None           # None or dict label -> list[(start_event, end_event, work)]
LAUNCHES = 0             # kernels launched through this module (counted, not estimated)


class _Timed:
    __slots__ = ("label", "work", "n", "s")

    def __init__(self, label, work=0.0, n=1):
        self.label, self.work, self.n = label, work, n

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += self.n
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            PROFILE.setdefault(self.label, []).append((self.s, e, self.work))
        return False


def num_sms() -> int:
    global _NUM_SMS
    if _NUM_SMS is None:
        _NUM_SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _NUM_SMS


def _p(t):
    return 0 if t is None else t.data_ptr()


def embed_pe(ids, emb, pe, out, scale, pos0=0, dropout_p=0.0, seed=0, site=0):
    with _Timed("embed", float(ids.numel() * emb.shape[1] * (4 + out.element_size())), 1):
        B, L = ids.shape
        V, d = emb.shape
        K.check(K.lib().smer_embed_pe_fwd(_p(ids), _p(emb), _p(pe), _p(out), K.dt(out), B, L, d, V, pos0, scale,
                                          dropout_p, seed, site, K.stream()), "embed_pe_fwd")


def embed_pe_packed(ids, pos, emb, pe, out, scale, dropout_p=0.0, seed=0, site=0):
    with _Timed("embed", float(ids.numel() * emb.shape[1] * (4 + out.element_size())), 1):
        V, d = emb.shape
        K.check(K.lib().smer_embed_pe_packed(_p(ids), _p(pos), _p(emb), _p(pe), _p(out), K.dt(out), ids.numel(), d, V, scale,
                                             dropout_p, seed, site, K.stream()), "embed_pe_packed")


def pack_rows(ids_padded, cu, rows_alloc, out_ids, out_pos):
    with _Timed("misc", 0.0, 1):
        B, L = ids_padded.shape
        K.check(K.lib().smer_pack_rows(_p(ids_padded), _p(cu), B, L, rows_alloc, _p(out_ids), _p(out_pos), K.stream()), "pack_rows")


def zero_tail_rows(buf, cu):
    """Zeroes the rows of the contiguous 2-D `buf` from cu[-1] (device value) on: the ghost rows of a packed batch."""
    with _Timed("misc", 0.0, 1):
        assert buf.is_contiguous()
        K.check(K.lib().smer_zero_tail_rows(_p(buf), buf.shape[1] * buf.element_size(), buf.shape[0], cu[-1:].data_ptr(),
                                            K.stream()), "zero_tail_rows")


def embed_bwd(ids, dout, demb, scale, dropout_p=0.0, seed=0, site=0):
    with _Timed("embed_bwd", float(ids.numel() * demb.shape[1] * (4 + dout.element_size())), 1):
        B, L = ids.shape
        V, d = demb.shape
        K.check(K.lib().smer_embed_bwd(_p(ids), _p(dout), K.dt(dout), _p(demb), B, L, d, V, scale, dropout_p, seed,
                                       site, K.stream()), "embed_bwd")


def layernorm_fwd(branch, resid, gamma, beta, z_out, y, mean, rstd, eps=1e-5, dropout_p=0.0, seed=0, site=0):
    with _Timed("layernorm_fwd", float(branch.numel() * branch.element_size() * (2 + (resid is not None) + (z_out is not None))), 1):
        rows, d = branch.shape
        K.check(K.lib().smer_layernorm_fwd(_p(branch), _p(resid), _p(gamma), _p(beta), _p(z_out), _p(y), _p(mean),
                                           _p(rstd), K.dt(branch), rows, d, eps, dropout_p, seed, site, K.stream()),
                "layernorm_fwd")


def layernorm_bwd(dy, z, mean, rstd, gamma, dz, dbranch, dgamma, dbeta, dropout_p=0.0, seed=0, site=0, dbias=None):
    with _Timed("layernorm_bwd", float(dy.numel() * dy.element_size() * (3 + (dbranch is not None))), 1):
        rows, d = dy.shape
        K.check(K.lib().smer_layernorm_bwd(_p(dy), _p(z), _p(mean), _p(rstd), _p(gamma), _p(dz), _p(dbranch),
                                           _p(dgamma), _p(dbeta), _p(dbias), K.dt(dy), rows, d, dropout_p, seed, site, K.stream()),
                "layernorm_bwd")


def _tc_ok(a_dtype, N, ldc, *pitches):
    return (_TC_GEMM == "tc" and a_dtype == torch.bfloat16 and N % 8 == 0 and ldc % 8 == 0
            and all(p % 8 == 0 for p in pitches))


def gemm_nt(A, W, out, bias=None, resid=None, flags=0, dropout_p=0.0, seed=0, site=0):
    """out[M,N] = epi(A[M,K] @ W[N,K]^T) -- nn.Linear forward.  A/out may be strided row views."""
    with _Timed("gemm", 2.0 * A.shape[0] * A.shape[1] * W.shape[0], 1):
        M, Kd = A.shape
        N = W.shape[0]
        lda, ldw, ldc = A.stride(0), W.stride(0), out.stride(0)
        ldr = resid.stride(0) if resid is not None else 0
        if _tc_ok(A.dtype, N, ldc, lda, ldw, ldr):
            K.check(K.lib().smer_gemm_bf16_tc(_p(A), lda, 1, _p(W), ldw, 1, _p(out), ldc, K.dt(out), M, N, Kd, _p(bias),
                                              _p(resid), ldr, flags, dropout_p, seed, site, 1, 0, K.stream()), "gemm_tc(nt)")
        else:
            K.check(K.lib().smer_gemm_simt(_p(A), lda, 1, _p(W), ldw, 1, _p(out), ldc, K.dt(A), K.dt(out), M, N, Kd,
                                           _p(bias), _p(resid), ldr, flags, dropout_p, seed, site, 1, K.stream()),
                    "gemm_simt(nt)")


def gemm_dx(dY, W, out, resid=None, flags=0, dropout_p=0.0, colsum_out=None):
    """out[M,Kin] = epi(dY[M,N] @ W[N,Kin]) -- input gradient of nn.Linear.  `colsum_out` (fp32 [Kin], gate
    epilogue): += column sums of `out`, i.e. the bias gradient of the Linear that produced the gated activation."""
    fused = (colsum_out is not None and (flags & K.EPI_GATE) and out.dtype == torch.bfloat16
             and _tc_ok(dY.dtype, W.shape[1], out.stride(0), dY.stride(0), W.stride(0), resid.stride(0) if resid is not None else 0))
    with _Timed("gemm", 2.0 * dY.shape[0] * dY.shape[1] * W.shape[1], 1):
        M, N = dY.shape
        Kin = W.shape[1]
        ldy, ldw, ldc = dY.stride(0), W.stride(0), out.stride(0)
        ldr = resid.stride(0) if resid is not None else 0
        if _tc_ok(dY.dtype, Kin, ldc, ldy, ldw, ldr):
            K.check(K.lib().smer_gemm_bf16_tc(_p(dY), ldy, 1, _p(W), ldw, 0, _p(out), ldc, K.dt(out), M, Kin, N, 0,
                                              _p(resid), ldr, flags, dropout_p, 0, 0, 1, _p(colsum_out) if fused else 0,
                                              K.stream()), "gemm_tc(dx)")
        else:
            K.check(K.lib().smer_gemm_simt(_p(dY), ldy, 1, _p(W), 1, ldw, _p(out), ldc, K.dt(dY), K.dt(out), M, Kin, N,
                                           0, _p(resid), ldr, flags, dropout_p, 0, 0, 1, K.stream()), "gemm_simt(dx)")
    if colsum_out is not None and not fused:
        colsum(out, colsum_out)


def _dw_split(N, Kin, M):
    """K-split of the weight-gradient GEMM (K = M tokens).  With 256x256 CTA-pair tiles: the split count whose
    item count fills whole rounds of SM pairs (>= 16 k-blocks per item, fewest splits among the best)."""
    kb = (M + 63) // 64
    pairs = max(1, num_sms() // 2)
    if N >= 256 and Kin >= 256:
        t = ((N + 255) // 256) * ((Kin + 255) // 256)
        best = None
        for s in range(1, max(1, kb // 16) + 1):
            items = t * s
            if items * 20 < pairs * 19:
                continue
            eff = items / (-(-items // pairs) * pairs)
            if best is None or eff > best[0]:
                best = (eff, s)
            if eff >= 0.95 or s * t > 6 * pairs:
                break
        if best is not None:
            return best[1]
    tiles = ((N + 127) // 128) * ((Kin + 127) // 128)
    return max(1, min(kb, (2 * num_sms()) // tiles))


def gemm_dw(dY, X, out):
    """out[N,Kin] += dY[M,N]^T @ X[M,Kin] (fp32, `out` pre-zeroed) -- weight gradient of nn.Linear."""
    with _Timed("gemm_dw", 2.0 * dY.shape[0] * dY.shape[1] * X.shape[1], 1):
        M, N = dY.shape
        Kin = X.shape[1]
        ldy, ldx, ldc = dY.stride(0), X.stride(0), out.stride(0)
        if _tc_ok(dY.dtype, Kin, ldc, ldy, ldx):
            split = _dw_split(N, Kin, M)
            K.check(K.lib().smer_gemm_bf16_tc(_p(dY), ldy, 0, _p(X), ldx, 0, _p(out), ldc, K.F32, N, Kin, M, 0, 0, 0,
                                              K.EPI_ATOMIC, 0.0, 0, 0, split, 0, K.stream()), "gemm_tc(dw)")
        else:
            tiles = ((N + 63) // 64) * ((Kin + 63) // 64)
            split = max(1, min((M + 63) // 64, (2 * num_sms()) // tiles))
            K.check(K.lib().smer_gemm_simt(_p(dY), 1, ldy, _p(X), 1, ldx, _p(out), ldc, K.dt(dY), K.F32, N, Kin, M, 0, 0,
                                           0, K.EPI_ATOMIC, 0.0, 0, 0, split, K.stream()), "gemm_simt(dw)")


def colsum(x, out):
    with _Timed("colsum", float(x.numel() * x.element_size()), 1):
        rows, cols = x.shape
        K.check(K.lib().smer_colsum(_p(x), K.dt(x), x.stride(0), _p(out), rows, cols, K.stream()), "colsum")


def attn_args(q, k, v, o, B, H, Lq, Lk, dh, *, lse=None, causal=False, q_pos0=0, key_pad=None, kv_len=None,
              add_mask=None, dropout_p=0.0, seed=0, site=0, dout=None, dq=None, dk=None, dv=None, dsum=None,
              dbq=None, dbk=None, dbv=None, dq_accum=None, cu_q=None, cu_k=None):
    """dbq/dbk/dbv (fp32 [H*dh], backward): += column sums of dq/dk/dv = the in-projection's bias gradient."""
    a = K.AttnArgs()
    a.dbq, a.dbk, a.dbv = _p(dbq), _p(dbk), _p(dbv)
    a._db, a._dqkv = (dbq, dbk, dbv), (dq, dk, dv)      # kept for the CUDA-core path, which sums them separately
    a.q, a.k, a.v, a.o = _p(q), _p(k), _p(v), _p(o)
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    a.dout, a.dq, a.dk, a.dv = _p(dout), _p(dq), _p(dk), _p(dv)
    a.lddo = dout.stride(0) if dout is not None else 0
    a.lddq = dq.stride(0) if dq is not None else 0
    a.lddk = dk.stride(0) if dk is not None else 0
    a.lddv = dv.stride(0) if dv is not None else 0
    a.lse, a.dsum = _p(lse), _p(dsum)
    a.key_pad, a.kv_len = _p(key_pad), _p(kv_len)
    a.add_mask = _p(add_mask)
    a.ld_mask = add_mask.stride(0) if add_mask is not None else 0
    a.B, a.H, a.Lq, a.Lk, a.dh = B, H, Lq, Lk, dh
    a.dtype = K.dt(q)
    a.causal, a.q_pos0 = int(causal), q_pos0
    a.scale = 1.0 / math.sqrt(dh)
    a.dropout_p, a.seed, a.site = dropout_p, seed, site
    a.dq_accum = _p(dq_accum)
    a._keep = dq_accum
    # padding-free layout: per-sequence row ranges of the packed q / k,v buffers (B, Lq, Lk then mean: sequences,
    # longest query sequence, longest key sequence)
    a.cu_q, a.cu_k = _p(cu_q), _p(cu_k)
    a.q_rows, a.k_rows = (q.shape[0], k.shape[0]) if cu_q is not None else (0, 0)
    return a


def _attn_tc_ok(a) -> bool:
    return (_TC_ATTN == "tc" and a.dtype == K.BF16 and a.dh == 64 and not a.add_mask and a.q_pos0 == 0
            and (not a.causal or a.Lq == a.Lk) and a.Lk <= 16384 and a.dropout_p <= 0.45
            and all(x % 8 == 0 for x in (a.ldq, a.ldk, a.ldv, a.ldo))
            and all((x or 0) % 16 == 0 for x in (a.q, a.k, a.v, a.o)))


def _attn_flops(a):
    """Algorithmic flops of one attention forward: 4*Lq*Lk*dh per (b,h); causal counts half."""
    f = 4.0 * a.B * a.H * a.Lq * a.Lk * a.dh
    return f * 0.5 if (a.causal and a.Lq == a.Lk) else f


def attn_fwd(a):
    if a.cu_q and not _attn_tc_ok(a):
        raise RuntimeError("packed-row attention needs the tcgen05 kernels (bf16, head dim 64, no additive mask)")
    with _Timed("attn_fwd", _attn_flops(a)):
        if _attn_tc_ok(a) and ATTN_TC_FWD:
            K.check(K.lib().smer_attn_fwd_tc(C.byref(a), K.stream()), "attn_fwd_tc")
        else:
            K.check(K.lib().smer_attn_fwd_simt(C.byref(a), K.stream()), "attn_fwd_simt")


def attn_bwd(a):
    tc = _attn_tc_ok(a) and ATTN_TC_BWD
    if a.cu_q and not tc:
        raise RuntimeError("packed-row attention needs the tcgen05 kernels (bf16, head dim 64, no additive mask)")
    if tc and not a.dq_accum:
        # fp32 accumulation buffer of the fused backward (dQ partial sums of the key-tile CTAs)
        rows = a.q_rows if a.cu_q else a.B * a.Lq
        a._keep = torch.empty(rows, a.H * a.dh, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        a.dq_accum = a._keep.data_ptr()
    with _Timed("attn_bwd", 2.0 * _attn_flops(a), 3 if tc else 3):      # tc: D = rowsum(dO o O), fused dQ/dK/dV, dQ convert
        if tc:
            K.check(K.lib().smer_attn_bwd_tc(C.byref(a), K.stream()), "attn_bwd_tc")
        else:
            K.check(K.lib().smer_attn_bwd_simt(C.byref(a), K.stream()), "attn_bwd_simt")
    if not tc:
        for g, db in zip(a._dqkv, a._db):
            if db is not None:
                colsum(g, db)


def attn_weights(a, w):
    with _Timed("attn_weights"):
        K.check(K.lib().smer_attn_weights(C.byref(a), _p(w), w.stride(-2), K.stream()), "attn_weights")


# Which tcgen05 attention kernels this build provides (attn_tc.cu); the dispatcher above is a
# capability table of the one library, not a backend switch.
ATTN_TC_FWD = True
ATTN_TC_BWD = True


def token_accuracy(logits, targets, class_of, ncls, counts, argmax_out=None):
    with _Timed("accuracy", float(logits.shape[0] * logits.shape[1] * 4), 1):
        K.check(K.lib().smer_token_accuracy(_p(logits), logits.stride(0), _p(targets), _p(class_of), ncls, _p(counts),
                                            _p(argmax_out), logits.shape[0], logits.shape[1], K.stream()), "token_accuracy")


def xent_fwd(logits, targets, W, Cw, category, ncat, lse, sums, V):
    with _Timed("xent_fwd", float(logits.shape[0] * V * 4), 1):
        rows = logits.shape[0]
        K.check(K.lib().smer_xent_fwd(_p(logits), logits.stride(0), _p(targets), _p(W), _p(Cw), _p(category), ncat,
                                      _p(lse), _p(sums), rows, V, K.stream()), "xent_fwd")


def xent_denominator(targets, Cw, sums, V):
    with _Timed("xent_fwd", 0.0, 1):
        K.check(K.lib().smer_xent_denominator(_p(targets), _p(Cw), _p(sums), targets.numel(), V, K.stream()), "xent_denominator")


def xent_bwd(logits, targets, W, lse, sums, dlogits, V, grad_scale=1.0, grad_scale_dev=None):
    with _Timed("xent_bwd", float(logits.shape[0] * V * (4 + dlogits.element_size())), 1):
        rows = logits.shape[0]
        K.check(K.lib().smer_xent_bwd(_p(logits), logits.stride(0), _p(targets), _p(W), _p(lse), _p(sums), _p(dlogits),
                                      K.dt(dlogits), dlogits.stride(0), rows, V, dlogits.shape[1], grad_scale,
                                      _p(grad_scale_dev), K.stream()), "xent_bwd")


def adam_step(p, g, m, v, shadow, step, lr, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0, step_dev=None):
    with _Timed("adam", float(p.numel() * (28 + (2 if shadow is not None else 0))), 1):
        if step_dev is not None:
            K.check(K.lib().smer_adam_step_dev(_p(p), _p(g), _p(m), _p(v), _p(shadow), p.numel(), _p(step_dev), lr, b1,
                                               b2, eps, grad_scale, K.stream()), "adam_step_dev")
        else:
            K.check(K.lib().smer_adam_step(_p(p), _p(g), _p(m), _p(v), _p(shadow), p.numel(), step, lr, b1, b2, eps,
                                           grad_scale, K.stream()), "adam_step")


def adam_multi(table_dev, n_tensors, max_n, total_n, step, lr, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0):
    """One launch over a device table of (p, g, m, v, n) entries (40 bytes each, smer_adam_tensor)."""
    with _Timed("adam", float(total_n * 28), 1):
        K.check(K.lib().smer_adam_multi(_p(table_dev), n_tensors, max_n, step, lr, b1, b2, eps, grad_scale, K.stream()),
                "adam_multi")


def cast2d(src, dst, cols=None):
    with _Timed("cast", 0.0, 1):
        rows = src.shape[0]
        cols = src.shape[1] if cols is None else cols
        K.check(K.lib().smer_cast2d(_p(src), K.dt(src), src.stride(0), _p(dst), K.dt(dst), dst.stride(0), rows, cols,
                                    dst.shape[1], K.stream()), "cast2d")


def cast_f32_to_bf16(src, dst):
    with _Timed("cast", float(src.numel() * 6), 1):
        K.check(K.lib().smer_cast_f32_to_bf16(_p(src), _p(dst), src.numel(), K.stream()), "cast_f32_to_bf16")


def kv_len_from_pad(pad_u8, out):
    with _Timed("misc", 0.0, 1):
        B, L = pad_u8.shape
        K.check(K.lib().smer_kv_len_from_pad(_p(pad_u8), _p(out), B, L, K.stream()), "kv_len_from_pad")


def classify_mask(mask2d, flags3):
    with _Timed("misc", 0.0, 1):
        T = mask2d.shape[0]
        K.check(K.lib().smer_classify_mask(_p(mask2d), mask2d.stride(0), T, _p(flags3), K.stream()), "classify_mask")


def decode_linear(a, w, out, bias=None, resid=None, relu=False, ln=None, ln_out=None, ln2=None, eps=1e-5):
    """Small-M linear of the decode step (csrc/decode_linear.cu): out = epi(LN2?(LN?(a)) @ w^T + bias)."""
    with _Timed("gemm", 2.0 * a.shape[0] * a.shape[1] * w.shape[0], 1):
        M, Kd = a.shape
        N = w.shape[0]
        K.check(K.lib().smer_decode_linear(_p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(resid),
                                           resid.stride(0) if resid is not None else 0, _p(out), out.stride(0), K.dt(out), M, N, Kd,
                                           int(relu), _p(ln[0]) if ln else 0, _p(ln[1]) if ln else 0, _p(ln_out),
                                           ln_out.stride(0) if ln_out is not None else 0,
                                           _p(ln2[0]) if ln2 else 0, _p(ln2[1]) if ln2 else 0, eps, K.stream()), "decode_linear")
