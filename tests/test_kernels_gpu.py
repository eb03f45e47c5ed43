"""Per-kernel parity through the C ABI (ctypes) against a plain torch fp32 restatement of the
same op on the same seeded inputs.  GPU only."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from smer_music_generation_b200 import _capi as K
    K.require_cuda_device()
    return torch.device("cuda:0")


def _ops():
    from smer_music_generation_b200 import ops, _capi as K
    return ops, K


def rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-30)


# ---------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K_", [(128, 128, 64), (256, 512, 512), (300, 320, 512), (1, 512, 512), (77, 1536, 512),
                                    (1024, 2048, 512), (640, 512, 2048), (130, 64, 32), (48, 96, 64)])
def test_gemm_tc_nt(dev, M, N, K_):
    ops, K = _ops()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    A = (torch.randn(M, K_, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K_, generator=g) * 0.1).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.float() @ W.float().t() + bias
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops.gemm_nt(A, W, out, bias=bias)
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-2
    out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    ops.gemm_nt(A, W, out32, bias=bias)
    assert rel(out32, ref) < 1e-5
    # relu + residual epilogues
    res = torch.randn(M, N, generator=g).to(dev).bfloat16()
    ops.gemm_nt(A, W, out, bias=bias, flags=K.EPI_RELU)
    assert rel(out, torch.relu(ref)) < 1e-2
    ops.gemm_nt(A, W, out, bias=bias, resid=res)
    assert rel(out, ref + res.float()) < 1e-2


@pytest.mark.parametrize("M,N,K_", [(8192, 1536, 1024), (8292, 512, 1024), (10000, 320, 2048), (16384, 2048, 1088)])
def test_gemm_tc_cta_pair_shapes(dev, M, N, K_):
    """Shapes large enough for the cta_group::2 (CTA-pair, 256x256 tile) variant: forward with every
    specialised epilogue, input gradient (B MN-major) and split-K weight gradient (A and B MN-major)."""
    ops, K = _ops()
    g = torch.Generator(device="cpu").manual_seed(M + N)
    A = (torch.randn(M, K_, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K_, generator=g) * 0.1).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.float() @ W.float().t() + bias
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops.gemm_nt(A, W, out, bias=bias)
    assert rel(out, ref) < 1e-2
    ops.gemm_nt(A, W, out, bias=bias, flags=K.EPI_RELU)
    assert rel(out, torch.relu(ref)) < 1e-2
    out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    ops.gemm_nt(A, W, out32, bias=bias)
    assert rel(out32, ref) < 1e-5
    # dropout epilogue: kept entries equal relu(ref)/(1-p), keep rate ~ 0.9
    ops.gemm_nt(A, W, out, bias=bias, flags=K.EPI_RELU, dropout_p=0.1, seed=5, site=9)
    pos = torch.relu(ref) > 1e-3
    kept = (out.float() != 0) & pos
    assert abs(kept.sum().item() / pos.sum().item() - 0.9) < 5e-3
    assert rel(out.float()[kept], (torch.relu(ref) / 0.9)[kept]) < 1e-2
    # input gradient and weight gradient
    dY = (torch.randn(M, N, generator=g) * 0.3).to(dev).bfloat16()
    res = torch.randn(M, K_, generator=g).to(dev).bfloat16()
    dx = torch.empty(M, K_, dtype=torch.bfloat16, device=dev)
    ops.gemm_dx(dY, W, dx, resid=res)
    assert rel(dx, dY.float() @ W.float() + res.float()) < 1e-2
    dw = torch.zeros(N, K_, dtype=torch.float32, device=dev)
    ops.gemm_dw(dY, A, dw)
    assert rel(dw, dY.float().t() @ A.float()) < 5e-5


def test_gemm_tc_strided_views(dev):
    ops, K = _ops()
    g = torch.Generator().manual_seed(3)
    M, d = 200, 64
    buf = (torch.randn(M, 3 * d, generator=g)).to(dev).bfloat16()
    W = (torch.randn(2 * d, d, generator=g) * 0.2).to(dev).bfloat16()
    out = torch.zeros(M, 4 * d, dtype=torch.bfloat16, device=dev)
    ops.gemm_nt(buf[:, d:2 * d], W, out[:, d:3 * d])
    ref = buf[:, d:2 * d].float() @ W.float().t()
    assert rel(out[:, d:3 * d], ref) < 1e-2
    assert out[:, :d].abs().max().item() == 0 and out[:, 3 * d:].abs().max().item() == 0


@pytest.mark.parametrize("M,N,Kin", [(256, 512, 512), (300, 320, 512), (1000, 2048, 512), (513, 512, 2048), (64, 1536, 512)])
def test_gemm_tc_dx_dw(dev, M, N, Kin):
    ops, K = _ops()
    g = torch.Generator().manual_seed(N + Kin)
    dY = (torch.randn(M, N, generator=g) * 0.3).to(dev).bfloat16()
    W = (torch.randn(N, Kin, generator=g) * 0.1).to(dev).bfloat16()
    X = (torch.randn(M, Kin, generator=g) * 0.5).to(dev).bfloat16()
    res = torch.randn(M, Kin, generator=g).to(dev).bfloat16()
    dx = torch.empty(M, Kin, dtype=torch.bfloat16, device=dev)
    ops.gemm_dx(dY, W, dx, resid=res)
    assert rel(dx, dY.float() @ W.float() + res.float()) < 1e-2
    # gate epilogue: backward of dropout(relu(.)) with p=0
    act = torch.relu(torch.randn(M, Kin, generator=g)).to(dev).bfloat16()
    ops.gemm_dx(dY, W, dx, resid=act, flags=K.EPI_GATE)
    assert rel(dx, (dY.float() @ W.float()) * (act.float() > 0)) < 1e-2
    # ... with the bias gradient (column sums of the gated result) accumulated by the same epilogue
    cs = torch.full((Kin,), 0.25, dtype=torch.float32, device=dev)
    dx2 = torch.empty_like(dx)
    ops.gemm_dx(dY, W, dx2, resid=act, flags=K.EPI_GATE, colsum_out=cs)
    assert torch.equal(dx2, dx)
    want = ((dY.float() @ W.float()) * (act.float() > 0)).sum(0) + 0.25
    assert (cs - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
    dw = torch.zeros(N, Kin, dtype=torch.float32, device=dev)
    ops.gemm_dw(dY, X, dw)
    assert rel(dw, dY.float().t() @ X.float()) < 2e-5
    db = torch.zeros(N, dtype=torch.float32, device=dev)
    ops.colsum(dY, db)
    assert rel(db, dY.float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,N,K_", [(37, 309, 32), (64, 64, 64), (130, 96, 50)])
def test_gemm_simt_fp32(dev, M, N, K_):
    ops, K = _ops()
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, K_, generator=g).to(dev)
    W = torch.randn(N, K_, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.empty(M, N, device=dev)
    ops.gemm_nt(A, W, out, bias=bias)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    assert rel(out, ref) < 1e-5
    dY = torch.randn(M, N, generator=g).to(dev)
    dx = torch.empty(M, K_, device=dev)
    ops.gemm_dx(dY, W, dx)
    assert rel(dx, (dY.double() @ W.double()).float()) < 1e-5
    dw = torch.zeros(N, K_, device=dev)
    ops.gemm_dw(dY, A, dw)
    assert rel(dw, (dY.double().t() @ A.double()).float()) < 1e-5


# ---------------------------------------------------------------------------------------- LN / embed
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d", [(5, 32), (1000, 512), (33, 768)])
def test_layernorm(dev, dtype, rows, d):
    ops, K = _ops()
    g = torch.Generator().manual_seed(rows + d)
    br = torch.randn(rows, d, generator=g).to(dev).to(dtype)
    rs = torch.randn(rows, d, generator=g).to(dev).to(dtype)
    gam = (1 + 0.1 * torch.randn(d, generator=g)).to(dev)
    bet = (0.1 * torch.randn(d, generator=g)).to(dev)
    z = torch.empty_like(br)
    y = torch.empty_like(br)
    mean = torch.empty(rows, device=dev)
    rstd = torch.empty(rows, device=dev)
    ops.layernorm_fwd(br, rs, gam, bet, z, y, mean, rstd)
    zr = (br.float() + rs.float()).to(dtype).float().requires_grad_(True)
    gr = gam.clone().requires_grad_(True)
    br_ = bet.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(zr, (d,), gr, br_, 1e-5)
    tol = 1e-5 if dtype == torch.float32 else 1.5e-2
    assert rel(y, yr) < tol
    dy = torch.randn(rows, d, generator=g).to(dev).to(dtype)
    yr.backward(dy.float())
    dz = torch.empty_like(br)
    dg = torch.zeros(d, device=dev)
    db = torch.zeros(d, device=dev)
    dbias = torch.zeros(d, device=dev)
    ops.layernorm_bwd(dy, z, mean, rstd, gam, dz, None, dg, db, dbias=dbias)
    assert rel(dz, zr.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert rel(dbias, zr.grad.sum(0)) < (1e-4 if dtype == torch.float32 else 2e-2) or zr.grad.sum(0).abs().max() < 1e-3
    assert rel(dg, gr.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
    assert rel(db, br_.grad) < 1e-4


def test_layernorm_dropout_consistency(dev):
    """Backward regenerates exactly the forward's Philox mask; keep-rate matches p."""
    ops, K = _ops()
    rows, d, p = 512, 512, 0.1
    br = torch.ones(rows, d, device=dev)
    rs = torch.zeros(rows, d, device=dev)
    gam, bet = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    z, y = torch.empty_like(br), torch.empty_like(br)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    ops.layernorm_fwd(br, rs, gam, bet, z, y, mean, rstd, dropout_p=p, seed=1234, site=7)
    keep = (z != 0)
    assert abs(keep.float().mean().item() - (1 - p)) < 5e-3
    assert torch.allclose(z[keep], torch.full_like(z[keep], 1 / (1 - p)))
    dz, dbr = torch.empty_like(br), torch.empty_like(br)
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.layernorm_bwd(torch.randn(rows, d, device=dev), z, mean, rstd, gam, dz, dbr, dg, db, dropout_p=p, seed=1234, site=7)
    assert torch.equal(dbr != 0, keep & (dz != 0))
    z2 = torch.empty_like(br)
    ops.layernorm_fwd(br, rs, gam, bet, z2, y, mean, rstd, dropout_p=p, seed=1234, site=8)
    assert not torch.equal(z2 != 0, keep)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed(dev, dtype):
    ops, K = _ops()
    from smer_music_generation_b200.model import _PositionalEncoding
    g = torch.Generator().manual_seed(0)
    B, L, d, V = 3, 40, 64, 309
    ids = torch.randint(0, V, (B, L), generator=g).to(dev)
    emb = torch.randn(V, d, generator=g).to(dev)
    pe = _PositionalEncoding(d, 100).pe.to(dev).view(-1, d)
    out = torch.empty(B * L, d, dtype=dtype, device=dev)
    ops.embed_pe(ids, emb, pe, out, math.sqrt(d), 5)
    ref = emb[ids] * math.sqrt(d) + pe[5:5 + L][None]
    assert rel(out.view(B, L, d), ref) < (1e-6 if dtype == torch.float32 else 1e-2)
    dout = torch.randn(B * L, d, generator=g).to(dev).to(dtype)
    demb = torch.zeros(V, d, device=dev)
    ops.embed_bwd(ids, dout, demb, math.sqrt(d))
    refg = torch.zeros(V, d, device=dev).index_add_(0, ids.view(-1), dout.float() * math.sqrt(d))
    assert rel(demb, refg) < 1e-5


# ---------------------------------------------------------------------------------------- xent
def test_xent_vs_oracle(dev, oracle):
    ops, K = _ops()
    from smer_music_generation_b200 import SmerLoss
    g = torch.Generator().manual_seed(2)
    N, V = 777, 309
    logits = (torch.randn(N, V, generator=g) * 2).requires_grad_(True)
    tgt = torch.randint(0, V, (N,), generator=g)
    tgt[::5] = 0
    for eos_w in (1.0, 0.8):
        W, Cw = oracle.loss_weights(eos_w)
        loss_ref, parts_ref, denom_ref = oracle.smer_loss(logits, tgt, W, Cw)
        (gref,) = torch.autograd.grad(loss_ref, logits)
        crit = SmerLoss(V, eos_w).to(dev)
        lg = logits.detach().to(dev).requires_grad_(True)
        loss, parts, denom = crit(lg, tgt.to(dev))
        loss.backward()
        assert abs(loss.item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
        assert rel(parts.cpu(), parts_ref) < 1e-5
        assert abs(denom.item() - denom_ref.item()) < 1e-3
        assert rel(lg.grad.cpu(), gref) < 1e-4
    # padded-pitch (B,T,V) view, as the model returns it
    full = torch.zeros(7, 111, 320, device=dev)
    full[:, :, :V] = logits.detach().to(dev).view(7, 111, V)
    view = full[:, :, :V].requires_grad_(True)
    loss2, _, _ = crit(view, tgt.to(dev).view(7, 111))
    assert abs(loss2.item() - loss.item()) < 1e-6


# ---------------------------------------------------------------------------------------- attention
def _ref_attn(q, k, v, H, causal, pad, q_pos0=0, add_mask=None):
    B, Lq, d = q.shape
    Lk = k.shape[1]
    dh = d // H
    qh = q.view(B, Lq, H, dh).transpose(1, 2) / math.sqrt(dh)
    kh = k.view(B, Lk, H, dh).transpose(1, 2)
    vh = v.view(B, Lk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if causal:
        i = torch.arange(Lq, device=q.device)[:, None] + q_pos0
        j = torch.arange(Lk, device=q.device)[None]
        s = s.masked_fill(j > i, float("-inf"))
    if add_mask is not None:
        s = s + add_mask
    if pad is not None:
        s = s.masked_fill(pad[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, d), p


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,dh,Lq,Lk,causal,use_pad", [(2, 16, 24, 24, True, True), (8, 64, 200, 200, True, False),
                                                      (8, 64, 150, 333, False, True), (4, 32, 70, 70, False, True),
                                                      (8, 64, 256, 256, True, True)])
def test_attention_fwd_bwd(dev, dtype, H, dh, Lq, Lk, causal, use_pad):
    ops, K = _ops()
    g = torch.Generator().manual_seed(Lq + Lk)
    B, d = 3, H * dh
    qkv_q = torch.randn(B * Lq, d, generator=g).to(dev).to(dtype)
    kv = torch.randn(B * Lk, 2 * d, generator=g).to(dev).to(dtype)
    pad = None
    kv_len = None
    if use_pad:
        lens = torch.tensor([Lk, max(1, Lk - 7), max(1, Lk // 2)])
        pad = (torch.arange(Lk)[None] >= lens[:, None]).to(torch.uint8).to(dev)
        kv_len = lens.to(torch.int32).to(dev)
    o = torch.empty(B * Lq, d, dtype=dtype, device=dev)
    lse = torch.empty(B, H, Lq, device=dev)
    a = ops.attn_args(qkv_q, kv[:, :d], kv[:, d:], o, B, H, Lq, Lk, dh, lse=lse, causal=causal, key_pad=pad, kv_len=kv_len)
    ops.attn_fwd(a)
    qr = qkv_q.float().view(B, Lq, d).requires_grad_(True)
    kr = kv[:, :d].float().reshape(B, Lk, d).requires_grad_(True)
    vr = kv[:, d:].float().reshape(B, Lk, d).requires_grad_(True)
    ref, p = _ref_attn(qr, kr, vr, H, causal, pad)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert rel(o.view(B, Lq, d), ref) < tol
    # head-averaged probabilities
    w = torch.empty(B, Lq, Lk, device=dev)
    ops.attn_weights(a, w)
    assert rel(w, p.mean(1)) < tol
    do = torch.randn(B * Lq, d, generator=g).to(dev).to(dtype)
    ref.backward(do.float().view(B, Lq, d))
    dq = torch.empty_like(qkv_q)
    dkv = torch.empty_like(kv)
    dsum = torch.empty(B, H, Lq, device=dev)
    a2 = ops.attn_args(qkv_q, kv[:, :d], kv[:, d:], o, B, H, Lq, Lk, dh, lse=lse, causal=causal, key_pad=pad,
                       kv_len=kv_len, dout=do, dq=dq, dk=dkv[:, :d], dv=dkv[:, d:], dsum=dsum)
    ops.attn_bwd(a2)
    btol = 1e-4 if dtype == torch.float32 else 3e-2
    assert rel(dq.view(B, Lq, d), qr.grad) < btol
    assert rel(dkv[:, :d].reshape(B, Lk, d), kr.grad) < btol
    assert rel(dkv[:, d:].reshape(B, Lk, d), vr.grad) < btol


@pytest.mark.parametrize("Lq,Lk,causal,use_pad,drop", [(128, 128, False, False, 0.0), (256, 256, True, False, 0.0),
                                                       (200, 200, True, True, 0.0), (1024, 1024, True, True, 0.0),
                                                       (100, 777, False, True, 0.0), (384, 1024, False, True, 0.1),
                                                       (512, 512, True, True, 0.1), (1, 130, False, False, 0.0),
                                                       (2304, 2304, True, True, 0.1), (640, 2100, False, True, 0.0)])
def test_attention_tc_vs_simt_and_torch(dev, Lq, Lk, causal, use_pad, drop):
    """tcgen05 forward == CUDA-core forward on the same bf16 inputs (same dropout mask: both use
    the counter hash of common.cuh) and == torch fp32 when dropout is off."""
    ops, K = _ops()
    assert ops.ATTN_TC_FWD
    g = torch.Generator().manual_seed(Lq * 3 + Lk)
    B, H, dh = 2, 8, 64
    d = H * dh
    qkv = torch.randn(B * Lq, 3 * d, generator=g).to(dev).bfloat16()
    kvb = torch.randn(B * Lk, 2 * d, generator=g).to(dev).bfloat16()
    q, k, v = qkv[:, :d], kvb[:, :d], kvb[:, d:]
    pad = kv_len = None
    if use_pad:
        lens = torch.tensor([Lk, max(1, Lk - 37)])
        pad = (torch.arange(Lk)[None] >= lens[:, None]).to(torch.uint8).to(dev)
        if not causal:
            pad[1, Lk // 3] = 1                  # a non-suffix masked key: honoured through the per-key mask
        # the loop bound comes from the library's own helper: its sign says whether the mask is a pure suffix (row 0)
        # or has holes (row 1 of the non-causal cases) -- include/smer_b200.h
        kv_len = torch.empty(B, dtype=torch.int32, device=dev)
        ops.kv_len_from_pad(pad, kv_len)
        assert kv_len.tolist() == [int(lens[0]), int(lens[1]) if causal else -int(lens[1])]
    outs = {}
    for path in ("tc", "simt"):
        ops._TC_ATTN = path
        o = torch.full((B * Lq, d), float("nan"), dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, H, Lq, device=dev)
        a = ops.attn_args(q, k, v, o, B, H, Lq, Lk, dh, lse=lse, causal=causal, key_pad=pad, kv_len=kv_len,
                          dropout_p=drop, seed=77, site=3)
        ops.attn_fwd(a)
        torch.cuda.synchronize()
        outs[path] = (o, lse)
    ops._TC_ATTN = "tc"
    assert torch.isfinite(outs["tc"][0].float()).all()
    assert rel(outs["tc"][0], outs["simt"][0]) < 2e-2
    assert (outs["tc"][1] - outs["simt"][1]).abs().max().item() < 2e-3
    if drop == 0.0:
        ref, _ = _ref_attn(q.float().reshape(B, Lq, d), k.float().reshape(B, Lk, d), v.float().reshape(B, Lk, d), H,
                           causal, pad)
        assert rel(outs["tc"][0].view(B, Lq, d), ref) < 2e-2
    # ---- backward: tcgen05 kernels vs CUDA-core kernels from the same forward state
    assert ops.ATTN_TC_BWD
    o, lse = outs["simt"]
    do = torch.randn(B * Lq, d, generator=g).to(dev).bfloat16()
    grads = {}
    for path in ("tc", "simt"):
        ops._TC_ATTN = path
        dqkv = torch.full((B * Lq, 3 * d), float("nan"), dtype=torch.bfloat16, device=dev)
        dkv = torch.full((B * Lk, 2 * d), float("nan"), dtype=torch.bfloat16, device=dev)
        dsum = torch.empty(B, H, Lq, device=dev)
        db = torch.full((3 * d,), 0.5, dtype=torch.float32, device=dev)      # bias-gradient accumulators (+=)
        a = ops.attn_args(q, k, v, o, B, H, Lq, Lk, dh, lse=lse, causal=causal, key_pad=pad, kv_len=kv_len,
                          dropout_p=drop, seed=77, site=3, dout=do, dq=dqkv[:, :d], dk=dkv[:, :d], dv=dkv[:, d:], dsum=dsum,
                          dbq=db[:d], dbk=db[d:2 * d], dbv=db[2 * d:])
        ops.attn_bwd(a)
        torch.cuda.synchronize()
        grads[path] = (dqkv[:, :d].clone(), dkv[:, :d].clone(), dkv[:, d:].clone())
        # in-projection bias gradient = column sums of dq | dk | dv (fused into the tcgen05 kernels' epilogues)
        want = torch.cat([x.float().sum(0) for x in grads[path]]) + 0.5
        tol = 2e-2 * max(1.0, want.abs().max().item())
        assert (db - want).abs().max().item() < tol, (path, (db - want).abs().max().item(), tol)
    ops._TC_ATTN = "tc"
    for name, x, y in zip(("dq", "dk", "dv"), grads["tc"], grads["simt"]):
        assert torch.isfinite(x.float()).all(), name
        assert rel(x, y) < 3e-2, (name, rel(x, y))
    if drop == 0.0:
        qr = q.float().reshape(B, Lq, d).requires_grad_(True)
        kr = k.float().reshape(B, Lk, d).requires_grad_(True)
        vr = v.float().reshape(B, Lk, d).requires_grad_(True)
        ref, _ = _ref_attn(qr, kr, vr, H, causal, pad)
        ref.backward(do.float().view(B, Lq, d))
        assert rel(grads["tc"][0].view(B, Lq, d), qr.grad) < 3e-2
        assert rel(grads["tc"][1].view(B, Lk, d), kr.grad) < 3e-2
        assert rel(grads["tc"][2].view(B, Lk, d), vr.grad) < 3e-2


def test_attention_dropout_mask_statistics(dev):
    """Keep-mask of the attention-probability dropout (counter hash in common.cuh): keep rate p,
    no correlation between the two keys of a pair, neighbouring pairs, neighbouring rows, or sites."""
    ops, K = _ops()
    B, H, dh, L, p = 2, 1, 64, 512, 0.1
    q = torch.zeros(B * L, dh, device=dev)                      # uniform attention: P = 1/L everywhere
    k = torch.zeros(B * L, dh, device=dev)
    v = torch.zeros(B * L, dh, device=dev)
    o = torch.empty_like(q)
    masks = []
    for site in (1, 2):
        lse = torch.empty(B, H, L, device=dev)
        a = ops.attn_args(q, k, v, o, B, H, L, L, dh, lse=lse, dropout_p=p, seed=2024, site=site)
        ops.attn_fwd(a)
        w = torch.empty(B, L, L, device=dev)
        ops.attn_weights(a, w)
        masks.append((w > 0).float())
    m = masks[0]
    n = m.numel()
    assert abs(m.mean().item() - (1 - p)) < 3e-3
    assert abs(m[:, :, 0::2].mean().item() - (1 - p)) < 4e-3 and abs(m[:, :, 1::2].mean().item() - (1 - p)) < 4e-3

    def corr(x, y):
        x = x - x.mean()
        y = y - y.mean()
        return (x * y).mean().item() / (x.std().item() * y.std().item() + 1e-12)

    tol = 5.0 / (n / 2) ** 0.5                                  # ~5 sigma for independent bits
    assert abs(corr(m[:, :, 0::2], m[:, :, 1::2])) < tol        # even vs odd key of a pair
    assert abs(corr(m[:, :, :-2], m[:, :, 2:])) < tol           # neighbouring pairs
    assert abs(corr(m[:, :-1, :], m[:, 1:, :])) < tol           # neighbouring rows
    assert abs(corr(m[0], m[1])) < tol                          # batches
    assert abs(corr(masks[0], masks[1])) < tol                  # dropout sites
    # row keep counts follow Binomial(L, 1-p): variance check
    rows = m.sum(-1).flatten()
    assert abs(rows.var().item() / (L * p * (1 - p)) - 1) < 0.25


def test_attention_qpos_and_addmask(dev):
    ops, K = _ops()
    g = torch.Generator().manual_seed(9)
    B, H, dh, Lq, Lk = 1, 2, 16, 3, 20
    d = H * dh
    q = torch.randn(B * Lq, d, generator=g).to(dev)
    k = torch.randn(B * Lk, d, generator=g).to(dev)
    v = torch.randn(B * Lk, d, generator=g).to(dev)
    o = torch.empty_like(q)
    a = ops.attn_args(q, k, v, o, B, H, Lq, Lk, dh, causal=True, q_pos0=Lk - Lq)
    ops.attn_fwd(a)
    ref, _ = _ref_attn(q.view(B, Lq, d), k.view(B, Lk, d), v.view(B, Lk, d), H, True, None, q_pos0=Lk - Lq)
    assert rel(o.view(B, Lq, d), ref) < 2e-5
    am = torch.randn(Lq, Lk, generator=g).to(dev)
    a = ops.attn_args(q, k, v, o, B, H, Lq, Lk, dh, add_mask=am)
    ops.attn_fwd(a)
    ref, _ = _ref_attn(q.view(B, Lq, d), k.view(B, Lk, d), v.view(B, Lk, d), H, False, None, add_mask=am)
    assert rel(o.view(B, Lq, d), ref) < 2e-5


# ---------------------------------------------------------------------------------------- decode attention
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("splits", [1, 4])
def test_decode_attn(dev, dtype, splits):
    ops, K = _ops()
    g = torch.Generator().manual_seed(4)
    n, H, dh, L = 5, 8, 64, 96
    d = H * dh
    cache = torch.randn(n, L, 2 * d, generator=g).to(dev).to(dtype)
    qkv = torch.randn(n, 3 * d, generator=g).to(dev).to(dtype)
    kv_len = torch.tensor([0, 1, 17, 50, 95], dtype=torch.int32).to(dev)
    out = torch.empty(n, d, dtype=dtype, device=dev)
    ref_cache = cache.clone()
    a = K.DecodeAttnArgs()
    a.q, a.new_k, a.new_v = qkv.data_ptr(), qkv[:, d:].data_ptr(), qkv[:, 2 * d:].data_ptr()
    a.k_cache, a.v_cache, a.out = cache.data_ptr(), cache[:, :, d:].data_ptr(), out.data_ptr()
    ws = torch.empty(max(4, K.lib().smer_decode_attn_workspace_bytes(n, H, dh, splits)) // 4, device=dev)
    a.kv_len, a.key_pad, a.workspace = kv_len.data_ptr(), None, ws.data_ptr()
    a.ldq, a.ld_new, a.ldo, a.ld_cache, a.cache_stride, a.ld_pad = 3 * d, 3 * d, d, 2 * d, L * 2 * d, 0
    a.n_seq, a.H, a.dh, a.cache_len, a.splits, a.dtype = n, H, dh, L, splits, K.dt(qkv)
    a.scale = 1 / math.sqrt(dh)
    K.check(K.lib().smer_decode_attn(C.byref(a), K.stream()))
    for s in range(n):
        t = int(kv_len[s])
        ref_cache[s, t] = qkv[s, d:]
        kk = ref_cache[s, : t + 1, :d].float().view(1, t + 1, d)
        vv = ref_cache[s, : t + 1, d:].float().view(1, t + 1, d)
        r, _ = _ref_attn(qkv[s, :d].float().view(1, 1, d), kk, vv, H, False, None)
        assert rel(out[s], r.view(d)) < (2e-5 if dtype == torch.float32 else 2e-2)
        assert torch.equal(cache[s, t], ref_cache[s, t])


# ---------------------------------------------------------------------------------------- sampler
def _sample_probs(K, logits, raw_flags, lo, hi, mode, t=1.0, top_p=0.9, top_k=0):
    dev = logits.device
    n, V = logits.shape
    a = K.SampleArgs()
    probs = torch.empty(n, V, dtype=torch.float64, device=dev)
    tok = torch.empty(n, dtype=torch.int64, device=dev)
    rf = torch.tensor(raw_flags, dtype=torch.int32, device=dev)
    lo_t = torch.tensor(lo, dtype=torch.int32, device=dev)
    hi_t = torch.tensor(hi, dtype=torch.int32, device=dev)
    a.logits, a.ld, a.n_seq, a.V, a.mode = logits.data_ptr(), logits.stride(0), n, V, mode
    a.temperature, a.top_p, a.top_k, a.seed = t, top_p, top_k, 99
    a.raw_flags, a.raw_only_lo, a.raw_only_hi = rf.data_ptr(), lo_t.data_ptr(), hi_t.data_ptr()
    a.out_token, a.out_probs = tok.data_ptr(), probs.data_ptr()
    K.check(K.lib().smer_sample_masked(C.byref(a), K.stream()))
    torch.cuda.synchronize()
    return probs.cpu().numpy(), tok.cpu().numpy()


FLAG_BITS = dict(no_pitch=1, no_duration=2, no_rest=4, no_whole_duration=8, no_eos=16, no_continue=32, no_sep=64)
ONLY = dict(is_density=(242, 251), is_occupation=(262, 271), is_polyphony=(252, 261), is_tensile=(296, 307))
GOLDEN_SETS = {
    "in_sep": dict(no_rest=True, no_sep=True, no_eos=True, no_whole_duration=True),
    "in_continue": dict(no_rest=True, no_sep=True, no_duration=True, no_continue=True, no_eos=True),
    "in_pitch_nwd0": dict(no_rest=True, no_sep=True, no_continue=True, no_eos=True),
    "in_pitch_nwd1": dict(no_rest=True, no_sep=True, no_continue=True, no_whole_duration=True, no_eos=True),
    "in_rest_nwd0": dict(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_eos=True),
    "in_rest_nwd1": dict(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_whole_duration=True, no_eos=True),
    "first_r": dict(no_duration=True), "first_d": dict(is_density=True), "first_o": dict(is_occupation=True),
    "first_p": dict(is_polyphony=True), "first_t": dict(is_tensile=True),
    "free_nwd0": dict(), "free_nwd1": dict(no_whole_duration=True),
}


def test_sampler_distributions_vs_reference_golden(dev, golden_dir):
    """The device sampler's masked / nucleus distributions equal the ones generation.sampling
    produced (tests/golden/sampling.npz was captured from the reference itself)."""
    import os
    ops, K = _ops()
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    logits = torch.from_numpy(g["logits"]).to(dev)
    n = logits.shape[0]
    for name, kw in GOLDEN_SETS.items():
        bits = sum(FLAG_BITS[k] for k in kw if k in FLAG_BITS)
        lo, hi = -1, -1
        for k in kw:
            if k in ONLY:
                lo, hi = ONLY[k]
        for t in (1.0, 0.7):
            probs, _ = _sample_probs(K, logits, [bits] * n, [lo] * n, [hi] * n, K.SAMPLE_MULTINOMIAL, t=t)
            for r in range(n):
                np.testing.assert_allclose(probs[r], g[f"{name}/{r}/t{t}/probs"], rtol=1e-6, atol=1e-300)
        probs, tok = _sample_probs(K, logits, [bits] * n, [lo] * n, [hi] * n, K.SAMPLE_TOP_P, top_p=0.9)
        for r in range(n):
            np.testing.assert_allclose(probs[r], g[f"{name}/{r}/nucleus0.9"], rtol=1e-6, atol=1e-300)
            assert probs[r][tok[r]] > 0
        probs, tok = _sample_probs(K, logits, [bits] * n, [lo] * n, [hi] * n, K.SAMPLE_GREEDY)
        for r in range(n):
            assert tok[r] == int(np.argmax(g[f"{name}/{r}/t1.0/probs"]))


def test_sampler_empirical_histogram(dev, oracle):
    """Draws follow the masked distribution (total-variation distance on 20k draws)."""
    ops, K = _ops()
    rng = np.random.RandomState(0)
    row = (rng.randn(309) * 1.5).astype(np.float32)
    n = 20000
    logits = torch.from_numpy(np.tile(row, (n, 1))).to(dev)
    bits = FLAG_BITS["no_rest"] | FLAG_BITS["no_sep"] | FLAG_BITS["no_eos"]
    probs, tok = _sample_probs(K, logits, [bits] * n, [-1] * n, [-1] * n, K.SAMPLE_MULTINOMIAL)
    q = oracle.masked_probs(row, oracle.Flags(no_rest=True, no_sep=True, no_eos=True))
    hist = np.bincount(tok, minlength=309) / n
    assert 0.5 * np.abs(hist - q).sum() < 0.06
    assert hist[3:146].sum() == 0


# ---------------------------------------------------------------------------------------- decode: small-M linear
@pytest.mark.parametrize("M,N,K,mode", [(5, 512, 512, "plain"), (130, 1536, 512, "ln"), (64, 320, 512, "ln_f32"),
                                        (128, 512, 2048, "resid"), (200, 2048, 512, "ln_relu"), (1, 512, 512, "ln_resid"),
                                        (77, 310, 512, "ln_ln_f32")])
def test_decode_linear_vs_torch(dev, M, N, K, mode):
    """csrc/decode_linear.cu: out = epi(LN?(a) @ w^T + b) for the decode step's few rows, every epilogue / prologue."""
    ops, Kc = _ops()
    g = torch.Generator().manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev).bfloat16() if "resid" in mode else None
    gamma, beta = (torch.rand(K, generator=g) + 0.5).to(dev), (torch.randn(K, generator=g) * 0.1).to(dev)
    ln = (gamma, beta) if mode.startswith("ln") else None
    out = torch.empty(M, N, dtype=torch.float32 if mode.endswith("f32") else torch.bfloat16, device=dev)
    y = torch.empty(M, K, dtype=torch.bfloat16, device=dev) if ln else None
    ln2 = ((torch.rand(K, generator=g) + 0.5).to(dev), (torch.randn(K, generator=g) * 0.1).to(dev)) if "ln_ln" in mode else None
    ops.decode_linear(a, w, out, bias=bias, resid=resid, relu="relu" in mode, ln=ln, ln_out=y, ln2=ln2)
    x = a.float()
    if ln:
        x = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5)
        assert rel(y, x) < 1e-2
        x = x.bfloat16().float()                    # the kernel feeds the tensor cores the bf16-rounded normalised rows
    if ln2:                                         # norm3 of the last layer, then the decoder's final norm
        x = torch.nn.functional.layer_norm(x, (K,), ln2[0], ln2[1], 1e-5).bfloat16().float()
    ref = x @ w.float().t() + bias
    if "relu" in mode:
        ref = ref.relu()
    if resid is not None:
        ref = ref + resid.float()
    assert rel(out, ref) < (2e-3 if mode == "ln_f32" else 1e-2)


def test_gemm_with_reserved_sms_is_unchanged(dev):
    """smer_set_reserved_sms only shortens the persistent grids (data-parallel runs leave SMs to the collectives): the work
    list is re-dealt over fewer CTAs / CTA pairs, the product is bit-identical."""
    ops, Kc = _ops()
    g = torch.Generator().manual_seed(9)
    a = torch.randn(4096 + 40, 512, generator=g).to(dev).bfloat16()
    w = (torch.randn(1536, 512, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(1536, generator=g).to(dev)
    out0, out1 = torch.empty(a.shape[0], 1536, dtype=torch.bfloat16, device=dev), torch.empty(a.shape[0], 1536, dtype=torch.bfloat16, device=dev)
    ops.gemm_nt(a, w, out0, bias=bias)
    assert Kc.lib().smer_set_reserved_sms(16) == 0
    try:
        ops.gemm_nt(a, w, out1, bias=bias)
    finally:
        Kc.lib().smer_set_reserved_sms(0)
    torch.cuda.synchronize()
    assert torch.equal(out0, out1)
    assert rel(out0, a.float() @ w.float().t() + bias) < 1e-2


def test_decode_chain_programmatic_launch_is_ordered(dev):
    """The decode kernels are programmatic dependent launches (csrc/common.cuh smer_launch_pdl): a kernel may start while
    its predecessor drains but touches activations only after griddepcontrol.wait.  A long dependent chain through ONE
    buffer pair, replayed from a CUDA graph, must give exactly what a synchronised launch-by-launch run gives."""
    ops, Kc = _ops()
    g = torch.Generator().manual_seed(5)
    M, Kd = 96, 512
    w = [(torch.randn(Kd, Kd, generator=g) * 0.06).to(dev).bfloat16() for _ in range(3)]
    bias = [torch.randn(Kd, generator=g).to(dev) * 0.1 for _ in range(3)]
    gamma, beta = (torch.rand(Kd, generator=g) + 0.5).to(dev), (torch.randn(Kd, generator=g) * 0.1).to(dev)
    x0 = torch.randn(M, Kd, generator=g).to(dev).bfloat16()
    bufs = [torch.empty_like(x0) for _ in range(2)]

    def chain(sync):
        bufs[0].copy_(x0)
        for i in range(24):
            src, dst = bufs[i % 2], bufs[(i + 1) % 2]
            ops.decode_linear(src, w[i % 3], dst, bias=bias[i % 3], resid=src if i % 2 else None, relu=(i % 2 == 0),
                              ln=(gamma, beta))
            if sync:
                torch.cuda.synchronize()
        return bufs[0]

    want = chain(True).clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    Kc.lib().smer_set_pdl(1)                        # as the small-batch decode step does around its launches
    try:
        with torch.cuda.stream(side):
            chain(False)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                chain(False)
    finally:
        Kc.lib().smer_set_pdl(0)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(bufs[0], want)
