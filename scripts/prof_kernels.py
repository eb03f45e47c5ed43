#!/usr/bin/env python
"""Launches each hot kernel once (after one warm-up launch) at the C2 shapes (B32 x S1024, d512,
H8, ff2048) for `ncu --set full`.  Prints CUDA-event timings of the second launch."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from smer_music_generation_b200 import ops, _capi as K  # noqa: E402

dev = torch.device("cuda:0")
B, L, d, H, ff = 32, 1024, 512, 8, 2048
M = B * L
g = torch.Generator().manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(dev).bfloat16()
x = bf(M, d)
w_qkv, w_o, w1, w2 = bf(3 * d, d), bf(d, d), bf(ff, d), bf(d, ff)
b_qkv, b_o, b1, b2 = (torch.zeros(n, device=dev) for n in (3 * d, d, ff, d))
qkv, proj, h, f = (torch.empty(M, n, dtype=torch.bfloat16, device=dev) for n in (3 * d, d, ff, d))
lens = torch.randint(768, 1025, (B,), generator=g)
pad = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
kv_len = lens.to(torch.int32).to(dev)
o = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B, H, L, device=dev)
dsum = torch.empty(B, H, L, device=dev)
do = bf(M, d)
dqkv = torch.empty(M, 3 * d, dtype=torch.bfloat16, device=dev)
dw = torch.zeros(3 * d, d, device=dev)
gam, bet = torch.ones(d, device=dev), torch.zeros(d, device=dev)
z, y = torch.empty_like(x), torch.empty_like(x)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
dbias = torch.zeros(3 * d, device=dev)


def attn(causal, drop):
    return ops.attn_args(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, B, H, L, L, 64, lse=lse, causal=causal, key_pad=pad,
                         kv_len=kv_len, dropout_p=drop, seed=5, site=1, dout=do, dq=dqkv[:, :d], dk=dqkv[:, d:2 * d],
                         dv=dqkv[:, 2 * d:], dsum=dsum)


cases = [
    ("gemm_qkv      NT N1536 K512 bias", 2.0 * M * d * 3 * d, lambda: ops.gemm_nt(x, w_qkv, qkv, bias=b_qkv)),
    ("gemm_outproj  NT N512 K512 bias", 2.0 * M * d * d, lambda: ops.gemm_nt(o, w_o, proj, bias=b_o)),
    ("gemm_ffn1     NT N2048 K512 relu+dropout", 2.0 * M * d * ff, lambda: ops.gemm_nt(x, w1, h, bias=b1, flags=K.EPI_RELU, dropout_p=0.1, seed=3, site=2)),
    ("gemm_ffn2     NT N512 K2048 bias", 2.0 * M * d * ff, lambda: ops.gemm_nt(h, w2, f, bias=b2)),
    ("gemm_dx_ffn2  dY[M,512].W2 -> [M,2048] gate", 2.0 * M * d * ff, lambda: ops.gemm_dx(f, w2, h, resid=h, flags=K.EPI_GATE, dropout_p=0.1)),
    ("gemm_dx_qkv   dY[M,1536].Wqkv -> [M,512] +resid", 2.0 * M * d * 3 * d, lambda: ops.gemm_dx(dqkv, w_qkv, proj, resid=x)),
    ("gemm_dw_qkv   dY^T X split-K", 2.0 * M * d * 3 * d, lambda: ops.gemm_dw(dqkv, x, dw)),
    ("attn_fwd enc  full+pad dropout", 4.0 * B * H * L * L * 64, lambda: ops.attn_fwd(attn(False, 0.1))),
    ("attn_fwd dec  causal+pad dropout", 2.0 * B * H * L * L * 64, lambda: ops.attn_fwd(attn(True, 0.1))),
    ("attn_bwd enc  full+pad dropout", 8.0 * B * H * L * L * 64, lambda: ops.attn_bwd(attn(False, 0.1))),
    ("attn_bwd dec  causal+pad dropout", 4.0 * B * H * L * L * 64, lambda: ops.attn_bwd(attn(True, 0.1))),
    ("layernorm_fwd resid+dropout", 0.0, lambda: ops.layernorm_fwd(proj, x, gam, bet, z, y, mean, rstd, dropout_p=0.1, seed=1, site=4)),
    ("layernorm_bwd dropout", 0.0, lambda: ops.layernorm_bwd(do, z, mean, rstd, gam, y, proj, dg, db, dropout_p=0.1, seed=1, site=4)),
    ("colsum [M,1536]", 0.0, lambda: ops.colsum(dqkv, dbias)),
]
ops.gemm_nt(x, w_qkv, qkv, bias=b_qkv)
sel = sys.argv[1] if len(sys.argv) > 1 else ""
if "attn" in sel:
    # realistic attention inputs: unit-variance q / k / v (scores of standard deviation 8 before the 1/8 scale), as a
    # trained or freshly initialised model produces them -- the x.W^T above has a standard deviation of 5.6, which makes
    # every softmax row one-hot and the running maximum jump by more than 2^8 on most tiles
    qkv.copy_(torch.randn(M, 3 * d, generator=g).to(dev).bfloat16())
ops.attn_fwd(attn(False, 0.0))
torch.cuda.synchronize()
for name, flops, fn in cases:
    if sel and sel not in name:
        continue
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(2e6))      # ~1 ms head start so the launch is queued before the event fires
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name:48s} {ms * 1e3:9.1f} us" + (f"  {flops / ms / 1e9:8.1f} TFLOP/s" if flops else ""), flush=True)
