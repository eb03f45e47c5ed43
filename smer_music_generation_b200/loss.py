"""Class-weighted softmax cross-entropy of the reference's training step in ONE pass.

train.py:555-642 builds 7-12 nn.CrossEntropyLoss(weight=one-hot range, ignore_index=0,
reduction='none') modules; train.py:726-780 evaluates each on the same logits, divides each
sum by sum(ce_weight_all[target]) and adds them.  Algebraically
    loss = sum_i W[y_i] * (lse(x_i) - x_i[y_i]) / sum_i C[y_i]
with W the sum of the active range weights and C = ce_weight_all.  The
kernels (csrc/xent.cu) also return the per-category numerators the reference logs.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import nn

from . import _capi as K
from . import ops

# name, first id, last id (inclusive): the reference's criteria in train.py:555-642 order
CATEGORIES = (
    ("meta", 1, 1), ("time_signature", 7, 10), ("program", 18, 145), ("tempo", 11, 17),
    ("structure", 3, 6), ("pitch", 146, 233), ("duration", 234, 241),
    ("tensile", 296, 307), ("key", 272, 295), ("density", 242, 251),
    ("occupation", 262, 271), ("polyphony", 252, 261),
)
CONTROL_NAMES = ("tensile", "key", "density", "occupation", "polyphony")


def loss_tables(vocab_size: int = 309, eos_weight: float = 1.0,
                control_list: Sequence[str] = CONTROL_NAMES) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(W, C, category) on the CPU: W[v] numerator weight, C[v] normaliser weight
    (train.py:646-650 ce_weight_all), category[v] index into CATEGORIES or -1."""
    W = torch.zeros(vocab_size)
    cat = torch.full((vocab_size,), -1, dtype=torch.int32)
    for k, (name, lo, hi) in enumerate(CATEGORIES):
        if name in CONTROL_NAMES and name not in control_list:
            continue
        W[lo:hi + 1] += 1.0
        cat[lo:hi + 1] = k
    W[1] = eos_weight
    C = torch.ones(vocab_size)
    C[0] = 0.0
    C[2] = 0.0
    C[vocab_size - 1] = 0.0
    C[1] = eos_weight
    return W, C, cat


def _rows(logits: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """View (..., V) logits as rows with one uniform pitch, without copying when possible."""
    V = logits.shape[-1]
    if logits.dim() == 2 and logits.stride(1) == 1:
        return logits, logits.shape[0], logits.stride(0)
    if logits.dim() == 3 and logits.stride(2) == 1 and logits.stride(0) == logits.shape[1] * logits.stride(1):
        B, T = logits.shape[:2]
        flat = logits.as_strided((B * T, V), (logits.stride(1), 1), logits.storage_offset())
        return flat, B * T, logits.stride(1)
    c = logits.reshape(-1, V).contiguous()
    return c, c.shape[0], V


class _XentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, W, C, cat, ncat, grad_scale_holder):
        flat, rows, ld = _rows(logits)
        if flat.dtype != torch.float32:
            raise TypeError("smer loss expects fp32 logits (the fc GEMM writes fp32)")
        V = logits.shape[-1]
        tg = targets.reshape(-1).to(torch.int64).contiguous()
        lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
        sums = torch.empty(K.XENT_MAX_SUMS, dtype=torch.float64, device=logits.device)
        ops.xent_fwd(flat, tg, W, C, cat, ncat, lse, sums, V)
        ctx.save_for_backward(flat, tg, W, lse, sums)
        ctx.shape, ctx.V, ctx.holder = logits.shape, V, grad_scale_holder
        denom = sums[1]
        loss = (sums[0] / denom).to(torch.float32)
        parts = (sums[2:2 + ncat] / denom).to(torch.float32)
        ctx.mark_non_differentiable(parts)
        return loss, parts, denom.to(torch.float32)

    @staticmethod
    def backward(ctx, gloss, gparts, gdenom):
        flat, tg, W, lse, sums = ctx.saved_tensors
        V = ctx.V
        dl = torch.empty(flat.shape[0], V, dtype=torch.float32, device=flat.device)
        g = gloss.reshape(1).to(torch.float32).contiguous()      # upstream scale stays on the device
        ops.xent_bwd(flat, tg, W, lse, sums, dl, V, 1.0, g)
        return dl.view(ctx.shape), None, None, None, None, None, None


class SmerLoss(nn.Module):
    """loss, parts, denom = SmerLoss(...)(logits (B,T,V) | (N,V), targets).  `parts` are the
    per-category terms the reference logs (loss == parts.sum() when every category is active)."""

    def __init__(self, vocab_size: int = 309, eos_weight: float = 1.0, control_list: Sequence[str] = CONTROL_NAMES):
        super().__init__()
        W, C, cat = loss_tables(vocab_size, eos_weight, control_list)
        self.register_buffer("W", W, persistent=False)
        self.register_buffer("C", C, persistent=False)
        self.register_buffer("cat", cat, persistent=False)
        self.ncat = len(CATEGORIES)
        self._holder = {"unit_grad": False}

    def set_eos_weight(self, w: float) -> None:      # train.py:670-673 switches 0.8 -> 1 after epoch 1
        self.W[1] = w
        self.C[1] = w

    def forward(self, logits, targets):
        K.require_cuda_device()
        return _XentFn.apply(logits, targets, self.W, self.C, self.cat, self.ncat, self._holder)


# ---------------------------------------------------------------------------------------
# Token accuracy by target class: the metric loop of train.py:988-1034 on the device.
# ---------------------------------------------------------------------------------------
# class name, first id, last id (inclusive) of vocab.token_class_ranges (vocab.py:159-300) for the
# default vocabulary (control list key/tensile/density/polyphony/occupation); ids 0 (pad) and 2 (m_0)
# have no class.  Pinned against the real WordVocab by tests/golden/metrics_small.pt.
TOKEN_CLASSES = (
    ("density", 242, 251), ("duration", 234, 241), ("eos", 1, 1), ("key", 272, 295), ("occupation", 262, 271),
    ("pitch", 146, 233), ("polyphony", 252, 261), ("program", 18, 145), ("structure", 3, 6), ("tempo", 11, 17),
    ("tensile", 296, 307), ("time_signature", 7, 10), ("unk", 308, 308),
)


def token_class_table(vocab_size: int = 309) -> torch.Tensor:
    """class_of[v]: index into TOKEN_CLASSES or -1 (int32, CPU)."""
    t = torch.full((vocab_size,), -1, dtype=torch.int32)
    for k, (_, lo, hi) in enumerate(TOKEN_CLASSES):
        t[lo:min(hi, vocab_size - 1) + 1] = k
    return t


class SmerAccuracy(nn.Module):
    """acc = SmerAccuracy()(logits (B,T,V) | (N,V), targets): the dict `accuracy(outputs, targets, vocab)[0]`
    of train.py:988-1034 returns -- per target class and 'total', classes without tokens stay 0 -- from one
    kernel launch and one small D2H instead of a per-token Python loop with an `.item()` per token.
    `update()` / `compute()` accumulate over several batches (one D2H at `compute()`);
    `first_sample_argmax` holds the predicted ids of batch element 0 (the reference's `generated_output`)."""

    def __init__(self, vocab_size: int = 309):
        super().__init__()
        self.names = [n for n, _, _ in TOKEN_CLASSES]
        self.register_buffer("class_of", token_class_table(vocab_size), persistent=False)
        self.register_buffer("counts", torch.zeros(2 * (len(self.names) + 1), dtype=torch.int64), persistent=False)
        self.first_sample_argmax: Optional[torch.Tensor] = None

    def reset(self) -> None:
        self.counts.zero_()

    @torch.no_grad()
    def update(self, logits: torch.Tensor, targets: torch.Tensor) -> None:
        K.require_cuda_device()
        lg = logits.detach()
        if lg.dtype != torch.float32:
            lg = lg.float()
        flat, rows, _ = _rows(lg)
        tg = targets.reshape(-1).contiguous()
        if tg.dtype != torch.int64:
            tg = tg.long()
        if tg.numel() != rows:
            raise ValueError("targets do not match the logits rows")
        am = torch.empty(rows, dtype=torch.int64, device=lg.device)
        ops.token_accuracy(flat, tg, self.class_of, len(self.names), self.counts, am)
        per_sample = rows // logits.shape[0] if logits.dim() == 3 else rows
        self.first_sample_argmax = am[:per_sample]

    def compute(self) -> dict:
        c = self.counts.cpu().tolist()
        n = len(self.names)
        out = {}
        for k, name in enumerate(self.names + ["total"]):
            seen = c[n + 1 + k]
            out[name] = c[k] / seen if seen else 0
        return out

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> dict:
        self.reset()
        self.update(logits, targets)
        return self.compute()
