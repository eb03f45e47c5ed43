"""CPU oracle for the SMER transformer compute path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (plain torch fp32 / numpy float64, no custom kernels)
of the reference's algorithm on the hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The product package ``smer_music_generation_b200`` never does.

Parity status: PINNED.  The reference ships no tests or golden vectors,
so ``oracle/make_golden.py`` imports the real reference modules from /root/reference
in the build container, runs them on seeded inputs and commits the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those vectors.

Where the arithmetic lives: the reference delegates to third-party PyTorch
(``nn.MultiheadAttention`` -> ``F.multi_head_attention_forward`` need_weights branch,
``nn.LayerNorm``, ``nn.Linear``, ``nn.CrossEntropyLoss``, ``torch.optim.Adam``); torch is
unpinned in the reference's requirements.txt.  The restatement below spells that
published algorithm out with matmul/softmax primitives and is keyed directly on the
reference's ``state_dict`` names.

Every function cites the reference file:line it follows (paths relative to the
reference repo root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# A10 -- vocabulary constants (vocab.py:114-310, mode 0 / SMER), verified by make_golden.py
# --------------------------------------------------------------------------------------
V = 309
PAD, EOS, M0 = 0, 1, 2
BAR = 3
TRACK0 = 4                     # track_0..track_2 = 4..6
STRUCTURE = range(3, 7)
TIME_SIG = range(7, 11)
TEMPO = range(11, 18)
PROGRAM = range(18, 146)
PITCH = range(146, 234)
DURATION_ONLY = range(234, 239)   # whole, half, quarter, eighth, sixteenth
WHOLE = 234
REST, SEP, CONTINUE = 239, 240, 241
DENSITY = range(242, 252)
POLYPHONY = range(252, 262)
OCCUPATION = range(262, 272)
KEY = range(272, 296)
TENSILE = range(296, 308)
UNK = 308

LOSS_CATEGORIES = (
    # name, first id, last id (inclusive) -- train.py:555-642
    ("meta", 1, 1), ("time_signature", 7, 10), ("program", 18, 145), ("tempo", 11, 17),
    ("structure", 3, 6), ("pitch", 146, 233), ("duration", 234, 241),
    ("tensile", 296, 307), ("key", 272, 295), ("density", 242, 251),
    ("occupation", 262, 271), ("polyphony", 252, 261),
)


# --------------------------------------------------------------------------------------
# A1 -- sinusoidal positional table (model.py:110-121)
# --------------------------------------------------------------------------------------
def positional_table(max_len: int, d_model: int) -> torch.Tensor:
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(1)          # (max_len, 1, d) -- the buffer saved as `pos_enc.pe`


# --------------------------------------------------------------------------------------
# A11 -- the causal ("nopeek") additive mask (generation.py:193-206, train.py:1356-1369)
# --------------------------------------------------------------------------------------
def nopeek_mask(length: int) -> torch.Tensor:
    m = torch.zeros(length, length)
    m.masked_fill_(torch.triu(torch.ones(length, length, dtype=torch.bool), diagonal=1), float("-inf"))
    return m


# --------------------------------------------------------------------------------------
# A6 -- multi-head attention, need_weights branch of F.multi_head_attention_forward
# --------------------------------------------------------------------------------------
def _mha(sd: Dict[str, torch.Tensor], prefix: str, q_in: torch.Tensor, kv_in: torch.Tensor,
         nhead: int, attn_mask: Optional[torch.Tensor], key_padding_mask: Optional[torch.Tensor],
         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """q_in (B,Lq,d), kv_in (B,Lk,d) -> (out (B,Lq,d), head-averaged probs (B,Lq,Lk)).

    Packed in-projection rows are Q|K|V (transformer.py:360,423-424 construct
    nn.MultiheadAttention; cross-attention uses rows [0:d] on the query stream and
    [d:3d] on memory).  q is scaled by 1/sqrt(dh) before the score product; masks
    are additive (-inf); softmax; P.V; out-projection; P averaged over heads.
    """
    W, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    Wo, bo = sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"]
    B, Lq, d = q_in.shape
    Lk = kv_in.shape[1]
    dh = d // nhead
    q = F.linear(q_in, W[:d], b[:d])
    k = F.linear(kv_in, W[d:2 * d], b[d:2 * d])
    v = F.linear(kv_in, W[2 * d:], b[2 * d:])
    q = q.view(B, Lq, nhead, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, Lk, nhead, dh).transpose(1, 2)
    v = v.view(B, Lk, nhead, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)                             # (B,H,Lq,Lk)
    if attn_mask is not None:
        s = s + attn_mask
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, Wo, bo), p.mean(dim=1)


def _ln(sd, prefix, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def _ffn(sd, prefix, x):
    # transformer.py:393 / 467: linear2(dropout(relu(linear1(x)))) -- eval mode, dropout = id
    return F.linear(F.relu(F.linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])),
                    sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])


def count_layers(sd: Dict[str, torch.Tensor], which: str) -> int:
    n = 0
    while f"transformer.{which}.layers.{n}.norm1.weight" in sd:
        n += 1
    return n


# --------------------------------------------------------------------------------------
# A2-A5 -- ScoreTransformer.forward in eval mode (model.py:85-106, transformer.py:69-127,
#          258-277, 303-335, 378-396, 444-470).  Batch-first internally; the reference is
#          sequence-first, which only permutes storage.
# --------------------------------------------------------------------------------------
def encode(sd, src, nhead, src_key_padding_mask=None):
    d = sd["embedding.weight"].shape[1]
    pe = sd["pos_enc.pe"][:, 0, :]
    x = sd["embedding.weight"][src] * math.sqrt(d) + pe[: src.shape[1]]
    for i in range(count_layers(sd, "encoder")):
        p = f"transformer.encoder.layers.{i}."
        a, _ = _mha(sd, p + "self_attn.", x, x, nhead, None, src_key_padding_mask)
        x = _ln(sd, p + "norm1.", x + a)
        x = _ln(sd, p + "norm2.", x + _ffn(sd, p, x))
    return _ln(sd, "transformer.encoder.norm.", x)


def decode(sd, tgt, memory, nhead, tgt_mask2d=None, tgt_key_padding_mask=None,
           memory_key_padding_mask=None):
    d = sd["embedding.weight"].shape[1]
    pe = sd["pos_enc.pe"][:, 0, :]
    y = sd["embedding.weight"][tgt] * math.sqrt(d) + pe[: tgt.shape[1]]
    weights = []
    for i in range(count_layers(sd, "decoder")):
        p = f"transformer.decoder.layers.{i}."
        a, _ = _mha(sd, p + "self_attn.", y, y, nhead, tgt_mask2d, tgt_key_padding_mask)
        y = _ln(sd, p + "norm1.", y + a)
        a, w = _mha(sd, p + "multihead_attn.", y, memory, nhead, None, memory_key_padding_mask)
        weights.append(w)
        y = _ln(sd, p + "norm2.", y + a)
        y = _ln(sd, p + "norm3.", y + _ffn(sd, p, y))
    y = _ln(sd, "transformer.decoder.norm.", y)
    logits = F.linear(y, sd["fc.weight"], sd["fc.bias"])
    return logits, torch.stack(weights, dim=1)              # (B,T,V), (B,Ld,T,S)


def score_transformer_forward(sd, src, tgt, nhead, src_key_padding_mask=None,
                              tgt_key_padding_mask=None, memory_key_padding_mask=None,
                              tgt_mask=None):
    """model.py:85-106.  `tgt_mask` is the (B,T,T) tensor the callers pass; only [0] is used."""
    mem = encode(sd, src, nhead, src_key_padding_mask)
    m2 = None if tgt_mask is None else tgt_mask[0]
    return decode(sd, tgt, mem, nhead, m2, tgt_key_padding_mask, memory_key_padding_mask)


# --------------------------------------------------------------------------------------
# A9 -- training loss (train.py:555-642 definition, 726-780 use)
# --------------------------------------------------------------------------------------
def loss_weights(eos_weight: float = 1.0,
                 control_list: Sequence[str] = ("key", "tensile", "density", "polyphony", "occupation"),
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (W, C): W = sum of the per-category one-hot-range weight vectors that are
    active, C = ce_weight_all (the normaliser weights)."""
    W = torch.zeros(V)
    for name, lo, hi in LOSS_CATEGORIES:
        if name in ("tensile", "key", "density", "occupation", "polyphony") and name not in control_list:
            continue
        W[lo:hi + 1] += 1.0
    W[EOS] = eos_weight
    C = torch.ones(V)
    C[PAD] = 0.0
    C[M0] = 0.0
    C[UNK] = 0.0
    C[EOS] = eos_weight
    return W, C


def smer_loss(logits: torch.Tensor, tgt_out: torch.Tensor, W: torch.Tensor, C: torch.Tensor):
    """logits (N,V) fp32, tgt_out (N,) int64.  Returns (loss, per-category sums (12,), denom).

    Each nn.CrossEntropyLoss(weight=w_k, ignore_index=0, reduction='none') row is
    w_k[y]*(lse(x)-x[y]); the reference sums each, divides by sum(C[y]) and adds them.
    """
    lse = torch.logsumexp(logits.float(), dim=-1)
    nll = lse - logits.float().gather(1, tgt_out[:, None])[:, 0]
    nll = torch.where(tgt_out == PAD, torch.zeros_like(nll), nll)
    denom = C[tgt_out].sum()
    cats = []
    for _, lo, hi in LOSS_CATEGORIES:
        sel = (tgt_out >= lo) & (tgt_out <= hi)
        cats.append((nll * sel * W[tgt_out]).sum() / denom)
    loss = (nll * W[tgt_out]).sum() / denom
    return loss, torch.stack(cats), denom


# --------------------------------------------------------------------------------------
# A13 -- sampling arithmetic (generation.py:11-95)
# --------------------------------------------------------------------------------------
@dataclass
class Flags:
    no_pitch: bool = False
    no_duration: bool = False
    no_rest: bool = False
    no_whole_duration: bool = False
    no_eos: bool = False
    no_continue: bool = False
    no_sep: bool = False
    is_density: bool = False
    is_polyphony: bool = False
    is_occupation: bool = False
    is_tensile: bool = False
    no_control: bool = False          # a no-op in the reference (generation.py:85-87)


def allowed_mask(f: Flags) -> np.ndarray:
    """Boolean (V,) -- True where the logit survives (is not overwritten with -100).
    Order of application follows generation.py:46-87; the result is order-independent."""
    ok = np.ones(V, dtype=bool)
    if f.no_pitch:
        ok[146:234] = False
    if f.no_duration:
        ok[234:239] = False
    if f.no_continue:
        ok[CONTINUE] = False
    if f.no_rest:
        ok[REST] = False
    if f.no_sep:
        ok[SEP] = False
    if f.no_whole_duration:
        ok[WHOLE] = False
    if f.no_eos:
        ok[EOS] = False
    for on, rng in ((f.is_density, DENSITY), (f.is_occupation, OCCUPATION),
                    (f.is_polyphony, POLYPHONY), (f.is_tensile, TENSILE)):
        if on:
            keep = np.zeros(V, dtype=bool)
            keep[rng.start:rng.stop] = True
            ok &= keep
    ok[3:146] = False                # generation.py:82-84, always
    return ok


def masked_probs(logit_row: np.ndarray, f: Flags, t: float = 1.0) -> np.ndarray:
    """float64 softmax-with-temperature of the -100-masked logits, no max shift
    (generation.py:28-30, 43-89)."""
    l = np.where(allowed_mask(f), logit_row.astype(np.float64), -100.0)
    e = np.exp(l / t)
    return e / np.sum(e)


def nucleus_probs(probs: np.ndarray, p: float) -> np.ndarray:
    """Distribution np.random.choice is called with in `nucleus` (generation.py:11-25),
    scattered back to vocabulary order."""
    probs = probs / (np.sum(probs) + 1e-5)
    order = np.argsort(probs)[::-1]
    cs = np.cumsum(probs[order])
    after = cs > p
    last = (np.where(after)[0][0] + 1) if after.any() else len(order)
    out = np.zeros_like(probs)
    out[order[:last]] = probs[order[:last]]
    return out / out.sum()


def resample_closed_form(q: np.ndarray, accept: Optional[np.ndarray]) -> np.ndarray:
    """Distribution of the rejection loop at generation.py:553-562 etc.: up to 12 i.i.d.
    draws, the first of d0..d10 inside A, else d11."""
    if accept is None:
        return q
    r = q[~accept].sum()
    out = np.where(accept, q * ((1 - r ** 12) / (1 - r) if r < 1 else 12.0), q * r ** 11)
    return out / out.sum()


# --------------------------------------------------------------------------------------
# A14 -- grammar state machine of generation_all (generation.py:538-683)
# --------------------------------------------------------------------------------------
ST_FREE, ST_SEP, ST_CONT, ST_PITCH, ST_REST = 0, 1, 2, 3, 4


@dataclass
class SpanState:
    in_pitch: bool = False
    in_rest: bool = False
    in_sep: bool = False
    in_continue: bool = False

    def flags(self, span_len: int, target: str, nwd: bool) -> Tuple[Flags, Optional[np.ndarray]]:
        """(flag set, accept set) used for the next draw; priority at generation.py:547-652."""
        acc = np.zeros(V, dtype=bool)
        if self.in_sep:
            f = Flags(no_rest=True, no_sep=True, no_eos=True, no_whole_duration=True, no_control=True)
            acc[:] = True
            acc[[REST, EOS, WHOLE]] = False
            return f, acc
        if self.in_continue:
            f = Flags(no_rest=True, no_sep=True, no_duration=True, no_continue=True, no_eos=True, no_control=True)
            acc[146:234] = True
            return f, acc
        if self.in_pitch:
            f = Flags(no_rest=True, no_sep=True, no_continue=True, no_whole_duration=nwd, no_eos=True, no_control=True)
            acc[146:239] = True
            return f, acc
        if self.in_rest:
            f = Flags(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_whole_duration=nwd,
                      no_eos=True, no_control=True)
            acc[234:239] = True
            return f, acc
        if span_len == 1:
            if target == "d":
                return Flags(is_density=True), None
            if target == "o":
                return Flags(is_occupation=True), None
            if target == "p":
                return Flags(is_polyphony=True), None
            if target == "t":
                return Flags(is_tensile=True), None
            acc[:] = True
            acc[234:239] = False
            return Flags(no_duration=True, no_control=True), acc
        return Flags(no_whole_duration=nwd, no_control=True), None

    def update(self, idx: int) -> None:
        """generation.py:654-671 (all five tests are applied in order)."""
        if idx == CONTINUE:
            self.in_continue, self.in_sep = True, False
        if 146 <= idx <= 233:
            self.in_pitch, self.in_sep, self.in_continue = True, False, False
        if 234 <= idx <= 238:
            self.in_rest, self.in_pitch = False, False
        if idx == SEP:
            self.in_sep = True
        if idx == REST:
            self.in_rest = True


def mask_targets(n_bars: int, tracks: Sequence[int], n_tracks: int) -> List[str]:
    """generation.py:485-492."""
    out: List[str] = []
    for _ in range(n_bars):
        for tr in tracks:
            out.extend(["r", "d", "o", "p"])
            if tr == n_tracks - 1:
                out.append("t")
    return out


def mask_bar_and_track_ids(ids: Sequence[int], mask_tracks: Sequence[int], mask_bars: Sequence[int],
                           n_tracks: int) -> np.ndarray:
    """Id-level restatement of generation.mask_bar_and_track (generation.py:248-341) for
    control_mode-2 layouts: for each selected (bar, track) the content span and each trailing
    control token (3 track controls + a tensile token when the track is last) become one m_0."""
    ids = list(ids)
    marks = [i for i, t in enumerate(ids) if t == BAR or TRACK0 <= t < TRACK0 + n_tracks]
    marks.append(len(ids))
    bars: List[List[Tuple[int, int]]] = []
    cur: List[int] = []
    for i, pos in enumerate(marks[1:]):
        if i % (n_tracks + 1) == 0:
            cur = [pos]
        else:
            cur.append(pos)
            if i % (n_tracks + 1) == n_tracks:
                bars.append([(cur[j] + 1, cur[j + 1]) for j in range(len(cur) - 1)])
    pairs: List[Tuple[int, int]] = []
    for b in mask_bars:
        for tpos, (ts, te) in enumerate(bars[b]):
            if tpos in mask_tracks:
                tensile_end = 1 if ids[te - 1] in TENSILE else 0
                tok_start = ts + 3
                tok_end = te - 3 - tensile_end
                pairs.append((tok_start, tok_end))
                for i in range(3 + tensile_end):
                    pairs.append((tok_end + i, tok_end + 1 + i))
    out = list(ids)
    for a, b in pairs[::-1]:
        del out[a:b]
        out.insert(a, M0)
    return np.asarray(out, dtype=np.int64)


@dataclass
class DecodeTrace:
    tokens: List[int] = field(default_factory=list)         # final tgt_inp (decoder input stream)
    step_logits: List[np.ndarray] = field(default_factory=list)
    step_probs: List[np.ndarray] = field(default_factory=list)
    step_prefix_len: List[int] = field(default_factory=list)
    generated: int = 0


def infill_decode(sd, src_ids: np.ndarray, targets: Sequence[str], nhead: int,
                  all_controls: Sequence[int] = tuple(range(242, 308)), nwd: bool = False,
                  mode: str = "greedy", t: float = 1.0, top_p: Optional[float] = None,
                  rng: Optional[np.random.Generator] = None, max_span: int = 100,
                  keep_trace: bool = False) -> DecodeTrace:
    """Uncached decode loop of generation_all (generation.py:523-687): every token re-runs the
    decoder over the whole prefix (the encoder output is recomputed too in the reference; it is
    input-invariant, so it is hoisted here -- same values).  mode="greedy": argmax of the
    masked distribution, lowest id on ties, no rejection loop.  mode="sample": the rejection
    loop with `rng`.
    """
    ctrl = set(int(c) for c in all_controls)
    src = torch.as_tensor(src_ids, dtype=torch.long)[None]
    tr = DecodeTrace()
    with torch.no_grad():
        mem = encode(sd, src, nhead)
        tgt_inp: List[int] = []
        for target in targets:
            span = [M0]
            st = SpanState()
            while span[-1] != EOS and len(span) < max_span:
                seq = tgt_inp + span
                tgt = torch.as_tensor(seq, dtype=torch.long)[None]
                logits, _ = decode(sd, tgt, mem, nhead, nopeek_mask(len(seq)))
                row = logits[0, -1].numpy()
                f, acc = st.flags(len(span), target, nwd)
                q = masked_probs(row, f, t)
                if top_p is not None:
                    q = nucleus_probs(q, top_p)
                if mode == "greedy":
                    idx = int(np.argmax(q))
                else:
                    idx = int(rng.choice(V, p=q))
                    n = 0
                    while acc is not None and not acc[idx]:
                        idx = int(rng.choice(V, p=q))
                        n += 1
                        if n > 10:
                            break
                if keep_trace:
                    tr.step_logits.append(row.copy())
                    tr.step_probs.append(q)
                    tr.step_prefix_len.append(len(seq))
                st.update(idx)
                span.append(idx)
                tr.generated += 1
                if idx in ctrl:
                    span.append(EOS)
                    tr.generated += 1
            tgt_inp.extend(span[:-1])            # generation.py:686 (drops <eos> or the capped token)
    tr.tokens = tgt_inp
    return tr


# --------------------------------------------------------------------------------------
# Training-step arithmetic (train.py:722-786) on the oracle forward, via torch autograd.
# --------------------------------------------------------------------------------------
# The 13 target classes of vocab.token_class_ranges (vocab.py:159-300, control list of train.py), sorted by name;
# ids 0 (pad) and 2 (m_0) have none.  Pinned by tests/golden/metrics_small.pt.
TOKEN_CLASSES = ("density", "duration", "eos", "key", "occupation", "pitch", "polyphony", "program", "structure",
                 "tempo", "tensile", "time_signature", "unk")


def token_accuracy(logits: torch.Tensor, targets: torch.Tensor, class_of: np.ndarray, classes=TOKEN_CLASSES):
    """train.py:988-1034 `accuracy()`: argmax (first maximum) of every row against the target, pad targets
    skipped, accumulated per target class and in total; classes never seen keep the value 0."""
    lg = logits.reshape(-1, logits.shape[-1]).double().numpy()
    y = targets.reshape(-1).numpy()
    am = lg.argmax(axis=1)                                     # numpy, like torch: first occurrence of the maximum
    correct = {c: 0 for c in classes}
    seen = {c: 0 for c in classes}
    correct["total"] = seen["total"] = 0
    for a, t in zip(am, y):
        if t == 0:
            continue
        k = int(class_of[t])
        if k >= 0:
            correct[classes[k]] += int(a == t)
            seen[classes[k]] += 1
        correct["total"] += int(a == t)
        seen["total"] += 1
    acc = {c: (correct[c] / seen[c] if seen[c] else 0) for c in correct}
    return acc, correct, seen, am.reshape(targets.shape)


def train_step_grads(sd: Dict[str, torch.Tensor], src, tgt_in, tgt_out, src_pad, tgt_pad, nhead, W, C):
    """Returns (loss, {name: grad}) for one batch in eval-mode arithmetic (dropout 0)."""
    leaf = {k: v.detach().clone().requires_grad_(k != "pos_enc.pe") for k, v in sd.items()}
    T = tgt_in.shape[1]
    logits, _ = score_transformer_forward(leaf, src, tgt_in, nhead, src_pad, tgt_pad, src_pad,
                                          nopeek_mask(T)[None])
    loss, cats, denom = smer_loss(logits.reshape(-1, logits.shape[-1]), tgt_out.reshape(-1), W, C)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items() if v.grad is not None}, logits.detach(), cats.detach()


def adam_step(p, g, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (train.py:264): no weight decay, no amsgrad."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# Synthetic SMER pieces -- used by tests and bench on both arms.
# --------------------------------------------------------------------------------------
def synth_piece(seed: int = 0, n_bars: int = 16, n_tracks: int = 3, events_per_track_bar: int = 4,
                time_sig: int = 7) -> List[int]:
    """Control-mode-2 layout (encode.py:559-804): header [ts, tempo, key, d*N, o*N, y*N, i*N];
    per bar `bar, s_*`; per track `track_n, d,o,y, <notes>, d,o,y`; trailing s_* after the
    last track of each bar."""
    import random
    r = random.Random(seed)
    ids = [time_sig, 14, 272]          # time signature ids 7..10 = '4/4', '3/4', '2/4', '6/8' (vocab.py:26)
    ids += [DENSITY.start + r.randrange(10) for _ in range(n_tracks)]
    ids += [OCCUPATION.start + r.randrange(10) for _ in range(n_tracks)]
    ids += [POLYPHONY.start + r.randrange(10) for _ in range(n_tracks)]
    ids += [18 + p for p in (0, 32, 48)[:n_tracks]]
    for _ in range(n_bars):
        ids += [BAR, TENSILE.start + r.randrange(12)]
        for tr in range(n_tracks):
            c = [DENSITY.start + r.randrange(10), OCCUPATION.start + r.randrange(10),
                 POLYPHONY.start + r.randrange(10)]
            ids += [TRACK0 + tr] + c
            for _ in range(events_per_track_bar):
                if r.random() < 0.2:
                    ids += [REST, 236]
                else:
                    ids += [146 + (r.randint(48, 84) - 21), 236]
            ids += c
            if tr == n_tracks - 1:
                ids.append(TENSILE.start + r.randrange(12))
    return ids


def synth_batch(B: int, S: int, T: int, seed: int = 1234, min_frac: float = 0.75):
    """Fixed-shape training batch with suffix padding: random valid lengths U[min_frac*L, L]
    (SURVEY.md §8d); ids drawn from the SMER id ranges so every loss category is hit."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(3, V - 1, (B, S), generator=g)
    tgt = torch.randint(3, V - 1, (B, T + 1), generator=g)
    sl = torch.randint(int(min_frac * S), S + 1, (B,), generator=g)
    tl = torch.randint(max(2, int(min_frac * T)), T + 1, (B,), generator=g)
    src_pad = torch.arange(S)[None] >= sl[:, None]
    tgt_pad = torch.arange(T)[None] >= tl[:, None]
    src = src.masked_fill(src_pad, PAD)
    tgt_in = tgt[:, :T].clone()
    tgt_in[:, 0] = M0
    tgt_out = tgt[:, 1:].clone()
    last = (tl - 1).clamp(min=0)
    tgt_out[torch.arange(B), last] = EOS
    tgt_in = tgt_in.masked_fill(tgt_pad, PAD)
    tgt_out = tgt_out.masked_fill(tgt_pad, PAD)
    return src, tgt_in, tgt_out, src_pad, tgt_pad


def random_state_dict(d=512, nhead=8, le=4, ld=4, ff=2048, max_len=2400, seed=0, std=None):
    """Random weights in the reference's state_dict layout, xavier-normal on
    matrices as train.py:261-263 does."""
    g = torch.Generator().manual_seed(seed)

    def mat(o, i):
        s = math.sqrt(2.0 / (o + i)) if std is None else std
        return torch.randn(o, i, generator=g) * s

    def vec(n, s=0.02):
        return torch.randn(n, generator=g) * s

    sd = {"embedding.weight": mat(V, d), "pos_enc.pe": positional_table(max_len, d)}

    def attn(p):
        sd[p + "in_proj_weight"] = mat(3 * d, d)
        sd[p + "in_proj_bias"] = vec(3 * d)
        sd[p + "out_proj.weight"] = mat(d, d)
        sd[p + "out_proj.bias"] = vec(d)

    def ffn_ln(p, n_ln):
        sd[p + "linear1.weight"] = mat(ff, d)
        sd[p + "linear1.bias"] = vec(ff)
        sd[p + "linear2.weight"] = mat(d, ff)
        sd[p + "linear2.bias"] = vec(d)
        for j in range(1, n_ln + 1):
            sd[p + f"norm{j}.weight"] = 1.0 + vec(d)
            sd[p + f"norm{j}.bias"] = vec(d)

    for i in range(le):
        p = f"transformer.encoder.layers.{i}."
        attn(p + "self_attn.")
        ffn_ln(p, 2)
    sd["transformer.encoder.norm.weight"] = 1.0 + vec(d)
    sd["transformer.encoder.norm.bias"] = vec(d)
    for i in range(ld):
        p = f"transformer.decoder.layers.{i}."
        attn(p + "self_attn.")
        attn(p + "multihead_attn.")
        ffn_ln(p, 3)
    sd["transformer.decoder.norm.weight"] = 1.0 + vec(d)
    sd["transformer.decoder.norm.bias"] = vec(d)
    sd["fc.weight"] = mat(V, d)
    sd["fc.bias"] = vec(V)
    return sd
