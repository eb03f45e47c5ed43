"""The reference arm of bench.py: the UNMODIFIED reference modules (model.py, transformer.py, vocab.py,
generation.py and what they import) driven through their own public API on the host cores.

The reference is a flat directory of scripts without a setup.py, so `pip install --target baseline/_ref` has nothing
to install; `__graft_entry__.build()` stages the needed files, byte for byte, under the git-ignored `baseline/_ref/`
 so that they travel to the GPU box, where /root/reference does not exist.  Nothing of this
repo's model, kernels or engine is on this path.  train.py itself cannot be imported (argparse / wandb.login /
coloredlogs at import time, train.py:8,25), so the step below is the loop body of train.py:702-797 written against
the reference's own `model.ScoreTransformer`, `nn.CrossEntropyLoss` criteria (train.py:555-642) and
`torch.optim.Adam` (train.py:264)."""
import os
import sys
import time
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
FILES = ("model.py", "transformer.py", "vocab.py", "vocab_control.py", "generation.py", "encode.py", "tension_calculation.py")


def reference_dir():
    if all(os.path.exists(os.path.join(STAGED, f)) for f in FILES):
        return STAGED
    if os.path.isdir("/root/reference"):
        return "/root/reference"
    return None


def load():
    """-> (model module, vocab module, generation module) of the reference, or None when it is not available."""
    d = reference_dir()
    if d is None:
        return None
    sys.dont_write_bytecode = True
    if d not in sys.path:
        sys.path.insert(0, d)
    for name in ("pretty_midi", "music21", "coloredlogs"):       # absent here, untouched on this path
        sys.modules.setdefault(name, types.ModuleType(name))
    import warnings
    warnings.filterwarnings("ignore")
    saved = sys.modules.pop("model", None)                        # never the drop-in shim
    try:
        import model as ref_model
    finally:
        if saved is not None and "model" not in sys.modules:
            sys.modules["model"] = saved
    import vocab as ref_vocab
    import generation as ref_gen
    assert os.path.dirname(os.path.abspath(ref_model.__file__)) == os.path.abspath(d), ref_model.__file__
    return ref_model, ref_vocab, ref_gen


def build_model(ref_model, cfg, dropout, device="cpu", seed=0):
    torch.manual_seed(seed)
    m = ref_model.ScoreTransformer(cfg["vocab"], cfg["d"], cfg["nhead"], cfg["le"], cfg["ld"], cfg["ff"], cfg["max_len"],
                                   dropout, dropout)
    for p in m.parameters():                                      # train.py:261-263
        if p.dim() > 1:
            nn.init.xavier_normal_(p)
    return m.to(device)


def criteria(vocab, eos_weight, device):
    """train.py:555-642 with control_number 5: (list of criteria, ce_weight_all)."""
    Vn = vocab.vocab_size

    def ce(lo, hi):
        w = torch.zeros(Vn)
        w[lo:hi] = 1
        return nn.CrossEntropyLoss(ignore_index=0, weight=w.to(device), reduction="none")

    meta_w = torch.zeros(Vn)
    meta_w[1] = eos_weight
    crit = [nn.CrossEntropyLoss(ignore_index=0, weight=meta_w.to(device), reduction="none"),
            ce(7, 11), ce(18, 146), ce(11, 18), ce(3, 7), ce(146, 234), ce(234, 234 + len(vocab.duration_indices))]
    for k in ("tensile", "key", "density", "occupation", "polyphony"):
        r = vocab.control_indices[k]
        crit.append(ce(r[0], r[-1] + 1))
    ce_all = torch.ones(Vn)
    ce_all[0] = 0
    ce_all[2] = 0
    ce_all[-1] = 0
    ce_all[1] = eos_weight
    return crit, ce_all.to(device)


def train_step(model, optim, crit, ce_all, gen_nopeek_mask, batch, device):
    """Loop body of train.py:702-797 (mask build + H2D, forward, 12 criteria, backward, Adam, loss.item())."""
    from einops import rearrange
    src, tgt_inp, tgt_out, src_pad, tgt_pad = (t.to(device) for t in batch)
    mem_pad = src_pad.clone()
    tgt_mask = gen_nopeek_mask(tgt_inp.shape[1])
    tgt_mask = torch.tensor(np.repeat(np.expand_dims(tgt_mask, 0), mem_pad.shape[0], axis=0)).float().to(device)
    optim.zero_grad()
    outputs, _ = model(src, tgt_inp, src_pad, tgt_pad, mem_pad, tgt_mask)
    x = rearrange(outputs, "b t v -> (b t) v")
    y = rearrange(tgt_out, "b o -> (b o)")
    denom = ce_all[y].sum()
    loss = sum(torch.sum(c(x, y)) / denom for c in crit)
    loss.backward()
    optim.step()
    return loss.item()


def time_train(cfg, batches, steps, warmup, threads, device="cpu", dropout=0.1):
    """tokens/s of the reference training step on `device`; batches: list of (src, tgt_in, tgt_out, src_pad, tgt_pad)."""
    mods = load()
    if mods is None:
        return None
    ref_model, ref_vocab, ref_gen = mods
    if device == "cpu":
        torch.set_num_threads(threads)
    vocab = ref_vocab.WordVocab(0, ["key", "tensile", "density", "polyphony", "occupation"])
    m = build_model(ref_model, cfg, dropout, device).train()
    optim = torch.optim.Adam(m.parameters(), lr=1e-4)              # train.py:264
    crit, ce_all = criteria(vocab, 0.8, device)
    times, toks, loss = [], 0, None
    for it in range(warmup + steps):
        b = batches[it % len(batches)]
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = train_step(m, optim, crit, ce_all, ref_gen.gen_nopeek_mask, b, device)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            toks += int((~b[3]).sum() + (~b[4]).sum())
    return {"tokens_per_s": toks / sum(times), "s_per_step": sum(times) / len(times), "loss": loss, "steps": len(times)}


def run_generation_all(model, piece_ids, tracks, bars, device, greedy=True, all_controls=tuple(range(242, 308))):
    """generation.generation_all (generation.py:468-696) unchanged, on one piece given as token ids.  greedy: the
    reference has no greedy mode; `weighted_sampling` is swapped for an argmax over the masked
    probabilities, which is how configs[0] defines it.  -> (restored ids, generated tokens, model calls, seconds)."""
    mods = load()
    ref_model, ref_vocab, ref_gen = mods
    vocab = getattr(run_generation_all, "_vocab", None)
    if vocab is None:
        vocab = ref_vocab.WordVocab(0, ["key", "tensile", "density", "polyphony", "occupation"])
        run_generation_all._vocab = vocab
    events = [vocab.index2char(int(i)) for i in piece_ids]
    calls = {"n": 0, "last": None}
    orig_mg, orig_ws = ref_gen.model_generate, ref_gen.weighted_sampling

    def counting_mg(mdl, src, tgt, dev, return_weights=False):
        calls["n"] += 1
        calls["last"] = list(tgt)
        return orig_mg(mdl, src, tgt, dev, return_weights=return_weights)

    class _Log:
        def info(self, *a, **k):
            pass

    ref_gen.model_generate = counting_mg
    if greedy:
        ref_gen.weighted_sampling = lambda probs: int(np.argmax(probs))
    old_tqdm = ref_gen.tqdm
    ref_gen.tqdm = lambda it, **k: it
    t0 = time.perf_counter()
    try:
        res = ref_gen.generation_all(model, events, device, vocab, _Log(), list(all_controls), list(tracks), list(bars))
    finally:
        ref_gen.model_generate, ref_gen.weighted_sampling, ref_gen.tqdm = orig_mg, orig_ws, old_tqdm
    dt = time.perf_counter() - t0
    if res is None:
        raise RuntimeError("generation_all returned None (it swallows exceptions, generation.py:695-696)")
    restored = [vocab.char2index(e) if e in vocab._char2idx else -1 for e in res[0]]
    return restored, calls["n"], calls["last"], dt
