"""Generate tests/golden/* by running the REAL reference modules (read-only import from
/root/reference) on seeded inputs.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference ships no tests or golden vectors; these files are what pins
oracle/smer_oracle.py (and through it the CUDA path) to the reference's behaviour.
Nothing here is imported at test time; tests read only the committed fixture files.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
for name in ("pretty_midi", "music21", "coloredlogs"):       # absent, unused on this path
    sys.modules.setdefault(name, types.ModuleType(name))

import warnings
warnings.filterwarnings("ignore")

import model as ref_model          # noqa: E402  (reference model.py)
import vocab as ref_vocab          # noqa: E402
import generation as ref_gen       # noqa: E402
import smer_oracle as O            # noqa: E402
from torch import nn               # noqa: E402
from einops import rearrange       # noqa: E402


def build_ref(d, h, le, ld, ff, maxlen, seed, dropout=0.0):
    torch.manual_seed(seed)
    m = ref_model.ScoreTransformer(309, d, h, le, ld, ff, maxlen, dropout, dropout)
    for p in m.parameters():                       # train.py:261-263
        if p.dim() > 1:
            nn.init.xavier_normal_(p)
    # make biases / LN affine non-trivial so that parity exercises them
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    return m


def reference_loss(vocab, outputs, tgt_out, eos_weight):
    """Literal restaging of train.py:555-642 + 726-780 with control_number 5."""
    Vn = vocab.vocab_size

    def ce(lo, hi):
        w = torch.zeros(Vn)
        w[lo:hi] = 1
        return nn.CrossEntropyLoss(ignore_index=0, weight=w, reduction="none")

    meta_w = torch.zeros(Vn)
    meta_w[1] = eos_weight
    crit = [nn.CrossEntropyLoss(ignore_index=0, weight=meta_w, reduction="none"),
            ce(7, 11), ce(18, 146), ce(11, 18), ce(3, 7), ce(146, 234),
            ce(234, 234 + len(vocab.duration_indices))]
    for k in ("tensile", "key", "density", "occupation", "polyphony"):
        r = vocab.control_indices[k]
        crit.append(ce(r[0], r[-1] + 1))
    ce_all = torch.ones(Vn)
    ce_all[0] = 0
    ce_all[2] = 0
    ce_all[-1] = 0
    ce_all[1] = eos_weight
    x = rearrange(outputs, "b t v -> (b t) v")
    y = rearrange(tgt_out, "b o -> (b o)")
    parts = [torch.sum(c(x, y)) / ce_all[y].sum() for c in crit]
    return sum(parts), torch.stack(parts)


def golden_vocab(vocab):
    d = {k: np.asarray(getattr(vocab, k)) for k in (
        "pitch_indices", "duration_only_indices", "rest_indices", "sep_indices", "program_indices",
        "structure_indices", "time_signature_indices", "tempo_indices", "density_indices",
        "occupation_indices", "polyphony_indices", "tensile_indices", "mask_indices", "duration_indices")}
    d["continue_index"] = np.asarray(vocab.continue_index)
    d["eos_index"] = np.asarray(vocab.eos_index)
    d["pad_index"] = np.asarray(vocab.pad_index)
    d["key_indices"] = np.asarray(vocab.control_indices["key"])
    d["vocab_size"] = np.asarray(vocab.vocab_size)
    d["bar"] = np.asarray(vocab.char2index("bar"))
    d["track_0"] = np.asarray(vocab.char2index("track_0"))
    d["unk"] = np.asarray(vocab.char2index("unk"))
    np.savez(os.path.join(OUT, "vocab.npz"), **d)


def golden_forward(vocab):
    cfg = dict(d=32, h=2, le=2, ld=2, ff=64, maxlen=96)
    m = build_ref(seed=7, **cfg)
    m.train()                                        # dropout p = 0: train == eval arithmetic
    src, tgt_in, tgt_out, src_pad, tgt_pad = O.synth_batch(3, 40, 24, seed=11)
    T = tgt_in.shape[1]
    tgt_mask = ref_gen.gen_nopeek_mask(T)[None].repeat(3, 1, 1)
    out = {}
    for eos_w in (1.0, 0.8):
        m.zero_grad()
        logits, attn = m(src, tgt_in, src_pad, tgt_pad, src_pad.clone(), tgt_mask)
        loss, parts = reference_loss(vocab, logits, tgt_out, eos_w)
        loss.backward()
        out[f"loss_{eos_w}"] = loss.detach()
        out[f"parts_{eos_w}"] = parts.detach()
        out[f"grads_{eos_w}"] = {n: p.grad.clone() for n, p in m.named_parameters()}
    out.update(cfg=cfg, state_dict={k: v.clone() for k, v in m.state_dict().items()},
               src=src, tgt_in=tgt_in, tgt_out=tgt_out, src_pad=src_pad, tgt_pad=tgt_pad,
               logits=logits.detach(), attn=attn.detach())
    # no-mask batch-1 call as generation.model_generate makes it
    m.eval()
    with torch.no_grad():
        lg1, at1 = m(src[:1, :30], tgt_in[:1, :9], None, None, None, ref_gen.gen_nopeek_mask(9)[None])
    out["b1_logits"] = lg1
    out["b1_attn"] = at1
    # one Adam step exactly as train.py:264,786
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    opt.step()
    out["after_adam"] = {n: p.detach().clone() for n, p in m.named_parameters()}
    torch.save(out, os.path.join(OUT, "fwd_small.pt"))


FLAG_SETS = {
    "in_sep": dict(no_rest=True, no_sep=True, no_eos=True, no_whole_duration=True, no_control=True),
    "in_continue": dict(no_rest=True, no_sep=True, no_duration=True, no_continue=True, no_eos=True, no_control=True),
    "in_pitch_nwd0": dict(no_rest=True, no_sep=True, no_continue=True, no_whole_duration=False, no_eos=True, no_control=True),
    "in_pitch_nwd1": dict(no_rest=True, no_sep=True, no_continue=True, no_whole_duration=True, no_eos=True, no_control=True),
    "in_rest_nwd0": dict(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_whole_duration=False, no_eos=True, no_control=True),
    "in_rest_nwd1": dict(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_whole_duration=True, no_eos=True, no_control=True),
    "first_r": dict(no_duration=True, no_control=True),
    "first_d": dict(is_density=True),
    "first_o": dict(is_occupation=True),
    "first_p": dict(is_polyphony=True),
    "first_t": dict(is_tensile=True),
    "free_nwd0": dict(no_whole_duration=False, no_control=True),
    "free_nwd1": dict(no_whole_duration=True, no_control=True),
}


def golden_sampling(vocab):
    rng = np.random.RandomState(5)
    logits = (rng.randn(4, 309) * 3.0).astype(np.float32)
    res = {"logits": logits}
    captured = {}
    orig_ws, orig_choice = ref_gen.weighted_sampling, np.random.choice

    def cap_ws(probs):
        captured["probs"] = probs.copy()
        return int(np.argmax(probs))

    def cap_choice(a, size=None, p=None):
        captured["nuc_idx"] = np.asarray(a).copy()
        captured["nuc_p"] = np.asarray(p).copy()
        return np.asarray([a[0]])

    ref_gen.weighted_sampling = cap_ws
    try:
        for name, kw in FLAG_SETS.items():
            for r in range(logits.shape[0]):
                for t in (1.0, 0.7):
                    ref_gen.sampling(torch.tensor(logits[r])[None], vocab, t=t, **kw)
                    res[f"{name}/{r}/t{t}/probs"] = captured["probs"]
            np.random.choice = cap_choice
            try:
                for r in range(logits.shape[0]):
                    ref_gen.sampling(torch.tensor(logits[r])[None], vocab, p=0.9, **kw)
                    full = np.zeros(309)
                    full[captured["nuc_idx"]] = captured["nuc_p"]
                    res[f"{name}/{r}/nucleus0.9"] = full
            finally:
                np.random.choice = orig_choice
    finally:
        ref_gen.weighted_sampling = orig_ws
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), **res)


def golden_decode(vocab, name="decode_greedy.pt", all_controls=tuple(range(242, 308)), tracks=(1, 2), bars=(1, 2),
                  time_sig=7, model_seed=21):
    cfg = dict(d=32, h=2, le=2, ld=2, ff=64, maxlen=700)
    m = build_ref(seed=model_seed, **cfg).eval()
    if time_sig != 7:
        # make the whole-note token (id 234) the likeliest continuation wherever the grammar allows it, so that the
        # no_whole_duration rule of generation.py:504-507 decides tokens of the greedy stream
        with torch.no_grad():
            m.fc.bias[234] += 6.0
    ids = O.synth_piece(seed=3, n_bars=4, n_tracks=3, events_per_track_bar=3, time_sig=time_sig)
    events = [vocab.index2char(i) for i in ids]
    steps = []
    orig_mg, orig_ws = ref_gen.model_generate, ref_gen.weighted_sampling

    def rec_mg(model, src, tgt, device, return_weights=False):
        out = orig_mg(model, src, tgt, device, return_weights=return_weights)
        o = out[0] if return_weights else out
        steps.append((list(tgt), o[-1].numpy().copy()))
        return out

    class L:
        def info(self, *a, **k):
            pass

    ref_gen.model_generate = rec_mg
    ref_gen.weighted_sampling = lambda probs: int(np.argmax(probs))     # greedy over masked probs
    ref_gen.tqdm = lambda it, **k: it
    try:
        res = ref_gen.generation_all(m, events, "cpu", vocab, L(), list(all_controls), list(tracks), list(bars))
    finally:
        ref_gen.model_generate, ref_gen.weighted_sampling = orig_mg, orig_ws
    assert res is not None
    restored, tnames, bnames = res
    src, _, _ = ref_gen.mask_bar_and_track(events, vocab, list(tracks), list(bars))
    # the decoder stream after the last step = last recorded prefix + its argmax bookkeeping
    out = dict(cfg=cfg, state_dict={k: v.clone() for k, v in m.state_dict().items()},
               piece_ids=np.asarray(ids), src=np.asarray(src),
               step_prefix=[np.asarray(s[0]) for s in steps],
               step_logits=np.stack([s[1] for s in steps]).astype(np.float32),
               restored=[vocab.char2index(e) if e in vocab._char2idx else -1 for e in restored],
               mask_tracks=tnames, mask_bars=bnames, all_controls=list(all_controls),
               tracks=list(tracks), bars=list(bars))
    torch.save(out, os.path.join(OUT, name))
    print("decode steps:", len(steps), "S =", len(src))


def reference_function(path, name, namespace):
    """Compiles ONE function of a reference script without importing the script (train.py cannot be
    imported here: coloredlogs / wandb.login at import time, train.py:8,25) and returns it."""
    import ast
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    code = compile(ast.Module(body=[fn], type_ignores=[]), path, "exec")
    exec(code, namespace)
    return namespace[name]


def golden_metrics(vocab):
    """train.py:988-1034 `accuracy()` run on the golden forward's logits: per-class and total token
    accuracy, plus the class table it looks targets up in (vocab.py:159-300)."""
    accuracy = reference_function(os.path.join(REF, "train.py"), "accuracy", {"torch": torch})
    fw = torch.load(os.path.join(OUT, "fwd_small.pt"))
    logits, tgt_out = fw["logits"], fw["tgt_out"]
    # make the argmax interesting: push half of the positions to their target, keep a few exact ties
    g = torch.Generator().manual_seed(5)
    lg = logits.clone()
    hit = torch.rand(tgt_out.shape, generator=g) < 0.5
    lg.scatter_add_(2, tgt_out[..., None], (hit.float() * 8.0)[..., None])
    lg[0, 3, 200] = lg[0, 3].max() + 1.0
    lg[0, 3, 150] = lg[0, 3, 200]                       # tie: torch.argmax takes the first maximum (150)
    res, gen, tgt = accuracy(lg, tgt_out, vocab)
    classes = sorted(set(vocab.token_class_ranges.values()))
    class_of = np.full(vocab.vocab_size, -1, dtype=np.int32)
    for idx, c in vocab.token_class_ranges.items():
        class_of[idx] = classes.index(c)
    torch.save({"logits": lg, "tgt_out": tgt_out, "classes": classes, "class_of": torch.from_numpy(class_of),
                "accuracy": {k: float(v) for k, v in res.items()},
                "first_generated": [vocab.char2index(t) for t in gen], "first_target": [vocab.char2index(t) for t in tgt]},
               os.path.join(OUT, "metrics_small.pt"))


def golden_spans(vocab):
    """generation.mask_bar_and_track (generation.py:248-341) and generation.restore_marked_input
    (generation.py:417-465) on synthetic pieces, as ids."""
    cases = []
    rng = np.random.RandomState(9)
    for seed, (n_bars, n_tracks, ev) in enumerate([(8, 3, 4), (16, 3, 6), (6, 2, 3), (5, 1, 5)]):
        ids = O.synth_piece(seed=seed, n_bars=n_bars, n_tracks=n_tracks, events_per_track_bar=ev)
        events = [vocab.index2char(int(i)) for i in ids]
        for tracks, bars in [(list(range(n_tracks)), list(range(n_bars // 2, n_bars // 2 + 2))), ([n_tracks - 1], [1]),
                             ([0], [0, n_bars - 1])]:
            src, tnames, bnames = ref_gen.mask_bar_and_track(list(events), vocab, tracks, bars)
            src_tokens = [vocab.index2char(int(i)) for i in src]
            n_mask = int((np.asarray(src) == vocab.char2index("m_0")).sum())
            gen = []
            for k in range(n_mask if seed % 2 == 0 else max(1, n_mask - 1)):      # odd seeds: one span fewer than masks
                gen.append("m_0")
                gen.extend(vocab.index2char(int(t)) for t in rng.randint(146, 242, size=rng.randint(0, 6)))
            restored = ref_gen.restore_marked_input(src_tokens, gen)
            cases.append({"ids": np.asarray(ids), "tracks": tracks, "bars": bars, "src": np.asarray(src),
                          "track_names": list(tnames), "bar_names": list(bnames),
                          "generated": np.asarray([vocab.char2index(t) for t in gen]),
                          "restored": np.asarray([vocab.char2index(str(t)) for t in restored])})
    torch.save(cases, os.path.join(OUT, "spans.pt"))


def golden_checkpoint():
    """A checkpoint file exactly as train.py:967-973 writes it (model + torch.optim.Adam state after
    two steps), for the load / resume / save compatibility tests."""
    cfg = dict(d=32, h=2, le=2, ld=2, ff=64, maxlen=96)
    m = build_ref(seed=21, **cfg)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)                     # train.py:264
    src, tgt_in, tgt_out, src_pad, tgt_pad = O.synth_batch(2, 24, 16, seed=3)
    tgt_mask = ref_gen.gen_nopeek_mask(tgt_in.shape[1])[None].repeat(2, 1, 1)
    loss = None
    for _ in range(2):
        opt.zero_grad()
        logits, _ = m(src, tgt_in, src_pad, tgt_pad, src_pad.clone(), tgt_mask)
        loss = nn.functional.cross_entropy(logits.reshape(-1, 309), tgt_out.reshape(-1), ignore_index=0)
        loss.backward()
        opt.step()
    # gradients of a third step, and the parameters torch.optim.Adam produces from them (resume parity)
    opt.zero_grad()
    logits, _ = m(src, tgt_in, src_pad, tgt_pad, src_pad.clone(), tgt_mask)
    nn.functional.cross_entropy(logits.reshape(-1, 309), tgt_out.reshape(-1), ignore_index=0).backward()
    ckpt = {"model_state_dict": {k: v.clone() for k, v in m.state_dict().items()},
            "optimizer_state_dict": __import__("copy").deepcopy(opt.state_dict()), "epoch": 3, "loss": float(loss)}
    grads3 = {n: p.grad.clone() for n, p in m.named_parameters()}
    opt.step()
    torch.save({"cfg": cfg, "checkpoint": ckpt, "grads_step3": grads3,
                "params_after_step3": {n: p.detach().clone() for n, p in m.named_parameters()}},
               os.path.join(OUT, "ckpt_ref_small.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    v = ref_vocab.WordVocab(0, ["key", "tensile", "density", "polyphony", "occupation"])
    if "--nwd-only" in sys.argv:         # just the fixtures added for the no_whole_duration rule
        golden_decode(v, name="decode_greedy_34.pt", tracks=(0, 2), bars=(1,), time_sig=8)
        golden_decode(v, name="decode_greedy_68.pt", tracks=(1,), bars=(2,), time_sig=10)
        sys.exit(0)
    golden_vocab(v)
    golden_forward(v)
    golden_sampling(v)
    golden_decode(v)
    golden_decode(v, name="decode_greedy_cap.pt", all_controls=(), tracks=(2,), bars=(1,))
    # 3/4 and 6/8 pieces: no_whole_duration = True (generation.py:504-507)
    golden_decode(v, name="decode_greedy_34.pt", tracks=(0, 2), bars=(1,), time_sig=8)
    golden_decode(v, name="decode_greedy_68.pt", tracks=(1,), bars=(2,), time_sig=10)
    golden_metrics(v)
    golden_spans(v)
    golden_checkpoint()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
