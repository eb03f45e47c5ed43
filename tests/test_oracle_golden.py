"""Pins oracle/smer_oracle.py against vectors produced by the real reference
(oracle/make_golden.py -> tests/golden/).  CPU only."""
import os

import numpy as np
import pytest
import torch


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_vocab_constants(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "vocab.npz"))
    O = oracle
    assert int(g["vocab_size"]) == O.V
    assert list(g["pitch_indices"]) == list(O.PITCH)
    assert list(g["duration_only_indices"]) == list(O.DURATION_ONLY)
    assert list(g["rest_indices"]) == [O.REST] and list(g["sep_indices"]) == [O.SEP]
    assert int(g["continue_index"]) == O.CONTINUE and int(g["eos_index"]) == O.EOS and int(g["pad_index"]) == O.PAD
    assert list(g["program_indices"]) == list(O.PROGRAM)
    assert list(g["structure_indices"]) == list(O.STRUCTURE)
    assert list(g["time_signature_indices"]) == list(O.TIME_SIG)
    assert list(g["tempo_indices"]) == list(O.TEMPO)
    assert list(g["density_indices"]) == list(O.DENSITY)
    assert list(g["occupation_indices"]) == list(O.OCCUPATION)
    assert list(g["polyphony_indices"]) == list(O.POLYPHONY)
    assert list(g["tensile_indices"]) == list(O.TENSILE)
    assert list(g["key_indices"]) == list(O.KEY)
    assert list(g["mask_indices"]) == [O.M0]
    assert list(g["duration_indices"]) == list(range(234, 242))
    assert int(g["bar"]) == O.BAR and int(g["track_0"]) == O.TRACK0 and int(g["unk"]) == O.UNK


def test_positional_table(oracle, golden_dir):
    g = _load(golden_dir, "fwd_small.pt")
    pe = g["state_dict"]["pos_enc.pe"]
    mine = oracle.positional_table(pe.shape[0], pe.shape[2])
    assert torch.equal(mine, pe)


def test_forward_logits_and_attn(oracle, golden_dir):
    g = _load(golden_dir, "fwd_small.pt")
    T = g["tgt_in"].shape[1]
    mask = oracle.nopeek_mask(T)[None].repeat(3, 1, 1)
    logits, attn = oracle.score_transformer_forward(
        g["state_dict"], g["src"], g["tgt_in"], g["cfg"]["h"], g["src_pad"], g["tgt_pad"], g["src_pad"], mask)
    valid = ~g["tgt_pad"]
    torch.testing.assert_close(logits[valid], g["logits"][valid], rtol=1e-5, atol=1e-5)
    a_ref = g["attn"]
    assert attn.shape == a_ref.shape
    torch.testing.assert_close(attn[valid[:, None].expand(-1, attn.shape[1], -1)],
                               a_ref[valid[:, None].expand(-1, attn.shape[1], -1)], rtol=1e-5, atol=1e-6)


def test_forward_batch1_no_masks(oracle, golden_dir):
    g = _load(golden_dir, "fwd_small.pt")
    lg, at = oracle.score_transformer_forward(g["state_dict"], g["src"][:1, :30], g["tgt_in"][:1, :9],
                                              g["cfg"]["h"], None, None, None, oracle.nopeek_mask(9)[None])
    torch.testing.assert_close(lg, g["b1_logits"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(at, g["b1_attn"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("eos_w", [1.0, 0.8])
def test_loss_and_grads(oracle, golden_dir, eos_w):
    g = _load(golden_dir, "fwd_small.pt")
    W, C = oracle.loss_weights(eos_w)
    loss, grads, logits, cats = oracle.train_step_grads(
        g["state_dict"], g["src"], g["tgt_in"], g["tgt_out"], g["src_pad"], g["tgt_pad"], g["cfg"]["h"], W, C)
    torch.testing.assert_close(loss, g[f"loss_{eos_w}"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cats, g[f"parts_{eos_w}"], rtol=1e-5, atol=1e-6)
    for n, gr in g[f"grads_{eos_w}"].items():
        torch.testing.assert_close(grads[n], gr, rtol=2e-4, atol=2e-6, msg=lambda m, n=n: f"{n}: {m}")


def test_adam_step(oracle, golden_dir):
    g = _load(golden_dir, "fwd_small.pt")
    for n, after in g["after_adam"].items():
        p = g["state_dict"][n].clone()
        gr = g["grads_0.8"][n]
        oracle.adam_step(p, gr, torch.zeros_like(p), torch.zeros_like(p), 1, lr=1e-4)
        torch.testing.assert_close(p, after, rtol=1e-6, atol=1e-7)


def test_sampling_distributions(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "sampling.npz"))
    logits = g["logits"]
    F = oracle.Flags
    sets = {
        "in_sep": F(no_rest=True, no_sep=True, no_eos=True, no_whole_duration=True, no_control=True),
        "in_continue": F(no_rest=True, no_sep=True, no_duration=True, no_continue=True, no_eos=True, no_control=True),
        "in_pitch_nwd0": F(no_rest=True, no_sep=True, no_continue=True, no_eos=True, no_control=True),
        "in_pitch_nwd1": F(no_rest=True, no_sep=True, no_continue=True, no_whole_duration=True, no_eos=True, no_control=True),
        "in_rest_nwd0": F(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_eos=True, no_control=True),
        "in_rest_nwd1": F(no_pitch=True, no_rest=True, no_sep=True, no_continue=True, no_whole_duration=True, no_eos=True, no_control=True),
        "first_r": F(no_duration=True, no_control=True),
        "first_d": F(is_density=True), "first_o": F(is_occupation=True),
        "first_p": F(is_polyphony=True), "first_t": F(is_tensile=True),
        "free_nwd0": F(no_control=True), "free_nwd1": F(no_whole_duration=True, no_control=True),
    }
    for name, f in sets.items():
        for r in range(logits.shape[0]):
            for t in (1.0, 0.7):
                q = oracle.masked_probs(logits[r], f, t)
                np.testing.assert_allclose(q, g[f"{name}/{r}/t{t}/probs"], rtol=1e-12, atol=0)
            nq = oracle.nucleus_probs(oracle.masked_probs(logits[r], f, 1.0), 0.9)
            np.testing.assert_allclose(nq, g[f"{name}/{r}/nucleus0.9"], rtol=1e-10, atol=1e-300)


def test_state_flags_match_appendix_counts(oracle):
    """Allowed-id counts of SURVEY.md appendix A (derived from the reference's sampling)."""
    O = oracle
    s = O.SpanState(in_sep=True)
    assert O.allowed_mask(s.flags(3, "r", False)[0]).sum() == 162
    s = O.SpanState(in_continue=True)
    assert O.allowed_mask(s.flags(3, "r", False)[0]).sum() == 157
    s = O.SpanState(in_pitch=True)
    assert O.allowed_mask(s.flags(3, "r", False)[0]).sum() == 162
    assert O.allowed_mask(s.flags(3, "r", True)[0]).sum() == 161
    s = O.SpanState(in_rest=True)
    assert O.allowed_mask(s.flags(3, "r", False)[0]).sum() == 74
    assert O.allowed_mask(s.flags(3, "r", True)[0]).sum() == 73
    s = O.SpanState()
    assert O.allowed_mask(s.flags(1, "r", False)[0]).sum() == 161
    assert O.allowed_mask(s.flags(1, "d", False)[0]).sum() == 10
    assert O.allowed_mask(s.flags(1, "t", False)[0]).sum() == 12
    assert O.allowed_mask(s.flags(2, "r", False)[0]).sum() == 166
    assert O.allowed_mask(s.flags(2, "r", True)[0]).sum() == 165


@pytest.mark.parametrize("name", ["decode_greedy.pt", "decode_greedy_cap.pt", "decode_greedy_34.pt", "decode_greedy_68.pt"])
def test_greedy_decode_matches_reference(oracle, golden_dir, name):
    g = _load(golden_dir, name)
    O = oracle
    src = O.mask_bar_and_track_ids(g["piece_ids"], g["tracks"], g["bars"], 3)
    assert np.array_equal(src, g["src"])
    targets = O.mask_targets(len(g["bars"]), g["tracks"], 3)
    nwd = int(g["piece_ids"][0]) != 7          # generation.py:504-507: only N/4 with N >= 4 allows the whole note
    tr = O.infill_decode(g["state_dict"], src, targets, g["cfg"]["h"], all_controls=g["all_controls"],
                         nwd=nwd, mode="greedy", keep_trace=True)
    assert len(tr.step_logits) == len(g["step_logits"])
    for i, (pref, row) in enumerate(zip(g["step_prefix"], g["step_logits"])):
        assert tr.step_prefix_len[i] == len(pref)
        np.testing.assert_allclose(tr.step_logits[i], row, rtol=1e-4, atol=2e-5)
    # final decoder stream: last reference prefix is a prefix of ours
    last = list(g["step_prefix"][-1])
    assert tr.tokens[: len(last)] == last or tr.tokens == last[: len(tr.tokens)]
