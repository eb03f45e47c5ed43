"""CPU checks of the two "next" rows of SURVEY.md §8(f) that sit beside the hot path:
* the oracle's token-accuracy restatement (train.py:988-1034) against the fixture produced by the
  reference's own `accuracy()` and WordVocab (oracle/make_golden.py::golden_metrics);
* checkpoint compatibility (train.py:266-303, 967-973): a checkpoint written by the reference loads
  into ScoreTransformer / FusedAdam, and what we save loads back into torch.optim.Adam."""
import copy
import io
import os

import torch


def test_oracle_token_accuracy_matches_reference_fixture(oracle, golden_dir):
    g = torch.load(os.path.join(golden_dir, "metrics_small.pt"), weights_only=False)
    assert tuple(g["classes"]) == tuple(oracle.TOKEN_CLASSES)
    acc, correct, seen, am = oracle.token_accuracy(g["logits"], g["tgt_out"], g["class_of"].numpy())
    assert set(acc) == set(g["accuracy"])
    for k, v in g["accuracy"].items():
        assert abs(acc[k] - v) < 1e-12, (k, acc[k], v)
    assert am[0].tolist() == g["first_generated"]            # incl. the exact tie at position 3 (first maximum wins)
    assert g["tgt_out"][0].tolist() == g["first_target"]
    assert seen["total"] == int((g["tgt_out"] != 0).sum())


def test_token_class_table_matches_reference_vocab(golden_dir):
    from smer_music_generation_b200.loss import TOKEN_CLASSES, token_class_table
    g = torch.load(os.path.join(golden_dir, "metrics_small.pt"), weights_only=False)
    assert [n for n, _, _ in TOKEN_CLASSES] == list(g["classes"])
    assert torch.equal(token_class_table(309), g["class_of"])


def test_reference_checkpoint_loads_and_round_trips(golden_dir):
    from smer_music_generation_b200 import ScoreTransformer
    from smer_music_generation_b200.trainer import FusedAdam
    fx = torch.load(os.path.join(golden_dir, "ckpt_ref_small.pt"), weights_only=False)
    c, ck = fx["cfg"], fx["checkpoint"]
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "epoch", "loss"}          # train.py:967-973
    m = ScoreTransformer(309, c["d"], c["h"], c["le"], c["ld"], c["ff"], c["maxlen"], 0.1, 0.1)
    missing, unexpected = m.load_state_dict(ck["model_state_dict"], strict=True)
    assert not missing and not unexpected
    for k, v in m.state_dict().items():
        assert torch.equal(v, ck["model_state_dict"][k]), k
    # resume the optimizer from the reference's torch.optim.Adam state (train.py:291-295)
    opt = FusedAdam(m.parameters(), lr=1e-4)
    opt.load_state_dict(copy.deepcopy(ck["optimizer_state_dict"]))
    ref_state = ck["optimizer_state_dict"]["state"]
    for i, p in enumerate(m.parameters()):
        st = opt.state[p]
        assert float(st["step"]) == float(ref_state[i]["step"]) == 2.0
        assert st["exp_avg"].shape == p.shape and torch.equal(st["exp_avg"], ref_state[i]["exp_avg"])
        assert torch.equal(st["exp_avg_sq"], ref_state[i]["exp_avg_sq"])
    # save exactly as train.py:967-973 does, reload, and hand the result to the reference's optimizer class
    buf = io.BytesIO()
    torch.save({"model_state_dict": m.state_dict(), "optimizer_state_dict": opt.state_dict(), "epoch": 4, "loss": 1.0}, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert list(back["model_state_dict"].keys()) == list(ck["model_state_dict"].keys())
    for k, v in back["model_state_dict"].items():
        assert v.dtype == ck["model_state_dict"][k].dtype and torch.equal(v, ck["model_state_dict"][k]), k
    m2 = ScoreTransformer(309, c["d"], c["h"], c["le"], c["ld"], c["ff"], c["maxlen"], 0.1, 0.1)
    m2.load_state_dict(back["model_state_dict"])
    adam = torch.optim.Adam(m2.parameters(), lr=1e-4)                                        # train.py:264
    adam.load_state_dict(back["optimizer_state_dict"])
    # ... and torch.optim.Adam continues from it to the parameters the reference run reached (CPU arithmetic)
    for n, p in m2.named_parameters():
        p.grad = fx["grads_step3"][n].clone()
    adam.step()
    for n, p in m2.named_parameters():
        assert torch.allclose(p.detach(), fx["params_after_step3"][n], rtol=0, atol=1e-7), n


def test_span_surgery_matches_reference_fixture(golden_dir, oracle):
    """spans.mask_bar_and_track / restore_marked_input against generation.mask_bar_and_track (generation.py:248-341)
    and generation.restore_marked_input (generation.py:417-465) run on the reference itself (tests/golden/spans.pt),
    and against the oracle's id-level restatement on more pieces."""
    import numpy as np
    from smer_music_generation_b200 import spans
    cases = torch.load(os.path.join(golden_dir, "spans.pt"), weights_only=False)
    assert len(cases) >= 12
    for c in cases:
        src, tn, bn = spans.mask_bar_and_track(c["ids"], c["tracks"], c["bars"])
        assert np.array_equal(src, c["src"]) and tn == c["track_names"] and bn == c["bar_names"]
        assert np.array_equal(spans.restore_marked_input(src, c["generated"]), c["restored"])
    for seed in range(6):
        ids = np.asarray(oracle.synth_piece(seed=seed, n_bars=16, n_tracks=3, events_per_track_bar=6))
        for tr, br in (([0, 1, 2], [4, 5, 6, 7]), ([2], [15]), ([0, 2], [0, 9])):
            src, _, _ = spans.mask_bar_and_track(ids, tr, br)
            assert np.array_equal(src, oracle.mask_bar_and_track_ids(ids, tr, br, 3))
            assert spans.mask_targets(ids, tr, br) == oracle.mask_targets(len(br), tr, 3)
    # nothing selected / nothing generated: identity
    ids = np.asarray(oracle.synth_piece(seed=1, n_bars=4, n_tracks=3, events_per_track_bar=2))
    assert np.array_equal(spans.mask_bar_and_track(ids, [], [1])[0], ids)
    assert np.array_equal(spans.restore_marked_input(ids, []), ids)
