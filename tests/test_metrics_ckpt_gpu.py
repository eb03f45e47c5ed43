"""GPU parity of the §8(f) rows beside the hot path: device-side token accuracy (train.py:988-1034)
against the reference fixture and the oracle, and resuming from a reference checkpoint with FusedAdam."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_accuracy_matches_reference_fixture(golden_dir):
    from smer_music_generation_b200 import SmerAccuracy
    g = torch.load(os.path.join(golden_dir, "metrics_small.pt"), weights_only=False)
    acc = SmerAccuracy().to(DEV)
    out = acc(g["logits"].to(DEV), g["tgt_out"].to(DEV))
    assert set(out) == set(g["accuracy"])
    for k, v in g["accuracy"].items():
        assert abs(out[k] - v) < 1e-12, (k, out[k], v)            # ratios of the same integers
    assert acc.first_sample_argmax.cpu().tolist() == g["first_generated"]        # incl. the tie: first maximum


def test_accuracy_matches_oracle_at_full_size_and_accumulates(oracle):
    from smer_music_generation_b200 import SmerAccuracy
    from smer_music_generation_b200.loss import token_class_table
    gen = torch.Generator().manual_seed(3)
    B, T, V = 32, 1024, 309
    _, _, tgt, _, _ = oracle.synth_batch(B, 64, T, seed=17)
    logits = torch.randn(B, T, V, generator=gen)
    hit = torch.rand(B, T, generator=gen) < 0.6
    logits.scatter_add_(2, tgt[..., None], (hit.float() * 9.0)[..., None])
    ref, correct, seen, am = oracle.token_accuracy(logits, tgt, token_class_table(V).numpy())
    acc = SmerAccuracy().to(DEV)
    out = acc(logits.to(DEV), tgt.to(DEV))
    for k, v in ref.items():
        assert abs(out[k] - v) < 1e-12, (k, out[k], v)
    assert torch.equal(acc.first_sample_argmax.cpu(), torch.from_numpy(am[0]))
    # two half batches accumulated == the whole batch; padded (strided) logits rows are read in place
    acc.reset()
    wide = torch.zeros(B, T, 320, device=DEV)
    wide[..., :V] = logits.to(DEV)
    acc.update(wide[: B // 2, :, :V], tgt[: B // 2].to(DEV))
    acc.update(wide[B // 2:, :, :V], tgt[B // 2:].to(DEV))
    out2 = acc.compute()
    assert out2 == out


def test_resume_from_reference_checkpoint(golden_dir):
    """load (train.py:266-303) -> one FusedAdam step on the reference run's next gradients == the parameters
    torch.optim.Adam reached in the reference run."""
    from smer_music_generation_b200 import ScoreTransformer
    from smer_music_generation_b200.trainer import FusedAdam
    fx = torch.load(os.path.join(golden_dir, "ckpt_ref_small.pt"), weights_only=False)
    c, ck = fx["cfg"], fx["checkpoint"]
    m = ScoreTransformer(309, c["d"], c["h"], c["le"], c["ld"], c["ff"], c["maxlen"], 0.1, 0.1, compute_dtype="fp32").to(DEV)
    m.load_state_dict(ck["model_state_dict"])
    opt = FusedAdam(m.parameters(), lr=1e-4)
    opt.load_state_dict(copy.deepcopy(ck["optimizer_state_dict"]))
    for n, p in m.named_parameters():
        p.grad = fx["grads_step3"][n].to(DEV)
    opt.step()
    torch.cuda.synchronize()
    for n, p in m.named_parameters():
        want = fx["params_after_step3"][n]
        err = (p.detach().cpu() - want).abs().max().item()
        assert err < 2e-7 + 1e-6 * want.abs().max().item(), (n, err)
    # what we save now is loadable by torch.optim.Adam again (state layout unchanged by the step)
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == 3.0
    adam = torch.optim.Adam(m.parameters(), lr=1e-4)
    adam.load_state_dict(sd)
