"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
module mirrors the reference's state_dict layout, host-side tables equal the oracle's, and the
product path refuses to run without a B200 (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from smer_music_generation_b200 import build, _capi
    build.build()
    return _capi.lib()


def test_library_exports_every_declared_symbol(lib):
    from smer_music_generation_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "smer_b200.h")).read()
    declared = set(re.findall(r"^(?:int|long long|const char\*)\s+(smer_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 25
    assert declared == set(_capi.EXPORTED), declared ^ set(_capi.EXPORTED)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.smer_version() == 100
    assert lib.smer_device_ok() == 0            # no GPU in the build container


def test_library_knobs_validate_their_arguments(lib):
    """The process-wide knobs of the C ABI need no device: the SM reservation for data-parallel runs takes an even count in
    [0, 64] and reports anything else through smer_last_error; the decode step's launch-mode toggle accepts 0 / 1."""
    assert lib.smer_set_reserved_sms(3) != 0
    assert b"even count" in lib.smer_last_error()
    assert lib.smer_set_reserved_sms(66) != 0 and lib.smer_set_reserved_sms(-2) != 0
    assert lib.smer_set_reserved_sms(8) == 0 and lib.smer_set_reserved_sms(0) == 0
    assert lib.smer_set_pdl(1) == 0 and lib.smer_set_pdl(0) == 0


def test_ctypes_structs_match_header_sizes(lib):
    """sizeof of the argument structs as nvcc laid them out == ctypes' layout."""
    import ctypes as C
    from smer_music_generation_b200 import _capi
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "smer_b200.h"
        int main(){printf("%zu %zu %zu\\n", sizeof(smer_attn_args), sizeof(smer_decode_attn_args), sizeof(smer_sample_args));return 0;}
    """)
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_capi.AttnArgs), C.sizeof(_capi.DecodeAttnArgs), C.sizeof(_capi.SampleArgs)]


def test_no_cpu_fallback():
    from smer_music_generation_b200 import ScoreTransformer, SmerLoss
    m = ScoreTransformer(309, 32, 2, 1, 1, 64, 32, 0.0, 0.0)
    src = torch.randint(3, 300, (1, 8))
    with pytest.raises(RuntimeError):
        m(src, src, None, None, None, "causal")
    with pytest.raises(RuntimeError):
        SmerLoss()(torch.zeros(4, 309), torch.zeros(4, dtype=torch.long))


def test_state_dict_layout(golden_dir):
    from smer_music_generation_b200 import ScoreTransformer
    g = torch.load(os.path.join(golden_dir, "fwd_small.pt"), weights_only=False)
    c = g["cfg"]
    m = ScoreTransformer(309, c["d"], c["h"], c["le"], c["ld"], c["ff"], c["maxlen"], 0.1, 0.1)
    ref = g["state_dict"]
    assert list(m.state_dict().keys()) == list(ref.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref)
    assert torch.equal(m.pos_enc.pe, ref["pos_enc.pe"])
    # default model: 128 entries, 29,744,437 parameters
    big = ScoreTransformer(309, 512, 8, 4, 4, 2048, 2400, 0.1, 0.1)
    assert len(big.state_dict()) == 128
    assert sum(p.numel() for p in big.parameters()) == 29744437
    for p in big.parameters():                       # train.py:261-263 re-initialises like this
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)


def test_loss_tables_equal_oracle(oracle):
    from smer_music_generation_b200 import loss_tables
    for eos_w in (1.0, 0.8):
        for cl in (("key", "tensile", "density", "polyphony", "occupation"), ("key",), ()):
            W, C, cat = loss_tables(309, eos_w, cl)
            Wo, Co = oracle.loss_weights(eos_w, cl)
            assert torch.equal(W, Wo) and torch.equal(C, Co)
    W, C, cat = loss_tables()
    for k, (name, lo, hi) in enumerate(oracle.LOSS_CATEGORIES):
        assert (cat[lo:hi + 1] == k).all()


def test_grad_arena_layout():
    from smer_music_generation_b200 import ScoreTransformer
    from smer_music_generation_b200.model import GradArena
    m = ScoreTransformer(309, 64, 4, 2, 2, 128, 32, 0.0, 0.0)
    a = GradArena(m)
    names = [n for n, _ in m.named_parameters()]
    assert sorted(a.order) == sorted(names)
    assert a.order[0].startswith("fc.") and a.order[-1] == "embedding.weight"
    seen = torch.zeros(a.total, dtype=torch.int32)
    for n in names:
        o, k = a.offsets[n]
        assert o % 64 == 0
        seen[o:o + k] += 1
        assert a.grads[n].shape == dict(m.named_parameters())[n].shape
    assert seen.max() == 1
    b, e = a.span("transformer.decoder.layers.1.")
    assert 0 < b < e <= a.total


def test_gradient_bucket_layouts_cover_the_arena():
    """trainer.GradBuckets: the default three buckets (decoder stack | encoder layers but the first | first encoder layer +
    embedding), its degenerate form for a one-layer encoder, and the per-layer layout are contiguous, disjoint and complete."""
    from smer_music_generation_b200 import ScoreTransformer
    from smer_music_generation_b200.trainer import GradArena, GradBuckets
    for ne, nd in ((1, 2), (3, 1), (2, 2)):
        m = ScoreTransformer(309, 32, 2, ne, nd, 64, 64, 0.0, 0.0)
        arena = GradArena(m)
        for nb in (0, -1, 2):
            gb = GradBuckets(m, arena, nb, None, comm_stream=None)
            cover = torch.zeros(arena.total, dtype=torch.int32)
            for _, b, e in gb.buckets:
                assert b < e
                cover[b:e] += 1
            assert int(cover.min()) == 1 and int(cover.max()) == 1, (ne, nd, nb)
            flat = [g for c, _, _ in gb.buckets for g in c]
            assert flat == gb.groups                      # buckets follow the order in which backward signals the groups
        gd = GradBuckets(m, arena, 0, None, comm_stream=None)
        assert gd.buckets[-1][0] == ["transformer.encoder.layers.0.", "embedding."]
        assert len(gd.buckets) == (2 if ne < 2 else 3)
        assert gd.buckets[0][0][0] == "fc." and all("encoder.layers" not in g for g in gd.buckets[0][0])


def test_install_as_reference_modules():
    import sys
    import smer_music_generation_b200 as pkg
    old = sys.modules.get("model")
    try:
        pkg.install_as_reference_modules()
        from model import ScoreTransformer
        assert ScoreTransformer is pkg.ScoreTransformer
    finally:
        if old is not None:
            sys.modules["model"] = old
        else:
            sys.modules.pop("model", None)
