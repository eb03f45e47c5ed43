"""B200-native implementation of the SMER transformer compute path (reference:
ruiguo-bio/smer_music_generation -- model.py / transformer.py / generation.py).

    from smer_music_generation_b200 import ScoreTransformer      # same API as model.ScoreTransformer

All arithmetic runs in libsmer_b200.so (hand-written sm_100a CUDA, C ABI in include/smer_b200.h);
there is no CPU or eager-PyTorch fallback.
"""
from .model import ScoreTransformer          # noqa: F401
from .loss import SmerLoss, SmerAccuracy, loss_tables      # noqa: F401
from .decode import InfillDecoder            # noqa: F401
from . import spans                          # noqa: F401

__all__ = ["ScoreTransformer", "SmerLoss", "SmerAccuracy", "loss_tables", "InfillDecoder", "install_as_reference_modules"]


def install_as_reference_modules() -> None:
    """Makes `from model import ScoreTransformer` (train.py:19, evaluation.py:21) resolve to this
    package, so the reference's scripts run unchanged on the B200 path."""
    import sys
    import types
    from . import model as _m
    shim = types.ModuleType("model")
    shim.ScoreTransformer = _m.ScoreTransformer
    shim.PositionalEncoding = _m._PositionalEncoding
    sys.modules["model"] = shim
