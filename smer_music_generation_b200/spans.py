"""Host-side span surgery around the batched infilling decode, on token ids and vectorised with numpy:
what generation.py does with Python list pops / inserts on event strings before and after the decode loop.

* mask_bar_and_track  -- generation.py:248-341: in every selected (bar, track) the note content and each trailing
  track-control token (3 controls, plus the tensile token after the last track) become one `m_0` each;
* mask_targets        -- the per-span target kinds generation_all derives from the same selection (generation.py:487-494);
* restore_marked_input -- generation.py:417-465: every `m_0` of the masked source is replaced by the tokens generated
  for it (the decoder stream is `m_0 span m_0 span ...`).

Token ids are the fixed WordVocab contract (vocab.py:114-310; pinned by tests/golden/vocab.npz)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

M0, BAR, TRACK0, N_TRACK_TOKENS = 2, 3, 4, 3
TENSILE_LO, TENSILE_HI = 296, 307
TRACK_CONTROLS = 3                      # density, occupation, polyphony after every track's notes (control mode 2)


TIME_SIG_4_4 = 7                        # vocab.py:26: time_signature_token = ['4/4', '3/4', '2/4', '6/8'] -> ids 7..10


def no_whole_duration(ids) -> bool:
    """generation.py:504-507: a whole note is allowed only when the piece's time signature (its first token) is
    N/4 with N >= 4 -- of the vocabulary's four signatures that is '4/4' alone."""
    return int(np.asarray(ids).reshape(-1)[0]) != TIME_SIG_4_4


def _track_count(ids: np.ndarray) -> int:
    present = np.unique(ids[(ids >= TRACK0) & (ids < TRACK0 + N_TRACK_TOKENS)])
    return int(present.size)


def _span_pairs(ids: np.ndarray, mask_tracks: Sequence[int], mask_bars: Sequence[int]) -> Tuple[List[Tuple[int, int]], List[int], List[int]]:
    n_tracks = _track_count(ids)
    marks = np.flatnonzero((ids == BAR) | ((ids >= TRACK0) & (ids < TRACK0 + n_tracks)))
    marks = np.append(marks, ids.size)
    # generation.py:270-288: groups of (n_tracks + 1) marks after the first one: [first track .. next bar / end]
    tail = marks[1:]
    n_bars = tail.size // (n_tracks + 1)
    grid = tail[: n_bars * (n_tracks + 1)].reshape(n_bars, n_tracks + 1)
    starts, ends = grid[:, :-1] + 1, grid[:, 1:]
    pairs: List[Tuple[int, int]] = []
    track_names: List[int] = []
    bar_names: List[int] = []
    mask_tracks = set(int(t) for t in mask_tracks)
    for b in mask_bars:
        for tpos in range(n_tracks):
            if tpos not in mask_tracks:
                continue
            ts, te = int(starts[b, tpos]), int(ends[b, tpos])
            bar_names.append(int(b))
            track_names.append(tpos)
            tensile_end = 1 if TENSILE_LO <= ids[te - 1] <= TENSILE_HI else 0
            tok_start = ts + TRACK_CONTROLS
            tok_end = te - TRACK_CONTROLS - tensile_end
            pairs.append((tok_start, tok_end))
            for i in range(TRACK_CONTROLS + tensile_end):
                pairs.append((tok_end + i, tok_end + 1 + i))
    return pairs, track_names, bar_names


def mask_bar_and_track(ids, mask_tracks: Sequence[int], mask_bars: Sequence[int]):
    """-> (src ids with the selected spans collapsed to m_0, mask_track_names, mask_bar_names), the return value of
    generation.mask_bar_and_track (generation.py:341) on ids instead of event strings."""
    ids = np.asarray(ids, dtype=np.int64)
    pairs, track_names, bar_names = _span_pairs(ids, mask_tracks, mask_bars)
    if not pairs:
        return ids.copy(), track_names, bar_names
    pairs.sort()
    keep = np.ones(ids.size, dtype=bool)
    out = ids.copy()
    for a, b in pairs:                     # a handful of spans per piece; the per-token work below is vectorised
        if b > a:
            keep[a + 1:b] = False
            out[a] = M0
        # an empty content span (a == b) still gets its m_0: inserted below
    empties = np.asarray([a for a, b in pairs if b == a], dtype=np.int64)
    res = out[keep]
    if empties.size:
        # positions in the compacted array: number of kept tokens before each empty span
        pos = np.cumsum(keep)[empties - 1] if empties.min() > 0 else np.asarray([int(keep[:a].sum()) for a in empties])
        res = np.insert(res, pos, M0)
    return res, track_names, bar_names


def mask_targets(ids, tracks_to_generate: Sequence[int], bars_to_generate: Sequence[int]) -> List[str]:
    """generation.py:487-494: kinds of the spans in decode order ('r' notes, 'd' / 'o' / 'p' controls, 't' tensile)."""
    n_tracks = _track_count(np.asarray(ids, dtype=np.int64))
    out: List[str] = []
    for _ in bars_to_generate:
        for track in tracks_to_generate:
            out.extend(["r", "d", "o", "p"])
            if track == n_tracks - 1:
                out.append("t")
    return out


def restore_marked_input(src_ids, generated_ids) -> np.ndarray:
    """generation.py:417-465 on ids: the k-th m_0 of `src_ids` is replaced by the tokens that follow the k-th m_0 of
    `generated_ids` (up to the next m_0 / the end).  Source masks beyond the generated spans stay."""
    src = np.asarray(src_ids, dtype=np.int64)
    gen = np.asarray(generated_ids, dtype=np.int64)
    gpos = np.flatnonzero(gen == M0)
    if gpos.size == 0:
        return src.copy()
    spans = np.split(gen, gpos)[1:]                         # each starts with its m_0
    spos = np.flatnonzero(src == M0)
    n = min(len(spans), spos.size)
    pieces, last = [], 0
    for k in range(n):
        pieces.append(src[last:spos[k]])
        pieces.append(spans[k][1:])
        last = spos[k] + 1
    pieces.append(src[last:])
    return np.concatenate(pieces)
