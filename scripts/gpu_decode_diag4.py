"""Device time of every 16-step check interval of the decode loop (how the step time grows with the self cache)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import load_oracle
from smer_music_generation_b200 import ScoreTransformer, InfillDecoder
O = load_oracle()
torch.manual_seed(1234)
dev = torch.device("cuda:0")
model = ScoreTransformer(309, 512, 8, 4, 4, 2048, 2400, 0.1, 0.1, compute_dtype="bf16").to(dev)
for p in model.parameters():
    if p.dim() > 1:
        torch.nn.init.xavier_normal_(p)
model.eval()
n = 1024
pieces, targets = [], []
for i in range(n):
    ids = O.synth_piece(seed=i, n_bars=16, n_tracks=3, events_per_track_bar=6)
    pieces.append(O.mask_bar_and_track_ids(ids, [0, 1, 2], [4, 5, 6, 7], 3))
    targets.append(O.mask_targets(4, [0, 1, 2], 3))
dec = InfillDecoder(model, mode="top_p", top_p=0.9, seed=7, max_len=512)
dec.trace_intervals = True
for rep in range(3):
    res = dec.generate(pieces, targets)
    iv = dec.interval_ms
    print("rep", rep, "device_ms", round(res["device_ms"], 1), "per-step ms by interval:",
          " ".join(f"{x / 16:.2f}" for x in iv), flush=True)
