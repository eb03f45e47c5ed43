"""Times the phases of InfillDecoder.generate() for 1024 pieces (encoder / cross-KV / decode loop)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import load_oracle
from smer_music_generation_b200 import ScoreTransformer, InfillDecoder
from smer_music_generation_b200 import decode as D
O = load_oracle()
torch.manual_seed(1234)
dev = torch.device("cuda:0")
model = ScoreTransformer(309, 512, 8, 4, 4, 2048, 2400, 0.1, 0.1, compute_dtype="bf16").to(dev)
for p in model.parameters():
    if p.dim() > 1:
        torch.nn.init.xavier_normal_(p)
model.eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pieces, targets = [], []
for i in range(n):
    ids = O.synth_piece(seed=i, n_bars=16, n_tracks=3, events_per_track_bar=6)
    pieces.append(O.mask_bar_and_track_ids(ids, [0, 1, 2], [4, 5, 6, 7], 3))
    targets.append(O.mask_targets(4, [0, 1, 2], 3))
for rep in range(2):
    dec = InfillDecoder(model, mode="top_p", top_p=0.9, seed=7, max_len=512, use_graph=True)
    torch.cuda.synchronize(); t0 = time.time()
    res = dec.generate(pieces, targets)
    torch.cuda.synchronize(); t1 = time.time()
    print("rep", rep, "wall ms", (t1 - t0) * 1e3, "device_ms", res["device_ms"], "steps", res["steps"],
          "generated", sum(res["generated"]), flush=True)
# encoder alone
src = torch.zeros(n, 848, dtype=torch.long, device=dev)
for i, p_ in enumerate(pieces):
    src[i, :len(p_)] = torch.tensor(p_, device=dev)
pad = src == 0
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    mem = D._encode(model, src, pad) if hasattr(D, "_encode") else None
    torch.cuda.synchronize(); print("encode ms", (time.time() - t0) * 1e3, flush=True)
# same decoder object reused (what bench.py does)
dec = InfillDecoder(model, mode="top_p", top_p=0.9, seed=7, max_len=512)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    res = dec.generate(pieces, targets)
    torch.cuda.synchronize(); t1 = time.time()
    print("reuse rep", rep, "wall ms", (t1 - t0) * 1e3, "device_ms", res["device_ms"], "steps", res["steps"], flush=True)
