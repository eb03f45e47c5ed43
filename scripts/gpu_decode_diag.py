import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import load_oracle, CFG
from smer_music_generation_b200 import ScoreTransformer, InfillDecoder
O = load_oracle()
torch.manual_seed(1234)
dev = torch.device("cuda:0")
model = ScoreTransformer(309, 512, 8, 4, 4, 2048, 2400, 0.1, 0.1, compute_dtype="bf16").to(dev)
for p in model.parameters():
    if p.dim() > 1:
        torch.nn.init.xavier_normal_(p)
model.eval()
pieces, targets = [], []
for i in range(64):
    ids = O.synth_piece(seed=i, n_bars=16, n_tracks=3, events_per_track_bar=6)
    pieces.append(O.mask_bar_and_track_ids(ids, [0, 1, 2], [4, 5, 6, 7], 3))
    targets.append(O.mask_targets(4, [0, 1, 2], 3))
for mode in ("top_p", "multinomial", "greedy"):
    for graph in (True, False):
        dec = InfillDecoder(model, mode=mode, top_p=0.9, seed=7, max_len=512, use_graph=graph)
        res = dec.generate(pieces, targets)
        g = np.array(res["generated"]); l = np.array([len(s) for s in res["streams"]])
        print(mode, "graph" if graph else "eager", "steps", res["steps"], "gen mean", g.mean(), "min", g.min(), "max", g.max(),
              "len mean", l.mean(), "done", sum(res["done"]), "first stream head", res["streams"][0][:12], flush=True)
