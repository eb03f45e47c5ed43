#!/usr/bin/env python
"""Secondary bar of SURVEY.md §8(d): the same training step written with stock PyTorch ops (the oracle's
restatement of model.py / transformer.py / the loss of train.py + torch.optim.Adam) ON THE SAME B200, in fp32
and under bf16 autocast.  Eval-mode arithmetic (no dropout RNG), which favours this baseline.  Not part of
bench.py; prints one JSON line per precision."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_oracle  # noqa: E402

O = load_oracle()
dev = torch.device("cuda:0")
B, S, T = int(os.environ.get("B", 32)), 1024, 1024
steps, warm = 5, 2
src, tin, tout, sp, tp = (t.to(dev) for t in O.synth_batch(B, S, T, seed=1234))
W, C = (t.to(dev) for t in O.loss_weights(0.8))
ntok = int((~sp).sum() + (~tp).sum())
mask = O.nopeek_mask(T)[None].to(dev)
for mode in ("fp32", "bf16_autocast"):
    sd = {k: v.to(dev) for k, v in O.random_state_dict(512, 8, 4, 4, 2048, 2400, seed=0).items()}
    params = {k: torch.nn.Parameter(v) for k, v in sd.items() if k != "pos_enc.pe"}
    leaf = dict(params, **{"pos_enc.pe": sd["pos_enc.pe"]})
    opt = torch.optim.Adam(params.values(), lr=1e-4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(warm + steps):
        if it == warm:
            torch.cuda.synchronize()
            e0.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
            logits, _ = O.score_transformer_forward(leaf, src, tin, 8, sp, tp, sp, mask)
        loss, _, _ = O.smer_loss(logits.reshape(-1, logits.shape[-1]).float(), tout.reshape(-1), W, C)
        loss.backward()
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"impl": "stock PyTorch ops on the same B200 (oracle restatement, eval-mode arithmetic)", "precision": mode,
                      "workload": f"B{B} x S{S} (+T{T}) train step (fwd + loss + bwd + torch.optim.Adam)", "ms_per_step": ms,
                      "tokens_per_s": ntok / (ms * 1e-3), "loss": float(loss), "torch": torch.__version__,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
    del sd, params, leaf, opt, logits, loss
    torch.cuda.empty_cache()
