#!/usr/bin/env python
"""Benchmark of the SMER transformer compute path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|decode] [--impl reference]

A "step" is one teacher-forced training step (forward + class-weighted cross-entropy + backward
+ Adam, dropout 0.1 on) over one synthetic batch of SMER tokens: BASELINE.json configs[1]
(B32 x S1024 (+T1024), bf16) per GPU.  For N > 1 (torchrun, one rank per GPU) every rank runs
that batch (weak scaling) and gradients are all-reduced over NCCL, overlapped with backward.
`--workload decode` measures batched KV-cached infilling (configs[3]) instead.
`--impl reference` times the CPU restatement of the reference path (oracle/) on the host cores.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC_TRAIN = "train_tokens_per_sec"
METRIC_DECODE = "infill_decode_tokens_per_sec"
CFG = dict(d=512, nhead=8, le=4, ld=4, ff=2048, max_len=2400, vocab=309)


def load_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("smer_oracle", os.path.join(ROOT, "oracle", "smer_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["smer_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (or None)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")))
        return t[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def train_flops(B, S, T, d=512, ff=2048, le=4, ld=4, V=309):
    """Algorithmic FLOPs of one forward; fwd+bwd = 3x."""
    enc_gemm = le * B * S * (8 * d * d + 4 * d * ff)
    enc_attn = le * B * 4 * S * S * d
    dec_gemm = ld * (B * T * (8 * d * d + 4 * d * d + 4 * d * ff) + B * S * 4 * d * d)
    dec_self = ld * B * 2 * T * T * d
    dec_cross = ld * B * 4 * T * S * d
    fc = 2 * B * T * d * V
    return enc_gemm + enc_attn + dec_gemm + dec_self + dec_cross + fc


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's restatement of train.py:722-786 on the host cores
# ------------------------------------------------------------------------------------------
def cpu_train_step_rate(B, S, T, steps, warmup, threads):
    O = load_oracle()
    torch.set_num_threads(threads)
    sd = O.random_state_dict(CFG["d"], CFG["nhead"], CFG["le"], CFG["ld"], CFG["ff"], CFG["max_len"], seed=0)
    W, C = O.loss_weights(0.8)
    src, tgt_in, tgt_out, sp, tp = O.synth_batch(B, S, T, seed=1234)
    ntok = int((~sp).sum() + (~tp).sum())
    names = [k for k in sd if k != "pos_enc.pe"]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss, grads, _, _ = O.train_step_grads(sd, src, tgt_in, tgt_out, sp, tp, CFG["nhead"], W, C)
        for k in names:
            O.adam_step(sd[k], grads[k], m[k], v[k], it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return ntok * len(times) / sum(times), sum(times) / len(times), ntok


def cpu_decode_rate(threads, spans=52):
    """The reference's decode loop (generation.py:523-687: the whole decoder re-run per token, no KV cache) as
    restated by the oracle, on ONE configs[3] piece, bounded to `spans` of its 52 masked spans
    (top-p 0.9 sampling like the GPU arm, spans capped at 24 tokens)."""
    O = load_oracle()
    torch.set_num_threads(threads)
    sd = O.random_state_dict(CFG["d"], CFG["nhead"], CFG["le"], CFG["ld"], CFG["ff"], CFG["max_len"], seed=0)
    ids = O.synth_piece(seed=0, n_bars=16, n_tracks=3, events_per_track_bar=6)
    src = O.mask_bar_and_track_ids(ids, [0, 1, 2], [4, 5, 6, 7], 3)
    targets = O.mask_targets(4, [0, 1, 2], 3)[:spans]
    t0 = time.perf_counter()
    import numpy as np
    tr = O.infill_decode(sd, src, targets, CFG["nhead"], mode="sample", top_p=0.9, rng=np.random.default_rng(7), max_span=24)
    dt = time.perf_counter() - t0
    return tr.generated / dt, dt, tr.generated, len(src)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.workload == "decode":
        rate, sec, ntok, S = cpu_decode_rate(threads)
        sample = (f"oracle port of generation.py:523-687 (uncached: whole decoder per token), one piece (S={S}), all 52 spans "
                  f"(<= 24 tokens each, encoder output computed once -- the reference re-encodes per token) = {ntok} tokens in {sec:.1f} s, top-p 0.9, fp32, {threads} torch threads")
        line = {"impl": "reference", "metric": METRIC_DECODE, "value": rate, "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": 1, "warmup": 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[3]: KV-cached infilling of 1024 pieces", "reference_sample": "one piece, 52 spans, on CPU"},
                "cpu_baseline": {"value": rate, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": rate, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    B, S, T = 2, args.seq, args.tgt
    rate, sec, ntok = cpu_train_step_rate(B, S, T, args.steps, min(args.warmup, 1), threads)
    sample = (f"oracle port of train.py:722-786 (fwd+loss+bwd+Adam, fp32, eval-mode arithmetic: no dropout RNG), "
              f"B{B} x S{S} (+T{T}) per step, {threads} torch threads")
    line = {"impl": "reference", "metric": METRIC_TRAIN, "value": rate, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: default SMER transformer, teacher-forced train step, B32 x S{S} (+T{T})",
                       "reference_sample": f"B{B} x S{S} (+T{T}) on CPU"},
            "cpu_baseline": {"value": rate, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def summarize_profile(prof):
    out = {}
    for label, evs in prof.items():
        ms = [s.elapsed_time(e) for s, e, _ in evs]
        work = sum(w for _, _, w in evs)
        out[label] = {"launch_groups": len(evs), "ms": sum(ms), "work": work}
    return out


def run_train(args):
    from smer_music_generation_b200 import ScoreTransformer, ops
    from smer_music_generation_b200.trainer import TrainEngine
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    O = load_oracle()
    B, S, T = args.batch, args.seq, args.tgt
    torch.manual_seed(1234)
    model = ScoreTransformer(CFG["vocab"], CFG["d"], CFG["nhead"], CFG["le"], CFG["ld"], CFG["ff"], CFG["max_len"], 0.1, 0.1,
                             compute_dtype=args.dtype).to(dev)
    for p in model.parameters():                         # train.py:261-263
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    model.train()
    eng = TrainEngine(model, lr=1e-4, eos_weight=0.8, process_group=pg)
    nb = 4
    host = [O.synth_batch(B, S, T, seed=1234 + 17 * rank + i) for i in range(nb)]
    host = [tuple(t.pin_memory() for t in b) for b in host]
    devb = [tuple(t.to(dev) for t in b) for b in host]
    ntok = [int((~b[3]).sum() + (~b[4]).sum()) for b in host]
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    use_graph = not args.no_graph
    launches_per_step = None
    if use_graph:
        l0 = ops.LAUNCHES
        try:
            eng.capture(B, S, T)                          # warm-up step + capture (two passes of launches)
            launches_per_step = (ops.LAUNCHES - l0) // 2 + 3  # + arena memset, loss-sum memset, counter bump
        except Exception as e:                            # e.g. a collective that refuses capture: run eagerly
            if rank == 0:
                print(f"[bench] graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            use_graph = False
    if world > 1:
        flag = torch.tensor([1 if use_graph else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_graph = bool(flag.item())
    if use_graph:
        step = lambda b: eng.step_graph(*b)
    else:
        step = lambda b: eng.step(*b)

    # ---- device-resident ("value") ----
    for i in range(args.warmup):
        step(devb[i % nb])
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    toks = 0
    for i in range(args.steps):
        step(devb[i % nb])
        toks += ntok[i % nb]
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    if use_graph:
        launches = launches_per_step * args.steps         # kernels in the captured step x replays
    else:
        launches = ops.LAUNCHES - l0 + 2 * args.steps      # + arena memset + loss-sum memset per step
    clk = clocks.stop() if clocks else None
    total_toks = sum_over_ranks(toks)
    value = total_toks / (ms * 1e-3)
    loss = eng.loss_value()

    # ---- end to end: pinned host buffers -> H2D -> step -> loss D2H, every step ----
    barrier()
    e0.record()
    toks2 = 0
    loss_host = torch.zeros(args.steps, 16, dtype=torch.float64).pin_memory()
    for i in range(args.steps):
        if use_graph:
            eng.step_graph(*host[i % nb])                # pinned host batch -> H2D into the step's input buffers
        else:
            eng.step(*tuple(t.to(dev, non_blocking=True) for t in host[i % nb]))
        loss_host[i].copy_(eng.sums, non_blocking=True)  # D2H of the step's 16 loss sums (128 B), every step
        toks2 += ntok[i % nb]
    e1.record()
    barrier()                                            # all losses have landed on the host here
    assert bool(torch.isfinite(loss_host[:, 0] / loss_host[:, 1]).all())
    ms2 = max_over_ranks(e0.elapsed_time(e1))
    e2e = sum_over_ranks(toks2) / (ms2 * 1e-3)

    # ---- per-kernel-family timing pass (events around every launch; not part of `value`) ----
    if use_graph:
        eng.release_graph()
    prof = None
    for _ in range(3):                   # three profiled steps, per family the fastest (a starved device inflates a pass)
        ops.PROFILE = {}
        torch.cuda._sleep(int(8e7))      # ~40 ms head start: the host enqueues the whole step ahead of the device, so the
        eng.step(*devb[0])               # events bracket device execution only, not launch latency
        torch.cuda.synchronize()
        one = summarize_profile(ops.PROFILE)
        prof = one if prof is None else {k: (one[k] if one[k]["ms"] < prof[k]["ms"] else prof[k]) for k in one}
    ops.PROFILE = None
    pk = peaks()
    tot_ms = sum(v["ms"] for v in prof.values())
    tensor_fams = {"gemm", "gemm_dw", "attn_fwd", "attn_bwd"}
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    if dom in tensor_fams:
        ach = d["work"] / (d["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"]}
    else:
        ach = d["work"] / (d["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    default_shape = (CFG["d"], CFG["le"], B, S, T) == (512, 4, 32, 1024, 1024)
    roof.update({"kernel": dom, "traffic": ncu_traffic(dom) if default_shape else None, "peak_source": pk["src"] + " (sustained)",
                 "share_of_step": d["ms"] / tot_ms, "launches_per_step": d["launch_groups"],
                 "avg_launch_ms": d["ms"] / d["launch_groups"]})
    step_flops = 3.0 * train_flops(B, S, T, CFG['d'], CFG['ff'], CFG['le'], CFG['ld'])
    fam = {k: {"ms": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4),
               **({"tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1)} if k in tensor_fams else
                  {"gbs": round(v["work"] / (v["ms"] * 1e-3) / 1e9, 1)} if v["work"] else {})}
           for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- the unchanged-train.py call pattern: module forward, SmerLoss, loss.backward(), FusedAdam.step() ----
    module_api = None
    if world == 1 and not args.no_module_api:
        from smer_music_generation_b200 import SmerLoss
        from smer_music_generation_b200.trainer import FusedAdam
        torch.manual_seed(7)
        m2 = ScoreTransformer(CFG["vocab"], CFG["d"], CFG["nhead"], CFG["le"], CFG["ld"], CFG["ff"], CFG["max_len"], 0.1, 0.1,
                              compute_dtype=args.dtype).to(dev)
        for p_ in m2.parameters():
            if p_.dim() > 1:
                torch.nn.init.xavier_normal_(p_)
        m2.train()
        crit = SmerLoss(CFG["vocab"], 0.8).to(dev)
        opt = FusedAdam(m2.parameters(), lr=1e-4)
        nsteps = 5

        def module_step(b):
            src, tin, tout, sp, tp = (t.to(dev, non_blocking=True) for t in b)
            opt.zero_grad(set_to_none=True)
            logits, _ = m2(src, tin, sp, tp, sp, "causal")
            loss, parts, denom = crit(logits, tout)
            loss.backward()
            opt.step()
            return loss

        for i in range(2):
            module_step(host[i % nb])
        torch.cuda.synchronize()
        e0.record()
        tk = 0
        for i in range(nsteps):
            l_ = module_step(host[i % nb])
            _ = l_.item()                                # train.py reads the loss every step (train.py:788-797)
            tk += ntok[i % nb]
        e1.record()
        torch.cuda.synchronize()
        mms = e0.elapsed_time(e1)
        module_api = {"value": tk / (mms * 1e-3), "unit": "tokens/s", "ms_per_step": mms / nsteps,
                      "api": "ScoreTransformer.forward + SmerLoss + loss.backward() + FusedAdam.step(), eager launches, "
                             "host batch in, loss.item() every step (the reference train loop's call pattern)"}
        del m2, opt, crit

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec, n = cpu_train_step_rate(2, S, T, 1, 1, threads)
        cpu = {"value": rate, "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": f"1 step of B2 x S{S} (+T{T}) after 1 warm-up, oracle port (fp32, no dropout RNG), {sec:.1f} s/step"}
    if rank == 0:
        line = {"metric": METRIC_TRAIN, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"{'configs[1]: default' if (CFG['d'], CFG['le'], B, S, T) == (512, 4, 32, 1024, 1024) else 'variant of configs[1]:'} SMER transformer (d{CFG['d']} h{CFG['nhead']} {CFG['le']}+{CFG['ld']} layers ff{CFG['ff']} V309), teacher-forced "
                                       f"train step fwd+loss+bwd+Adam, dropout 0.1, B{B}/GPU x S{S} (+T{T}), suffix padding "
                                       f"U[0.75L,L], tokens counted = non-pad src+tgt",
                           "l2": "working set per step (~3.5 GB activations) >> 126 MB L2; 4 rotating input batches",
                           "parallelism": f"dp{world}", "global_batch": B * world,
                           "launch": "whole step captured in one CUDA graph, replayed per step" if use_graph else "eager launches"},
                "clocks": clk,
                "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 128,
                        "ms_per_step": ms2 / args.steps, "api": "TrainEngine.step_graph/step over ScoreTransformer (pinned host batch in, loss sums out, async D2H every step)"},
                "e2e_module_api": module_api,
                "gpu_launches": launches,
                "roofline": roof,
                "cpu_baseline": cpu,
                "model_flops_per_step": step_flops,
                "model_tflops": step_flops * world / (ms / args.steps * 1e-3) / 1e12,
                "model_frac_of_bf16_sustained": step_flops / (ms / args.steps * 1e-3) / 1e12 / pk["tf_sust"],
                "kernel_families": fam, "loss": loss}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_decode(args):
    from smer_music_generation_b200 import ScoreTransformer, InfillDecoder, ops
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    O = load_oracle()
    torch.manual_seed(1234)
    model = ScoreTransformer(CFG["vocab"], CFG["d"], CFG["nhead"], CFG["le"], CFG["ld"], CFG["ff"], CFG["max_len"], 0.1, 0.1,
                             compute_dtype=args.dtype).to(dev)
    for p in model.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)
    model.eval()
    n_total = args.pieces
    per = n_total // world
    pieces, targets = [], []
    for i in range(per):
        ids = O.synth_piece(seed=rank * per + i, n_bars=16, n_tracks=3, events_per_track_bar=6)
        pieces.append(O.mask_bar_and_track_ids(ids, [0, 1, 2], [4, 5, 6, 7], 3))
        targets.append(O.mask_targets(4, [0, 1, 2], 3))
    dec = InfillDecoder(model, mode="top_p", top_p=0.9, seed=7, max_len=args.decode_len, splits=args.splits)
    dec.trace_intervals = True                           # device time of every 16-step graph launch (events only)
    res = None
    times, dev_times = [], []
    gens = []
    clocks = None
    for it in range(args.warmup + args.steps):
        if it == args.warmup and rank == 0:
            clocks = ClockSampler(local)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = dec.generate(pieces, targets, seq_base=rank * per)
        e1.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(e0.elapsed_time(e1))           # whole call: host packing, H2D, encoder, decode, D2H, unpacking
            dev_times.append(res["device_ms"])          # encoder + cross K/V + graph capture + decode loop
            gens.append(sum(res["generated"]))
    clk = clocks.stop() if clocks else None
    intervals = [round(x / 16, 3) for x in getattr(dec, "interval_ms", [])]      # of the last timed generate()
    ms = sum(times)
    ms_dev = sum(dev_times)
    toks = float(sum(gens))
    kl = dec.kernel_launches
    if world > 1:
        tm = torch.tensor([ms, ms_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = torch.tensor([toks], dtype=torch.float64, device=dev)
        dist.all_reduce(ts)
        ms, ms_dev, toks = float(tm[0].item()), float(tm[1].item()), float(ts.item())
    # roofline of the dominant kernel (attention over the cross-attention K/V): one extra eager step
    dec.use_graph = False
    dec.generate(pieces, targets, seq_base=rank * per, max_steps=64, check_every=64)
    pr = dec.profile_step()
    pk = peaks()
    ach = pr["cross"]["bytes"] / (pr["cross"]["ms"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "decode_attn (cross-attention K/V)", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
            "frac": ach / pk["hbm"], "traffic": ncu_traffic("decode_attn (cross-attention K/V)") if n_total == 1024 and world == 1 else None,
            "peak_source": pk["src"],
            "avg_launch_ms": pr["cross"]["ms"] / max(1, pr["cross"]["launches"]),
            "share_of_step": pr["cross"]["ms"] / pr["step_ms"], "eager_step_ms": pr["step_ms"],
            "self_attn_gbs": pr["self"]["bytes"] / (pr["self"]["ms"] * 1e-3) / 1e9}
    if rank == 0:
        S = dec.S
        line = {"metric": METRIC_DECODE, "value": toks / (ms_dev * 1e-3), "unit": "tokens/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"configs[3]: {n_total} independent 16-bar pieces (S<={S}), 4 bars x 3 tracks masked "
                                       f"(52 spans), KV cache, grammar-masked top-p 0.9 sampling, stream cap {args.decode_len}",
                           "timed": "value: pieces resident on the device -> encoder, cross-KV, decode loop (CUDA events; the 16-step CUDA graph is "
                                    "captured in the warm-up call and replayed by the timed calls: same shapes, same weights); "
                                    "e2e: whole InfillDecoder.generate() incl. host packing, H2D, D2H of the token streams"},
                "e2e": {"value": toks / (ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": dec.h2d_bytes,
                        "d2h_bytes_per_step": dec.d2h_bytes, "ms_per_step": ms / args.steps},
                "clocks": clk, "gpu_launches": kl, "decode_steps": res["steps"], "roofline": roof,
                "step_ms_graph": ms_dev / args.steps / max(1, res["steps"]),
                "step_ms_by_interval": intervals[:40],
                "hbm_bytes_per_step_algorithmic": pr["cross"]["bytes"] + pr["self"]["bytes"]}
        line["cpu_baseline"] = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, sec, ntok, S0 = cpu_decode_rate(threads)
            line["cpu_baseline"] = {"value": rate, "unit": "tokens/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle port of generation.py:523-687 (uncached), one piece (S={S0}), 52 spans (encoder hoisted) = "
                                              f"{ntok} tokens in {sec:.1f} s, top-p 0.9, fp32"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "decode"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--tgt", type=int, default=1024)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--pieces", type=int, default=1024)
    ap.add_argument("--decode-len", type=int, default=512)
    ap.add_argument("--splits", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-module-api", action="store_true")
    ap.add_argument("--d-model", type=int, default=512)
    ap.add_argument("--nhead", type=int, default=8)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--ff", type=int, default=2048)
    args = ap.parse_args()
    CFG.update(d=args.d_model, nhead=args.nhead, le=args.layers, ld=args.layers, ff=args.ff,
               max_len=max(2400, args.seq, args.tgt))
    if args.warmup < 3 and args.impl == "ours" and args.workload == "train":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "decode":
        return run_decode(args)
    return run_train(args)


if __name__ == "__main__":
    main()
