#!/usr/bin/env python
"""Benchmark of the SMER transformer compute path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|decode] [--impl reference]

A "step" is one teacher-forced training step (forward + class-weighted cross-entropy + backward
+ Adam, dropout 0.1 on) over one synthetic batch of SMER tokens: BASELINE.json configs[1]
(B32 x S1024 (+T1024), bf16) per GPU.  For N > 1 (torchrun, one rank per GPU) every rank runs
that batch (weak scaling) and gradients are all-reduced over NCCL, overlapped with backward.
The same JSON line carries sub-records for the other configurations BASELINE.json names:
  "decode"  configs[3]: batched KV-cached infilling of 1024 pieces sharded over the N GPUs (value, e2e,
            whole-step HBM fraction, dominant-kernel roofline, clocks);
  "c3"      configs[2]'s per-GPU shape B64 x S2048 (+T2048), with and without the gradient all-reduce;
  "c5_attention"  configs[4]: d768 / 12 heads, S = T = 4096 attention forward+backward sweep, B/GPU in {1,2,4,8};
  "c1"      configs[0]: the reference's own generation_all (unchanged) driving the drop-in module on the GPU;
  "torch_gpu_baseline"  the unmodified reference modules through stock PyTorch on the same B200.
`--workload decode` prints the decode record as the top-level line instead.
`--impl reference` times the UNMODIFIED reference (baseline/_ref, see baseline/reference_arm.py) on the host cores.
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
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")))
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
# process context (one process per GPU; torchrun sets RANK / LOCAL_RANK / WORLD_SIZE)
# ------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, need_cuda=True):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.pg = None
        self.dev = None
        if need_cuda:
            torch.cuda.set_device(self.local)
            self.dev = torch.device("cuda", self.local)
            if self.world > 1:
                import torch.distributed as dist
                # (NCCL's CTA count is left at its default: capping it to 2 / 4 / 8 CTAs made the bucket all-reduces too slow to
                #  hide behind the backward kernels -- exposed time 0.76 / 0.72 / 0.47 ms against 0.30 ms at 16, 2 GPUs)
                # The collectives run on a high-priority stream: the persistent GEMM / attention grids hold every SM, so a
                # default-priority all-reduce kernel only gets its CTAs placed when a compute kernel happens to leave room.
                opts = None
                if os.environ.get("SMER_NCCL_HIGH_PRIORITY", "1") != "0":
                    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                dist.init_process_group("nccl", device_id=self.dev, pg_options=opts)
                self.pg = dist.group.WORLD

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t)
        return float(t.item())

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def load_reference_arm():
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_arm", os.path.join(ROOT, "baseline", "reference_arm.py"))
    mod = sys.modules.get("reference_arm")
    if mod is None:
        mod = importlib.util.module_from_spec(spec)
        sys.modules["reference_arm"] = mod
        spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (baseline/_ref) when it is staged, else the oracle's port of it
# ------------------------------------------------------------------------------------------
def cpu_train_rate(cfg, B, S, T, steps, warmup, threads, budget_s=150.0):
    """-> dict(value tokens/s, kind, sample).  One step = loop body of train.py:702-797 on B x S (+T) tokens."""
    O = load_oracle()
    R = load_reference_arm()
    batches = [O.synth_batch(B, S, T, seed=1234 + i) for i in range(2)]
    if R.reference_dir() is not None:
        r = R.time_train(cfg, batches, steps, warmup, threads, "cpu", dropout=0.1)
        return {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": threads, "kind": "reference",
                "sample": f"{r['steps']} step(s) of B{B} x S{S} (+T{T}) after {warmup} warm-up, the unmodified reference "
                          f"model.ScoreTransformer + its 12 nn.CrossEntropyLoss criteria + torch.optim.Adam (train.py:702-797), "
                          f"fp32, dropout 0.1 on, {r['s_per_step']:.1f} s/step, {threads} torch threads",
                "s_per_step": r["s_per_step"]}
    torch.set_num_threads(threads)
    sd = O.random_state_dict(cfg["d"], cfg["nhead"], cfg["le"], cfg["ld"], cfg["ff"], cfg["max_len"], seed=0)
    W, C = O.loss_weights(0.8)
    names = [k for k in sd if k != "pos_enc.pe"]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    times, toks = [], 0
    for it in range(warmup + steps):
        src, tgt_in, tgt_out, sp, tp = batches[it % 2]
        t0 = time.perf_counter()
        loss, grads, _, _ = O.train_step_grads(sd, src, tgt_in, tgt_out, sp, tp, cfg["nhead"], W, C)
        for k in names:
            O.adam_step(sd[k], grads[k], m[k], v[k], it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            toks += int((~sp).sum() + (~tp).sum())
    return {"value": toks / sum(times), "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{len(times)} step(s) of B{B} x S{S} (+T{T}) after {warmup} warm-up, oracle port of train.py:722-786 "
                      f"(fp32, no dropout RNG), {sum(times) / len(times):.1f} s/step, {threads} torch threads",
            "s_per_step": sum(times) / len(times)}


def c1_piece():
    """configs[0]: one 16-bar 3-track piece, bars 4..7 x all tracks masked (52 spans)."""
    O = load_oracle()
    return O.synth_piece(seed=0, n_bars=16, n_tracks=3, events_per_track_bar=4), [0, 1, 2], [4, 5, 6, 7]


def cpu_decode_rate(cfg, threads, bars=(4, 5, 6, 7)):
    """configs[0] on the host cores: the reference's generation_all, unchanged (generation.py:468-696: whole encoder
    and decoder re-run per token, NumPy sampling), greedy.  Falls back to the oracle's port of that loop."""
    O = load_oracle()
    R = load_reference_arm()
    ids, tracks, _ = c1_piece()
    torch.set_num_threads(threads)
    if R.reference_dir() is not None:
        mods = R.load()
        m = R.build_model(mods[0], cfg, 0.1).eval()
        with torch.no_grad():
            restored, calls, last, dt = R.run_generation_all(m, ids, tracks, list(bars), "cpu")
        return {"value": calls / dt, "unit": "tokens/s", "cores": threads, "kind": "reference",
                "sample": f"generation.generation_all unchanged (no KV cache: encoder + decoder per token), one {len(ids)}-token piece, "
                          f"{len(bars)} bars x 3 tracks masked, greedy, {calls} model calls (= sampled tokens) in {dt:.1f} s, fp32, {threads} torch threads",
                "seconds": dt, "tokens": calls}
    sd = O.random_state_dict(cfg["d"], cfg["nhead"], cfg["le"], cfg["ld"], cfg["ff"], cfg["max_len"], seed=0)
    src = O.mask_bar_and_track_ids(ids, tracks, list(bars), 3)
    targets = O.mask_targets(len(bars), tracks, 3)
    t0 = time.perf_counter()
    tr = O.infill_decode(sd, src, targets, cfg["nhead"], mode="greedy")
    dt = time.perf_counter() - t0
    return {"value": tr.generated / dt, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"oracle port of generation.py:523-687 (uncached decoder per token, encoder hoisted), one piece, "
                      f"{len(targets)} spans, greedy, {tr.generated} tokens in {dt:.1f} s, fp32, {threads} torch threads",
            "seconds": dt, "tokens": tr.generated}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = dict(CFG)
    if args.workload == "decode":
        r = cpu_decode_rate(cfg, threads)
        line = {"impl": "reference", "metric": METRIC_DECODE, "value": r["value"], "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": 1, "warmup": 0, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[3]: KV-cached infilling of 1024 pieces", "reference_sample": "configs[0]: one piece, 52 spans, on CPU"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    B, S, T = 2, args.seq, args.tgt
    probe = cpu_train_rate(cfg, B, S, T, 1, 0, threads)             # sizes the sample: the run must end within minutes
    steps = max(1, min(args.steps, int(150.0 / max(probe["s_per_step"], 1e-3))))
    warm = 1 if steps * probe["s_per_step"] < 100.0 else 0
    r = cpu_train_rate(cfg, B, S, T, steps, warm, threads)
    line = {"impl": "reference", "metric": METRIC_TRAIN, "value": r["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: default SMER transformer, teacher-forced train step, B32 x S{S} (+T{T})",
                       "reference_sample": f"B{B} x S{S} (+T{T}) per step on CPU ({steps} of the requested {args.steps} steps: bounded sample)",
                       "same_config": False},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:
        d = cpu_decode_rate(cfg, threads)
        line["decode"] = {"metric": METRIC_DECODE, "impl": "reference", "value": d["value"], "unit": "tokens/s",
                          "cpu_baseline": {k: d[k] for k in ("value", "unit", "cores", "kind", "sample")}}
    except Exception as e:
        line["decode"] = {"error": f"{type(e).__name__}: {e}"}
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


def attn_executed_fraction(src_pad, tgt_pad, S, T):
    """Share of the nominal attention FLOPs (4*Lq*Lk*dh per (b,h), causal = half) the kernels execute when they stop
    at kv_len: encoder self Lq*len, causal decoder self sum_i min(i+1, len), cross T*len_src; all query rows are run."""
    ls = (~src_pad).sum(1).double()
    lt = (~tgt_pad).sum(1).double()
    B = src_pad.shape[0]
    nominal = B * (S * S + 0.5 * T * T + T * S)
    causal = (lt * (lt + 1) / 2 + (T - lt) * lt).sum()
    executed = (S * ls).sum() + causal + (T * ls).sum()
    return float(executed / nominal)


def build_model(cfg, dtype, dev, dropout=0.1, seed=1234, world=1):
    from smer_music_generation_b200 import ScoreTransformer
    torch.manual_seed(seed)
    model = ScoreTransformer(cfg["vocab"], cfg["d"], cfg["nhead"], cfg["le"], cfg["ld"], cfg["ff"], cfg["max_len"], dropout,
                             dropout, compute_dtype=dtype).to(dev)
    for p in model.parameters():                         # train.py:261-263
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)
    if world > 1:
        import torch.distributed as dist
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    return model


def timed_train(ctx, cfg, dtype, B, S, T, steps, warmup, pg, use_graph=True, e2e=False, profile=False, clocks=False,
                packed=True):
    """One training workload on every rank: `steps` timed steps after `warmup`, CUDA events, max over ranks."""
    from smer_music_generation_b200 import ops
    from smer_music_generation_b200.trainer import TrainEngine
    O = load_oracle()
    dev, rank = ctx.dev, ctx.rank
    world = ctx.world if pg is not None else 1
    model = build_model(cfg, dtype, dev, world=world).train()
    eng = TrainEngine(model, lr=1e-4, eos_weight=0.8, process_group=pg, n_buckets=int(os.environ.get("SMER_DP_BUCKETS", "0")))
    nb = 4
    host = [O.synth_batch(B, S, T, seed=1234 + 17 * rank + i) for i in range(nb)]
    host = [tuple(t.pin_memory() for t in b) for b in host]
    devb = [tuple(t.to(dev) for t in b) for b in host]
    ntok = [int((~b[3]).sum() + (~b[4]).sum()) for b in host]
    out = {"h2d": sum(t.numel() * t.element_size() for t in host[0])}
    packed = packed and dtype == "bf16" and cfg["d"] // cfg["nhead"] == 64
    out["packed"] = packed
    lens = [((~b[3]).sum(1).tolist(), (~b[4]).sum(1).tolist()) for b in host]     # HOST lengths, as a collate step has them
    if packed:
        from smer_music_generation_b200.model import PackedBatch
        up = lambda n: (n + 127) // 128 * 128
        rows_s, rows_t = max(up(sum(l[0])) for l in lens), max(up(sum(l[1])) for l in lens)
        out["h2d"] = sum(t.numel() * t.element_size() for t in host[0][:3]) + 8 * (B + 1)
    launches_per_step = None
    if use_graph:
        l0 = ops.LAUNCHES
        try:
            if packed:
                eng.capture_packed(B, rows_s, rows_t, S, T)
            else:
                eng.capture(B, S, T)                      # warm-up pass + capture (two passes of launches)
            launches_per_step = (ops.LAUNCHES - l0 - 1) // 2 + 3  # + arena memset, loss-sum memset, counter bump
        except Exception as e:                            # e.g. a collective that refuses capture: run eagerly
            if rank == 0:
                print(f"[bench] graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            use_graph = False
    if ctx.world > 1:
        import torch.distributed as dist
        flag = torch.tensor([1 if use_graph else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_graph = bool(flag.item())
    if packed:
        if use_graph:
            stepi = lambda bt, i: eng.step_graph_packed(bt[0], bt[1], bt[2], lens[i][0], lens[i][1])
        else:
            stepi = lambda bt, i: eng.step_packed(PackedBatch.pack(bt[0].to(dev, non_blocking=True), bt[1].to(dev, non_blocking=True),
                                                                   bt[2].to(dev, non_blocking=True), lens[i][0], lens[i][1]))
    else:
        stepi = (lambda bt, i: eng.step_graph(*bt)) if use_graph else (lambda bt, i: eng.step(*tuple(t.to(dev, non_blocking=True) for t in bt)))
    step = lambda b: stepi(b[0], b[1])
    devb = [(b, i) for i, b in enumerate(devb)]
    for i in range(warmup):
        step(devb[i % nb])
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if (clocks and rank == 0) else None
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    toks = 0
    for i in range(steps):
        step(devb[i % nb])
        toks += ntok[i % nb]
    e1.record()
    ctx.barrier()
    ms = ctx.max(e0.elapsed_time(e1))
    out["launches"] = launches_per_step * steps if use_graph else ops.LAUNCHES - l0 + 2 * steps
    out["clocks"] = sampler.stop() if sampler else None
    total = ctx.sum(toks) if pg is not None else float(toks)
    out.update(ms=ms, ms_per_step=ms / steps, value=total / (ms * 1e-3), loss=eng.loss_value(), use_graph=use_graph,
               exec_frac=attn_executed_fraction(host[0][3], host[0][4], S, T))
    tl = os.environ.get("SMER_TIMELINE")
    if tl:
        # kernel timeline of three more steps (CUPTI through torch.profiler; nsys is not in the image): every rank runs the
        # steps, rank 0 writes {name, stream, start us, duration us} per kernel -- untimed, after `value`
        from torch.profiler import profile as _tprofile, ProfilerActivity
        ctx.barrier()
        if rank == 0:
            with _tprofile(activities=[ProfilerActivity.CUDA]) as tp:
                for i in range(3):
                    step(devb[i % nb])
                torch.cuda.synchronize()
            evs = [dict(name=e.name, stream=getattr(e, "device_resource_id", None), ts=e.time_range.start, dur=e.time_range.elapsed_us())
                   for e in tp.events() if e.device_type == torch.autograd.DeviceType.CUDA]
            with open(tl, "w") as f:
                json.dump(evs, f)
        else:
            for i in range(3):
                step(devb[i % nb])
            torch.cuda.synchronize()
        ctx.barrier()
    if e2e:
        # end to end: pinned host buffers -> H2D -> step -> loss D2H, every step
        ctx.barrier()
        e0.record()
        toks2 = 0
        loss_host = torch.zeros(steps, 16, dtype=torch.float64).pin_memory()
        for i in range(steps):
            stepi(host[i % nb], i % nb)                   # pinned host batch -> H2D into the step's input buffers
            loss_host[i].copy_(eng.sums, non_blocking=True)   # D2H of the step's 16 loss sums (128 B), every step
            toks2 += ntok[i % nb]
        e1.record()
        ctx.barrier()                                      # all losses have landed on the host here
        assert bool(torch.isfinite(loss_host[:, 0] / loss_host[:, 1]).all())
        ms2 = ctx.max(e0.elapsed_time(e1))
        out.update(e2e_value=ctx.sum(toks2) / (ms2 * 1e-3), e2e_ms_per_step=ms2 / steps)
    if use_graph:
        eng.release_graph()
    if profile:
        # per-kernel-family timing pass (events around every launch; not part of `value`)
        prof = None
        for _ in range(3):               # three profiled steps, per family the fastest (a starved device inflates a pass)
            ops.PROFILE = {}
            torch.cuda._sleep(int(8e7))  # ~40 ms head start: the host enqueues the whole step ahead of the device, so the
            if packed:                   # events bracket device execution only, not launch latency
                eng.step_packed(PackedBatch.pack(devb[0][0][0], devb[0][0][1], devb[0][0][2], lens[0][0], lens[0][1]))
            else:
                eng.step(*devb[0][0])
            torch.cuda.synchronize()
            one = summarize_profile(ops.PROFILE)
            prof = one if prof is None else {k: (one[k] if one[k]["ms"] < prof[k]["ms"] else prof[k]) for k in one}
        ops.PROFILE = None
        out["profile"] = prof
    out["host"], out["ntok"] = host, ntok
    del eng, model, devb
    torch.cuda.empty_cache()
    return out


def module_api_record(ctx, cfg, dtype, host, ntok):
    """The unchanged-train.py call pattern: module forward, SmerLoss, loss.backward(), FusedAdam.step(), loss.item()."""
    from smer_music_generation_b200 import SmerLoss
    from smer_music_generation_b200.trainer import FusedAdam
    dev = ctx.dev
    m2 = build_model(cfg, dtype, dev, seed=7).train()
    crit = SmerLoss(cfg["vocab"], 0.8).to(dev)
    opt = FusedAdam(m2.parameters(), lr=1e-4)
    nsteps, nb = 5, len(host)

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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tk, first, last = 0, None, None
    for i in range(nsteps):
        l_ = module_step(host[0])
        last = l_.item()                                 # train.py reads the loss every step (train.py:788-797)
        first = last if first is None else first
        tk += ntok[0]
    e1.record()
    torch.cuda.synchronize()
    mms = e0.elapsed_time(e1)
    del m2, opt, crit
    torch.cuda.empty_cache()
    return {"value": tk / (mms * 1e-3), "unit": "tokens/s", "ms_per_step": mms / nsteps, "loss_first": first, "loss_last": last,
            "api": "ScoreTransformer.forward + SmerLoss + loss.backward() + FusedAdam.step(), eager launches, "
                   "host batch in, loss.item() every step (the reference train loop's call pattern)"}


def torch_gpu_record(ctx, cfg, B, S, T):
    """The secondary bar of SURVEY 8(d): the unmodified reference modules (stock PyTorch ops: cuBLAS GEMMs, unfused
    softmax / dropout / LayerNorm, 12 CE passes, torch.optim.Adam) on the same B200, fp32 and bf16 autocast."""
    R = load_reference_arm()
    if R.reference_dir() is None:
        return {"unavailable": "baseline/_ref is not staged"}
    O = load_oracle()
    mods = R.load()
    vocab = mods[1].WordVocab(0, ["key", "tensile", "density", "polyphony", "occupation"])
    batches = [O.synth_batch(B, S, T, seed=1234 + i) for i in range(2)]
    ntok = [int((~b[3]).sum() + (~b[4]).sum()) for b in batches]
    out = {"workload": f"B{B} x S{S} (+T{T}) train step, unmodified reference model.py/transformer.py + 12 CE criteria + "
                       f"torch.optim.Adam, dropout 0.1 on, stock PyTorch {torch.__version__}"}
    for mode in ("fp32", "bf16_autocast"):
        try:
            m = R.build_model(mods[0], cfg, 0.1, ctx.dev).train()
            optim = torch.optim.Adam(m.parameters(), lr=1e-4)
            crit, ce_all = R.criteria(vocab, 0.8, ctx.dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps, warm, tk = 3, 2, 0
            for it in range(warm + steps):
                if it == warm:
                    torch.cuda.synchronize()
                    e0.record()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
                    R.train_step(m, optim, crit, ce_all, mods[2].gen_nopeek_mask, batches[it % 2], ctx.dev)
                if it >= warm:
                    tk += ntok[it % 2]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"ms_per_step": ms, "tokens_per_s": tk / steps / (ms * 1e-3),
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            del m, optim, crit
        except Exception as e:
            out[mode] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()
    return out


def c1_record(ctx, cfg):
    """configs[0] through the drop-in boundary: the reference's generation_all, UNCHANGED (generation.py:468-696),
    driving this repo's ScoreTransformer on the GPU (fp32 path; the KV cache lives inside forward)."""
    R = load_reference_arm()
    if R.reference_dir() is None:
        return {"unavailable": "baseline/_ref is not staged"}
    ids, tracks, bars = c1_piece()
    m = build_model(cfg, "fp32", ctx.dev, dropout=0.1, seed=3).eval()
    with torch.no_grad():
        R.run_generation_all(m, ids, tracks, bars[:1], str(ctx.dev))      # warm-up: lazy module loads, allocator
        m._decode_cache = None
        torch.cuda.synchronize()
        restored, calls, last, dt = R.run_generation_all(m, ids, tracks, bars, str(ctx.dev))
        # the same run with the reference's own model on this GPU (stock PyTorch, no cache)
        rm = R.build_model(R.load()[0], cfg, 0.1, ctx.dev, seed=3).eval()
        rm.load_state_dict(m.state_dict())
        _, calls_ref, last_ref, dt_ref = R.run_generation_all(rm, ids, tracks, bars, str(ctx.dev))
    same = last == last_ref
    del m, rm
    torch.cuda.empty_cache()
    return {"workload": f"configs[0]: generation.generation_all unchanged, one {len(ids)}-token piece, 4 bars x 3 tracks = 52 spans, greedy, batch 1",
            "value": calls / dt, "unit": "tokens/s", "tokens": calls, "seconds": dt, "dtype": "f32",
            "reference_model_same_gpu": {"value": calls_ref / dt_ref, "unit": "tokens/s", "seconds": dt_ref, "tokens": calls_ref},
            "greedy_stream_identical_to_reference_model": bool(same)}


def c5_attention_record(ctx):
    """configs[4]: d768 / 12 heads of 64, S = T = 4096 attention forward + backward of the three variants (encoder
    full, decoder causal, cross), B/GPU in {1,2,4,8}, every rank an independent replica (max over ranks)."""
    from smer_music_generation_b200 import ops
    dev = ctx.dev
    H, dh, L = 12, 64, 4096
    d = H * dh
    pk = peaks()
    rows = []
    g = torch.Generator().manual_seed(3)
    for B in (1, 2, 4, 8):
        M = B * L
        qkv = (torch.randn(M, 3 * d, generator=g) * 0.5).to(dev).bfloat16()
        do = (torch.randn(M, d, generator=g) * 0.5).to(dev).bfloat16()
        o = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        dqkv = torch.empty_like(qkv)
        lse = torch.empty(B, H, L, device=dev)
        dsum = torch.empty(B, H, L, device=dev)
        lens = torch.randint(3 * L // 4, L + 1, (B,), generator=g)
        pad = (torch.arange(L)[None] >= lens[:, None]).to(torch.uint8).to(dev)
        kvl = lens.to(torch.int32).to(dev)
        db = torch.zeros(3 * d, device=dev)
        for name, causal in (("encoder_full", False), ("decoder_causal", True), ("cross", False)):
            a = ops.attn_args(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, B, H, L, L, dh, lse=lse, causal=causal,
                              key_pad=pad, kv_len=kvl, dropout_p=0.1, seed=11, site=5, dout=do, dq=dqkv[:, :d],
                              dk=dqkv[:, d:2 * d], dv=dqkv[:, 2 * d:], dsum=dsum, dbq=db[:d], dbk=db[d:2 * d], dbv=db[2 * d:])
            ms = []
            for fn in (ops.attn_fwd, ops.attn_bwd):
                for _ in range(2):
                    fn(a)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn(a)
                e1.record()
                torch.cuda.synchronize()
                ms.append(ctx.max(e0.elapsed_time(e1) / 5))
            fl = ops._attn_flops(a)
            rows.append({"variant": name, "B_per_gpu": B, "fwd_ms": round(ms[0], 4), "bwd_ms": round(ms[1], 4),
                         "fwd_tflops": round(fl / ms[0] / 1e9, 1), "bwd_tflops": round(2 * fl / ms[1] / 1e9, 1),
                         "fwd_bwd_frac_of_bf16_sustained": round(3 * fl / (ms[0] + ms[1]) / 1e9 / pk["tf_sust"], 4)})
        del qkv, do, o, dqkv
    torch.cuda.empty_cache()
    return {"workload": "configs[4]: H12 x dh64 (d768), S = T = 4096, bf16, dropout 0.1, suffix padding U[0.75L, L]; "
                        "nominal FLOPs 4*Lq*Lk*dh per (b,h), causal half, backward = 2x forward",
            "n_gpus": ctx.world, "points": rows}


def run_train(args, ctx):
    from smer_music_generation_b200 import ops
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    B, S, T = args.batch, args.seq, args.tgt
    cfg = dict(CFG)
    r = timed_train(ctx, cfg, args.dtype, B, S, T, args.steps, args.warmup, ctx.pg, use_graph=not args.no_graph, e2e=True,
                    profile=True, clocks=True, packed=not args.padded)
    prof = r["profile"]
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
    default_shape = (cfg["d"], cfg["le"], B, S, T) == (512, 4, 32, 1024, 1024)
    roof.update({"kernel": dom, "traffic": ncu_traffic(dom) if default_shape else None, "peak_source": pk["src"] + " (sustained)",
                 "share_of_step": d["ms"] / tot_ms, "launches_per_step": d["launch_groups"],
                 "avg_launch_ms": d["ms"] / d["launch_groups"]})
    if dom.startswith("attn"):
        roof["achieved_kv_len"] = ach * r["exec_frac"]
        roof["note"] = ("achieved counts the nominal 4*Lq*Lk*dh FLOPs; achieved_kv_len counts only keys below kv_len "
                        "(the kernels stop there; batches are U[0.75L, L])")
    step_flops = 3.0 * train_flops(B, S, T, cfg['d'], cfg['ff'], cfg['le'], cfg['ld'])
    fam = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4)}
        if k in tensor_fams:
            e["tflops"] = round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1)
            if k.startswith("attn"):
                e["tflops_kv_len"] = round(e["tflops"] * r["exec_frac"], 1)
        elif v["work"]:
            e["gbs"] = round(v["work"] / (v["ms"] * 1e-3) / 1e9, 1)
        fam[k] = e
    ms_step = r["ms_per_step"]
    line = {"metric": METRIC_TRAIN, "value": r["value"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{'configs[1]: default' if default_shape else 'variant of configs[1]:'} SMER transformer (d{cfg['d']} h{cfg['nhead']} {cfg['le']}+{cfg['ld']} layers ff{cfg['ff']} V309), teacher-forced "
                                   f"train step fwd+loss+bwd+Adam, dropout 0.1, B{B}/GPU x S{S} (+T{T}), suffix padding "
                                   f"U[0.75L,L], tokens counted = non-pad src+tgt",
                       "l2": "working set per step (~3.5 GB activations) >> 126 MB L2; 4 rotating input batches",
                       "layout": ("padding-free: packed rows + cu_seqlens, the collate (pad removal) runs on the GPU inside the step"
                                  if r["packed"] else "padded (B, L) batches"),
                       "parallelism": f"dp{world}", "global_batch": B * world,
                       "launch": "whole step captured in one CUDA graph, replayed per step" if r["use_graph"] else "eager launches"},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e_value"], "unit": "tokens/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 128,
                    "ms_per_step": r["e2e_ms_per_step"], "api": "TrainEngine.step_graph/step over ScoreTransformer (pinned host batch in, loss sums out, async D2H every step)"},
            "gpu_launches": r["launches"],
            "roofline": roof,
            "model_flops_per_step": step_flops,
            "model_tflops": step_flops * world / (ms_step * 1e-3) / 1e12,
            "model_frac_of_bf16_sustained": step_flops / (ms_step * 1e-3) / 1e12 / pk["tf_sust"],
            "attention_flops_executed_fraction": r["exec_frac"],
            "kernel_families": fam, "loss": r["loss"]}

    def sub(name, fn, cond=True):
        if not cond or name in args.skip:
            return
        try:
            t0 = time.perf_counter()
            rec = fn()
            if isinstance(rec, dict):
                rec["bench_seconds"] = round(time.perf_counter() - t0, 1)
            line[name] = rec
        except Exception as e:
            import traceback
            line[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            if rank == 0:
                traceback.print_exc(file=sys.stderr)
        torch.cuda.empty_cache()

    sub("e2e_module_api", lambda: module_api_record(ctx, cfg, args.dtype, r["host"], r["ntok"]), world == 1 and not args.no_module_api)

    def padded():
        q = timed_train(ctx, cfg, args.dtype, B, S, T, 5, 3, ctx.pg, use_graph=not args.no_graph, packed=False)
        return {"workload": "the same batches in the padded (B, L) layout (12.5 % of the rows are padding)", "value": q["value"],
                "unit": "tokens/s", "ms_per_step": q["ms_per_step"], "loss": q["loss"]}

    sub("padded_layout", padded, r["packed"])

    def c3():
        # configs[2]'s per-GPU shape, with the gradient all-reduce (N > 1) and without it (the same rank alone)
        c3cfg = dict(cfg, max_len=max(cfg["max_len"], 2048))
        dp = timed_train(ctx, c3cfg, args.dtype, 64, 2048, 2048, 5, 3, ctx.pg, use_graph=not args.no_graph)
        rec = {"workload": "configs[2]: B64/GPU x S2048 (+T2048) data-parallel train step, bf16, dropout 0.1", "n_gpus": world,
               "value": dp["value"], "unit": "tokens/s", "ms_per_step": dp["ms_per_step"], "steps": 5, "warmup": 3, "loss": dp["loss"],
               "model_frac_of_bf16_sustained": 3.0 * train_flops(64, 2048, 2048) / (dp["ms_per_step"] * 1e-3) / 1e12 / pk["tf_sust"]}
        if world > 1:
            solo = timed_train(ctx, c3cfg, args.dtype, 64, 2048, 2048, 5, 3, None, use_graph=not args.no_graph)
            rec.update(ms_per_step_without_allreduce=solo["ms_per_step"], allreduce_exposed_ms=dp["ms_per_step"] - solo["ms_per_step"],
                       efficiency_vs_same_rank_alone=solo["ms_per_step"] / dp["ms_per_step"])
        return rec

    sub("c3", c3, default_shape)
    if world > 1 and default_shape and "dp_overhead" not in args.skip:
        try:
            solo = timed_train(ctx, cfg, args.dtype, B, S, T, args.steps, args.warmup, None, use_graph=not args.no_graph)
            line["dp_overhead"] = {"ms_per_step_without_allreduce": solo["ms_per_step"],
                                   "allreduce_exposed_ms": ms_step - solo["ms_per_step"],
                                   "efficiency_vs_same_rank_alone": solo["ms_per_step"] / ms_step}
        except Exception as e:
            line["dp_overhead"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    sub("decode", lambda: decode_record(args, ctx, cpu_baseline=False), default_shape)
    sub("c5_attention", lambda: c5_attention_record(ctx), default_shape and args.dtype == "bf16")
    sub("c1", lambda: c1_record(ctx, cfg), world == 1 and default_shape and rank == 0)
    sub("torch_gpu_baseline", lambda: torch_gpu_record(ctx, cfg, 8, S, T), world == 1 and default_shape and rank == 0)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu = cpu_train_rate(cfg, 2, S, T, 1, 1, threads)
        cpu.pop("s_per_step", None)
        if isinstance(line.get("decode"), dict) and "error" not in line["decode"]:
            try:
                dcpu = cpu_decode_rate(cfg, threads)
                line["decode"]["cpu_baseline"] = {k: dcpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                line["decode"]["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    line["cpu_baseline"] = cpu
    if rank == 0:
        print(json.dumps(line), flush=True)


def decode_record(args, ctx, cpu_baseline=True):
    from smer_music_generation_b200 import InfillDecoder
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    O = load_oracle()
    cfg = dict(CFG)
    model = build_model(cfg, args.dtype, dev, seed=1234).eval()
    n_total = args.pieces
    per = n_total // world
    pieces, targets = [], []
    for i in range(per):
        ids = O.synth_piece(seed=rank * per + i, n_bars=16, n_tracks=3, events_per_track_bar=6)
        pieces.append(O.mask_bar_and_track_ids(ids, [0, 1, 2], [4, 5, 6, 7], 3))
        targets.append(O.mask_targets(4, [0, 1, 2], 3))
    dec = InfillDecoder(model, mode="top_p", top_p=0.9, seed=7, max_len=args.decode_len, splits=args.splits)
    dec.trace_intervals = True                           # device time of every graph launch between two done-checks (events only)
    res = None
    times, dev_times, gens = [], [], []
    clocks = None
    n_timed, n_warm = max(1, min(args.steps, 3)), 1
    for it in range(n_warm + n_timed):
        if it == n_warm and rank == 0:
            clocks = ClockSampler(ctx.local)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = dec.generate(pieces, targets, seq_base=rank * per)
        e1.record()
        torch.cuda.synchronize()
        if it >= n_warm:
            times.append(e0.elapsed_time(e1))           # whole call: host packing, H2D, encoder, decode, D2H, unpacking
            dev_times.append(res["device_ms"])          # encoder + cross K/V + decode loop
            gens.append(sum(res["generated"]))
    clk = clocks.stop() if clocks else None
    ms, ms_dev, toks = ctx.max(sum(times)), ctx.max(sum(dev_times)), ctx.sum(float(sum(gens)))
    kl = dec.kernel_launches
    pk = peaks()
    iv = list(getattr(dec, "interval_ms", []))
    live = list(getattr(dec, "interval_live", []))
    ce = dec.check_every_used
    # whole-step HBM fraction: algorithmic bytes of every decode step of the last timed call / its decode-loop device time
    d, esz, nl = cfg["d"], (2 if args.dtype == "bf16" else 4), cfg["ld"]
    wbytes = (nl * (6 * d * d + 2 * d * cfg["ff"]) + d * cfg["vocab"]) * esz
    mean_src = float(dec.src_len.float().mean().item())
    alg = 0.0
    for k, nlive in enumerate(live):
        pos = (k + 0.5) * ce                              # cached positions of a live piece in this interval (upper bound: catch-up steps)
        alg += ce * (nlive * nl * 2 * d * esz * (mean_src + min(pos, args.decode_len)) + wbytes)
    loop_ms = sum(iv)
    whole_frac = alg / (loop_ms * 1e-3) / 1e9 / pk["hbm"] if loop_ms > 0 else None
    # roofline of the dominant kernel: one extra eager step with events around the attention launches
    dec.use_graph = False
    dec.generate(pieces, targets, seq_base=rank * per, max_steps=64, check_every=64)
    pr = dec.profile_step()
    dom = max(("cross", "self"), key=lambda k_: pr[k_]["ms"])
    ach = pr[dom]["bytes"] / (pr[dom]["ms"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": f"decode attention over the {dom}-attention K/V", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
            "frac": ach / pk["hbm"], "traffic": ncu_traffic("decode_attn (cross-attention K/V)") if (n_total == 1024 and world == 1 and dom == "cross") else None,
            "peak_source": pk["src"], "avg_launch_ms": pr[dom]["ms"] / max(1, pr[dom]["launches"]),
            "share_of_step": pr[dom]["ms"] / pr["step_ms"], "eager_step_ms": pr["step_ms"],
            "kinds": {k_: {"gbs": pr[k_]["bytes"] / max(pr[k_]["ms"], 1e-9) / 1e6, "ms": pr[k_]["ms"]} for k_ in ("cross", "self")}}
    S = dec.S
    rec = {"metric": METRIC_DECODE, "value": toks / (ms_dev * 1e-3), "unit": "tokens/s", "n_gpus": world,
           "steps": n_timed, "warmup": n_warm, "ms_per_step": ms_dev / n_timed, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
           "config": {"workload": f"configs[3]: {n_total} independent 16-bar pieces (S<={S}) sharded {per}/GPU, 4 bars x 3 tracks masked "
                                  f"(52 spans), KV cache, grammar-masked top-p 0.9 sampling, stream cap {args.decode_len}",
                      "timed": "value: pieces resident on the device -> encoder, cross-KV, decode loop (CUDA events; the decode-step CUDA graph is "
                               "captured in the warm-up call and replayed by the timed calls: same shapes, same weights); "
                               "e2e: whole InfillDecoder.generate() incl. host packing, H2D, D2H of the token streams"},
           "e2e": {"value": toks / (ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": dec.h2d_bytes,
                   "d2h_bytes_per_step": dec.d2h_bytes, "ms_per_step": ms / n_timed},
           "clocks": clk, "gpu_launches": kl, "launches_per_decode_step": dec.launches_per_step, "decode_steps": res["steps"],
           "roofline": roof,
           "whole_step_hbm": {"frac": whole_frac, "algorithmic_bytes": alg, "decode_loop_ms": loop_ms, "peak": pk["hbm"],
                              "formula": "sum over check intervals of steps * (live pieces * Ld * 2*d*esz * (mean src len + cached positions) + weight bytes)"},
           "step_ms_graph": loop_ms / max(1, res["steps"]),
           "step_ms_by_interval": [round(x / ce, 3) for x in iv][:48], "live_by_interval": live[:48]}
    if cpu_baseline and world == 1 and rank == 0 and not args.no_cpu_baseline:
        dcpu = cpu_decode_rate(cfg, os.cpu_count() or 1)
        rec["cpu_baseline"] = {k: dcpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    del dec, model
    torch.cuda.empty_cache()
    return rec


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
    ap.add_argument("--padded", action="store_true", help="headline in the padded (B, L) layout instead of packed rows")
    ap.add_argument("--skip", default="", help="comma list of sub-records to skip: decode,c3,c5_attention,c1,torch_gpu_baseline,dp_overhead")
    ap.add_argument("--d-model", type=int, default=512)
    ap.add_argument("--nhead", type=int, default=8)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--ff", type=int, default=2048)
    args = ap.parse_args()
    args.skip = set(x for x in args.skip.replace("+", ",").split(",") if x)
    CFG.update(d=args.d_model, nhead=args.nhead, le=args.layers, ld=args.layers, ff=args.ff,
               max_len=max(2400, args.seq, args.tgt))
    if args.warmup < 3 and args.impl == "ours" and args.workload == "train":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    ctx = Ctx()
    try:
        if args.workload == "decode":
            rec = decode_record(args, ctx)
            if ctx.rank == 0:
                print(json.dumps(rec), flush=True)
        else:
            run_train(args, ctx)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
