"""Data-parallel host logic on CPU with gloo, world_size 2:
 * the bucketed all-reduce over the flat gradient arena covers every parameter exactly once and
   fires in backward-completion order;
 * all-reducing the loss kernel's (numerator, normaliser) sums BEFORE the loss backward makes the
   per-rank gradients add up to the single-process gradient on the concatenated batch
   (train.py:736 normalises by the batch-global sum)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smer_music_generation_b200 import ScoreTransformer
        from smer_music_generation_b200.model import GradArena
        from smer_music_generation_b200.trainer import GradBuckets
        import importlib.util
        spec = importlib.util.spec_from_file_location("smer_oracle", os.path.join(ROOT, "oracle", "smer_oracle.py"))
        O = importlib.util.module_from_spec(spec)
        sys.modules["smer_oracle"] = O
        spec.loader.exec_module(O)
        torch.manual_seed(0)
        m = ScoreTransformer(309, 32, 2, 2, 2, 64, 64, 0.0, 0.0)
        arena = GradArena(m)
        # default layout: decoder stack | encoder layers but the first | first encoder layer + embedding (the exposed tail)
        gd = GradBuckets(m, arena, 0, dist.group.WORLD, comm_stream=None)
        assert [c for c, _, _ in gd.buckets] == [
            ["fc.", "transformer.decoder.norm.", "transformer.decoder.layers.1.", "transformer.decoder.layers.0."],
            ["transformer.encoder.norm.", "transformer.encoder.layers.1."],
            ["transformer.encoder.layers.0.", "embedding."]]
        # per-layer layout: the embedding table alone in the last bucket
        gl = GradBuckets(m, arena, -1, dist.group.WORLD, comm_stream=None)
        assert gl.buckets[-1][0] == ["embedding."] and len(gl.buckets) == 2 + 2 + 1
        assert gl.buckets[0][0] == ["fc.", "transformer.decoder.norm.", "transformer.decoder.layers.1."]
        b_emb, e_emb = gl.buckets[-1][1:]
        assert e_emb - b_emb == (309 * 32 + 63) // 64 * 64
        for gb in (gd, gl, GradBuckets(m, arena, 4, dist.group.WORLD, comm_stream=None)):
            # every arena element belongs to exactly one bucket
            cover = torch.zeros(arena.total, dtype=torch.int32)
            for _, b, e in gb.buckets:
                cover[b:e] += 1
            assert int(cover.min()) == 1 and int(cover.max()) == 1
        gb = gd
        # rank-dependent gradients, signalled in the order _Run.backward signals them
        for n, v in arena.views.items():
            v.fill_(float(rank + 1))
        order = ["fc.", "transformer.decoder.norm."] + [f"transformer.decoder.layers.{i}." for i in (1, 0)] + \
                ["transformer.encoder.norm."] + [f"transformer.encoder.layers.{i}." for i in (1, 0)] + ["embedding."]
        for g in order:
            gb.ready(g)
        gb.finish()
        assert gb.launch_order == sorted(gb.launch_order)
        for n, v in arena.grads.items():
            assert torch.all(v == 3.0), n                       # 1 + 2

        # global-normaliser semantics with the oracle's loss on a split batch
        sd = O.random_state_dict(32, 2, 1, 1, 64, 64, seed=1)
        src, tin, tout, sp, tp = O.synth_batch(4, 24, 16, seed=2)
        W, C = O.loss_weights(0.8)
        full_loss, full_grads, _, _ = O.train_step_grads(sd, src, tin, tout, sp, tp, 2, W, C)
        sl = slice(rank * 2, rank * 2 + 2)
        leaf = {k: v.detach().clone().requires_grad_(k != "pos_enc.pe") for k, v in sd.items()}
        logits, _ = O.score_transformer_forward(leaf, src[sl], tin[sl], 2, sp[sl], tp[sl], sp[sl], O.nopeek_mask(16)[None])
        lg = logits.reshape(-1, 309)
        y = tout[sl].reshape(-1)
        lse = torch.logsumexp(lg, -1)
        nll = torch.where(y == 0, torch.zeros_like(lse), lse - lg.gather(1, y[:, None])[:, 0])
        sums = torch.stack([(nll * W[y]).sum().detach().double(), C[y].sum().double()])
        dist.all_reduce(sums)                                    # what TrainEngine does with the kernel's sums
        ((nll * W[y]).sum() / sums[1].float()).backward()         # local numerator / GLOBAL normaliser
        g = leaf["fc.weight"].grad.clone()
        dist.all_reduce(g)
        torch.testing.assert_close(g, full_grads["fc.weight"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close((sums[0] / sums[1]).float(), full_loss, rtol=1e-5, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:                                       # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_dp_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
