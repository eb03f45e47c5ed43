"""TrainEngine (fused step: forward + loss + backward + Adam over flat arenas) against the oracle's
restatement of train.py:722-786, and its CUDA-graph replay mode.  GPU only."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(O, mode, dropout=0.0, seed=5):
    from smer_music_generation_b200 import ScoreTransformer
    cfg = dict(d=64, h=4, le=2, ld=2, ff=128, maxlen=64)
    sd = O.random_state_dict(cfg["d"], cfg["h"], cfg["le"], cfg["ld"], cfg["ff"], cfg["maxlen"], seed=seed)
    m = ScoreTransformer(309, cfg["d"], cfg["h"], cfg["le"], cfg["ld"], cfg["ff"], cfg["maxlen"], dropout, dropout,
                         compute_dtype=mode).to(DEV)
    m.load_state_dict(sd)
    return m.train(), sd, cfg


def test_engine_step_matches_oracle_fp32(oracle):
    from smer_music_generation_b200.trainer import TrainEngine
    O = oracle
    m, sd, cfg = _model(O, "fp32")
    src, tin, tout, sp, tp = O.synth_batch(3, 32, 24, seed=9)
    W, C = O.loss_weights(0.8)
    ref_loss, ref_grads, _, ref_parts = O.train_step_grads(sd, src, tin, tout, sp, tp, cfg["h"], W, C)
    eng = TrainEngine(m, lr=1e-4, eos_weight=0.8)
    args = [t.to(DEV) for t in (src, tin, tout, sp, tp)]
    eng.step(*args, update=False)
    assert abs(eng.loss_value() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    s = eng.sums.cpu()
    torch.testing.assert_close((s[2:14] / s[1]).float(), ref_parts, rtol=1e-4, atol=1e-6)
    for n, g in ref_grads.items():
        mine = eng.grads.grads[n].cpu()
        err = (mine - g).abs().max().item() / max(g.abs().max().item(), 1e-12)
        assert err < 1e-3, (n, err)
    # state_dict still has the reference layout after the parameters moved into the arena
    assert list(m.state_dict().keys()) == list(sd.keys())
    # one full step == oracle Adam on the oracle gradients
    eng2_m, sd2, _ = _model(O, "fp32")
    eng2 = TrainEngine(eng2_m, lr=1e-3, eos_weight=0.8)
    eng2.step(*args)
    for n, g in ref_grads.items():
        p = sd2[n].clone()
        O.adam_step(p, g, torch.zeros_like(p), torch.zeros_like(p), 1, lr=1e-3)
        mine = dict(eng2_m.named_parameters())[n].detach().cpu()
        stable = g.abs() > 1e-5                 # where |g| ~ eps the Adam direction is ill-conditioned
        torch.testing.assert_close(mine[stable], p[stable], rtol=2e-3, atol=5e-5)
        assert (mine - sd2[n]).abs().max().item() <= 1.01e-3      # no entry moves by more than lr


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_engine_graph_replay_trains(oracle, mode):
    """Graph-captured step: loss on a fixed batch goes down, stays finite, dropout masks and the
    Adam step number advance across replays (device counter)."""
    from smer_music_generation_b200.trainer import TrainEngine
    O = oracle
    m, sd, cfg = _model(O, mode, dropout=0.1)
    src, tin, tout, sp, tp = O.synth_batch(4, 32, 24, seed=3)
    eng = TrainEngine(m, lr=2e-3, eos_weight=0.8)
    eng.capture(4, 32, 24)
    batch = [t.pin_memory() for t in (src, tin, tout, sp, tp)]
    losses = []
    for _ in range(30):
        eng.step_graph(*batch)
        losses.append(eng.loss_value())
    assert all(l == l and l < 1e4 for l in losses)
    assert sum(losses[-5:]) / 5 < 0.8 * sum(losses[:5]) / 5
    assert int(eng._ctr.item()) == eng.step_count
    # with lr = 0 two replays differ only through the dropout seed
    eng.release_graph()
    eng.lr = 0.0
    eng.capture(4, 32, 24)
    eng.step_graph(*batch)
    a = eng.loss_value()
    eng.step_graph(*batch)
    b = eng.loss_value()
    assert a != b and abs(a - b) < 0.2 * abs(a)
    eng.release_graph()
    # bf16 shadows follow the fp32 masters after Adam
    if mode == "bf16":
        sh = eng.arena.shadow.float()
        assert (sh - eng.arena.flat).abs().max().item() <= 1e-2 * eng.arena.flat.abs().max().item()


def test_fused_adam_bf16_weights_follow_the_update(oracle):
    """Module-API training in bf16: FusedAdam updates through raw pointers, so it must bump the parameter versions --
    otherwise the bf16 weight shadows (and the decode cache) keep serving the initial weights."""
    from smer_music_generation_b200 import SmerLoss
    from smer_music_generation_b200.trainer import FusedAdam
    O = oracle
    m, sd, cfg = _model(O, "bf16")
    src, tin, tout, sp, tp = [t.to(DEV) for t in O.synth_batch(3, 32, 24, seed=9)]
    crit = SmerLoss(309, 0.8).to(DEV)
    opt = FusedAdam(m.parameters(), lr=5e-2)
    logits0, _ = m(src, tin, sp, tp, sp, "causal")
    v0 = m.transformer.encoder.layers[0].linear1.weight._version
    loss, _, _ = crit(logits0, tout)
    loss.backward()
    opt.step()
    w = m.transformer.encoder.layers[0].linear1.weight
    assert w._version > v0
    with torch.no_grad():
        logits1, _ = m(src, tin, sp, tp, sp, "causal")
    assert (logits1 - logits0.detach()).abs().max().item() > 1e-2          # the forward reads the updated weights
    sh = m._w._shadow["transformer.encoder.layers.0.linear1.weight"]
    assert torch.equal(sh, w.detach().bfloat16())
    # same step with the weights in fp32 mode gives (nearly) the same logits: the shadows really are the new weights
    m32, _, _ = _model(O, "fp32")
    m32.load_state_dict(m.state_dict())
    with torch.no_grad():
        ref1, _ = m32(src, tin, sp, tp, sp, "causal")
    assert (logits1 - ref1).abs().max().item() < 5e-2 * ref1.abs().max().item()
    # eval-mode decode cache is invalidated by the step as well
    m.eval()
    with torch.no_grad():
        a, _ = m(src[:1], tin[:1, :5], None, None, None, "causal")
    m.train()
    logits2, _ = m(src, tin, sp, tp, sp, "causal")
    crit(logits2, tout)[0].backward()
    opt.step()
    m.eval()
    with torch.no_grad():
        b, _ = m(src[:1], tin[:1, :5], None, None, None, "causal")
    assert not torch.equal(a, b)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_capture_leaves_parameters_and_optimizer_state_untouched(oracle, mode):
    from smer_music_generation_b200.trainer import TrainEngine
    O = oracle
    m, sd, cfg = _model(O, mode, dropout=0.1)
    eng = TrainEngine(m, lr=1e-2, eos_weight=0.8)
    batch = [t.to(DEV) for t in O.synth_batch(4, 32, 24, seed=3)]
    eng.step(*batch)                                                  # some real state first
    snap = [t.clone() for t in (eng.arena.flat, eng.arena.m, eng.arena.v)]
    shadow = eng.arena.shadow.clone() if eng.arena.shadow is not None else None
    count = eng.step_count
    eng.capture(4, 32, 24)
    assert eng.step_count == count and int(eng._ctr.item()) == count
    for a, b in zip(snap, (eng.arena.flat, eng.arena.m, eng.arena.v)):
        assert torch.equal(a, b)
    if shadow is not None:
        assert torch.equal(shadow, eng.arena.shadow)
    eng.step_graph(*batch)
    assert eng.step_count == count + 1 and int(eng._ctr.item()) == count + 1
    assert not torch.equal(snap[0], eng.arena.flat)
    eng.release_graph()


def test_engine_optimizer_state_dict_roundtrip(oracle):
    """TrainEngine's flat Adam state <-> torch.optim.Adam's state_dict layout (train.py:967-973), and shadows that
    follow model.load_state_dict()."""
    from smer_music_generation_b200.trainer import TrainEngine
    O = oracle
    batch = [t.to(DEV) for t in O.synth_batch(3, 32, 24, seed=9)]
    m, sd, cfg = _model(O, "bf16")
    eng = TrainEngine(m, lr=1e-3, eos_weight=0.8)
    eng.step(*batch)
    eng.step(*batch)
    osd = eng.optimizer_state_dict()
    msd = {k: v.clone() for k, v in m.state_dict().items()}
    # the dict loads into torch.optim.Adam built over an ordinary copy of the parameters
    plain = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters()]
    topt = torch.optim.Adam(plain, lr=1e-3)
    topt.load_state_dict(osd)
    assert all(float(topt.state[p]["step"]) == 2.0 for p in plain)
    # resume in a fresh engine: third step equals the third step of the original
    m2, _, _ = _model(O, "bf16", seed=11)
    eng2 = TrainEngine(m2, lr=1e-3, eos_weight=0.8)
    m2.load_state_dict(msd)
    assert torch.equal(eng2.arena.shadow, eng.arena.shadow)           # post-hook refreshed the bf16 shadows
    eng2.load_optimizer_state_dict(osd)
    assert eng2.step_count == 2
    eng.step(*batch)
    eng2.step(*batch)
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        torch.testing.assert_close(a, b, rtol=0, atol=1e-6, msg=n)


def _dp_worker(rank, world, port, out_q):
    import os
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from conftest import load_oracle
    from smer_music_generation_b200 import ScoreTransformer
    from smer_music_generation_b200.trainer import TrainEngine
    O = load_oracle()
    cfg = dict(d=64, h=4, le=2, ld=2, ff=128, maxlen=64)
    sd = O.random_state_dict(cfg["d"], cfg["h"], cfg["le"], cfg["ld"], cfg["ff"], cfg["maxlen"], seed=5)
    m = ScoreTransformer(309, cfg["d"], cfg["h"], cfg["le"], cfg["ld"], cfg["ff"], cfg["maxlen"], 0.0, 0.0,
                         compute_dtype="fp32").to(dev)
    m.load_state_dict(sd)
    m.train()
    eng = TrainEngine(m, lr=1e-3, eos_weight=0.8, process_group=dist.group.WORLD)
    full = O.synth_batch(4 * world, 32, 24, seed=21)
    mine = [t[rank * 4:(rank + 1) * 4].to(dev) for t in full]
    eng.step(*mine, update=False)
    torch.cuda.synchronize()
    if rank == 0:
        out_q.put({n: g.cpu() for n, g in eng.grads.grads.items()})
        out_q.put(eng.sums.cpu())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_engine_grads_equal_single_gpu_on_concatenated_batch(oracle):
    """DP over real GPUs (NCCL): arena gradients after the bucketed all-reduce == the single-GPU gradients on the
    concatenated batch, loss normalised by the batch-global sum C[y] (train.py:736)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from smer_music_generation_b200.trainer import TrainEngine
    O = oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, 29517, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads2 = q.get(timeout=300)
    sums2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    m, sd, cfg = _model(O, "fp32")
    eng = TrainEngine(m, lr=1e-3, eos_weight=0.8)
    full = [t.to(DEV) for t in O.synth_batch(8, 32, 24, seed=21)]
    eng.step(*full, update=False)
    torch.testing.assert_close(eng.sums.cpu(), sums2, rtol=1e-9, atol=1e-9)
    for n, g in eng.grads.grads.items():
        ref = g.cpu()
        err = (grads2[n] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        assert err < 1e-4, (n, err)


def test_packed_rows_match_padded_batch_and_oracle(oracle):
    """Padding-free layout: a ragged batch run as packed rows + cu_seqlens gives the loss sums and
    gradients of the padded run of the same batch (same kernels, other tile boundaries: bf16 tolerance) and of the
    oracle; tokens are the same non-pad tokens.  Also through the captured step with the on-GPU collate."""
    from smer_music_generation_b200 import ScoreTransformer
    from smer_music_generation_b200.model import PackedBatch
    from smer_music_generation_b200.trainer import TrainEngine
    O = oracle
    cfg = dict(d=128, h=2, le=2, ld=2, ff=256, maxlen=512)
    sd = O.random_state_dict(cfg["d"], cfg["h"], cfg["le"], cfg["ld"], cfg["ff"], cfg["maxlen"], seed=5)
    src, tin, tout, sp, tp = O.synth_batch(5, 300, 280, seed=12, min_frac=0.3)
    ls, lt = (~sp).sum(1).tolist(), (~tp).sum(1).tolist()
    assert min(ls) < 200 < max(ls)
    W, C = O.loss_weights(0.8)
    ref_loss, ref_grads, _, _ = O.train_step_grads(sd, src, tin, tout, sp, tp, cfg["h"], W, C)

    def engine():
        m = ScoreTransformer(309, cfg["d"], cfg["h"], cfg["le"], cfg["ld"], cfg["ff"], cfg["maxlen"], 0.0, 0.0,
                             compute_dtype="bf16").to(DEV)
        m.load_state_dict(sd)
        return TrainEngine(m.train(), lr=1e-3, eos_weight=0.8)

    dev_b = [t.to(DEV) for t in (src, tin, tout, sp, tp)]
    e_pad = engine()
    e_pad.step(*dev_b, update=False)
    s_pad = e_pad.sums.cpu()
    e_pk = engine()
    pk = PackedBatch.pack(dev_b[0], dev_b[1], dev_b[2], ls, lt)
    assert pk.rows_s % 128 == 0 and pk.n_s == sum(ls) and pk.rows_s - pk.n_s < 128
    # the collate kernel moved exactly the non-pad tokens, in order
    assert pk.src_ids[: pk.n_s].cpu().tolist() == src[~sp].tolist()
    assert pk.tgt_out[: pk.n_t].cpu().tolist() == tout[~tp].tolist()
    assert pk.pos_t[: pk.n_t].cpu().tolist() == [i for n in lt for i in range(n)]
    e_pk.step_packed(pk, update=False)
    s_pk = e_pk.sums.cpu()
    torch.testing.assert_close(s_pk[1], s_pad[1], rtol=0, atol=0)                 # the normaliser counts the same targets
    assert abs(float(s_pk[0] / s_pk[1]) - float(s_pad[0] / s_pad[1])) < 2e-3 * abs(float(s_pad[0] / s_pad[1]))
    assert abs(float(s_pk[0] / s_pk[1]) - ref_loss.item()) < 2e-2 * abs(ref_loss.item())
    for n, g in e_pad.grads.grads.items():
        a, b, r = e_pk.grads.grads[n].cpu().flatten(), g.cpu().flatten(), ref_grads[n].flatten()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        assert cos > 0.9995, (n, cos)
        assert (a - b).abs().max().item() < 3e-2 * max(b.abs().max().item(), 1e-12), n
        assert torch.nn.functional.cosine_similarity(a, r, dim=0).item() > 0.995, n
    # captured packed step with the collate inside the graph: replay on two different ragged batches
    e_g = engine()
    e_g.lr = 0.0                                    # (the learning rate is baked into the captured Adam launch)
    e_g.capture_packed(5, 5 * 300, 5 * 280, 300, 280)
    for seed in (12, 13):
        b = O.synth_batch(5, 300, 280, seed=seed, min_frac=0.3)
        bl, tl = (~b[3]).sum(1).tolist(), (~b[4]).sum(1).tolist()
        e_g.lr = 0.0
        e_g.step_graph_packed(b[0].pin_memory(), b[1].pin_memory(), b[2].pin_memory(), bl, tl)
        e_ref = engine()
        e_ref.step(*[t.to(DEV) for t in b], update=False)
        sg, sr = e_g.sums.cpu(), e_ref.sums.cpu()
        assert float(sg[1]) == float(sr[1])
        assert abs(float(sg[0] / sg[1]) - float(sr[0] / sr[1])) < 2e-3 * abs(float(sr[0] / sr[1]))
    e_g.release_graph()
