"""CPU, world_size 2 over gloo: the host-side data-parallel logic of dp.py (the N > 1 path of the train step and of
batch-sharded inference).  The same code runs over NCCL on the B200 box (tests/test_gpu_dp.py, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world, out_dir)
    finally:
        dist.destroy_process_group()


def _spawn(fn_name, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn_name, str(tmp_path)), nprocs=world, join=True)


# ------------------------------------------------------------------------------------------------ workers
def _w_grad_exchange(rank, world, out_dir):
    from opensetgaitrecognition_pcaa_b200 import dp
    n = 1000
    rng = np.random.default_rng(100 + rank)
    flat = torch.from_numpy(rng.normal(0, 1, n).astype(np.float32))
    mine = flat.clone()
    x = dp.GradExchange(flat)
    assert (x.rank, x.world) == (rank, world) and x.grad_scale == 1.0 / world
    # decoder-like span first, then the rest, in buckets
    for lo, hi in dp.split_spans(400, 1000, 256):
        x.start(lo, hi)
    x.start(0, 400)
    x.finish()
    want = sum(torch.from_numpy(np.random.default_rng(100 + r).normal(0, 1, n).astype(np.float32)) for r in range(world))
    assert torch.allclose(flat, want, atol=1e-6), float((flat - want).abs().max())
    assert x.bytes_reduced == 4 * n
    # mean gradient == gradient of the global-batch mean loss for equal shards
    assert torch.allclose(flat * x.grad_scale, want / world)
    assert not torch.equal(mine, flat)
    torch.save(flat, os.path.join(out_dir, f"flat{rank}.pt"))


def _w_draws_and_shards(rank, world, out_dir):
    from opensetgaitrecognition_pcaa_b200 import dp
    B = 10
    np.random.seed(0)
    torch.manual_seed(0)
    z0, al = dp.global_draws(B, 32, rank, world)
    s, e = dp.shard_range(B, rank, world)
    assert z0.shape == (e - s, 32) and al.shape == (e - s, 1)
    torch.save((z0, al, s, e), os.path.join(out_dir, f"draw{rank}.pt"))
    # inference: ragged shards of a stream + the single gather of per-sample scores
    n = 11
    s, e = dp.shard_range(n, rank, world)
    local = torch.arange(s, e, dtype=torch.float64) * 0.5
    counts = [dp.shard_range(n, r, world)[1] - dp.shard_range(n, r, world)[0] for r in range(world)]
    allv = dp.gather_scores(local, counts)
    assert torch.equal(allv, torch.arange(n, dtype=torch.float64) * 0.5)


def _w_sharded_mean_gradient(rank, world, out_dir):
    """Batch-mean loss: the mean over ranks of per-shard gradients equals the global-batch gradient."""
    from opensetgaitrecognition_pcaa_b200 import dp
    torch.manual_seed(1)
    W = torch.randn(6, 3, dtype=torch.float64, requires_grad=True)
    X = torch.randn(8, 6, dtype=torch.float64)
    s, e = dp.shard_range(8, rank, world)
    loss = (X[s:e] @ W).pow(2).mean()
    (g,) = torch.autograd.grad(loss, W)
    flat = g.reshape(-1).clone()
    x = dp.GradExchange(flat)
    x.start(0, flat.numel())
    x.finish()
    (gg,) = torch.autograd.grad((X @ W).pow(2).mean(), W)
    assert torch.allclose(flat * x.grad_scale, gg.reshape(-1), atol=1e-12)


def _w_syncbn_math(rank, world, out_dir):
    """The SyncBN arithmetic of engine.BnSync / engine._bn_bwd_coefs in plain torch (float64): all-reduced sums + global row
    count give the concatenated batch's statistics; backward with LOCAL d gamma / d beta (the gradient exchange adds the
    ranks') and GLOBAL dy coefficients equals autograd through BatchNorm on the concatenated batch."""
    from opensetgaitrecognition_pcaa_b200 import dp, engine
    torch.manual_seed(7)
    R, C, eps = 12, 5, 1e-5
    y_all = torch.randn(world * R, C, dtype=torch.float64) * 1.3 + 0.2
    dz_all = torch.randn(world * R, C, dtype=torch.float64)
    gamma, beta = torch.rand(C, dtype=torch.float64) + 0.5, torch.randn(C, dtype=torch.float64)
    y, dz = y_all[rank * R:(rank + 1) * R], dz_all[rank * R:(rank + 1) * R]
    bn = engine.BnSync()
    assert bn.world == world
    st = torch.cat([y.sum(0), (y * y).sum(0)])
    bn.reduce(st)
    Rg = R * bn.world
    mean = st[:C] / Rg
    var = st[C:] / Rg - mean * mean
    invstd = 1.0 / torch.sqrt(var + eps)
    xh = (y - mean) * invstd
    st2 = torch.cat([dz.sum(0), (dz * xh).sum(0)])
    dbeta_local, dgamma_local = st2[:C].clone(), st2[C:].clone()          # BEFORE the reduction
    bn.reduce(st2)
    dy = gamma * invstd * (dz - st2[:C] / Rg - xh * st2[C:] / Rg)         # dy = c1*dz + c2*y + c3 with the global sums
    # reference: autograd through BatchNorm on the concatenated batch
    ya = y_all.clone().requires_grad_(True)
    ga, ba = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out = torch.nn.functional.batch_norm(ya, None, None, ga, ba, training=True, eps=eps)
    out.backward(dz_all)
    assert torch.allclose(dy, ya.grad[rank * R:(rank + 1) * R], atol=1e-12)
    flat = torch.cat([dgamma_local, dbeta_local])
    x = dp.GradExchange(flat)
    x.start(0, flat.numel())
    x.finish()
    assert torch.allclose(flat[:C], ga.grad, atol=1e-12) and torch.allclose(flat[C:], ba.grad, atol=1e-12)
    assert bn.calls == 2


# ------------------------------------------------------------------------------------------------ tests
def test_syncbn_sums_and_backward_world2(tmp_path):
    _spawn("_w_syncbn_math", tmp_path)


def test_grad_exchange_world2(tmp_path):
    _spawn("_w_grad_exchange", tmp_path)
    a, b = torch.load(tmp_path / "flat0.pt"), torch.load(tmp_path / "flat1.pt")
    assert torch.equal(a, b)                      # every rank ends with bit-identical reduced gradients


def test_global_draws_are_the_single_process_draws(tmp_path):
    _spawn("_w_draws_and_shards", tmp_path)
    np.random.seed(0)
    torch.manual_seed(0)
    z_full = torch.from_numpy(np.random.normal(0, 1, (10, 32))).float()       # PCAA_ablation.py:915-921
    a_full = torch.rand(size=(10, 1))                                          # PCAA_ablation.py:944-948
    got_z, got_a, cover = [], [], []
    for r in range(2):
        z0, al, s, e = torch.load(tmp_path / f"draw{r}.pt")
        got_z.append(z0), got_a.append(al), cover.append((s, e))
    assert cover == [(0, 5), (5, 10)]
    assert torch.equal(torch.cat(got_z), z_full) and torch.equal(torch.cat(got_a), a_full)


def test_sharded_mean_gradient_world2(tmp_path):
    _spawn("_w_sharded_mean_gradient", tmp_path)


def test_shard_range_properties():
    from opensetgaitrecognition_pcaa_b200 import dp
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)
    assert dp.split_spans(0, 20, 8) == [(0, 8), (8, 16), (16, 20)]
    assert dp.split_spans(3, 3, 8) == []
    assert dp.world_info() == (0, 1)


def test_peer_exchange_chunks_partition_the_span():
    """The chunk of the decoder span a rank reduces -- and, with the sharded optimizer, updates and serves to its peers:
    chunks tile [lo, hi) in rank order, start on multiples of 8 elements relative to lo (16-byte aligned in the fp32
    buffers AND in the bf16 operand copy) and may be empty at the tail."""
    from opensetgaitrecognition_pcaa_b200 import dp
    for world in (2, 3, 4, 8):
        px = object.__new__(dp.PeerExchange)
        px.world = world
        for lo, hi in ((0, 8), (16, 16 + 217_767_240), (24, 24 + 4484), (8, 8 + 40), (0, 12)):
            ch = px.chunks(lo, hi)
            assert len(ch) == world and ch[0][0] == lo and ch[-1][1] == hi
            assert all(ch[i][1] == ch[i + 1][0] for i in range(world - 1))
            assert all(b <= e for b, e in ch) and all((b - lo) % 8 == 0 for b, e in ch if e > b)
            sizes = [e - b for b, e in ch]
            assert max(sizes) <= ((hi - lo + world - 1) // world + 7) // 8 * 8
