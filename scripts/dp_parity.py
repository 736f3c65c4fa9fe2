"""Data-parallel parity on real GPUs: run under torchrun with N ranks.

Each rank runs ONE variant-4 step of the fused trainer on its shard of a global batch (local BatchNorm statistics, sum-all-reduce of the flat gradient -- copy engines over NVLink peer memory by default, NCCL with PCAA_DP_EXCHANGE=nccl --, 1/N folded into Adam).  Reference for every rank: a single-process trainer (world 1)
stepped on each shard separately from the same initial weights; the DP gradient must equal the MEAN of those per-shard
gradients and the DP weights the Adam update of that mean.  Prints max relative deviations; exits non-zero on failure.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_parity.py
"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensetgaitrecognition_pcaa_b200 import dp, synth
from opensetgaitrecognition_pcaa_b200.train import build_variant4

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
C, nmax, Bg = 4, 50, 8 * world


def syncbn_check():
    """SyncBN (PCAATrainer(sync_bn=True)): the N-rank iteration must equal ONE process stepping the concatenated global batch
    -- losses, the exchanged gradient / N vs the single-process gradient, BatchNorm running statistics, weights."""
    from opensetgaitrecognition_pcaa_b200 import engine
    # SyncBN runs the TCN layers as separate statistics / coefficient kernels (the all-reduce sits between them); the
    # single-process reference is switched to the same kernels, so that the comparison isolates the statistics exchange
    # (the fused TCN kernels take their column sums from the GEMM epilogue: another fp32 summation order, enough to flip
    # bf16 roundings downstream)
    engine.TCN_FUSED = False
    pcs, gt = synth.synth_batch(Bg, nmax, C, seed=11)
    np.random.seed(3); torch.manual_seed(3)
    z0_g, al_g = dp.global_draws(Bg, 32, 0, 1)
    s, e = dp.shard_range(Bg, rank, world)
    tr = build_variant4(C, nmax, seed=0, device=dev, sync_bn=True)
    assert tr.bn_sync is not None and tr.world == world
    dist.broadcast(tr.G.p, 0); dist.broadcast(tr.D.p, 0)
    for b in tr.enc.buffers():
        dist.broadcast(b, 0)
    tr.G.make_shadow(); tr._refresh_views()
    snap = tr.snapshot()
    out = tr.step(pcs[s:e].to(dev), gt[s:e].to(dev), z0_g[s:e].to(dev), al_g[s:e].to(dev))
    torch.cuda.synchronize()
    ref = build_variant4(C, nmax, seed=0, device=dev, process_group=dp.SINGLE)
    ref.restore(snap)
    oref = ref.step(pcs.to(dev), gt.to(dev), z0_g.to(dev), al_g.to(dev))
    torch.cuda.synchronize()
    g_red = tr.reduced_gradient()                 # (assembled from the owners' chunks under the sharded decoder update)
    rel_g = float((g_red / world - ref.G.g).norm() / ref.G.g.norm())
    rel_d = float((tr.D.g / world - ref.D.g).norm() / ref.D.g.norm())
    worst_t, worst_name = dp.per_tensor_relnorm(tr.G, g_red / world, ref.G.g)
    # noise floor: the single-process iteration repeated from the same state (atomics / bf16 rounding flips, see dp.graphed_step_parity)
    g_first = ref.G.g.clone()
    ref.restore(snap)
    ref.step(pcs.to(dev), gt.to(dev), z0_g.to(dev), al_g.to(dev))
    torch.cuda.synchronize()
    noise_g = float((ref.G.g - g_first).norm() / g_first.norm())
    noise_t, _ = dp.per_tensor_relnorm(ref.G, ref.G.g, g_first)
    frac_p = float(((tr.G.p - ref.G.p).abs() > 2e-6).float().mean())
    bn_dev = max(float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30)) for a, b in zip(tr.enc.buffers(), ref.enc.buffers()))
    losses = torch.stack([out[k] for k in ("rec_loss", "d_loss", "sup_loss", "loss_g")])
    dist.all_reduce(losses)
    losses /= world
    lref = torch.stack([oref[k] for k in ("rec_loss", "d_loss", "sup_loss", "loss_g")])
    rel_l = float(((losses - lref).abs() / lref.abs().clamp_min(1.0)).max())
    # (the repeat restored the snapshot before stepping again: ref.G.p is still one Adam step of the single-process gradient)
    # decisive check of the SyncBN machinery with IDENTICAL tiling on both sides: every rank steps the SAME shard, so the
    # all-reduced statistics equal the local ones (two identical halves) and each rank's gradient must equal the gradient of
    # a plain local-BatchNorm step on that shard -- any wrong factor in the reduced sums or row counts shows here at O(1)
    sh = (pcs[:Bg // world].to(dev), gt[:Bg // world].to(dev), z0_g[:Bg // world].to(dev), al_g[:Bg // world].to(dev))
    tr.restore(snap)
    tr.step(*sh)
    torch.cuda.synchronize()
    g_sync = tr.reduced_gradient() / world
    ref.restore(snap)
    ref.step(*sh)
    torch.cuda.synchronize()
    same_g = float((g_sync - ref.G.g).norm() / ref.G.g.norm())
    same_t, same_name = dp.per_tensor_relnorm(ref.G, g_sync, ref.G.g)
    ok = (same_g <= 1e-3 and same_t <= 2e-2 and rel_g <= 2e-2 and worst_t <= 8e-2 and rel_d < 1e-3 and bn_dev < 1e-3
          and rel_l < 1e-3)
    res = torch.tensor([rel_g, rel_d, frac_p, bn_dev, rel_l, worst_t, noise_g, noise_t, 0.0 if ok else 1.0, same_g, same_t], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dp_parity syncbn identical shards on all ranks vs local BatchNorm on that shard: ||g_sync - g_local|| / ||.|| = {float(res[9]):.2e}, "
              f"worst tensor {float(res[10]):.2e} ({same_name})")
        print(f"dp_parity syncbn world={world} [{tr.bn_sync.calls} statistics all-reduces]: ||g_dp/N - g_single|| / ||.|| = {float(res[0]):.2e} "
              f"(worst tensor {float(res[5]):.2e} at {worst_name}; run-to-run noise of the single process {float(res[6]):.2e} / {float(res[7]):.2e}; "
              f"critic {float(res[1]):.2e}), weights off Adam(single) by > 2e-6: {float(res[2]):.2e}, BatchNorm running statistics {float(res[3]):.2e}, "
              f"mean losses {float(res[4]):.2e} -> {'OK' if float(res[8]) == 0 else 'FAIL'}")
    return float(res[8]) == 0


if "--sync-bn" in sys.argv:
    good = syncbn_check()
    dist.destroy_process_group()
    sys.exit(0 if good else 1)

pcs, gt = synth.synth_batch(Bg, nmax, C, seed=11)
np.random.seed(3); torch.manual_seed(3)
z0_l, al_l = dp.global_draws(Bg, 32, rank, world)                 # global draws, sliced
np.random.seed(3); torch.manual_seed(3)
z0_g, al_g = dp.global_draws(Bg, 32, 0, 1)
s, e = dp.shard_range(Bg, rank, world)
assert torch.equal(z0_g[s:e], z0_l) and torch.equal(al_g[s:e], al_l)

tr = build_variant4(C, nmax, seed=0, device=dev, process_group=None)          # default group -> world ranks
assert tr.world == world
dist.broadcast(tr.G.p, 0); dist.broadcast(tr.D.p, 0)
for b in tr.enc.buffers():
    dist.broadcast(b, 0)
tr.G.make_shadow(); tr._refresh_views()
p0, d0 = tr.G.p.clone(), tr.D.p.clone()
bufs0 = [b.clone() for b in tr.enc.buffers()]
out = tr.step(pcs[s:e].to(dev), gt[s:e].to(dev), z0_l.to(dev), al_l.to(dev))
torch.cuda.synchronize()
g_dp, p_dp = tr.reduced_gradient(), tr.G.p.clone()                 # the all-reduced SUM

# single-process reference on every shard, from the same initial state (no process group: world 1)
import opensetgaitrecognition_pcaa_b200.dp as dpm
real_world_info = dpm.world_info
dpm.world_info = lambda group=None: (0, 1)
try:
    gsum = torch.zeros_like(g_dp)
    ref = build_variant4(C, nmax, seed=0, device=dev)
    for r in range(world):
        ref.G.p.copy_(p0); ref.D.p.copy_(d0)
        for b, b0 in zip(ref.enc.buffers(), bufs0):
            b.copy_(b0)
        ref.G.m.zero_(); ref.G.v.zero_(); ref.D.m.zero_(); ref.D.v.zero_(); ref.G.step = 0; ref.D.step = 0
        ref.G.make_shadow(); ref._refresh_views()
        rs, re = dp.shard_range(Bg, r, world)
        ref.step(pcs[rs:re].to(dev), gt[rs:re].to(dev), z0_g[rs:re].to(dev), al_g[rs:re].to(dev))
        gsum += ref.G.g
finally:
    dpm.world_info = real_world_info
torch.cuda.synchronize()
rel_g = float((g_dp - gsum).norm() / gsum.norm())
# Adam of the mean gradient (first step: p - lr * g/(|g| + eps) up to bias correction)
gm = gsum / world
upd = 1e-4 * gm / (gm.abs() + 1e-8)
dev_p = ((p_dp - p0) + upd).abs()
rel_p = float((dev_p > 2e-6).float().mean())     # entries whose gradient is rounding noise may take either Adam sign step
ok = rel_g < 2e-3 and rel_p < 1e-2
allp = [torch.empty_like(p_dp) for _ in range(world)]
dist.all_gather(allp, p_dp)
same = all(torch.equal(allp[0], t) for t in allp)
# the same check THROUGH step_graphed (what bench.py --gpus N times): captured graphs + the selected exchange, iteration 3+
inp = (pcs[s:e].to(dev), gt[s:e].to(dev), z0_l.to(dev), al_l.to(dev))
for _ in range(3):
    tr.step_graphed(*inp)
gp = dp.graphed_step_parity(tr, inp, lambda: build_variant4(C, nmax, seed=0, device=dev, process_group=dp.SINGLE))
ok = ok and gp["ok"]
if rank == 0:
    print("dp_parity graphed: " + ", ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in gp.items()))
if rank == 0:
    xch = "peer copies (copy engines, symmetric memory)" if tr.G.peer is not None else "NCCL all-reduce"
    if tr.shard_adam:
        xch += ", sharded decoder update"
    print(f"dp_parity world={world} [{xch}, {tr.xG.bytes_reduced / 1e6:.1f} MB reduced]: ||g_dp - sum_shards g|| / ||.|| = {rel_g:.2e}, fraction of weights off Adam(mean grad) by > 2e-6 = {rel_p:.2e}, "
          f"replicas identical after the step: {same} -> {'OK' if ok and same else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if ok and same else 1)
