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
g_dp, p_dp = tr.G.g.clone(), tr.G.p.clone()                       # g holds the all-reduced SUM

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
if rank == 0:
    xch = "peer copies (copy engines, symmetric memory)" if tr.G.peer is not None else "NCCL all-reduce"
    print(f"dp_parity world={world} [{xch}, {tr.xG.bytes_reduced / 1e6:.1f} MB reduced]: ||g_dp - sum_shards g|| / ||.|| = {rel_g:.2e}, fraction of weights off Adam(mean grad) by > 2e-6 = {rel_p:.2e}, "
          f"replicas identical after the step: {same} -> {'OK' if ok and same else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if ok and same else 1)
