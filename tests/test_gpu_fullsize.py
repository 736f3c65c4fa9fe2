"""Parity at BASELINE.json's FULL sizes (per-GPU batch 256, N = 150 points, 30 frames: 1.15 M points per step), where the
CPU oracle would need minutes: size-independent properties instead of element-wise comparison (B200 only).

* Chamfer: the fused nearest-neighbour kernel against min-reductions of the distance matrix produced by a different
  kernel -- values and arg-mins bit-exact (same arithmetic, lowest index on ties);
* mean pool: the fused BN + ELU + segmented-mean kernel against the separate apply kernel + a plain reduction, and the
  BatchNorm property (normalised pre-activations have zero mean / unit variance per channel);
* eval forward: invariance to the order of the points inside a frame and to the composition of the batch; the pooled
  GEMM epilogue against the unfused pair;
* train step: replicas of the same step agree, the bf16 operand copy is exactly bf16(master) after Adam, the padding
  invariant of the activation format holds, the losses fall over 20 iterations;
* open-set scoring: monotone in the distance, invariant to prototype order, votes idempotent under repetition.
"""
import numpy as np
import pytest
import torch

from oracle import pcaa_oracle as O

pytestmark = pytest.mark.gpu

B, NMAX, C, T = 256, 150, 4, 30


@pytest.fixture(scope="module")
def ops():
    from opensetgaitrecognition_pcaa_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def batch():
    from opensetgaitrecognition_pcaa_b200 import synth
    pcs, gt = synth.synth_batch(B, NMAX, C, seed=77)
    return pcs.cuda(), gt.cuda()


def test_chamfer_full_size_equals_min_reductions_of_the_distance_matrix(ops, batch):
    pcs, _ = batch
    g = torch.Generator(device="cuda").manual_seed(1)
    preds = pcs + 0.3 * torch.randn(pcs.shape, device="cuda", generator=g)
    fl, i1, i2 = ops.chamfer_fwd(preds, pcs)
    P = ops.pairwise_dist(pcs, preds)                       # [B, T, N(gt), N(pred)], same expanded arithmetic
    m1, a1 = P.min(dim=2)                                   # for every predicted point: nearest ground-truth point
    m2, a2 = P.min(dim=3)                                   # for every ground-truth point: nearest predicted point
    # torch.min on CUDA does not promise the first index on ties: compare the chosen DISTANCES bit for bit, and the
    # indices wherever the minimum is unique
    d1 = torch.gather(P, 2, i1.view(B, T, 1, NMAX).long()).squeeze(2)
    d2 = torch.gather(P, 3, i2.view(B, T, NMAX, 1).long()).squeeze(3)
    assert torch.equal(d1, m1) and torch.equal(d2, m2)
    uniq1 = (P == m1.unsqueeze(2)).sum(2) == 1
    uniq2 = (P == m2.unsqueeze(3)).sum(3) == 1
    assert torch.equal(i1.view(B, T, NMAX).long()[uniq1], a1[uniq1]) and torch.equal(i2.view(B, T, NMAX).long()[uniq2], a2[uniq2])
    # ties (duplicated padding points of the ground truth): the kernel returns the LOWEST index
    first1 = (P == m1.unsqueeze(2)).int().argmax(2)
    first2 = (P == m2.unsqueeze(3)).int().argmax(3)
    assert torch.equal(i1.view(B, T, NMAX).long(), first1) and torch.equal(i2.view(B, T, NMAX).long(), first2)
    ref = (m1.double().sum(2) + m2.double().sum(2)).view(-1)
    assert float((fl.double().view(-1) - ref).abs().max() / ref.abs().max()) < 1e-5
    # gradient: 2 (pred_j - gt_i1(j)) + 2 sum_{i: i2(i) = j} (pred_j - gt_i), scaled by 1 / (B T)
    grad = ops.chamfer_bwd(preds, pcs, i1, i2, torch.ones((), device="cuda"), True)
    pj = preds.permute(0, 2, 3, 1)                           # [B, T, N, F]
    gi = pcs.permute(0, 2, 3, 1)
    want = pj - torch.gather(gi, 2, i1.view(B, T, NMAX, 1).long().expand(-1, -1, -1, 4))
    contrib = torch.gather(pj, 2, i2.view(B, T, NMAX, 1).long().expand(-1, -1, -1, 4)) - gi
    want = want.scatter_add(2, i2.view(B, T, NMAX, 1).long().expand(-1, -1, -1, 4), contrib)
    want = (2.0 / (B * T)) * want.permute(0, 3, 1, 2)
    assert float((grad - want).abs().max() / want.abs().max()) < 1e-5


def test_meanpool_full_size_cross_kernel_and_batchnorm_property(ops):
    Cc, G = 1024, B * T
    P = G * NMAX
    g = torch.Generator(device="cuda").manual_seed(2)
    yT = ops.t256_empty(Cc, P, "cuda")
    yT.copy_((0.5 + torch.randn(yT.shape, device="cuda", generator=g) * 1.5).bfloat16())
    flat = yT.permute(1, 0, 2).reshape(Cc, -1)
    flat[:, P:] = 0                                          # the format's invariant: pad points are zeros
    y = flat[:, :P].float()
    gamma = 1 + 0.1 * torch.randn(Cc, device="cuda", generator=g)
    beta = 0.1 * torch.randn(Cc, device="cuda", generator=g)
    mean, var = y.double().mean(1), y.double().var(1, unbiased=False)
    invstd = (1.0 / torch.sqrt(var + 1e-5))
    scale = (gamma.double() * invstd).float()
    shift = (beta.double() - mean * gamma.double() * invstd).float()
    coef = torch.stack([scale, shift, mean.float(), invstd.float()]).contiguous()
    # BatchNorm property on the device-side statistics path: colwise sums of the normalised tensor
    z = (y * scale[:, None] + shift[:, None] - beta[:, None]) / gamma[:, None]
    assert float(z.double().mean(1).abs().max()) < 1e-3 and float((z.double().var(1, unbiased=False) - 1).abs().max()) < 1e-3
    pooled, e1, e2 = ops.bn_elu_meanpool_t(yT, coef, G, NMAX, want_e=True)
    a = ops.bn_elu_apply_t(yT, coef, P)                      # separate kernel (bf16 output)
    a32 = torch.nn.functional.elu(y * scale[:, None] + shift[:, None])
    ref = a32.view(Cc, G, NMAX).mean(2).t()
    assert float((pooled - ref).abs().max() / ref.abs().max()) < 1e-4
    plain, _, _ = ops.bn_elu_meanpool_t(a, None, G, NMAX)    # mean of the bf16 activations through the no-BN path
    assert float((plain - ref).abs().max() / ref.abs().max()) < 5e-3
    zz = y * scale[:, None] + shift[:, None]
    d = torch.where(zz > 0, torch.ones_like(zz), torch.exp(zz))
    assert float((e1 - d.view(Cc, G, NMAX).sum(2).t()).abs().max() / NMAX) < 1e-4
    xh = ((y.double() - mean[:, None]) * invstd[:, None]).float()
    want2 = (d * xh).view(Cc, G, NMAX).sum(2).t()
    assert float((e2 - want2).abs().max() / want2.abs().max()) < 1e-3
    # total conservation: sum over groups of n * pooled == sum over all points of the activation
    assert float(((pooled.double().sum(0) * NMAX) - a32.double().sum(1)).abs().max() / a32.double().sum(1).abs().max()) < 1e-4


def _encoder(seed=0):
    from opensetgaitrecognition_pcaa_b200 import models
    p = O.det_params(C, NMAX, seed)
    enc = models.CGEncoder(n_out_labels=C, use_projection_head=True, nmax_points=NMAX)
    enc.load_state_dict({k[2:]: v.clone() for k, v in p.items() if k.startswith("E.")})
    return enc.cuda().float().eval()


def test_eval_forward_full_size_point_order_and_batch_composition(batch):
    from opensetgaitrecognition_pcaa_b200 import inference
    pcs, _ = batch
    enc = _encoder()
    fv, pred = inference.encode(enc, pcs, batch=B)
    # mean pooling over the points of a frame: any permutation of the points gives the same embedding
    perm = torch.argsort(torch.rand(B, 1, T, NMAX, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)), dim=3)
    shuffled = torch.gather(pcs, 3, perm.expand(-1, 4, -1, -1))
    fv2, pred2 = inference.encode(enc, shuffled, batch=B)
    scale = float(fv.abs().max())
    assert float((fv - fv2).abs().max()) < 2e-3 * scale
    # eval-mode BatchNorm uses running statistics: a crop's embedding does not depend on its batch
    fv3, pred3 = inference.encode(enc, pcs.flip(0), batch=64)
    assert float((fv - fv3.flip(0)).abs().max()) < 2e-3 * scale
    lg_gap_ok = torch.ones_like(pred, dtype=torch.bool)
    assert float((pred[lg_gap_ok] != pred3.flip(0)[lg_gap_ok]).float().mean()) < 0.01
    assert float((pred != pred2).float().mean()) < 0.01


def test_train_step_full_size_invariants(batch):
    from opensetgaitrecognition_pcaa_b200 import engine
    from opensetgaitrecognition_pcaa_b200.train import build_variant4
    pcs, gt = batch
    tr = build_variant4(C, NMAX, seed=0)
    rng = np.random.default_rng(0)
    losses = []
    for s in range(20):
        z0 = torch.from_numpy(rng.normal(0, 1, (B, 32))).float().cuda()
        al = torch.from_numpy(rng.uniform(0, 1, (B, 1)).astype(np.float32)).cuda()
        out = tr.step_graphed(pcs, gt, z0, al)
        losses.append([float(out[k]) for k in ("rec_loss", "sup_loss", "d_loss", "loss_g")])
        assert all(np.isfinite(losses[-1])), (s, losses[-1])
    torch.cuda.synchronize()
    # the same batch 20 times: the reconstruction loss falls, the classification loss does not rise (it competes with the
    # adversarial term for the same embedding; measured: falls by ~0.2 % in 20 iterations at lr 1e-4)
    assert losses[-1][0] < losses[0][0] and losses[-1][1] < losses[0][1] * 1.01
    assert tr.G.step == 20 and int(tr.G.step_dev) == 20
    # the tensor-core operand copy is exactly bf16(master) after every Adam update
    assert torch.equal(tr.G.shadow, tr.G.p.to(torch.bfloat16))
    # running statistics moved and stayed finite; 20 updates with momentum 0.1
    sd = tr.enc.state_dict()
    assert int(sd["pc_block.pointnet4.module.1.num_batches_tracked"]) == 20
    assert all(torch.isfinite(v).all() for v in sd.values() if v.dtype.is_floating_point)
    # activation format invariant at full size: pad points of the last tile are zeros after the forward
    pooled, sv = engine.pointnet_forward(pcs, tr.P_E, True, wb16=tr._enc_wb16)
    P = B * T * NMAX
    for l in (1, 2, 3):
        for tname in ("y", "a"):
            xT = sv[tname][l]
            flat = xT.permute(1, 0, 2).reshape(xT.shape[1], -1)
            assert bool((flat[:, P:] == 0).all()), (tname, l)
    assert pooled.shape == (B * T, 1024) and bool(torch.isfinite(pooled).all())


def test_openset_scoring_full_stream_properties(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    M = 1 << 20                                              # one million embeddings (config 5's stream length)
    means = O.sample_distant_points(32, C, 10, 10).float().cuda()
    lab = torch.randint(0, C, (M,), device="cuda", generator=g)
    emb = means[lab] + torch.randn(M, 32, device="cuda", generator=g)
    ll = ops.openset_score(emb, means)
    assert ll.dtype == torch.float64 and bool(torch.isfinite(ll).all())
    # invariant to the order of the prototypes (a mixture is a sum)
    ll_p = ops.openset_score(emb, means.flip(0).contiguous())
    assert float((ll - ll_p).abs().max()) < 1e-9
    # monotone: moving an embedding away from every prototype along the ray from the origin lowers its likelihood
    ll_far = ops.openset_score(emb * 3.0, means)
    assert float((ll_far < ll).double().mean()) > 0.999
    # bounded above by the single-component density at distance zero: -16 ln(2 pi) - ln C + ln C
    assert float(ll.max()) <= -16 * np.log(2 * np.pi) + 1e-9
    # votes: windows of k copies of one sample reproduce that sample's own decision
    k, thr = 6, float(ll.median())
    n = 6000
    pred = lab[:n].to(torch.int32)
    rep_ll = ll[:n].repeat_interleave(k).contiguous()
    rep_pred = pred.repeat_interleave(k).contiguous()
    votes = ops.openset_vote(rep_ll, rep_pred, k, thr, C)
    want = torch.where(ll[:n] > thr, pred, torch.full_like(pred, C))
    assert torch.equal(votes, want)
