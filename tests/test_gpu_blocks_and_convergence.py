"""(1) The stand-alone forwards of the reference's building blocks (PointNetModule, DilTempConv1d, PointNetBlock,
TemporalConvolutionBlock: models.py:6-160) composed the way the reference's ORCEDEncoder composes them (models.py:446-495),
against the oracle in train and eval mode, outputs and gradients.  (2) A 200-iteration training run of the fused bf16 path
against the fp32 oracle from the same initial weights on the same batches and RNG draws: the loss curves and the final
validation predictions must track (B200 only).
"""
import numpy as np
import pytest
import torch

from oracle import pcaa_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(LR=1e-4, B1=0.9, B2=0.99, GP_WEIGHT=15, ADV_WEIGHT=1)


def relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def relnorm(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("training", [True, False])
def test_orced_style_trunk_from_standalone_blocks(training):
    """x -> PointNetBlock -> AvgPool2d((1, N)) -> TemporalConvolutionBlock -> AvgPool1d(T): ORCEDEncoder.forward's trunk
    (models.py:488-492) built from this repository's block classes with torch's own pooling modules in between."""
    from opensetgaitrecognition_pcaa_b200 import models
    B, N, C = 3, 50, 4
    p = O.det_params(C, N, 9)
    pc, tc = models.PointNetBlock(), models.TemporalConvolutionBlock()
    pc.load_state_dict({k[len("E.pc_block."):]: v.clone() for k, v in p.items() if k.startswith("E.pc_block.")})
    tc.load_state_dict({k[len("E.tc_block."):]: v.clone() for k, v in p.items() if k.startswith("E.tc_block.")})
    pc, tc = pc.cuda().float().train(training), tc.cuda().float().train(training)
    x, _ = O.synth_batch(B, N, C, seed=3)
    pool1, pool2 = torch.nn.AvgPool2d(kernel_size=(1, N)), torch.nn.AvgPool1d(kernel_size=30)
    x1 = pc(x.cuda())
    assert tuple(x1.shape) == (B, 1024, 30, N)
    x3 = tc(torch.squeeze(pool1(x1), dim=-1))
    assert tuple(x3.shape) == (B, 512, 30)
    out = torch.squeeze(pool2(x3), dim=-1)
    # oracle
    q = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not O.is_buffer(k)) for k, v in p.items() if k.startswith("E.")}
    upd = {}
    a4 = O.pointnet_block(q, x, training, upd)
    h = O.tcn_block(q, a4.reshape(B, 30, N, -1).mean(2), training, upd)
    ref = h.mean(1)
    e = relmax(out, ref)
    print(f"[blocks training={training}] trunk output relmax {e:.4f}")
    assert e < 3e-2
    w = torch.randn(B, 512, generator=torch.Generator().manual_seed(0))
    (out * w.cuda()).sum().backward()
    (ref * w).sum().backward()
    worst = 0.0
    for mod, pre in ((pc, "E.pc_block."), (tc, "E.tc_block.")):
        for k, v in mod.named_parameters():
            g_ref = q[pre + k].grad
            if training and (k.endswith("module.0.bias") or k.endswith("conv1d.bias")):
                continue                                             # cancelled by the train-mode BatchNorm: fp noise
            assert v.grad is not None and v.grad.shape == v.shape, k
            err = relnorm(v.grad, g_ref)
            worst = max(worst, err)
            assert err < 8e-2, (k, err)
    print(f"[blocks training={training}] worst parameter-gradient relnorm {worst:.4f}")
    if training:
        sd = pc.state_dict()
        assert int(sd["pointnet1.module.1.num_batches_tracked"]) == 1
        assert relmax(sd["pointnet2.module.1.running_mean"], upd["E.pc_block.pointnet2.module.1.running_mean"]) < 2e-2


def test_single_layer_modules_match_torch_layers():
    """PointNetModule / DilTempConv1d called directly on NCHW / NCT tensors == the torch.nn layers they contain."""
    from opensetgaitrecognition_pcaa_b200 import models
    torch.manual_seed(0)
    m = models.PointNetModule(64, 128).cuda().float().train()
    x = torch.randn(2, 64, 30, 20, device="cuda", requires_grad=True)
    bn_state = {k: v.clone() for k, v in m.module[1].state_dict().items()}
    y = m(x)
    m.module[1].load_state_dict(bn_state)                           # the torch pass below updates the running stats again
    xr = x.detach().clone().requires_grad_(True)
    yr = m.module(xr)                                               # the nn.Sequential itself: Conv2d, BatchNorm2d, ELU (fp32)
    assert relmax(y, yr) < 2e-2
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    assert relnorm(x.grad, xr.grad) < 3e-2
    d = models.DilTempConv1d(32, 48, dilation=2).cuda().float().train()
    h = torch.randn(3, 32, 30, device="cuda")
    out = d(h)
    want = d.activation(d.batch_norm(d.conv1d(h)[:, :, :-d.padding]))                  # models.py:73-79
    assert tuple(out.shape) == (3, 48, 30) and relmax(out, want) < 2e-2


def test_200_iterations_track_the_fp32_oracle():
    """Same initial weights, same 4 batches of 4 crops cycled, same z0 / alphas draws, 200 iterations: fused bf16 path vs fp32
    oracle.  Individual weights drift apart by Adam's sign steps on near-zero gradients, the optimisation must not: the loss
    curves (mean over windows of 20 iterations) stay within 5 % (reconstruction), 0.05 absolute (cross-entropy), 10 % of
    max(1, |.|) (critic); both learn (final CE < initial CE); eval-mode predictions on held-out crops agree wherever the oracle's
    top-2 logit gap exceeds 5 % of max |logit|."""
    from test_gpu_baseline_sizes import build_trainer
    B, nmax, C, steps = 4, 50, 2, 200
    p = O.det_params(C, nmax, 4)
    po = {k: v.clone() for k, v in p.items()}
    tr, means = build_trainer(p, C, nmax)
    data = []
    for i in range(4):
        xs, ys = [], []
        for c in range(C):
            x, _ = O.synth_batch(B // C, nmax, C, seed=300 + 10 * i + c, sigma_scale=O.subject_sigma_scale(3 * c))
            xs.append(x), ys.append(torch.full((B // C,), c, dtype=torch.int64))
        data.append((torch.cat(xs), torch.cat(ys)))
    rng = np.random.default_rng(77)
    curves = {"gpu": [], "ref": []}
    ost = {}
    ocfg = dict(CFG, NMAX=nmax)
    for s in range(steps):
        pcs, gt = data[s % 4]
        z0 = torch.from_numpy(rng.normal(0, 1, (B, 32))).float()
        al = torch.from_numpy(rng.uniform(0, 1, (B, 1)).astype(np.float32))
        ref = O.train_step_variant4(po, ost, pcs, gt, z0, al, means, ocfg)
        out = tr.step_graphed(pcs.cuda(), gt.cuda(), z0.cuda(), al.cuda())
        curves["ref"].append([float(ref[k]) for k in ("rec_loss", "sup_loss", "d_loss")])
        curves["gpu"].append([float(out[k]) for k in ("rec_loss", "sup_loss", "d_loss")])
    cg, cr = np.array(curves["gpu"]).reshape(10, 20, 3).mean(1), np.array(curves["ref"]).reshape(10, 20, 3).mean(1)
    for w in range(10):
        print(f"[200 it] window {w}: rec {cg[w,0]:.3f}/{cr[w,0]:.3f}  ce {cg[w,1]:.4f}/{cr[w,1]:.4f}  d {cg[w,2]:.3f}/{cr[w,2]:.3f}  (fused / oracle)")
    assert np.all(np.abs(cg[:, 0] - cr[:, 0]) <= 5e-2 * np.abs(cr[:, 0]))
    assert np.all(np.abs(cg[:, 1] - cr[:, 1]) <= 5e-2)
    assert np.all(np.abs(cg[:, 2] - cr[:, 2]) <= 1e-1 * np.maximum(1.0, np.abs(cr[:, 2])))
    assert cr[-1, 1] < cr[0, 1] and cg[-1, 1] < cg[0, 1] and cr[-1, 0] < cr[0, 0] and cg[-1, 0] < cg[0, 0]
    vx, vy = [], []
    for c in range(C):
        x, _ = O.synth_batch(8, nmax, C, seed=900 + c, sigma_scale=O.subject_sigma_scale(3 * c))
        vx.append(x), vy.append(torch.full((8,), c, dtype=torch.int64))
    vx, vy = torch.cat(vx), torch.cat(vy)
    _, _, pred = tr.evaluate(vx.cuda(), vy.cuda())
    with torch.no_grad():
        lg, _ = O.encoder_forward(po, vx, False, True)
    top2 = lg.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 5e-2 * float(lg.abs().max())
    agree = pred.cpu().long() == lg.argmax(1)
    print(f"[200 it] held-out predictions: {int(agree.sum())}/{len(agree)} agree, {int(decided.sum())} decided; "
          f"oracle accuracy {float((lg.argmax(1) == vy).float().mean()):.2f}, fused {float((pred.cpu().long() == vy).float().mean()):.2f}")
    assert bool(agree[decided].all())
