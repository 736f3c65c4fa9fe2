"""CPU experiment (no GPU needed): which bf16 rounding of the PointNet path costs how much gradient accuracy?

Runs the encoder + CE + a random embedding gradient of the oracle in fp32, then again with the roundings the sm_100a path
performs switched on one at a time (and all together):
  W    layer 2-4 weights rounded to bf16 (tensor-core operands)
  y    pre-BatchNorm GEMM outputs y_l stored as bf16 (statistics are taken from the fp32 accumulators)
  a    activations a_l = ELU(BN(y_l)) stored as bf16, l = 1..3
  dz   dz_l (gradient at the BatchNorm output) stored as bf16, l = 1..3
  dy   dy_l (gradient at the GEMM output) stored as bf16, l = 2..4
  tcn  TCN operands (im2col activations, weights, dy) rounded to bf16
and prints ||g - g_fp32|| / ||g_fp32|| per encoder tensor group.  Usage: python scripts/sim_bf16_rounding.py [B] [N]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcaa_oracle as O  # noqa: E402


def r(x):
    return x.bfloat16().float()


class RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return r(g)


class RoundSTE(torch.autograd.Function):
    """bf16 rounding in the forward, identity in the backward."""

    @staticmethod
    def forward(ctx, x):
        return r(x)

    @staticmethod
    def backward(ctx, g):
        return g


def bn_train(y_stat, y_used, gamma, beta):
    """BatchNorm with statistics from y_stat (the fp32 accumulators) applied to y_used (what was stored)."""
    mean = y_stat.mean(0)
    var = ((y_stat - mean) ** 2).mean(0)
    return (y_used - mean) / torch.sqrt(var + O.BN_EPS) * gamma + beta


def encoder(p, x, sw):
    B, C, T, N = x.shape
    a = x.permute(0, 2, 3, 1).reshape(B * T * N, C)
    for l in range(1, 5):
        k = f"E.pc_block.pointnet{l}.module."
        w = p[k + "0.weight"].reshape(p[k + "0.weight"].shape[0], -1)
        if "W" in sw and l > 1:
            w = RoundSTE.apply(w)
        y = a @ w.t() + p[k + "0.bias"]
        yu = RoundSTE.apply(y) if "y" in sw else y
        if "dy" in sw and l > 1:
            yu = RoundGrad.apply(yu)
        z = bn_train(y if "y" in sw else yu, yu, p[k + "1.weight"], p[k + "1.bias"])
        if "dz" in sw and l < 4:
            z = RoundGrad.apply(z)
        a = O.elu(z)
        if "a" in sw and l < 4:
            a = RoundSTE.apply(a)
    pooled = a.reshape(B, T, N, -1).mean(2)
    h = pooled
    for l in range(1, 7):
        k = f"E.tc_block.dtc{l}."
        w, hh = p[k + "conv1d.weight"], h
        if "tcn" in sw:
            w, hh = RoundSTE.apply(w), RoundSTE.apply(h)
        y = O.causal_dilated_conv(hh, w, p[k + "conv1d.bias"], O.DTC_DILATIONS[l - 1])
        if "tcn" in sw:
            y = RoundGrad.apply(y)
        Cc = y.shape[-1]
        y2 = y.reshape(B * T, Cc)
        z = bn_train(y2, y2, p[k + "batch_norm.weight"], p[k + "batch_norm.bias"])
        h = O.elu(z).reshape(B, T, Cc)
    g = h.mean(1)
    fv = O.elu(g @ p["E.MLP_sup1.0.weight"].t() + p["E.MLP_sup1.0.bias"])
    hh = O.elu(fv @ p["E.MLP_head.0.weight"].t() + p["E.MLP_head.0.bias"])
    logits = O.elu(hh @ p["E.MLP_sup2.0.weight"].t() + p["E.MLP_sup2.0.bias"])
    return logits, fv


def grads(p, x, gt, dfv, sw):
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items() if k.startswith("E.") and not O.is_buffer(k)}
    q = dict(p)
    q.update(leaves)
    logits, fv = encoder(q, x, sw)
    if dfv is None:
        # the generator loss of the real iteration (PCAA_ablation.py:985-1013): Chamfer(decoder(head(fv)), x) - mean D(fv) + CE
        C = logits.shape[1]
        onehot = torch.nn.functional.one_hot(gt, C).float()
        rec = O.decoder_forward(q, O.proj_head_forward(q, fv), x.shape[-1])
        loss = O.chamfer(rec, x)[0] - O.disc_forward(q, fv, onehot).mean() + O.cross_entropy(logits, gt)
    else:
        loss = O.cross_entropy(logits, gt) + (fv * dfv).sum()
    names = list(leaves)
    g = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    return {n: gg for n, gg in zip(names, g) if gg is not None}, fv.detach()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    C = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 4
    torch.manual_seed(0)
    p = O.det_params(C, N, 0)
    x, gt = O.synth_batch(B, N, C, seed=4321)
    dfv = None if "--full-loss" in sys.argv else torch.randn(B, 32) * 0.05
    ref, fv0 = grads(p, x, gt, dfv, set())
    groups = {"pointnet1-2 W": lambda n: "pointnet1.module.0.weight" in n or "pointnet2.module.0.weight" in n,
              "pointnet3-4 W": lambda n: "pointnet3.module.0.weight" in n or "pointnet4.module.0.weight" in n,
              "pointnet BN": lambda n: "pc_block" in n and "module.1." in n,
              "tcn W": lambda n: "conv1d.weight" in n, "tcn BN": lambda n: "batch_norm" in n, "heads": lambda n: "MLP" in n}
    print(f"B={B} N={N}: relative gradient error vs fp32 by rounding switch")
    print(f"{'switch':16s}" + "".join(f"{g:>15s}" for g in groups) + f"{'fv relmax':>12s}")
    for sw in (["W"], ["y"], ["a"], ["dz"], ["dy"], ["tcn"], ["W", "y", "a"], ["dz", "dy"], ["W", "y", "a", "dz", "dy", "tcn"]):
        g, fv = grads(p, x, gt, dfv, set(sw))
        row = []
        for gname, sel in groups.items():
            num = sum(float((g[n] - ref[n]).double().norm() ** 2) for n in ref if sel(n) and not n.endswith("0.bias"))
            den = sum(float(ref[n].double().norm() ** 2) for n in ref if sel(n) and not n.endswith("0.bias"))
            row.append((num / max(den, 1e-300)) ** 0.5)
        print(f"{'+'.join(sw):16s}" + "".join(f"{v:15.4f}" for v in row) + f"{float((fv - fv0).abs().max() / fv0.abs().max()):12.4f}")


if __name__ == "__main__":
    main()
