"""Parity of every C-ABI kernel against the CPU oracle / explicit fp32 torch-CPU formulas (B200 only)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import pcaa_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from opensetgaitrecognition_pcaa_b200 import ops as _ops
    return _ops


def cuda(t):
    return t.cuda().contiguous()


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ------------------------------------------------------------------------------------------------ CUDA-core GEMM
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 16), (130, 70, 33), (960, 48, 3072), (5, 18000, 77), (16, 96, 7680)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_simt(ops, M, N, K, ta, tb):
    g = torch.Generator().manual_seed(M * 131 + N * 7 + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    b = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (a.t() if ta else a) @ (b.t() if tb else b) + bias
    out = ops.gemm(cuda(a), cuda(b), trans_a=ta, trans_b=tb, bias=cuda(bias))
    assert rel_err(out, ref) < 2e-5          # fp32, different summation order only
    out2 = ops.gemm(cuda(a), cuda(b), trans_a=ta, trans_b=tb, bias=cuda(bias), act=ops.ACT_ELU, out=out.clone(),
                    accumulate=True)
    assert rel_err(out2, O.elu(ref) + ref) < 2e-5


def test_gemm_simt_bf16_inputs(ops):
    g = torch.Generator().manual_seed(3)
    a, b = bf16_round(torch.randn(100, 200, generator=g)), bf16_round(torch.randn(50, 200, generator=g))
    out = ops.gemm(cuda(a).bfloat16(), cuda(b).bfloat16(), trans_b=True)
    assert rel_err(out, a @ b.t()) < 2e-5


# ------------------------------------------------------------------------------------------------ tcgen05 GEMMs
TC_SHAPES = [(128, 256, 64), (6000, 512, 512), (4500, 1024, 512), (1000, 1024, 1024), (300, 512, 72), (129, 264, 520)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_tc_bias_stats(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = bf16_round(torch.randn(M, K, generator=g))
    w = bf16_round(torch.randn(N, K, generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    ref = a @ w.t() + bias
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    out = ops.gemm_tc_tn(cuda(a).bfloat16(), cuda(w).bfloat16(), ops._lib.TC_BIAS_STATS, bias=cuda(bias), stats=stats)
    torch.cuda.synchronize()
    # bf16 output rounding: relative 2^-8 per element; accumulation is fp32 (tolerance stated: 1e-2 of max |ref|)
    assert rel_err(out.float(), ref) < 1e-2
    s = stats.cpu()
    assert rel_err(s[:N], ref.double().sum(0)) < 1e-4
    assert rel_err(s[N:], (ref.double() ** 2).sum(0)) < 1e-4


@pytest.mark.parametrize("M,N,K", TC_SHAPES[:4])
def test_gemm_tc_bias_elu_and_plain(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K + 1)
    a = bf16_round(torch.randn(M, K, generator=g))
    w = bf16_round(torch.randn(N, K, generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    out = ops.gemm_tc_tn(cuda(a).bfloat16(), cuda(w).bfloat16(), ops._lib.TC_BIAS_ELU, bias=cuda(bias))
    assert rel_err(out.float(), O.elu(a @ w.t() + bias)) < 1e-2
    out = ops.gemm_tc_tn(cuda(a).bfloat16(), cuda(w).bfloat16(), ops._lib.TC_PLAIN)
    assert rel_err(out.float(), a @ w.t()) < 1e-2
    # the CUDA-core GEMM on the same device agrees too (full-size on-device comparator)
    cmp = ops.gemm(cuda(a).bfloat16(), cuda(w).bfloat16(), trans_b=True)
    assert rel_err(out.float(), cmp) < 1e-2


@pytest.mark.parametrize("M,N,K", TC_SHAPES[:4])
def test_gemm_tc_dgrad_elubn(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K + 2)
    dy = bf16_round(torch.randn(M, K, generator=g))
    wt = bf16_round(torch.randn(N, K, generator=g) / math.sqrt(K))
    yprev = bf16_round(torch.randn(M, N, generator=g))
    coef = torch.stack([1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g),
                        0.1 * torch.randn(N, generator=g), 1 + 0.1 * torch.rand(N, generator=g)])
    z = yprev * coef[0] + coef[1]
    dz = (dy @ wt.t()) * torch.where(z > 0, torch.ones_like(z), torch.exp(z))
    xh = (yprev - coef[2]) * coef[3]
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    out = ops.gemm_tc_tn(cuda(dy).bfloat16(), cuda(wt).bfloat16(), ops._lib.TC_DGRAD_ELUBN, stats=stats,
                         yprev=cuda(yprev).bfloat16(), coef=cuda(coef))
    assert rel_err(out.float(), dz) < 1e-2
    s = stats.cpu()
    assert rel_err(s[:N], dz.double().sum(0)) < 1e-3
    assert rel_err(s[N:], (dz.double() * xh.double()).sum(0)) < 1e-3


@pytest.mark.parametrize("K,N1,N2", [(64, 128, 256), (6000, 512, 512), (4500, 1024, 512), (20000, 1024, 1024), (333, 136, 264)])
def test_gemm_tc_wgrad(ops, K, N1, N2):
    g = torch.Generator().manual_seed(K + N1 + N2)
    a = bf16_round(torch.randn(K, N1, generator=g))
    b = bf16_round(torch.randn(K, N2, generator=g))
    dW = torch.zeros(N1, N2, device="cuda")
    ops.gemm_tc_nt_wgrad(cuda(a).bfloat16(), cuda(b).bfloat16(), dW)
    ref = a.double().t() @ b.double()
    assert rel_err(dW, ref) < 1e-4
    ops.gemm_tc_nt_wgrad(cuda(a).bfloat16(), cuda(b).bfloat16(), dW)      # accumulates
    assert rel_err(dW, 2 * ref) < 1e-4


# ------------------------------------------------------------------------------------------------ BatchNorm family
@pytest.mark.parametrize("R,C,dtype", [(960, 16, torch.float32), (6000, 512, torch.bfloat16), (1234, 1024, torch.bfloat16),
                                        (90, 64, torch.float32)])
def test_bn_forward_backward(ops, R, C, dtype):
    g = torch.Generator().manual_seed(R + C)
    y = torch.randn(R, C, generator=g) * 1.5 + 0.3
    if dtype == torch.bfloat16:
        y = bf16_round(y)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    rm, rv = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
    upd = {}
    yy = y.clone().requires_grad_(True)
    gg, bb = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = O.batchnorm_rows(yy, gg, bb, rm, rv, True, upd, "")
    a = O.elu(z)
    dout = torch.randn(R, C, generator=g)
    if dtype == torch.bfloat16:
        dout = bf16_round(dout)
    a.backward(dout)
    yd = cuda(y).to(dtype)
    stats = ops.colstats(yd)
    rmd, rvd = cuda(rm), cuda(rv)
    coef = ops.bn_finalize(stats, R, cuda(gamma), cuda(beta), rmd, rvd)
    out = ops.bn_elu_apply(yd, coef[0], coef[1], out_dtype=torch.float32)
    assert rel_err(out, a.detach()) < 2e-5
    assert rel_err(rmd, upd["running_mean"]) < 1e-5 and rel_err(rvd, upd["running_var"]) < 1e-5
    dz, st2 = ops.elu_bwd_colstats(cuda(dout).to(dtype), yd, coef, dz_dtype=dtype)
    c, dgamma, dbeta = ops.bn_bwd_finalize(st2, R, coef)
    dy = ops.bn_bwd_apply(dz, yd, c)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert rel_err(dy.float(), yy.grad) < tol
    assert rel_err(dgamma, gg.grad) < (5e-3 if dtype == torch.bfloat16 else 1e-4)
    assert rel_err(dbeta, bb.grad) < (5e-3 if dtype == torch.bfloat16 else 1e-4)
    # eval coefficients
    ce = ops.bn_eval_coeffs(cuda(gamma), cuda(beta), cuda(rm), cuda(rv))
    oe = ops.bn_elu_apply(yd, ce[0], ce[1], out_dtype=torch.float32)
    assert rel_err(oe, O.elu(O.batchnorm_rows(y, gamma, beta, rm, rv, False))) < 2e-5


@pytest.mark.parametrize("G,n,C", [(120, 50, 1024), (90, 150, 1024), (7, 70, 512)])
def test_bn_elu_meanpool_and_backward(ops, G, n, C):
    g = torch.Generator().manual_seed(G + n)
    y = bf16_round(torch.randn(G * n, C, generator=g))
    sc, sh = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ref = O.elu(y * sc + sh).reshape(G, n, C).mean(1)
    out = ops.bn_elu_meanpool(cuda(y).bfloat16(), cuda(sc), cuda(sh), n)
    assert rel_err(out, ref) < 1e-5
    gp = torch.randn(G, C, generator=g)
    coef = torch.stack([sc, sh, 0.05 * torch.randn(C, generator=g), 1 + 0.1 * torch.rand(C, generator=g)])
    dz, st2 = ops.elu_bwd_colstats(cuda(gp), cuda(y).bfloat16(), cuda(coef), pooled_n=n)
    z = y * sc + sh
    dz_ref = (gp / n).repeat_interleave(n, dim=0) * torch.where(z > 0, torch.ones_like(z), torch.exp(z))
    assert rel_err(dz.float(), dz_ref) < 1e-2
    assert rel_err(st2[:C], dz_ref.double().sum(0)) < 1e-3


# ------------------------------------------------------------------------------------------------ PointNet layer 1
@pytest.mark.parametrize("B,N", [(2, 50), (3, 150)])
def test_pointnet_l1(ops, B, N):
    x, _ = O.synth_batch(B, N, 2, seed=5)
    g = torch.Generator().manual_seed(1)
    w, b = torch.randn(512, 4, generator=g) * 0.5, torch.randn(512, generator=g)
    x2 = x.permute(0, 2, 3, 1).reshape(-1, 4)
    ref = x2 @ w.t() + b
    y, stats = ops.pointnet_l1_fwd(cuda(x), cuda(w), cuda(b))
    assert rel_err(y.float(), ref) < 1e-2
    assert rel_err(stats[:512], ref.double().sum(0)) < 1e-4
    assert rel_err(stats[512:], (ref.double() ** 2).sum(0)) < 1e-4
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    dW = ops.pointnet_l1_wgrad(cuda(x), cuda(dy).bfloat16())
    assert rel_err(dW, dy.t() @ x2) < 1e-4


# ------------------------------------------------------------------------------------------------ TCN pieces
@pytest.mark.parametrize("dil", [1, 2, 4])
def test_tcn_conv_as_gemm(ops, dil):
    g = torch.Generator().manual_seed(dil)
    B, T, Cin, Cout = 3, 30, 24, 16
    x = torch.randn(B, T, Cin, generator=g)
    w, b = torch.randn(Cout, Cin, 3, generator=g), torch.randn(Cout, generator=g)
    xx, ww = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = O.causal_dilated_conv(xx, ww, b, dil)
    col = ops.tcn_im2col(cuda(x), dil)
    y = ops.gemm(col, cuda(w).reshape(Cout, Cin * 3), trans_b=True, bias=cuda(b))
    assert rel_err(y, ref.detach().reshape(B * T, Cout)) < 1e-5
    dy = torch.randn(B, T, Cout, generator=g)
    ref.backward(dy)
    dcol = ops.gemm(cuda(dy).reshape(B * T, Cout), cuda(w).reshape(Cout, Cin * 3))
    dx = ops.tcn_col2im(dcol, B, T, Cin, dil)
    assert rel_err(dx, xx.grad) < 1e-5
    dW = ops.gemm(cuda(dy).reshape(B * T, Cout), col, trans_a=True)
    assert rel_err(dW.reshape(Cout, Cin, 3), ww.grad) < 1e-5


@pytest.mark.parametrize("dil", [1, 2, 4])
@pytest.mark.parametrize("B,Cin,Cout", [(5, 1024, 16), (3, 16, 32), (7, 256, 512), (64, 64, 128)])
def test_tcn_conv_on_tensor_cores(ops, dil, B, Cin, Cout):
    """The path engine.tcn_forward / tcn_backward actually run (models.py:59-76): bf16 im2col + tcgen05 GEMM (TC_PLAIN, fp32
    output) forward; weight gradient TC_WGRAD_ACC (k = the B*T rows) and data gradient TC_PLAIN + col2im backward; against
    autograd through oracle.causal_dilated_conv on bf16-rounded operands (so the comparison isolates the kernels: fp32
    accumulation order only -> 1e-4) and against the unrounded fp32 result (bf16 operand tolerance 1e-2)."""
    L = ops._lib
    g = torch.Generator().manual_seed(17 * dil + B + Cin)
    T = 30
    x = torch.randn(B, T, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, generator=g) / math.sqrt(3 * Cin)
    b = torch.randn(Cout, generator=g)
    xr, wr = bf16_round(x).requires_grad_(True), bf16_round(w).requires_grad_(True)
    ref = O.causal_dilated_conv(xr, wr, b, dil)
    R, K = B * T, Cin * 3
    col = ops.tcn_im2col(cuda(x), dil, torch.bfloat16)
    wb = ops.pack_bf16(cuda(w).reshape(Cout, K))
    y = ops.gemm_tc(col, wb, L.TC_PLAIN, R, Cout, K, bias=cuda(b), out_dtype=torch.float32)
    assert y.dtype == torch.float32 and tuple(y.shape) == (R, Cout)
    assert rel_err(y, ref.detach().reshape(R, Cout)) < 1e-4
    assert rel_err(y, O.causal_dilated_conv(x, w, b, dil).reshape(R, Cout)) < 1e-2
    dy = bf16_round(torch.randn(B, T, Cout, generator=g))
    ref.backward(dy)
    dyb = cuda(dy).reshape(R, Cout).bfloat16()
    dW = torch.zeros(Cout, K, device="cuda")
    ops.gemm_tc(dyb, col, L.TC_WGRAD_ACC, Cout, K, R, a_mn=L.OP_MN, b_mn=L.OP_MN, out=dW)
    assert rel_err(dW.reshape(Cout, Cin, 3), wr.grad) < 1e-4
    ops.gemm_tc(dyb, col, L.TC_WGRAD_ACC, Cout, K, R, a_mn=L.OP_MN, b_mn=L.OP_MN, out=dW)          # accumulates
    assert rel_err(dW.reshape(Cout, Cin, 3), 2 * wr.grad) < 1e-4
    dcol = ops.gemm_tc(dyb, wb, L.TC_PLAIN, R, K, Cout, b_mn=L.OP_MN, out_dtype=torch.float32)
    dx = ops.tcn_col2im(dcol, B, T, Cin, dil)
    assert rel_err(dx, xr.grad) < 1e-4


def test_small_helpers(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(11, 30, 40, generator=g)
    assert rel_err(ops.mean_rows(cuda(x)), x.mean(1)) < 1e-6
    gg = torch.randn(11, 40, generator=g)
    assert rel_err(ops.mean_rows_bwd(cuda(gg), 30), (gg / 30)[:, None, :].expand(11, 30, 40)) < 1e-6
    out = O.elu(torch.randn(50, 33, generator=g))
    dout = torch.randn(50, 33, generator=g)
    ref = dout * torch.where(out > 0, torch.ones_like(out), out + 1)
    assert rel_err(ops.elu_bwd_from_out(cuda(dout), cuda(out)), ref) < 1e-6
    assert rel_err(ops.colsum(cuda(dout)), dout.sum(0)) < 1e-5
    w = torch.randn(37, 1125, generator=g)
    pk = ops.pack_bf16(cuda(w), ld_out=1152)
    assert pk.shape == (37, 1152) and rel_err(pk[:, :1125].float(), bf16_round(w)) == 0 and float(pk[:, 1125:].float().abs().max()) == 0
    pt = ops.pack_bf16(cuda(w), ld_out=40, transpose=True)
    assert pt.shape == (1125, 40) and rel_err(pt[:, :37].float(), bf16_round(w.t())) == 0
    logits = torch.randn(9, 4, generator=g)
    gt = torch.randint(0, 4, (9,), generator=g)
    ll = logits.clone().requires_grad_(True)
    ce = O.cross_entropy(ll, gt)
    ce.backward()
    loss, dl, pred = ops.softmax_ce(cuda(logits), cuda(gt))
    assert abs(float(loss) - float(ce)) < 1e-5 and rel_err(dl, ll.grad) < 1e-5
    assert torch.equal(pred.cpu().long(), logits.argmax(1))


# ------------------------------------------------------------------------------------------------ Chamfer
def _check_idx(P, idx, ref_idx, dim):
    """argmins equal, or (listed exception class: fp near-ties) the chosen distances agree to 1e-4 relative; every
    exception is LISTED: (b, t, point) -> ours / reference index and the two distances."""
    idx, ref_idx = idx.cpu().long(), ref_idx.long()
    bad = idx != ref_idx
    if bad.any():
        d_ours = torch.gather(P, dim, idx.unsqueeze(dim)).squeeze(dim)
        d_ref = torch.gather(P, dim, ref_idx.unsqueeze(dim)).squeeze(dim)
        for b, t, n in torch.nonzero(bad).tolist()[:50]:
            print(f"  arg-min exception (dim {dim}) at b={b} t={t} point={n}: ours {int(idx[b, t, n])} (d={float(d_ours[b, t, n]):.7g}) "
                  f"vs reference {int(ref_idx[b, t, n])} (d={float(d_ref[b, t, n]):.7g})")
        assert float(((d_ours - d_ref).abs()[bad] / (d_ref.abs()[bad] + 1e-6)).max()) < 1e-4
    return int(bad.sum())


@pytest.mark.parametrize("B,N", [(2, 50), (3, 150), (1, 1), (2, 7), (1, 2), (2, 3), (2, 130), (1, 256)])
def test_chamfer(ops, B, N):
    gts, _ = O.synth_batch(B, N, 2, seed=N)
    g = torch.Generator().manual_seed(N)
    preds = torch.randn(gts.shape, generator=g) * 0.6
    pp = preds.clone().requires_grad_(True)
    loss, i1, i2 = O.chamfer(pp, gts)
    loss.backward()
    fl, j1, j2 = ops.chamfer_fwd(cuda(preds), cuda(gts))
    out = ops.chamfer_reduce(fl, True)
    assert abs(float(out) - float(loss)) / abs(float(loss)) < 1e-5
    per = ops.chamfer_reduce(fl, False)
    assert rel_err(per, O.chamfer(preds, gts, avg_out=False)[0]) < 1e-5
    P = O.pairwise_dist(gts, preds)
    nbad = _check_idx(P, j1, i1, 2) + _check_idx(P, j2, i2, 3)
    assert nbad <= 0.01 * i1.numel()           # duplicated (padded) gt points tie exactly; both sides pick the first
    grad = ops.chamfer_bwd(cuda(preds), cuda(gts), j1, j2, torch.ones((), device="cuda"), True)
    if nbad == 0:
        assert rel_err(grad, pp.grad) < 1e-5


def test_chamfer_module_gradients_for_both_clouds():
    """SeqChamferLoss through autograd (utils.py:98-107): gradient w.r.t. the predictions and, when asked for, w.r.t. the
    other cloud -- against autograd through the oracle's formulation."""
    from opensetgaitrecognition_pcaa_b200.utils import SeqChamferLoss
    g = torch.Generator().manual_seed(5)
    B, N = 2, 37
    p, q = torch.randn(B, 4, 30, N, generator=g), torch.randn(B, 4, 30, N, generator=g)
    for avg_out in (True, False):
        pr, qr = p.clone().requires_grad_(True), q.clone().requires_grad_(True)
        ref, _, _ = O.chamfer(pr, qr, avg_out)
        w = torch.ones_like(ref) if avg_out else torch.arange(1, B + 1).float()
        (ref * w).sum().backward()
        pc, qc = cuda(p).requires_grad_(True), cuda(q).requires_grad_(True)
        out = SeqChamferLoss()(pc, qc, avg_out=avg_out)
        (out * cuda(w)).sum().backward()
        assert rel_err(out.detach(), ref.detach()) < 1e-5
        assert rel_err(pc.grad, pr.grad) < 1e-5 and rel_err(qc.grad, qr.grad) < 1e-5


def test_chamfer_golden(ops, golden_dir):
    for name in ("n50_c2_b4", "n70_c4_b3"):
        gd = np.load(os.path.join(golden_dir, f"modules_{name}.npz"))
        B, nmax, C, seed = int(gd["B"]), int(gd["nmax"]), int(gd["C"]), int(gd["seed"])
        pcs, _ = O.synth_batch(B, nmax, C, seed=1234 + seed)
        rng = np.random.default_rng(77 + seed)
        pr = torch.from_numpy(rng.normal(0, 0.6, pcs.shape).astype(np.float32))
        fl, j1, j2 = ops.chamfer_fwd(cuda(pr), cuda(pcs))
        out = float(ops.chamfer_reduce(fl, True))
        assert abs(out - float(gd["chamfer2"])) / float(gd["chamfer2"]) < 1e-5
        P = O.pairwise_dist(pcs, pr)
        n1 = _check_idx(P, j1, torch.from_numpy(gd["chamfer2_idx_gt_for_pred"].astype(np.int64)), 2)
        n2 = _check_idx(P, j2, torch.from_numpy(gd["chamfer2_idx_pred_for_gt"].astype(np.int64)), 3)
        assert n1 + n2 <= 0.002 * j1.numel()


# ------------------------------------------------------------------------------------------------ critic
@pytest.mark.parametrize("B,C", [(4, 2), (16, 4), (37, 8), (300, 4)])
def test_wgangp_dstep(ops, B, C):
    g = torch.Generator().manual_seed(B + C)
    p = {k: v for k, v in O.det_params(C, 50, seed=B).items() if k.startswith("D.")}
    fv, z0 = torch.randn(B, 32, generator=g), torch.randn(B, 32, generator=g)
    means = O.sample_distant_points(32, C, 10, 10).float()
    gt = torch.randint(0, C, (B,), generator=g)
    alphas = torch.rand(B, 1, generator=g)
    oh = torch.nn.functional.one_hot(gt, C).float()
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    dl, gp = O.d_loss_fn(leaves, fv, z0 + oh @ means, oh, alphas, 15.0)
    names = list(leaves)
    ref = dict(zip(names, torch.autograd.grad(dl, [leaves[n] for n in names])))
    W = [cuda(p[f"D.model.{i}.{s}"]) for i in (0, 2, 4) for s in ("weight", "bias")]
    grads = [torch.zeros_like(w) for w in W]
    losses = ops.wgangp_dstep(cuda(fv), cuda(z0), cuda(means), cuda(gt), cuda(alphas), *W, 15.0, grads).cpu()
    assert abs(float(losses[0]) - float(dl)) < 2e-4 * max(1.0, abs(float(dl)))
    assert abs(float(losses[1]) - float(gp)) < 2e-4 * max(1.0, abs(float(gp)))
    for gr, (i, s) in zip(grads, [(i, s) for i in (0, 2, 4) for s in ("weight", "bias")]):
        r = ref[f"D.model.{i}.{s}"]
        assert float((gr.cpu() - r).abs().max()) < 2e-4 * (float(r.abs().max()) + 1e-3), (i, s)
    # critic forward + input gradient
    x = fv.clone().requires_grad_(True)
    o = O.disc_forward(p, x, oh)
    o.sum().backward()
    out, dx = ops.disc_fwd(cuda(fv), cuda(gt), *W, C, want_dx=True)
    assert rel_err(out, o.detach()) < 1e-5 and rel_err(dx, x.grad) < 1e-5


# ------------------------------------------------------------------------------------------------ Adam / scoring
def test_adam_flat(ops):
    g = torch.Generator().manual_seed(0)
    n = 100003
    p, m, v = torch.randn(n, generator=g), torch.zeros(n), torch.zeros(n)
    pd, md, vd = cuda(p), cuda(m), cuda(v)
    shadow = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    pr = [p.clone()]
    st = [dict()]
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * 10 ** float(torch.randint(-6, 1, (1,), generator=g))
        O.adam_update(pr, [gr], st, 1e-4, 0.9, 0.99)
        ops.adam_flat(pd, cuda(gr), md, vd, 1e-4, 0.9, 0.99, 1e-8, step, 1.0, shadow)
        assert float((pd.cpu() - pr[0]).abs().max()) < 2e-7
    assert torch.equal(shadow.cpu(), pd.cpu().bfloat16())


def test_openset_scoring(ops, golden_dir):
    gd = np.load(os.path.join(golden_dir, "scoring.npz"))
    emb, means = torch.from_numpy(gd["emb"]), torch.from_numpy(gd["means"])
    ll = ops.openset_score(cuda(emb), cuda(means)).cpu().numpy()
    ref = O.joint_log_likelihood(gd["emb"], gd["means"])
    assert np.max(np.abs(ll - ref)) < 1e-9                      # float64 on both sides
    lik = gd["lik"]
    nz = lik > 0
    assert np.max(np.abs(ll[nz] - np.log(lik[nz]))) < 1e-9      # == log of the reference's scipy pdf
    thr = float(gd["threshold"])
    preds = torch.from_numpy(gd["preds"].astype(np.int32))
    C = means.shape[0]
    for k in (1, 2, 4, 6):
        n = (len(lik) // k) * k
        votes = ops.openset_vote(cuda(torch.from_numpy(ll[:n])), cuda(preds[:n]), k, math.log(thr), C).cpu().numpy()
        assert np.array_equal(votes, gd[f"votes_k{k}"])          # bit-exact integer labels vs the reference


# ------------------------------------------------------------------------------------------------ decoder tensor-core path
@pytest.mark.parametrize("B,K,Nout", [(4, 64, 375), (32, 1125, 2250), (256, 375, 752), (130, 4500, 1000)])
def test_gemm_tc_decoder_modes(ops, B, K, Nout):
    from opensetgaitrecognition_pcaa_b200 import engine
    L = ops._lib
    g = torch.Generator().manual_seed(B + K + Nout)
    a = bf16_round(torch.randn(B, K, generator=g))
    W = torch.randn(Nout, K, generator=g) / math.sqrt(K)
    bias = torch.randn(Nout, generator=g)
    wb = ops.pack_bf16(cuda(W), ld_out=engine.pad8(K))
    Wr = bf16_round(W)
    ad = torch.zeros(B, engine.pad8(K), dtype=torch.bfloat16, device="cuda")
    ad[:, :K] = cuda(a).bfloat16()
    # forward: bias + ELU, bf16 out with padded leading dimension; and fp32 out
    out = ops.gemm_tc(ad, wb, L.TC_BIAS_ELU, B, Nout, K, bias=cuda(bias))
    ref = O.elu(a @ Wr.t() + bias)
    assert out.shape == (B, engine.pad8(Nout)) and rel_err(out[:, :Nout].float(), ref) < 1e-2
    if Nout % 4 == 0:
        o32 = ops.gemm_tc(ad, wb, L.TC_PLAIN, B, Nout, K, bias=cuda(bias), out_dtype=torch.float32)
        assert rel_err(o32, a @ Wr.t() + bias) < 1e-4
    # data gradient through the MN-major weight + ELU'(saved output)
    dz = bf16_round(torch.randn(B, Nout, generator=g))
    dzd = torch.zeros(B, engine.pad8(Nout), dtype=torch.bfloat16, device="cuda")
    dzd[:, :Nout] = cuda(dz).bfloat16()
    act_prev = bf16_round(O.elu(torch.randn(B, K, generator=g)))
    apd = torch.zeros(B, engine.pad8(K), dtype=torch.bfloat16, device="cuda")
    apd[:, :K] = cuda(act_prev).bfloat16()
    da = ops.gemm_tc(dzd, wb, L.TC_DGRAD_ELUOUT, B, K, Nout, b_mn=True, yprev=apd)
    ref = (dz @ Wr) * torch.where(act_prev > 0, torch.ones_like(act_prev), act_prev + 1)
    assert rel_err(da[:, :K].float(), ref) < 1e-2
    # weight gradient, plain stores, rows not 16-byte aligned when K % 4 != 0
    dW = torch.full((Nout, K), 7.0, device="cuda")
    ops.gemm_tc(dzd, apd, L.TC_WGRAD_STORE, Nout, K, B, a_mn=True, b_mn=True, out=dW)
    assert rel_err(dW, dz.double().t() @ act_prev.double()) < 1e-4
    assert rel_err(ops.colsum_ld(dzd, Nout), dz.sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------------ channel-major (T256) PointNet path
def _t256(ops, t):
    """[C, P] fp32 CPU -> T256 bf16 CUDA [tiles, C, 256] (pad points zero, the format's invariant)."""
    return ops.t256_pack(t.cuda())


def _un(ops, xT, P):
    return ops.t256_unpack(xT, P).float().cpu()


def _pad_is_zero(ops, xT, P):
    nt, C, _ = xT.shape
    flat = xT.permute(1, 0, 2).reshape(C, nt * 256)
    return bool((flat[:, P:] == 0).all())


T_SHAPES = [(512, 6000, 512), (1024, 4500, 512), (1024, 1000, 1024), (512, 300, 72), (264, 1003, 520), (128, 37, 64)]


@pytest.mark.parametrize("Cout,P,Cin", T_SHAPES)
def test_gemm_tc_t_bias_stats_and_affine(ops, Cout, P, Cin):
    g = torch.Generator().manual_seed(Cout + P + Cin)
    w = bf16_round(torch.randn(Cout, Cin, generator=g) / math.sqrt(Cin))
    aT = bf16_round(torch.randn(Cin, P, generator=g))
    bias = torch.randn(Cout, generator=g)
    ref = w @ aT + bias[:, None]
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    out = torch.full((ops.t256_tiles(P), Cout, 256), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.gemm_tc(cuda(w).bfloat16(), _t256(ops, aT), ops._lib.TC_T_BIAS_STATS, Cout, P, Cin, b_mn=ops._lib.OP_T256_MN, out=out,
                bias=cuda(bias), stats=stats)
    torch.cuda.synchronize()
    assert rel_err(_un(ops, out, P), ref) < 1e-2
    assert _pad_is_zero(ops, out, P)
    s = stats.cpu()
    assert rel_err(s[:Cout], ref.double().sum(1)) < 1e-4
    assert rel_err(s[Cout:], (ref.double() ** 2).sum(1)) < 1e-4
    # eval mode: BatchNorm (running statistics) + ELU applied in the epilogue
    sc, sh = 1 + 0.1 * torch.randn(Cout, generator=g), 0.1 * torch.randn(Cout, generator=g)
    out2 = ops.gemm_tc(cuda(w).bfloat16(), _t256(ops, aT), ops._lib.TC_T_AFFINE_ELU, Cout, P, Cin, b_mn=ops._lib.OP_T256_MN,
                       bias=cuda(bias), coef=cuda(torch.stack([sc, sh])))
    assert rel_err(_un(ops, out2, P), O.elu(ref * sc[:, None] + sh[:, None])) < 1e-2
    assert _pad_is_zero(ops, out2, P)


@pytest.mark.parametrize("Cout,G,n,Cin", [(1024, 60, 150, 1024), (256, 7, 32, 64), (384, 30, 50, 512), (128, 3, 300, 128)])
def test_gemm_tc_pooled_eval_epilogue(ops, Cout, G, n, Cin):
    """Eval-mode last PointNet layer with the mean pool over points fused into the GEMM epilogue: equals the unfused
    pair (affine + ELU GEMM, mean pool); groups straddle 32-column chunks, warp halves and 256-point tiles."""
    g = torch.Generator().manual_seed(Cout + G + n)
    P = G * n
    w = bf16_round(torch.randn(Cout, Cin, generator=g) / math.sqrt(Cin))
    aT = bf16_round(torch.randn(Cin, P, generator=g))
    bias = 0.1 * torch.randn(Cout, generator=g)
    coef = torch.stack([1 + 0.1 * torch.randn(Cout, generator=g), 0.1 * torch.randn(Cout, generator=g)])
    ref = O.elu((w @ aT + bias[:, None]) * coef[0][:, None] + coef[1][:, None])
    pooled = ops.gemm_tc_pooled(cuda(w).bfloat16(), _t256(ops, aT), Cout, P, Cin, n, bias=cuda(bias), coef=cuda(coef))
    torch.cuda.synchronize()
    assert pooled.shape == (G, Cout)
    assert rel_err(pooled, ref.reshape(Cout, G, n).mean(2).t()) < 2e-3
    # against the unfused kernels (same GEMM; the only difference is the bf16 rounding of the stored activation)
    a = ops.gemm_tc(cuda(w).bfloat16(), _t256(ops, aT), ops._lib.TC_T_AFFINE_ELU, Cout, P, Cin, b_mn=ops._lib.OP_T256_MN,
                    bias=cuda(bias), coef=cuda(coef))
    plain, _, _ = ops.bn_elu_meanpool_t(a, None, G, n)
    assert rel_err(pooled, plain) < 2e-3


@pytest.mark.parametrize("Cout,P,Cin", T_SHAPES)
def test_gemm_tc_t_dgrad_elubn(ops, Cout, P, Cin):
    g = torch.Generator().manual_seed(Cout + P + Cin + 5)
    w = bf16_round(torch.randn(Cout, Cin, generator=g) / math.sqrt(Cout))
    dyT = bf16_round(torch.randn(Cout, P, generator=g))
    yprev = bf16_round(torch.randn(Cin, P, generator=g))
    coef = torch.stack([1 + 0.1 * torch.randn(Cin, generator=g), 0.1 * torch.randn(Cin, generator=g),
                        0.1 * torch.randn(Cin, generator=g), 1 + 0.1 * torch.rand(Cin, generator=g)])
    z = yprev * coef[0][:, None] + coef[1][:, None]
    dz = (w.t() @ dyT) * torch.where(z > 0, torch.ones_like(z), torch.exp(z))
    xh = (yprev - coef[2][:, None]) * coef[3][:, None]
    stats = torch.zeros(2 * Cin, dtype=torch.float64, device="cuda")
    out = ops.gemm_tc(cuda(w).bfloat16(), _t256(ops, dyT), ops._lib.TC_T_DGRAD_ELUBN, Cin, P, Cout, a_mn=ops._lib.OP_MN,
                      b_mn=ops._lib.OP_T256_MN, stats=stats, yprev=_t256(ops, yprev), coef=cuda(coef))
    torch.cuda.synchronize()
    assert rel_err(_un(ops, out, P), dz) < 1e-2
    assert _pad_is_zero(ops, out, P)
    s = stats.cpu()
    assert rel_err(s[:Cin], dz.double().sum(1)) < 1e-3
    assert rel_err(s[Cin:], (dz.double() * xh.double()).sum(1)) < 1e-3


@pytest.mark.parametrize("Cout,P,Cin", T_SHAPES)
def test_gemm_tc_wgrad_points_k(ops, Cout, P, Cin):
    g = torch.Generator().manual_seed(Cout + P + Cin + 9)
    dyT = bf16_round(torch.randn(Cout, P, generator=g))
    aT = bf16_round(torch.randn(Cin, P, generator=g))
    dW = torch.zeros(Cout, Cin, device="cuda")
    K_, MN_ = ops._lib.OP_T256_K, ops._lib.OP_T256_MN
    ops.gemm_tc(_t256(ops, dyT), _t256(ops, aT), ops._lib.TC_WGRAD_ACC, Cout, Cin, P, a_mn=K_, b_mn=K_, out=dW)
    ref = dyT.double() @ aT.double().t()
    assert rel_err(dW, ref) < 1e-4
    ops.gemm_tc(_t256(ops, dyT), _t256(ops, aT), ops._lib.TC_WGRAD_ACC, Cout, Cin, P, a_mn=K_, b_mn=K_, out=dW)   # accumulates
    assert rel_err(dW, 2 * ref) < 1e-4


@pytest.mark.parametrize("B,N,C", [(2, 50, 512), (3, 7, 64), (1, 150, 1024), (5, 3, 32), (2, 8, 96), (3, 130, 64), (1, 256, 128), (1, 300, 64)])
def test_pointnet_t_kernels(ops, B, N, C):
    """Layer-1 conv, BN+ELU apply, mean pool (+ group sums), pooled backward and BN backward in the T256 layout against
    explicit fp32 formulas (P = B*30*N is not a multiple of 256; odd N and N < 8 take the general paths)."""
    g = torch.Generator().manual_seed(B * 100 + N)
    T = 30
    P, G = B * T * N, B * T
    x = torch.randn(B, 4, T, N, generator=g)
    w, b = torch.randn(C, 4, generator=g), torch.randn(C, generator=g)
    xr = x.permute(1, 0, 2, 3).reshape(4, P)                       # [f, p], p = (b, t, n)
    y_ref = w @ xr + b[:, None]
    yT, st = ops.pointnet_l1_fwd_t(cuda(x), cuda(w), cuda(b))
    assert yT.shape == ((P + 255) // 256, C, 256)
    assert rel_err(_un(ops, yT, P), y_ref) < 1e-2 and _pad_is_zero(ops, yT, P)
    s = st.cpu()
    assert rel_err(s[:C], y_ref.double().sum(1)) < 1e-5 and rel_err(s[C:], (y_ref.double() ** 2).sum(1)) < 1e-5
    coef = torch.stack([1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g),
                        0.1 * torch.randn(C, generator=g), 1 + 0.1 * torch.rand(C, generator=g)])
    sc, sh, mu, inv = (coef[i][:, None] for i in range(4))
    # eval-mode layer 1: activation directly
    aT, none = ops.pointnet_l1_fwd_t(cuda(x), cuda(w), cuda(b), coef=cuda(coef[:2]))
    assert none is None and rel_err(_un(ops, aT, P), O.elu(y_ref * sc + sh)) < 1e-2 and _pad_is_zero(ops, aT, P)
    # train-mode layer 1 with the BatchNorm statistics taken from the INPUT moments (y is linear in x): same coefficients as
    # the pass over y gives, running statistics updated the same way, and one kernel writes y and ELU(BN(y))
    mom = ops.input_moments(cuda(x)).cpu()
    xd = xr.double()
    want_mom = torch.cat([xd.sum(1), torch.stack([(xd[f] * xd[h]).sum() for f in range(4) for h in range(f, 4)])])
    assert rel_err(mom, want_mom) < 1e-6
    gam, bet = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    rm0, rv0 = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
    rm_a, rv_a, rm_b, rv_b = cuda(rm0), cuda(rv0), cuda(rm0), cuda(rv0)
    coef_stats = ops.bn_finalize(st, P, cuda(gam), cuda(bet), rm_a, rv_a)
    coef_mom = ops.bn_from_input_moments(ops.input_moments(cuda(x)), P, cuda(w), cuda(b), cuda(gam), cuda(bet), rm_b, rv_b)
    y64 = w.double() @ xd + b.double()[:, None]
    mean64, var64 = y64.mean(1), y64.var(1, unbiased=False)
    assert rel_err(coef_mom[2], mean64) < 1e-5 and rel_err(coef_mom[3], 1 / torch.sqrt(var64 + 1e-5)) < 1e-5
    assert rel_err(coef_mom, coef_stats) < 1e-4 and rel_err(rm_b, rm_a) < 1e-5 and rel_err(rv_b, rv_a) < 1e-4
    y2T, a2T = ops.pointnet_l1_fwd_bn_t(cuda(x), cuda(w), cuda(b), coef_mom)
    assert torch.equal(y2T, yT) and _pad_is_zero(ops, a2T, P)
    cm = coef_mom.cpu()
    assert rel_err(_un(ops, a2T, P), O.elu(y_ref * cm[0][:, None] + cm[1][:, None])) < 1e-2
    # from here on the bf16-rounded y is the common input
    yb = _un(ops, yT, P)
    a_ref = O.elu(yb * sc + sh)
    a = ops.bn_elu_apply_t(yT, cuda(coef), P)
    assert rel_err(_un(ops, a, P), a_ref) < 1e-2 and _pad_is_zero(ops, a, P)
    pooled, e1, e2 = ops.bn_elu_meanpool_t(yT, cuda(coef), G, N, want_e=True)
    z = yb * sc + sh
    d = torch.where(z > 0, torch.ones_like(z), torch.exp(z))
    xh = (yb - mu) * inv
    assert rel_err(pooled, a_ref.reshape(C, G, N).mean(2).t()) < 1e-5
    assert rel_err(e1, d.reshape(C, G, N).sum(2).t()) < 1e-5
    assert rel_err(e2, (d * xh).reshape(C, G, N).sum(2).t()) < 1e-4
    plain, _, _ = ops.bn_elu_meanpool_t(a, None, G, N)
    assert rel_err(plain, _un(ops, a, P).reshape(C, G, N).mean(2).t()) < 1e-5
    # backward of the pooled layer
    dpool = torch.randn(G, C, generator=g)
    dz_ref = (dpool.t().reshape(C, G, 1) / N).expand(C, G, N).reshape(C, P) * d
    st2 = ops.pool_bwd_stats(cuda(dpool), e1, e2, N).cpu()
    assert rel_err(st2[:C], dz_ref.double().sum(1)) < 1e-4
    assert rel_err(st2[C:], (dz_ref.double() * xh.double()).sum(1)) < 1e-4
    c = torch.stack([1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)])
    dy_ref = c[0][:, None] * dz_ref + c[1][:, None] * yb + c[2][:, None]
    dy = ops.pool_bwd_apply_t(cuda(dpool), yT, cuda(coef), cuda(c), N)
    assert rel_err(_un(ops, dy, P), dy_ref) < 1e-2 and _pad_is_zero(ops, dy, P)
    # generic BatchNorm backward (in place) and the fused layer-1 weight gradient
    dzb = _t256(ops, bf16_round(dz_ref))
    dzf = _un(ops, dzb, P)
    dW = ops.pointnet_l1_wgrad_t(cuda(x), dzb, yT, cuda(c))
    want = (c[0][:, None] * dzf + c[1][:, None] * yb + c[2][:, None]).double() @ xr.double().t()
    assert rel_err(dW, want) < 1e-4
    dW0 = ops.pointnet_l1_wgrad_t(cuda(x), dzb)
    assert rel_err(dW0, dzf.double() @ xr.double().t()) < 1e-4
    out = ops.bn_bwd_apply_t(dzb, yT, cuda(c), P, out=dzb)
    assert out.data_ptr() == dzb.data_ptr()
    assert rel_err(_un(ops, out, P), c[0][:, None] * dzf + c[1][:, None] * yb + c[2][:, None]) < 1e-2
    assert _pad_is_zero(ops, out, P)
