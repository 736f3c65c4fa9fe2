"""Drop-in replacements of the reference's ``models.py`` classes (same names, constructor arguments,
sub-module / ``state_dict`` keys and forward signatures), computing through libpcaa_sm100 on a B200.

Reference: models.py:6-34 (PointNetModule), 37-79 (DilTempConv1d), 82-160 (blocks), 232-292 (CGEncoder),
340-385 (CGDecoder), 405-421 (CGDiscriminator), 424-443 (GaussianMeanLearner).

The torch.nn layer objects are kept purely as parameter / buffer containers (so initialisation, ``.to()``,
``.parameters()``, ``state_dict()`` and checkpoints interchange with the reference); ``forward`` never calls
them -- it calls one autograd.Function per network whose forward/backward are sequences of C-ABI kernels
(engine.py).  CUDA only: there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import constants as _c
from . import engine, ops
from ._lib import EW_ADD, EW_ADD_ROWVEC, EW_ELU, EW_ELU_GRAD, EW_ELU_GRAD2, EW_MUL

constants = _c.get()


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: the PCAA B200 implementation runs on CUDA tensors only (no CPU fallback)")


def _named_tensors(module: torch.nn.Module):
    d = {k: v for k, v in module.named_parameters()}
    d.update({k: v for k, v in module.named_buffers()})
    return d


# ------------------------------------------------------------------------------------------------ stand-alone blocks
# CGEncoder never calls the layer objects below: its fused path (engine.encoder_forward) reads their parameters.  Their own
# forward()s exist for code that composes the blocks itself -- the reference's ORCEDEncoder does (models.py:446-495:
# pc_block -> AvgPool2d -> tc_block -> AvgPool1d) -- and run one shared autograd.Function: y = x W^T + b on the tensor cores
# (tcgen05, bf16 operands, fp32 accumulation; the CUDA-core GEMM when the inner dimension is not a multiple of 8, i.e. the
# K = 4 first layer), BatchNorm over the rows (batch statistics in training, running statistics in eval), ELU.
class _LinearBnEluFn(torch.autograd.Function):
    """a = ELU(BN(x W^T + b)) over rows: x [R, K] fp32, W [Cout, K] (a view of the conv weight), BatchNorm `bn`."""

    @staticmethod
    def forward(ctx, x, w2d, bias, gamma, beta, bn, training):
        from ._lib import TC_PLAIN
        x = x.contiguous()
        R, K = x.shape
        Cout = w2d.shape[0]
        tc = K % 8 == 0
        if tc:
            xb = ops.convert(x, torch.bfloat16)
            wb = ops.pack_bf16(w2d.contiguous())
            y = ops.gemm_tc(xb, wb, TC_PLAIN, R, Cout, K, bias=bias, out_dtype=torch.float32)
        else:
            xb, wb = x, None
            y = ops.gemm(x, w2d, trans_b=True, bias=bias)
        if training:
            coef = ops.bn_finalize(ops.colstats(y), R, gamma, beta, bn.running_mean, bn.running_var, engine.BN_MOMENTUM, engine.BN_EPS)
            bn.num_batches_tracked.add_(1)
        else:
            c2 = ops.bn_eval_coeffs(gamma, beta, bn.running_mean, bn.running_var, engine.BN_EPS)
            coef = torch.stack([c2[0], c2[1], bn.running_mean.float(), torch.rsqrt(bn.running_var.float() + engine.BN_EPS)])
        ctx.sv = (xb, wb, y, coef, w2d, tc, training)
        return ops.bn_elu_apply(y, coef[0], coef[1])

    @staticmethod
    def backward(ctx, dout):
        from ._lib import OP_MN, TC_PLAIN, TC_WGRAD_ACC
        xb, wb, y, coef, w2d, tc, training = ctx.sv
        R, Cout = y.shape
        K = w2d.shape[1]
        dz, st2 = ops.elu_bwd_colstats(dout.contiguous().float(), y, coef)
        if training:
            c, dgam, dbet = ops.bn_bwd_finalize(st2, R, coef)
        else:      # running statistics are constants: dy = scale * dz, d gamma = sum dz * xhat, d beta = sum dz
            c = torch.stack([coef[0], torch.zeros_like(coef[0]), torch.zeros_like(coef[0])])
            dgam, dbet = st2[Cout:].float(), st2[:Cout].float()
        if tc:
            dy = ops.bn_bwd_apply(dz, y, c, out_dtype=torch.bfloat16)
            dW = torch.zeros((Cout, K), device=y.device, dtype=torch.float32)
            ops.gemm_tc(dy, xb, TC_WGRAD_ACC, Cout, K, R, a_mn=OP_MN, b_mn=OP_MN, out=dW)
            dx = ops.gemm_tc(dy, wb, TC_PLAIN, R, K, Cout, b_mn=OP_MN, out_dtype=torch.float32) if ctx.needs_input_grad[0] else None
            dbias = ops.colsum_ld(dy, Cout)
        else:
            dy = ops.bn_bwd_apply(dz, y, c)
            dW = ops.gemm(dy, xb, trans_a=True)
            dx = ops.gemm(dy, w2d) if ctx.needs_input_grad[0] else None
            dbias = ops.colsum(dy)
        return dx, dW, dbias, dgam, dbet, None, None


def _only_elu(act, who):
    if not isinstance(act, torch.nn.ELU) or act.alpha != 1.0:
        raise NotImplementedError(f"{who}: the B200 kernels implement the reference's ELU(alpha=1) activation only")


class PointNetModule(torch.nn.Module):
    """Reference layout (models.py:6-34): module.0 = Conv2d(1x1), module.1 = BatchNorm2d, module.2 = activation."""

    def __init__(self, in_chs, out_chs, activation=torch.nn.ELU()):
        super().__init__()
        _only_elu(activation, "PointNetModule")
        self.module = torch.nn.Sequential(
            torch.nn.Conv2d(in_chs, out_chs, (1, 1), stride=1, padding="valid", dilation=1),
            torch.nn.BatchNorm2d(num_features=out_chs),
            activation,
        )

    def forward_rows(self, rows):
        """rows [R, Cin] fp32 (one row per point) -> [R, Cout]."""
        conv, bn = self.module[0], self.module[1]
        return _LinearBnEluFn.apply(rows, conv.weight.view(conv.weight.shape[0], -1), conv.bias, bn.weight, bn.bias, bn,
                                    self.training)

    def forward(self, x):
        """x (B, Cin, T, N) -> (B, Cout, T, N), as models.py:33-34."""
        _require_cuda(x, "PointNetModule")
        B, C, T, N = x.shape
        out = self.forward_rows(x.float().permute(0, 2, 3, 1).reshape(B * T * N, C))      # data movement only
        return out.view(B, T, N, -1).permute(0, 3, 1, 2)


class _Im2ColFn(torch.autograd.Function):
    """[B, T, Cin] -> [B*T, Cin*3]: the three causally shifted taps of a dilated k=3 convolution side by side."""

    @staticmethod
    def forward(ctx, h, dil):
        ctx.shape, ctx.dil = h.shape, dil
        return ops.tcn_im2col(h.contiguous(), dil, torch.float32)

    @staticmethod
    def backward(ctx, dcol):
        B, T, Cin = ctx.shape
        return ops.tcn_col2im(dcol.contiguous(), B, T, Cin, ctx.dil), None


class DilTempConv1d(torch.nn.Module):
    """Reference layout (models.py:37-79): conv1d (k=3, dilation d, padding 2d, last 2d outputs dropped), batch_norm, ELU."""

    def __init__(self, in_chs, out_chs, dilation, kernel_size=3, stride=1, use_bias=True, activation=torch.nn.ELU()):
        super().__init__()
        _only_elu(activation, "DilTempConv1d")
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("DilTempConv1d: the B200 kernels implement kernel_size=3, stride=1 (the reference's only use)")
        self.padding = int(np.floor((kernel_size - 1) * dilation))
        self.conv1d = torch.nn.Conv1d(in_chs, out_chs, kernel_size=kernel_size, stride=stride, padding=self.padding,
                                      dilation=dilation, bias=True)
        self.activation = activation
        self.batch_norm = torch.nn.BatchNorm1d(out_chs)

    def forward_btc(self, h):
        """h [B, T, Cin] fp32 (channels last) -> [B, T, Cout]."""
        B, T, _ = h.shape
        W = self.conv1d.weight                                           # (Cout, Cin, 3)
        # pcaa_tcn_im2col lays a row out as [Cin][3] (channel-major, tap inner) -- the memory order of Conv1d's (Cout, Cin, 3)
        # weight, so W.view(Cout, Cin*3) is the matching matrix (engine.tcn_forward relies on the same fact)
        w2d = W.view(W.shape[0], -1)
        col = _Im2ColFn.apply(h, int(self.conv1d.dilation[0]))
        a = _LinearBnEluFn.apply(col, w2d, self.conv1d.bias, self.batch_norm.weight, self.batch_norm.bias, self.batch_norm,
                                 self.training)
        return a.view(B, T, -1)

    def forward(self, x):
        """x (B, Cin, T) -> (B, Cout, T), as models.py:73-79."""
        _require_cuda(x, "DilTempConv1d")
        return self.forward_btc(x.float().permute(0, 2, 1)).permute(0, 2, 1)


class PointNetBlock(torch.nn.Module):
    def __init__(self):
        super().__init__()
        d = constants.POINTNET_OUT_DIM
        self.pointnet1 = PointNetModule(in_chs=constants.NFEATURES, out_chs=d // 2)
        self.pointnet2 = PointNetModule(in_chs=d // 2, out_chs=d // 2)
        self.pointnet3 = PointNetModule(in_chs=d // 2, out_chs=d)
        self.pointnet4 = PointNetModule(in_chs=d, out_chs=d)

    def forward(self, x):
        """x (B, 4, T, N) -> (B, 1024, T, N), as models.py:100-105 (the un-pooled activation: stand-alone use only; CGEncoder's
        fused path never materialises it)."""
        _require_cuda(x, "PointNetBlock")
        B, C, T, N = x.shape
        rows = x.float().permute(0, 2, 3, 1).reshape(B * T * N, C)
        for m in (self.pointnet1, self.pointnet2, self.pointnet3, self.pointnet4):
            rows = m.forward_rows(rows)
        return rows.view(B, T, N, -1).permute(0, 3, 1, 2)


class TemporalConvolutionBlock(torch.nn.Module):
    def __init__(self):
        super().__init__()
        f = constants.DTC_FILTERS
        chans = [constants.POINTNET_OUT_DIM] + list(f)
        for l, dil in enumerate(engine.DTC_DILATIONS, start=1):
            setattr(self, f"dtc{l}", DilTempConv1d(in_chs=chans[l - 1], out_chs=chans[l], dilation=dil, kernel_size=3))

    def forward(self, x):
        """x (B, 1024, T) -> (B, 512, T), as models.py:153-160."""
        _require_cuda(x, "TemporalConvolutionBlock")
        h = x.float().permute(0, 2, 1)
        for l in range(1, 7):
            h = getattr(self, f"dtc{l}").forward_btc(h)
        return h.permute(0, 2, 1)


# ------------------------------------------------------------------------------------------------ encoder
class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, names, *params):
        P = _named_tensors(module)
        training = module.training
        logits, fv, saved = engine.encoder_forward(x, P, training, module.use_projection_head)
        if training:
            for k, v in P.items():
                if k.endswith("num_batches_tracked"):
                    v.add_(1)
        ctx.saved, ctx.module, ctx.names, ctx.eval_mode = saved, module, names, not training
        ctx.set_materialize_grads(False)
        return logits, fv

    @staticmethod
    def backward(ctx, dlogits, dfv):
        P = _named_tensors(ctx.module)
        dlogits = None if dlogits is None else dlogits.contiguous()
        dfv = None if dfv is None else dfv.contiguous()
        if ctx.eval_mode:
            # the eval-mode forward folds BatchNorm into the GEMM epilogues and keeps no activations: there is nothing to
            # differentiate through (the reference's scripts only call the eval-mode encoder under no_grad():
            # PCAA_ablation.py:1045-1048, inference_PCAA.py:195-208, 239-251)
            raise RuntimeError("CGEncoder (B200): the eval-mode forward is inference-only and cannot be differentiated; "
                               "use .train(), or run it under torch.no_grad()")
        if ctx.saved is None:
            raise RuntimeError("CGEncoder (B200): backward through the same forward twice is not supported "
                               "(the saved activations are released after the first backward)")
        G = engine.encoder_backward(dlogits, dfv, ctx.saved, P)
        ctx.saved = None
        return (None, None, None) + tuple(G[n] for n in ctx.names)


class CGEncoder(torch.nn.Module):
    def __init__(self, n_out_labels, nmax_points=constants.NMAX, use_projection_head=False):
        super().__init__()
        self.use_projection_head = use_projection_head
        self.pc_block = PointNetBlock()
        self.glob_avg_pool1 = torch.nn.AvgPool2d(kernel_size=(1, nmax_points))
        self.tc_block = TemporalConvolutionBlock()
        self.glob_avg_pool2 = torch.nn.AvgPool1d(kernel_size=constants.NSTEPS)
        self.MLP_sup1 = torch.nn.Sequential(
            torch.nn.Linear(in_features=constants.DTC_FILTERS[-1], out_features=constants.SUP_LATENT_DIM),
            torch.nn.ELU(),
        )
        head_out = constants.SUP_LATENT_DIM if not use_projection_head else constants.SUP_LATENT_DIM // 2
        if self.use_projection_head:
            self.MLP_head = torch.nn.Sequential(
                torch.nn.Linear(in_features=constants.SUP_LATENT_DIM, out_features=head_out), torch.nn.ELU())
        self.MLP_sup2 = torch.nn.Sequential(
            torch.nn.Linear(in_features=head_out, out_features=n_out_labels), torch.nn.ELU())
        self.nmax_points = nmax_points

    def forward(self, x):
        _require_cuda(x, "CGEncoder")
        if x.shape[-1] != self.nmax_points or x.shape[2] != constants.NSTEPS:
            raise ValueError(f"CGEncoder expects (B,{constants.NFEATURES},{constants.NSTEPS},{self.nmax_points}), got {tuple(x.shape)}")
        x = x.float().contiguous()
        names = [k for k, _ in self.named_parameters()]
        out_classes, sup_fv = _EncoderFn.apply(x, self, names, *[p for _, p in self.named_parameters()])
        return out_classes, sup_fv


# ------------------------------------------------------------------------------------------------ decoder
class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, names, *params):
        P = _named_tensors(module)
        wb = engine.decoder_pack_weights(P, cache=module.__dict__.setdefault("_pcaa_wcache", {}))
        out, acts = engine.decoder_forward_tc(x, P, wb)
        ctx.acts, ctx.module, ctx.names, ctx.wb = acts, module, names, wb
        return out

    @staticmethod
    def backward(ctx, dout):
        P = _named_tensors(ctx.module)
        dx, G = engine.decoder_backward_tc(dout.contiguous(), ctx.acts, P, ctx.wb)
        ctx.acts = ctx.wb = None
        # bn1-4 are constructed but never applied (models.py:353-368 vs 373-385): their gradient stays None
        return (dx, None, None) + tuple(G.get(n) for n in ctx.names)


class CGDecoder(torch.nn.Module):
    def __init__(self, input_dim=constants.SUP_LATENT_DIM, nmax_points=constants.NMAX):
        super().__init__()
        self.decoder_mlp_size = constants.NSTEPS * constants.NFEATURES * nmax_points
        self.nmax_points = nmax_points
        self.activation = torch.nn.ELU()
        s = self.decoder_mlp_size
        self.dense1 = torch.nn.Linear(in_features=input_dim, out_features=s // 16)
        self.bn1 = torch.nn.BatchNorm1d(s // 16)
        self.dense2 = torch.nn.Linear(in_features=s // 16, out_features=s // 8)
        self.bn2 = torch.nn.BatchNorm1d(s // 8)
        self.dense3 = torch.nn.Linear(in_features=s // 8, out_features=s // 4)
        self.bn3 = torch.nn.BatchNorm1d(s // 4)
        self.dense4 = torch.nn.Linear(in_features=s // 4, out_features=s // 2)
        self.bn4 = torch.nn.BatchNorm1d(s // 2)
        self.dense5 = torch.nn.Linear(in_features=s // 2, out_features=s)

    def forward(self, x):
        _require_cuda(x, "CGDecoder")
        names = [k for k, _ in self.named_parameters()]
        out = _DecoderFn.apply(x.float().contiguous(), self, names, *[p for _, p in self.named_parameters()])
        return out.view(-1, constants.NFEATURES, constants.NSTEPS, self.nmax_points)


# ------------------------------------------------------------------------------------------------ critic
# The unmodified trainers call torch.autograd.grad(D(interp), interp, create_graph=True) and then backpropagate through
# that gradient (PCAA_ablation.py:955-973), so every Function below has a backward that is itself built from these
# Functions (closed under differentiation); each one is a single small CUDA kernel.  The fused trainer (train.py) uses
# the analytic single-kernel pcaa_wgangp_dstep instead.
class _MatMul(torch.autograd.Function):
    """op(a) @ op(b) through the CUDA-core GEMM."""

    @staticmethod
    def forward(ctx, a, b, ta, tb):
        ctx.save_for_backward(a, b)
        ctx.ta, ctx.tb = ta, tb
        return ops.gemm(a, b, trans_a=ta, trans_b=tb)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ta, tb = ctx.ta, ctx.tb
        g = g.contiguous()
        if not ta:
            da = _MatMul.apply(g, b, False, not tb)
        else:
            da = _MatMul.apply(b, g, tb, True)
        if not tb:
            db = _MatMul.apply(a, g, not ta, False)
        else:
            db = _MatMul.apply(g, a, True, ta)
        return da, db, None, None


class _ColSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.rows = x.shape[0]
        return ops.colsum(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return _BroadcastRows.apply(g, ctx.rows)


class _BroadcastRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, rows):
        return ops.ew(EW_ADD_ROWVEC, None, v.contiguous(), shape=(rows, v.numel()))

    @staticmethod
    def backward(ctx, g):
        return _ColSum.apply(g), None


class _AddRowVec(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, v):
        return ops.ew(EW_ADD_ROWVEC, x.contiguous(), v.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, _ColSum.apply(g)


class _Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return ops.ew(EW_MUL, a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return _Mul.apply(g, b), _Mul.apply(g, a)


class _EluD(torch.autograd.Function):
    """order-th derivative of ELU applied element-wise (order 0 = ELU itself)."""

    @staticmethod
    def forward(ctx, x, order):
        ctx.save_for_backward(x)
        ctx.order = order
        return ops.ew((EW_ELU, EW_ELU_GRAD, EW_ELU_GRAD2)[min(order, 2)], x.contiguous())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return _Mul.apply(g, _EluD.apply(x, ctx.order + 1)), None


class CGDiscriminator(torch.nn.Module):
    def __init__(self, n_in_labels):
        super().__init__()
        self.model = torch.nn.Sequential(
            torch.nn.Linear(constants.SUP_LATENT_DIM + n_in_labels, 64, bias=True),
            torch.nn.ELU(),
            torch.nn.Linear(64, 32, bias=True),
            torch.nn.ELU(),
            torch.nn.Linear(32, 1, bias=True),
        )

    def forward(self, x, label):
        _require_cuda(x, "CGDiscriminator")
        h = torch.cat([x, label], dim=-1).float()          # data movement only
        for i in (0, 2, 4):
            lin = self.model[i]
            h = _AddRowVec.apply(_MatMul.apply(h, lin.weight, False, True), lin.bias)
            if i < 4:
                h = _EluD.apply(h, 0)
        return h


# ------------------------------------------------------------------------------------------------ variant-1 learner
class GaussianMeanLearner(torch.nn.Module):
    """Learned class centroids of ablation variant 1 (models.py:424-443): 3 x (Linear, BatchNorm1d, ELU) + Linear."""

    def __init__(self, n_in_labels):
        super().__init__()
        self.model = torch.nn.Sequential(
            torch.nn.Linear(n_in_labels, 16, bias=True), torch.nn.BatchNorm1d(16), torch.nn.ELU(),
            torch.nn.Linear(16, 32, bias=True), torch.nn.BatchNorm1d(32), torch.nn.ELU(),
            torch.nn.Linear(32, 64, bias=True), torch.nn.BatchNorm1d(64), torch.nn.ELU(),
            torch.nn.Linear(64, constants.SUP_LATENT_DIM, bias=True),
        )

    def forward(self, x):
        _require_cuda(x, "GaussianMeanLearner")
        return _MeanLearnerFn.apply(x.float().contiguous(), self, *[p for _, p in self.named_parameters()])


class _MeanLearnerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, *params):
        m = module.model
        training = module.training
        sv = []
        h = x
        for i in (0, 3, 6):
            lin, bn = m[i], m[i + 1]
            y = ops.gemm(h, lin.weight, trans_b=True, bias=lin.bias)
            if training:
                coef = ops.bn_finalize(ops.colstats(y), y.shape[0], bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                       engine.BN_MOMENTUM, engine.BN_EPS)
                bn.num_batches_tracked.add_(1)
            else:
                c2 = ops.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, engine.BN_EPS)
                # (scale, shift, mean, invstd) as bn_finalize returns them, from the running statistics (16..64 floats)
                coef = torch.stack([c2[0], c2[1], bn.running_mean.float(), torch.rsqrt(bn.running_var.float() + engine.BN_EPS)])
            a = ops.bn_elu_apply(y, coef[0], coef[1])
            sv.append((h, y, coef))
            h = a
        out = ops.gemm(h, m[9].weight, trans_b=True, bias=m[9].bias)
        ctx.sv, ctx.h_last, ctx.module, ctx.training = sv, h, module, training
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.module.model
        dout = dout.contiguous()
        grads = {}
        grads["9.weight"] = ops.gemm(dout, ctx.h_last, trans_a=True)
        grads["9.bias"] = ops.colsum(dout)
        d = ops.gemm(dout, m[9].weight)
        for idx, i in reversed(list(enumerate((0, 3, 6)))):
            h_in, y, coef = ctx.sv[idx]
            dz, st2 = ops.elu_bwd_colstats(d, y, coef)
            if ctx.training:
                c, dgam, dbet = ops.bn_bwd_finalize(st2, y.shape[0], coef)
            else:
                # running statistics are constants: dy = scale * dz, d gamma = sum dz * xhat, d beta = sum dz
                Cc = y.shape[1]
                c = torch.stack([coef[0], torch.zeros_like(coef[0]), torch.zeros_like(coef[0])])
                dgam, dbet = st2[Cc:].float(), st2[:Cc].float()
            dy = ops.bn_bwd_apply(dz, y, c, out=dz)
            grads[f"{i + 1}.weight"], grads[f"{i + 1}.bias"] = dgam, dbet
            grads[f"{i}.weight"] = ops.gemm(dy, h_in, trans_a=True)
            grads[f"{i}.bias"] = ops.colsum(dy)
            d = ops.gemm(dy, m[i].weight)
        names = [k[len("model."):] for k, _ in ctx.module.named_parameters()]
        return (d, None) + tuple(grads[n] for n in names)


# ------------------------------------------------------------------------------------------------ dead classes
def _dead(name):
    class _Dead(torch.nn.Module):
        def __init__(self, *a, **k):
            # the reference's un-prefixed Encoder / Decoder / Discriminator read constants.UNSUP_LATENT_DIM, which
            # does not exist (models.py:202,301,393): constructing them raises AttributeError there too.
            raise AttributeError("module 'constants' has no attribute 'UNSUP_LATENT_DIM'")
    _Dead.__name__ = _Dead.__qualname__ = name
    return _Dead


Encoder = _dead("Encoder")
Decoder = _dead("Decoder")
Discriminator = _dead("Discriminator")
