"""Block-level forward / backward of the PCAA networks on top of the C-ABI kernels.

Everything here is orchestration: which kernel runs on which buffer.  Parameters come as dicts keyed with the
reference's ``state_dict`` names (relative to the module: e.g. ``pc_block.pointnet2.module.0.weight``), so the same
functions serve the drop-in ``nn.Module``s (models.py) and the fused trainer (train.py).

Layouts: point activations are channel-major bf16 in 256-point tiles ``[tiles, C, 256]`` (T256, include/pcaa.h); TCN /
head / decoder activations are fp32 ``[rows, C]``.  Reference: models.py:82-160, 232-292, 340-385.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import ops
from ._lib import (ACT_ELU, ACT_NONE, OP_MN, OP_T256_K, OP_T256_MN, TC_BIAS_ELU, TC_BIAS_STATS, TC_DGRAD_ELUBN, TC_DGRAD_ELUOUT, TC_PLAIN,
                   TC_T_AFFINE_ELU, TC_T_BIAS_STATS, TC_T_DGRAD_ELUBN, TC_WGRAD_ACC, TC_WGRAD_STORE)

# PCAA_TCN_FUSED=0 selects the unfused TCN layer pipeline (im2col, GEMM, colstats, finalize, apply as separate launches):
# kept for A/B timing of the fused one
TCN_FUSED = os.environ.get("PCAA_TCN_FUSED", "1") == "1"
T_STEPS = 30
DTC_DILATIONS = (1, 2, 4, 1, 2, 4)
BN_MOMENTUM = 0.1
BN_EPS = 1e-5

Params = Dict[str, torch.Tensor]
Grads = Dict[str, torch.Tensor]


def _out(gradbuf: Optional[Grads], name: str):
    return None if gradbuf is None else gradbuf.get(name)


def _zeros_like_param(gradbuf, name, ref):
    """A zeroed gradient tensor for `name`.  A gradient buffer that sets "__zeroed__" (the fused trainer: ONE fill of its
    flat encoder span per iteration) hands its slot out as it is."""
    t = _out(gradbuf, name)
    if t is None:
        return torch.zeros_like(ref)
    if not gradbuf.get("__zeroed__", False):
        t.zero_()
    return t


class BnSync:
    """SyncBN for data-parallel training: `reduce(t)` sums a statistics buffer over the ranks in place, `world`
    scales the per-rank row counts.  With it every BatchNorm of the encoder normalises with GLOBAL-batch statistics, so an
    N-rank iteration computes what one process would on the concatenated batch.  The reductions are NCCL all-reduces
    issued between kernels (20 small ones per iteration): an eager-step option, not capturable in the step's CUDA graph."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.calls = 0

    def reduce(self, t: torch.Tensor) -> torch.Tensor:
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        return t


def _bn_bwd_coefs(st2, R, Rg, coef, dgamma_out, dbeta_out, bn: Optional[BnSync]):
    """BatchNorm-backward coefficients c (for dy = c1*dz + c2*y + c3) and d gamma / d beta from the sums st2 = [sum dz,
    sum dz*xhat] of THIS rank's rows.  SyncBN: d gamma / d beta stay the local sums (the gradient exchange adds the ranks'),
    while c uses the sums and the row count of all ranks (the statistics couple every row of the global batch)."""
    if bn is None:
        return ops.bn_bwd_finalize(st2, R, coef, dgamma_out, dbeta_out)
    _, dgam, dbet = ops.bn_bwd_finalize(st2, R, coef, dgamma_out, dbeta_out)
    bn.reduce(st2)
    c, _, _ = ops.bn_bwd_finalize(st2, Rg, coef)
    return c, dgam, dbet


def _stats_arena(sizes, device):
    """One zero-filled double buffer for several [2*C] BatchNorm statistics accumulators (one fill instead of one each)."""
    buf = torch.zeros(2 * sum(sizes), device=device, dtype=torch.float64)
    out, o = [], 0
    for c in sizes:
        out.append(buf[o:o + 2 * c])
        o += 2 * c
    return out


# ====================================================================================================== PointNet
def _conv_w(P: Params, pre: str, l: int):
    W = P[f"{pre}pointnet{l}.module.0.weight"]
    return W.view(W.shape[0], W.shape[1])


def pointnet_forward(x: torch.Tensor, P: Params, training: bool, pre: str = "pc_block.", wb16: Optional[dict] = None,
                     bn: Optional[BnSync] = None):
    """x (B,4,T,N) fp32 -> pooled [B*T, 1024] fp32 (mean over the N points of ELU(BN(conv))), saved state.

    Activations are channel-major bf16 in 256-point tiles (T256 ``[tiles, C, 256]``): the tcgen05 GEMM of layer l
    computes ``y_l^T = W_l a_{l-1}^T`` with the channel as the accumulator row, so bias / BatchNorm coefficients are
    per-thread scalars and the batch statistics are in-thread sums of its epilogue.  ``wb16`` optionally maps the
    layer number to a ready bf16 copy of the [Cout, Cin] weight (the trainer's Adam-maintained shadow)."""
    B, F, T, N = x.shape
    R = B * T * N
    Rg = R * (bn.world if bn is not None else 1)          # rows behind the statistics (SyncBN: of all ranks)
    sv = {"x": x, "N": N, "R": R, "Rg": Rg, "G": B * T, "y": [None] * 5, "a": [None] * 5, "coef": [None] * 5, "wb": [None] * 5}

    def bn_coef(l, stats):
        k = f"{pre}pointnet{l}.module.1."
        if training:
            if bn is not None:
                bn.reduce(stats)
            return ops.bn_finalize(stats, Rg, P[k + "weight"], P[k + "bias"], P[k + "running_mean"], P[k + "running_var"],
                                   BN_MOMENTUM, BN_EPS)
        return ops.bn_eval_coeffs(P[k + "weight"], P[k + "bias"], P[k + "running_mean"], P[k + "running_var"], BN_EPS)

    b1 = P[f"{pre}pointnet1.module.0.bias"]
    if training:
        # y1 = W1 x + b1 is linear in the 4 input features: BatchNorm 1's batch statistics follow from the input moments
        # (14 numbers), so the coefficients exist BEFORE y1 does and one pass writes y1 and a1 = ELU(BN(y1)) -- y1 is not
        # re-read for the activation (-2.4 GB of HBM traffic at B = 256)
        k1 = f"{pre}pointnet1.module.1."
        mom = ops.input_moments(x)
        if bn is not None:
            bn.reduce(mom)
        coef = ops.bn_from_input_moments(mom, Rg, _conv_w(P, pre, 1), b1, P[k1 + "weight"], P[k1 + "bias"],
                                         P[k1 + "running_mean"], P[k1 + "running_var"], BN_MOMENTUM, BN_EPS)
        y, a = ops.pointnet_l1_fwd_bn_t(x, _conv_w(P, pre, 1), b1, coef)
        sv["y"][1], sv["coef"][1], sv["a"][1] = y, coef, a
        arena = dict(zip((2, 3, 4), _stats_arena([_conv_w(P, pre, l).shape[0] for l in (2, 3, 4)], x.device)))
    else:
        a, _ = ops.pointnet_l1_fwd_t(x, _conv_w(P, pre, 1), b1, coef=bn_coef(1, None))
    for l in (2, 3, 4):
        W = _conv_w(P, pre, l)
        Cout, Cin = W.shape
        wb = wb16[l] if wb16 is not None else ops.pack_bf16(W)
        bias = P[f"{pre}pointnet{l}.module.0.bias"]
        if training:
            st = arena[l]
            y = ops.gemm_tc(wb, a, TC_T_BIAS_STATS, Cout, R, Cin, b_mn=OP_T256_MN, bias=bias, stats=st)
            coef = bn_coef(l, st)
            sv["y"][l], sv["coef"][l], sv["wb"][l] = y, coef, wb
            if l < 4:
                a = ops.bn_elu_apply_t(y, coef, R)
                sv["a"][l] = a
        elif l == 4 and N >= 32:
            # eval: the mean pool over the N points of a frame runs in the epilogue of the last layer's GEMM; the
            # [1024, P] activation (the largest tensor of the forward) is never written or re-read
            return ops.gemm_tc_pooled(wb, a, Cout, R, Cin, N, bias=bias, coef=bn_coef(l, None)), sv
        else:
            a = ops.gemm_tc(wb, a, TC_T_AFFINE_ELU, Cout, R, Cin, b_mn=OP_T256_MN, bias=bias, coef=bn_coef(l, None))
    if training:
        pooled, e1, e2 = ops.bn_elu_meanpool_t(y, coef, B * T, N, want_e=True)
        sv["e1"], sv["e2"] = e1, e2
    else:
        pooled, _, _ = ops.bn_elu_meanpool_t(a, None, B * T, N)
    return pooled, sv


def pointnet_backward(gpool: torch.Tensor, sv, P: Params, gradbuf: Optional[Grads] = None,
                      pre: str = "pc_block.", side: Optional[torch.cuda.Stream] = None, bn: Optional[BnSync] = None,
                      pause_after: Optional[int] = None):
    """gpool [B*T, 1024] fp32 = d loss / d pooled.  Returns parameter gradients (written into gradbuf when given).

    With a `side` stream the weight-gradient GEMM of layer l (tensor bound, reads dy_l and a_{l-1}) runs there while the
    current stream goes on with the data gradient's successor, the HBM-bound BatchNorm-backward pass of layer l-1: the
    two are independent and have complementary bottlenecks.  The current stream waits for `side` before returning.

    `pause_after=l` (4, 3 or 2) returns ``(G, resume)`` once the weight AND data gradient of layer l are enqueued: at that
    point the gradients of layers >= l are final and W_l is no longer read, so a data-parallel trainer can start exchanging
    (and updating) them while ``resume()`` -- the rest of the backward -- runs."""
    G: Grads = {}
    steps = _pointnet_backward_steps(gpool, sv, P, gradbuf, pre, side, bn, G, pause_after)
    for l in steps:
        if pause_after is not None and l == pause_after:
            def resume():
                for _ in steps:
                    pass
                return G
            return G, resume
    if pause_after is not None:
        raise ValueError("pointnet_backward: pause_after must be 4, 3 or 2")
    return G


def _pointnet_backward_steps(gpool, sv, P, gradbuf, pre, side, bn, G, pause_after=None):
    """Generator behind pointnet_backward: yields l after the weight and data gradients of layer l (4, 3, 2) are enqueued."""
    R, N, Rg = sv["R"], sv["N"], sv["Rg"]
    if (bn is not None) != (Rg != R):
        raise RuntimeError("pointnet_backward: SyncBN must be used in both the forward and the backward")
    gpool = gpool.contiguous()
    main = torch.cuda.current_stream()
    # layer 4: the BatchNorm-backward statistics follow from the forward's group sums (no pass over y4), then ONE pass
    # forms dy4 = BN'(ELU'(pool'(gpool)))
    st2 = ops.pool_bwd_stats(gpool, sv["e1"], sv["e2"], N)
    kb = f"{pre}pointnet4.module.1."
    c, dgam, dbet = _bn_bwd_coefs(st2, R, Rg, sv["coef"][4], _out(gradbuf, kb + "weight"), _out(gradbuf, kb + "bias"), bn)
    G[kb + "weight"], G[kb + "bias"] = dgam, dbet
    dy = ops.pool_bwd_apply_t(gpool, sv["y"][4], sv["coef"][4], c, N)
    arena = dict(zip((4, 3, 2), _stats_arena([P[f"{pre}pointnet{l}.module.0.weight"].shape[1] for l in (4, 3, 2)], gpool.device)))
    for l in (4, 3, 2):
        kc = f"{pre}pointnet{l}.module.0."
        W = P[kc + "weight"]
        Cout, Cin = W.shape[0], W.shape[1]

        def wgrad(dy=dy, l=l, kc=kc, W=W, Cout=Cout, Cin=Cin):
            dW = _zeros_like_param(gradbuf, kc + "weight", W)
            # dW[Cout, Cin] += dyT [Cout, P] . a_{l-1}T [Cin, P]^T   (both operands K-major, K = points, split over the SMs)
            ops.gemm_tc(dy, sv["a"][l - 1], TC_WGRAD_ACC, Cout, Cin, R, a_mn=OP_T256_K, b_mn=OP_T256_K, out=dW.view(Cout, Cin))
            G[kc + "weight"] = dW
        if side is None:
            wgrad()
        else:
            side.wait_stream(main)                    # dy_l is complete
            with torch.cuda.stream(side):
                wgrad()
            dy.record_stream(side)                    # dy_l is released on this stream while `side` may still read it
        # the conv bias feeds a train-mode BatchNorm: its gradient is identically zero
        G[kc + "bias"] = _zeros_like_param(gradbuf, kc + "bias", P[kc + "bias"])
        # dz_{l-1}T [Cin, P] = (W^T dyT) * ELU'(BN(y_{l-1})) with the statistics of BatchNorm l-1's backward
        st2 = arena[l]
        dz = ops.gemm_tc(sv["wb"][l], dy, TC_T_DGRAD_ELUBN, Cin, R, Cout, a_mn=OP_MN, b_mn=OP_T256_MN, stats=st2,
                         yprev=sv["y"][l - 1], coef=sv["coef"][l - 1])
        kb = f"{pre}pointnet{l - 1}.module.1."
        c, dgam, dbet = _bn_bwd_coefs(st2, R, Rg, sv["coef"][l - 1], _out(gradbuf, kb + "weight"), _out(gradbuf, kb + "bias"), bn)
        G[kb + "weight"], G[kb + "bias"] = dgam, dbet
        if side is not None and l == pause_after:
            main.wait_stream(side)                    # the caller hands these gradients on: the weight gradients must be in
        yield l
        if l > 2:
            dy = ops.bn_bwd_apply_t(dz, sv["y"][l - 1], c, R, out=dz)
    kc = f"{pre}pointnet1.module.0."
    W1 = P[kc + "weight"]
    o = _out(gradbuf, kc + "weight")
    # layer 1: BatchNorm backward fused into the K = 4 weight gradient (dy1 is never materialised)
    dW1 = ops.pointnet_l1_wgrad_t(sv["x"], dz, sv["y"][1], c, None if o is None else o.view(W1.shape[0], 4))
    G[kc + "weight"] = dW1.view(W1.shape) if o is None else o
    G[kc + "bias"] = _zeros_like_param(gradbuf, kc + "bias", P[kc + "bias"])
    if side is not None:
        main.wait_stream(side)
        for l in (1, 2, 3):
            sv["a"][l].record_stream(side)


# ====================================================================================================== TCN
def _tcn_bn(P: Params, k: str):
    return (P[k + "batch_norm.weight"], P[k + "batch_norm.bias"], P[k + "batch_norm.running_mean"], P[k + "batch_norm.running_var"])


def tcn_forward(h: torch.Tensor, P: Params, training: bool, pre: str = "tc_block.", wb16: Optional[dict] = None,
                bn: Optional[BnSync] = None):
    """h [B, T, 1024] fp32 -> [B, T, 512]; causal dilated conv (bf16 im2col operand + tcgen05 GEMM, fp32 accumulate /
    output) + BatchNorm1d + ELU, six times.  ``wb16`` optionally maps the layer number to a ready bf16 copy of the
    [Cout, Cin*3] weight (the trainer's Adam-maintained shadow).

    Per layer two launches: the GEMM, whose epilogue also accumulates the BatchNorm column statistics (training), and one
    kernel that turns those into coefficients, applies BatchNorm + ELU and writes the result directly as the NEXT layer's
    im2col operand (the last layer writes the fp32 activation)."""
    B, T, _ = h.shape
    R = B * T
    sv = {"B": B, "T": T, "col": [], "y": [], "coef": [], "cin": [], "wb": [], "sync": bn is not None}
    if not TCN_FUSED or (bn is not None and training):
        # SyncBN: the statistics are all-reduced between the statistics kernel and the coefficient kernel of every layer
        return _tcn_forward_unfused(h, P, training, pre, wb16, sv, bn)
    chans = [P[f"{pre}dtc{l}.conv1d.weight"].shape[0] for l in range(1, 7)]
    arena = _stats_arena(chans, h.device) if training else [None] * 6
    col = ops.tcn_im2col(h, DTC_DILATIONS[0], torch.bfloat16)
    act = None
    for l in range(1, 7):
        k = f"{pre}dtc{l}."
        W = P[k + "conv1d.weight"]
        Cout, Cin, _ = W.shape
        wb = wb16[l] if wb16 is not None else ops.pack_bf16(W.view(Cout, Cin * 3))
        gamma, beta, rmean, rvar = _tcn_bn(P, k)
        dil_next = DTC_DILATIONS[l] if l < 6 else 0
        if training:
            y = ops.gemm_tc(col, wb, TC_BIAS_STATS, R, Cout, Cin * 3, bias=P[k + "conv1d.bias"], stats=arena[l - 1],
                            out_dtype=torch.float32)
            nxt, act, coef = ops.tcn_bn_elu_next(y, B, T, stats=arena[l - 1], gamma=gamma, beta=beta, running_mean=rmean,
                                                 running_var=rvar, momentum=BN_MOMENTUM, eps=BN_EPS, dil_next=dil_next,
                                                 want_act=(l == 6))
        else:
            y = ops.gemm_tc(col, wb, TC_PLAIN, R, Cout, Cin * 3, bias=P[k + "conv1d.bias"], out_dtype=torch.float32)
            coef = ops.bn_eval_coeffs(gamma, beta, rmean, rvar, BN_EPS)
            nxt, act, _ = ops.tcn_bn_elu_next(y, B, T, scale=coef[0], shift=coef[1], dil_next=dil_next, want_act=(l == 6))
        sv["col"].append(col), sv["y"].append(y), sv["coef"].append(coef), sv["cin"].append(Cin), sv["wb"].append(wb)
        col = nxt
    return act.view(B, T, -1), sv


def _tcn_forward_unfused(h, P, training, pre, wb16, sv, bn=None):
    B, T, _ = h.shape
    R = B * T
    Rg = R * (bn.world if bn is not None else 1)
    for l in range(1, 7):
        k = f"{pre}dtc{l}."
        W = P[k + "conv1d.weight"]
        Cout, Cin, _ = W.shape
        wb = wb16[l] if wb16 is not None else ops.pack_bf16(W.view(Cout, Cin * 3))
        col = ops.tcn_im2col(h, DTC_DILATIONS[l - 1], torch.bfloat16)
        y = ops.gemm_tc(col, wb, TC_PLAIN, R, Cout, Cin * 3, bias=P[k + "conv1d.bias"], out_dtype=torch.float32)
        if training:
            st = ops.colstats(y)
            if bn is not None:
                bn.reduce(st)
            coef = ops.bn_finalize(st, Rg, *_tcn_bn(P, k), BN_MOMENTUM, BN_EPS)
        else:
            coef = ops.bn_eval_coeffs(*_tcn_bn(P, k), BN_EPS)
        a = ops.bn_elu_apply(y, coef[0], coef[1])
        sv["col"].append(col), sv["y"].append(y), sv["coef"].append(coef), sv["cin"].append(Cin), sv["wb"].append(wb)
        h = a.view(B, T, Cout)
    return h, sv


def tcn_backward(dout: torch.Tensor, sv, P: Params, gradbuf: Optional[Grads] = None, pre: str = "tc_block.",
                 dout_is_frame_mean: bool = False, bn: Optional[BnSync] = None):
    """dout [B, T, 512] (or, with dout_is_frame_mean, the gradient [B, 512] of the mean over the T frames, whose broadcast
    is then formed inside the first kernel) -> (d input [B, T, 1024], parameter gradients).

    Per layer four launches: ELU' * d + BatchNorm-backward sums (d read straight from the layer above's d-im2col: the
    col2im is folded in), BatchNorm backward -> bf16 dy (coefficients and d gamma / d beta derived in the same kernel),
    weight-gradient GEMM, data-gradient GEMM."""
    G: Grads = {}
    B, T = sv["B"], sv["T"]
    R = B * T
    if (bn is not None) != sv.get("sync", False):
        raise RuntimeError("tcn_backward: SyncBN must be used in both the forward and the backward")
    if not TCN_FUSED or bn is not None:
        if dout_is_frame_mean:
            dout = ops.mean_rows_bwd(dout, T)
        return _tcn_backward_unfused(dout, sv, P, gradbuf, pre, bn)
    chans = [P[f"{pre}dtc{l}.conv1d.weight"].shape[0] for l in range(1, 7)]
    arena = _stats_arena(chans, dout.device)
    src, mode, dil_up = dout.reshape(-1, dout.shape[-1]).contiguous(), (2 if dout_is_frame_mean else 0), 0
    for l in range(6, 0, -1):
        k = f"{pre}dtc{l}."
        W = P[k + "conv1d.weight"]
        Cout, Cin, _ = W.shape
        y, coef, col, wb = sv["y"][l - 1], sv["coef"][l - 1], sv["col"][l - 1], sv["wb"][l - 1]
        if len(coef) < 4:
            raise RuntimeError("tcn_backward: the forward ran in eval mode (no batch statistics were saved)")
        dz = ops.tcn_elu_bwd_stats(src, mode, dil_up, y, coef, arena[l - 1], B, T)
        dy, dgam, dbet = ops.tcn_bn_bwd_apply(dz, y, arena[l - 1], coef, _out(gradbuf, k + "batch_norm.weight"),
                                              _out(gradbuf, k + "batch_norm.bias"))
        G[k + "batch_norm.weight"], G[k + "batch_norm.bias"] = dgam, dbet
        # dW [Cout, Cin*3] += dy^T col (k = the B*T rows, split over the SMs), dcol = dy W
        dW = _zeros_like_param(gradbuf, k + "conv1d.weight", W)
        ops.gemm_tc(dy, col, TC_WGRAD_ACC, Cout, Cin * 3, R, a_mn=OP_MN, b_mn=OP_MN, out=dW.view(Cout, Cin * 3))
        G[k + "conv1d.weight"] = dW
        # the conv bias feeds a train-mode BatchNorm: its gradient is identically zero (sum_rows dy = 0)
        G[k + "conv1d.bias"] = _zeros_like_param(gradbuf, k + "conv1d.bias", P[k + "conv1d.bias"])
        src = ops.gemm_tc(dy, wb, TC_PLAIN, R, Cin * 3, Cout, b_mn=OP_MN, out_dtype=torch.float32)
        mode, dil_up = 1, DTC_DILATIONS[l - 1]
    d = ops.tcn_col2im(src, B, T, sv["cin"][0], DTC_DILATIONS[0])
    return d.view(B, T, -1), G


def _tcn_backward_unfused(dout, sv, P, gradbuf, pre, bn=None):
    G: Grads = {}
    B, T = sv["B"], sv["T"]
    R = B * T
    Rg = R * (bn.world if bn is not None else 1)
    d = dout.reshape(R, -1)
    for l in range(6, 0, -1):
        k = f"{pre}dtc{l}."
        W = P[k + "conv1d.weight"]
        Cout, Cin, _ = W.shape
        y, coef, col, wb = sv["y"][l - 1], sv["coef"][l - 1], sv["col"][l - 1], sv["wb"][l - 1]
        dz, st2 = ops.elu_bwd_colstats(d, y, coef)
        c, dgam, dbet = _bn_bwd_coefs(st2, R, Rg, coef, _out(gradbuf, k + "batch_norm.weight"), _out(gradbuf, k + "batch_norm.bias"), bn)
        G[k + "batch_norm.weight"], G[k + "batch_norm.bias"] = dgam, dbet
        dy = ops.bn_bwd_apply(dz, y, c, out_dtype=torch.bfloat16)
        dW = _zeros_like_param(gradbuf, k + "conv1d.weight", W)
        ops.gemm_tc(dy, col, TC_WGRAD_ACC, Cout, Cin * 3, R, a_mn=OP_MN, b_mn=OP_MN, out=dW.view(Cout, Cin * 3))
        G[k + "conv1d.weight"] = dW
        G[k + "conv1d.bias"] = _zeros_like_param(gradbuf, k + "conv1d.bias", P[k + "conv1d.bias"])
        dcol = ops.gemm_tc(dy, wb, TC_PLAIN, R, Cin * 3, Cout, b_mn=OP_MN, out_dtype=torch.float32)
        d = ops.tcn_col2im(dcol, B, T, Cin, DTC_DILATIONS[l - 1]).view(R, Cin)
    return d.view(B, T, -1), G


# ====================================================================================================== Linear+ELU
def linear_forward(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor, act: int = ACT_ELU) -> torch.Tensor:
    return ops.gemm(x, W, trans_b=True, bias=b, act=act)


def linear_backward(dout: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor], W: torch.Tensor, wname: str,
                    bname: str, G: Grads, gradbuf: Optional[Grads], need_dx: bool = True, dx_out=None, dx_acc=False):
    """Backward of out = act(x W^T + b); `out` is the saved OUTPUT when the layer has an ELU, else None."""
    dz = ops.elu_bwd_from_out(dout, out) if out is not None else dout
    G[wname] = ops.gemm(dz, x, trans_a=True, out=_out(gradbuf, wname))
    G[bname] = ops.colsum(dz, _out(gradbuf, bname))
    if not need_dx:
        return None
    return ops.gemm(dz, W, out=dx_out, accumulate=dx_acc)


# ====================================================================================================== heads
def heads_forward(h6: torch.Tensor, P: Params, use_projection_head: bool):
    """h6 [B, T, 512] -> logits (B,C), sup_fv (B,32) (models.py:284-292)."""
    g = ops.mean_rows(h6)
    fv = linear_forward(g, P["MLP_sup1.0.weight"], P["MLP_sup1.0.bias"])
    hh = linear_forward(fv, P["MLP_head.0.weight"], P["MLP_head.0.bias"]) if use_projection_head else fv
    logits = linear_forward(hh, P["MLP_sup2.0.weight"], P["MLP_sup2.0.bias"])
    return logits, fv, {"g": g, "fv": fv, "hh": hh, "logits": logits, "T": h6.shape[1], "head": use_projection_head}


def heads_backward(dlogits: Optional[torch.Tensor], dfv_ext: Optional[torch.Tensor], sv, P: Params,
                   gradbuf: Optional[Grads] = None):
    """Returns (d g [B,512] = gradient of the frame mean of h6, grads).  dfv_ext is the gradient reaching sup_fv from
    outside the encoder."""
    G: Grads = {}
    fv = sv["fv"]
    dfv = None if dfv_ext is None else dfv_ext.clone()
    if dlogits is not None:
        dhh = linear_backward(dlogits, sv["hh"], sv["logits"], P["MLP_sup2.0.weight"], "MLP_sup2.0.weight",
                              "MLP_sup2.0.bias", G, gradbuf, dx_out=None if sv["head"] else dfv,
                              dx_acc=(not sv["head"]) and dfv is not None)
        if sv["head"]:
            dfv = linear_backward(dhh, fv, sv["hh"], P["MLP_head.0.weight"], "MLP_head.0.weight", "MLP_head.0.bias", G,
                                  gradbuf, dx_out=dfv, dx_acc=dfv is not None)
        else:
            dfv = dhh
    else:
        for n in ("MLP_sup2.0.weight", "MLP_sup2.0.bias") + (("MLP_head.0.weight", "MLP_head.0.bias") if sv["head"] else ()):
            G[n] = _zeros_like_param(gradbuf, n, P[n])
    dg = linear_backward(dfv, sv["g"], fv, P["MLP_sup1.0.weight"], "MLP_sup1.0.weight", "MLP_sup1.0.bias", G, gradbuf)
    return dg, G        # [B, 512]: the gradient of the mean over frames; tcn_backward broadcasts it (/ T) in its first kernel


# ====================================================================================================== encoder
def encoder_forward(x: torch.Tensor, P: Params, training: bool, use_projection_head: bool, wb16: Optional[dict] = None,
                    tcn_wb16: Optional[dict] = None, bn: Optional[BnSync] = None):
    B, F, T, N = x.shape
    pooled, sv_p = pointnet_forward(x, P, training, wb16=wb16, bn=bn if training else None)
    h6, sv_t = tcn_forward(pooled.view(B, T, -1), P, training, wb16=tcn_wb16, bn=bn if training else None)
    logits, fv, sv_h = heads_forward(h6, P, use_projection_head)
    return logits, fv, (sv_p, sv_t, sv_h)


def encoder_backward(dlogits, dfv, saved, P: Params, gradbuf: Optional[Grads] = None,
                     side: Optional[torch.cuda.Stream] = None, bn: Optional[BnSync] = None, pause_after: Optional[int] = None):
    """Backward of the whole encoder.  With `pause_after=l` returns ``(G, resume)`` after PointNet layer l (see
    pointnet_backward): heads, TCN and the PointNet layers >= l are then final; ``resume()`` finishes and returns G."""
    sv_p, sv_t, sv_h = saved
    dg, G = heads_backward(dlogits, dfv, sv_h, P, gradbuf)
    dpool, Gt = tcn_backward(dg, sv_t, P, gradbuf, dout_is_frame_mean=True, bn=bn)
    G.update(Gt)
    r = pointnet_backward(dpool.reshape(-1, dpool.shape[-1]), sv_p, P, gradbuf, side=side, bn=bn, pause_after=pause_after)
    if pause_after is None:
        G.update(r)
        return G
    Gp, resume_p = r
    G.update(Gp)

    def resume():
        G.update(resume_p())
        return G
    return G, resume


# ====================================================================================================== decoder
def decoder_forward(h: torch.Tensor, P: Params, pre: str = ""):
    """h [B, input_dim] -> [B, 4*30*nmax] (five Linear layers, ELU after the first four; models.py:373-382)."""
    acts = [h]
    x = h
    for l in range(1, 6):
        x = linear_forward(x, P[f"{pre}dense{l}.weight"], P[f"{pre}dense{l}.bias"], ACT_ELU if l < 5 else ACT_NONE)
        acts.append(x)
    return x, acts


def decoder_backward(dout: torch.Tensor, acts, P: Params, gradbuf: Optional[Grads] = None, pre: str = "",
                     need_dx: bool = True):
    G: Grads = {}
    d = dout
    for l in range(5, 0, -1):
        d = linear_backward(d, acts[l - 1], acts[l] if l < 5 else None, P[f"{pre}dense{l}.weight"],
                            f"{pre}dense{l}.weight", f"{pre}dense{l}.bias", G, gradbuf, need_dx=(l > 1 or need_dx))
    return d, G


# ------------------------------------------------------------------------------------------------ decoder, tensor cores
def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def decoder_pack_weights(P: Params, pre: str = "", cache: Optional[dict] = None):
    """bf16 operand copies of the five decoder weight matrices ([out, pad8(in)], zero padded).  With `cache`, a copy
    is only refreshed when the fp32 parameter changed (data_ptr / in-place version counter)."""
    wb = {}
    for l in range(1, 6):
        W = P[f"{pre}dense{l}.weight"]
        key = (W.data_ptr(), W._version)
        if cache is not None and cache.get(l, (None, None))[0] == key:
            wb[l] = cache[l][1]
            continue
        wb[l] = ops.pack_bf16(W, ld_out=pad8(W.shape[1]))
        if cache is not None:
            cache[l] = (key, wb[l])
    return wb


def decoder_forward_tc(h: torch.Tensor, P: Params, wb, pre: str = ""):
    """Tensor-core decoder: h [B, input_dim] fp32 -> rec [B, S] fp32.  Activations are bf16 [B, pad8(dim)];
    weights wb[l] bf16 [out, pad8(in)] (tcgen05 GEMM with fused bias + ELU, fp32 accumulation)."""
    B = h.shape[0]
    a = ops.convert(h, torch.bfloat16)
    acts = [a]
    for l in range(1, 6):
        W = P[f"{pre}dense{l}.weight"]
        Nout, K = W.shape
        if l < 5:
            a = ops.gemm_tc(a, wb[l], TC_BIAS_ELU, B, Nout, K, bias=P[f"{pre}dense{l}.bias"])
        else:
            a = ops.gemm_tc(a, wb[l], TC_PLAIN, B, Nout, K, bias=P[f"{pre}dense{l}.bias"], out_dtype=torch.float32)
        acts.append(a)
    return a, acts


def decoder_backward_tc(dout: torch.Tensor, acts, P: Params, wb, gradbuf: Optional[Grads] = None, pre: str = ""):
    """dout [B, S] fp32 -> (d h [B, input_dim] fp32, grads).  Data gradients: dz_{l-1} = (dz_l W_l) * ELU'(.) with the
    weight read MN-major straight from the forward's bf16 copy; weight gradients dW_l = dz_l^T a_{l-1} stored fp32."""
    G: Grads = {}
    B = dout.shape[0]
    dz = ops.convert(dout, torch.bfloat16)
    for l in range(5, 0, -1):
        W = P[f"{pre}dense{l}.weight"]
        Nout, K = W.shape
        wname, bname = f"{pre}dense{l}.weight", f"{pre}dense{l}.bias"
        dW = _out(gradbuf, wname)
        if dW is None:
            dW = torch.empty_like(W)
        ops.gemm_tc(dz, acts[l - 1], TC_WGRAD_STORE, Nout, K, B, a_mn=True, b_mn=True, out=dW)
        G[wname] = dW
        G[bname] = ops.colsum_ld(dz, Nout, _out(gradbuf, bname))
        if l > 1:
            dz = ops.gemm_tc(dz, wb[l], TC_DGRAD_ELUOUT, B, K, Nout, b_mn=True, yprev=acts[l - 1])
        else:
            dz = ops.gemm_tc(dz, wb[l], TC_PLAIN, B, K, Nout, b_mn=True, out_dtype=torch.float32)
    return dz, G
