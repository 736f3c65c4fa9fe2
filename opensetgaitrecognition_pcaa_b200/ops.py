"""Thin torch-tensor front-ends of the C-ABI kernels (device memory, streams: PyTorch; arithmetic: libpcaa_sm100).

Every function enqueues on torch's current CUDA stream and returns torch tensors it allocated with torch.empty
(so the caching allocator and CUDA graphs stay valid).  No function here computes with torch ops.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_ELU, ACT_NONE, BF16, F32, call

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype=None, contiguous=True) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("PCAA B200 ops need CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return t


# ------------------------------------------------------------------------------------------------ GEMM (CUDA cores)
def gemm(a: torch.Tensor, b: torch.Tensor, *, trans_a=False, trans_b=False, bias: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, out: Optional[torch.Tensor] = None, accumulate=False,
         out_dtype=torch.float32) -> torch.Tensor:
    """out = act(op(a) @ op(b) + bias); a, b 2-D (any strides), fp32 or bf16."""
    _chk(a, contiguous=False), _chk(b, contiguous=False)
    am = a.t() if trans_a else a
    bm = b.t() if trans_b else b
    M, K = am.shape
    K2, N = bm.shape
    if K != K2:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    call("pcaa_gemm_simt", _p(am), _DT[am.dtype], am.stride(0), am.stride(1), _p(bm), _DT[bm.dtype], bm.stride(0),
         bm.stride(1), _p(out), _DT[out.dtype], out.stride(0), out.stride(1), M, N, K, _p(bias), act,
         1 if accumulate else 0, _s())
    return out


# ------------------------------------------------------------------------------------------------ BatchNorm family
def colstats(y: torch.Tensor, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(y)
    R, Cc = y.shape
    if stats is None:
        stats = torch.zeros(2 * Cc, device=y.device, dtype=torch.float64)
    call("pcaa_colstats", _p(y), _DT[y.dtype], R, Cc, _p(stats), _s())
    return stats


def bn_finalize(stats, R, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5):
    Cc = gamma.numel()
    dev = gamma.device
    coef = torch.empty((4, Cc), device=dev, dtype=torch.float32)   # scale, shift, mean, invstd
    call("pcaa_bn_finalize", _p(stats), R, Cc, _p(gamma), _p(beta), _p(running_mean), _p(running_var), momentum, eps,
         _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(coef[3]), _s())
    return coef


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps=1e-5):
    Cc = gamma.numel()
    coef = torch.empty((2, Cc), device=gamma.device, dtype=torch.float32)
    call("pcaa_bn_eval_coeffs", _p(gamma), _p(beta), _p(running_mean), _p(running_var), eps, _p(coef[0]), _p(coef[1]),
         Cc, _s())
    return coef


def bn_elu_apply(y, scale, shift, out_dtype=None):
    _chk(y)
    R, Cc = y.shape
    out = torch.empty((R, Cc), device=y.device, dtype=out_dtype or y.dtype)
    call("pcaa_bn_elu_apply", _p(y), _DT[y.dtype], _p(scale), _p(shift), _p(out), _DT[out.dtype], R, Cc, _s())
    return out


def bn_elu_meanpool(y, scale, shift, n: int):
    _chk(y)
    R, Cc = y.shape
    G = R // n
    pooled = torch.empty((G, Cc), device=y.device, dtype=torch.float32)
    call("pcaa_bn_elu_meanpool", _p(y), _DT[y.dtype], _p(scale), _p(shift), _p(pooled), G, n, Cc, _s())
    return pooled


def elu_bwd_colstats(dout, y, coef, pooled_n: int = 0, dz_dtype=None):
    """dz = dout * ELU'(scale*y+shift) and stats2 = [sum dz, sum dz*xhat]; coef = bn_finalize output."""
    _chk(y), _chk(dout)
    R, Cc = y.shape
    dz = torch.empty((R, Cc), device=y.device, dtype=dz_dtype or y.dtype)
    stats2 = torch.zeros(2 * Cc, device=y.device, dtype=torch.float64)
    call("pcaa_elu_bwd_colstats", _p(dout), _DT[dout.dtype], pooled_n, _p(y), _DT[y.dtype], _p(coef[0]), _p(coef[1]),
         _p(coef[2]), _p(coef[3]), _p(dz), _DT[dz.dtype], _p(stats2), R, Cc, _s())
    return dz, stats2


def bn_bwd_finalize(stats2, R, coef, dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None):
    Cc = coef.shape[1]
    c = torch.empty((3, Cc), device=coef.device, dtype=torch.float32)
    if dgamma is None:
        dgamma = torch.empty(Cc, device=coef.device, dtype=torch.float32)
    if dbeta is None:
        dbeta = torch.empty(Cc, device=coef.device, dtype=torch.float32)
    call("pcaa_bn_bwd_finalize", _p(stats2), R, Cc, _p(coef[0]), _p(coef[2]), _p(coef[3]), _p(c[0]), _p(c[1]), _p(c[2]),
         _p(dgamma), _p(dbeta), _s())
    return c, dgamma, dbeta


def bn_bwd_apply(dz, y, c, out: Optional[torch.Tensor] = None, out_dtype=None):
    R, Cc = y.shape
    if out is None:
        out = torch.empty_like(dz) if out_dtype is None else torch.empty(dz.shape, device=dz.device, dtype=out_dtype)
    call("pcaa_bn_bwd_apply", _p(dz), _DT[dz.dtype], _p(y), _DT[y.dtype], _p(c[0]), _p(c[1]), _p(c[2]), _p(out),
         _DT[out.dtype], R, Cc, _s())
    return out


# ------------------------------------------------------------------------------------------------ helpers
def elu_bwd_from_out(dout, out):
    _chk(dout, torch.float32), _chk(out, torch.float32)
    dz = torch.empty_like(out)
    call("pcaa_elu_bwd_from_out", _p(dout), _p(out), _p(dz), out.numel(), _s())
    return dz


def colsum(x, out: Optional[torch.Tensor] = None):
    _chk(x, torch.float32)
    R, Cc = x.shape
    if out is None:
        out = torch.empty(Cc, device=x.device, dtype=torch.float32)
    call("pcaa_colsum", _p(x), R, Cc, _p(out), _s())
    return out


def convert(x, dtype):
    _chk(x)
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    call("pcaa_convert", _p(x), _DT[x.dtype], _p(out), _DT[dtype], x.numel(), _s())
    return out


def convert_into(x, out):
    """out[:] = x converted to out's dtype (same number of elements; in place, so views of `out` stay valid)."""
    _chk(x), _chk(out)
    if x.numel() != out.numel():
        raise ValueError("convert_into: element counts differ")
    call("pcaa_convert", _p(x), _DT[x.dtype], _p(out), _DT[out.dtype], x.numel(), _s())
    return out


def pack_bf16(w: torch.Tensor, ld_out: Optional[int] = None, transpose=False, out: Optional[torch.Tensor] = None):
    """bf16 copy of a 2-D fp32 matrix (optionally transposed) with the leading dimension padded to ld_out.  `w` may be a
    row-strided view (unit inner stride)."""
    _chk(w, torch.float32, contiguous=False)
    if w.dim() != 2 or w.stride(1) != 1:
        raise ValueError("pack_bf16: expected a 2-D matrix with unit inner stride")
    R, Cc = w.shape
    rows, cols = (Cc, R) if transpose else (R, Cc)
    if ld_out is None:
        ld_out = (cols + 7) // 8 * 8
    if out is None:
        out = torch.empty((rows, ld_out), device=w.device, dtype=torch.bfloat16)
    call("pcaa_pack_bf16", _p(w), R, Cc, w.stride(0), _p(out), ld_out, 1 if transpose else 0, _s())
    return out


def tcn_im2col(x, dil: int, dtype=torch.float32):
    _chk(x, torch.float32)
    B, T, Cin = x.shape
    col = torch.empty((B * T, Cin * 3), device=x.device, dtype=dtype)
    call("pcaa_tcn_im2col", _p(x), _p(col), _DT[dtype], B, T, Cin, dil, _s())
    return col


def tcn_col2im(dcol, B: int, T: int, Cin: int, dil: int):
    dx = torch.empty((B, T, Cin), device=dcol.device, dtype=torch.float32)
    call("pcaa_tcn_col2im", _p(dcol), _p(dx), B, T, Cin, dil, _s())
    return dx


def tcn_bn_elu_next(y, B: int, T: int, *, stats=None, gamma=None, beta=None, running_mean=None, running_var=None,
                    momentum=0.1, eps=1e-5, scale=None, shift=None, dil_next: int = 0, want_act: bool = False):
    """BatchNorm1d + ELU of a TCN layer output y [B*T, C] fp32 (statistics from the GEMM epilogue in training, scale / shift
    in eval).  Returns (col, act, coef): the next layer's bf16 im2col operand [B*T, C*3] (dil_next > 0), the fp32
    activation (want_act) and, in training, coef [4, C] = scale, shift, mean, invstd."""
    _chk(y, torch.float32)
    Cc = y.shape[1]
    col = torch.empty((B * T, Cc * 3), device=y.device, dtype=torch.bfloat16) if dil_next > 0 else None
    act = torch.empty((B * T, Cc), device=y.device, dtype=torch.float32) if want_act else None
    coef = torch.empty((4, Cc), device=y.device, dtype=torch.float32) if stats is not None else None
    call("pcaa_tcn_bn_elu_next", _p(y), _p(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var), momentum, eps,
         _p(scale), _p(shift), _p(coef), B, T, Cc, dil_next, _p(col), _p(act), _s())
    return col, act, coef


def tcn_elu_bwd_stats(src, src_mode: int, dil_up: int, y, coef, stats2, B: int, T: int):
    """dz [B*T, C] fp32 = d * ELU'(BN(y)) with d formed from src (mode 0: as is, 1: col2im of the layer above, 2: frame-mean
    broadcast); stats2 [2*C] double (pre-zeroed) += BatchNorm-backward sums."""
    _chk(y, torch.float32), _chk(src, torch.float32)
    Cc = y.shape[1]
    dz = torch.empty_like(y)
    call("pcaa_tcn_elu_bwd_stats", _p(src), src_mode, dil_up, _p(y), _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(coef[3]), _p(dz),
         _p(stats2), B, T, Cc, _s())
    return dz


def tcn_bn_bwd_apply(dz, y, stats2, coef, dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None):
    """dy bf16 [R, C] = BatchNorm1d backward of dz (coefficients derived from the completed stats2 inside the kernel)."""
    R, Cc = y.shape
    dy = torch.empty((R, Cc), device=y.device, dtype=torch.bfloat16)
    if dgamma is None:
        dgamma = torch.empty(Cc, device=y.device, dtype=torch.float32)
    if dbeta is None:
        dbeta = torch.empty(Cc, device=y.device, dtype=torch.float32)
    call("pcaa_tcn_bn_bwd_apply", _p(dz), _p(y), _p(stats2), _p(coef[0]), _p(coef[2]), _p(coef[3]), _p(dgamma), _p(dbeta), _p(dy),
         R, Cc, _s())
    return dy, dgamma, dbeta


def mean_rows(x):
    _chk(x, torch.float32)
    G, n, Cc = x.shape
    out = torch.empty((G, Cc), device=x.device, dtype=torch.float32)
    call("pcaa_mean_rows", _p(x), _p(out), G, n, Cc, _s())
    return out


def mean_rows_bwd(g, n: int):
    _chk(g, torch.float32)
    G, Cc = g.shape
    dx = torch.empty((G, n, Cc), device=g.device, dtype=torch.float32)
    call("pcaa_mean_rows_bwd", _p(g), _p(dx), G, n, Cc, _s())
    return dx


def softmax_ce(logits, gt, want_grad=True, gscale: float = 1.0):
    _chk(logits, torch.float32), _chk(gt, torch.int64)
    B, Cc = logits.shape
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    dl = torch.empty_like(logits) if want_grad else None
    pred = torch.empty(B, device=logits.device, dtype=torch.int32)
    call("pcaa_softmax_ce", _p(logits), _p(gt), _p(loss), _p(dl), gscale, _p(pred), B, Cc, _s())
    return loss, dl, pred


# ------------------------------------------------------------------------------------------------ PointNet layer 1
def pointnet_l1_fwd(x, w, bias, want_stats=True):
    _chk(x, torch.float32), _chk(w, torch.float32)
    B, F, T, N = x.shape
    Cout = w.shape[0]
    y = torch.empty((B * T * N, Cout), device=x.device, dtype=torch.bfloat16)
    stats = torch.zeros(2 * Cout, device=x.device, dtype=torch.float64) if want_stats else None
    call("pcaa_pointnet_l1_fwd", _p(x), _p(w), _p(bias), _p(y), _p(stats), B, T * N, Cout, _s())
    return y, stats


def pointnet_l1_wgrad(x, dy, out: Optional[torch.Tensor] = None):
    B, F, T, N = x.shape
    Cout = dy.shape[1]
    if out is None:
        out = torch.empty((Cout, 4), device=x.device, dtype=torch.float32)
    call("pcaa_pointnet_l1_wgrad", _p(x), _p(dy), _p(out), B, T * N, Cout, _s())
    return out


# ------------------------------------------------------------------------------------------------ channel-major PointNet
# T256 activation format: bf16 tensor [n_tiles, C, 256], element (c, p) at [p // 256, c, p % 256]; pad points are zeros.
def t256_tiles(P: int) -> int:
    return (P + 255) // 256


def t256_empty(C: int, P: int, device) -> torch.Tensor:
    return torch.empty((t256_tiles(P), C, 256), device=device, dtype=torch.bfloat16)


def t256_pack(x: torch.Tensor) -> torch.Tensor:
    """[C, P] -> T256 bf16 (torch glue for tests / tools; the hot path never converts)."""
    C, P = x.shape
    nt = t256_tiles(P)
    out = torch.zeros((C, nt * 256), device=x.device, dtype=torch.bfloat16)
    out[:, :P] = x.to(torch.bfloat16)
    return out.view(C, nt, 256).permute(1, 0, 2).contiguous()


def t256_unpack(xT: torch.Tensor, P: int) -> torch.Tensor:
    """T256 -> [C, P] (same dtype)."""
    nt, C, _ = xT.shape
    return xT.permute(1, 0, 2).reshape(C, nt * 256)[:, :P]


def pointnet_l1_fwd_t(x, w, bias, coef=None, want_stats=True):
    """x (B,4,T,N) fp32 -> yT T256 [tiles, Cout, 256] bf16 (+ row statistics); with coef = (scale, shift) the stored value
    is ELU(scale*y + shift) (eval mode) and no statistics are produced."""
    _chk(x, torch.float32), _chk(w, torch.float32)
    B, F, T, N = x.shape
    Cout = w.shape[0]
    yT = t256_empty(Cout, B * T * N, x.device)
    stats = torch.zeros(2 * Cout, device=x.device, dtype=torch.float64) if (want_stats and coef is None) else None
    sc, sh = (None, None) if coef is None else (coef[0], coef[1])
    call("pcaa_pointnet_l1_fwd_t", _p(x), _p(w), _p(bias), _p(sc), _p(sh), _p(yT), _p(stats), B, T * N, Cout, _s())
    return yT, stats


def input_moments(x):
    """double[14] = first and second moments (sums over all B*T*N points) of the 4 input features of x (B,4,T,N)."""
    _chk(x, torch.float32)
    B, F, T, N = x.shape
    if F != 4:
        raise ValueError("input_moments: x must have 4 feature planes")
    mom = torch.empty(14, device=x.device, dtype=torch.float64)
    call("pcaa_input_moments", _p(x), B, T * N, _p(mom), _s())
    return mom


def bn_from_input_moments(mom, R, w, bias, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5):
    """BatchNorm coefficients [4, C] (scale, shift, mean, invstd) of the K = 4 layer y = w x + bias from the input moments;
    updates the running statistics like bn_finalize."""
    Cc = gamma.numel()
    coef = torch.empty((4, Cc), device=gamma.device, dtype=torch.float32)
    call("pcaa_bn_from_input_moments", _p(mom), R, Cc, _p(w), _p(bias), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
         momentum, eps, _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(coef[3]), _s())
    return coef


def pointnet_l1_fwd_bn_t(x, w, bias, coef):
    """x (B,4,T,N) fp32 -> (yT, aT) T256 bf16: y = w x + bias and a = ELU(scale*y + shift) written in one pass."""
    _chk(x, torch.float32), _chk(w, torch.float32)
    B, F, T, N = x.shape
    Cout = w.shape[0]
    yT = t256_empty(Cout, B * T * N, x.device)
    aT = torch.empty_like(yT)
    call("pcaa_pointnet_l1_fwd_bn_t", _p(x), _p(w), _p(bias), _p(coef[0]), _p(coef[1]), _p(yT), _p(aT), B, T * N, Cout, _s())
    return yT, aT


def pointnet_l1_wgrad_t(x, dzT, yT=None, c=None, out: Optional[torch.Tensor] = None):
    """dW1 [Cout,4] = sum_p (c1*dzT + c2*yT + c3)(c,p) * x[f,p]  (yT / c None: dy = dzT)."""
    B, F, T, N = x.shape
    Cout = dzT.shape[1]
    if out is None:
        out = torch.empty((Cout, 4), device=x.device, dtype=torch.float32)
    c1, c2, c3 = (None, None, None) if c is None else (c[0], c[1], c[2])
    call("pcaa_pointnet_l1_wgrad_t", _p(x), _p(dzT), _p(yT), _p(c1), _p(c2), _p(c3), _p(out), B, T * N, Cout, _s())
    return out


def bn_elu_apply_t(yT, coef, P: int):
    out = torch.empty_like(yT)
    call("pcaa_bn_elu_apply_t", _p(yT), _p(coef[0]), _p(coef[1]), _p(out), P, yT.shape[1], _s())
    return out


def bn_bwd_apply_t(dzT, yT, c, P: int, out: Optional[torch.Tensor] = None):
    if out is None:
        out = torch.empty_like(dzT)
    call("pcaa_bn_bwd_apply_t", _p(dzT), _p(yT), _p(c[0]), _p(c[1]), _p(c[2]), _p(out), P, yT.shape[1], _s())
    return out


def bn_elu_meanpool_t(yT, coef, G: int, n: int, want_e=False):
    """pooled [G, C] fp32 = mean over each group of n points of ELU(scale*yT+shift) (coef None: plain mean); with
    want_e also the group sums e1 = sum ELU'(z), e2 = sum ELU'(z)*xhat used by the backward statistics."""
    Cc = yT.shape[1]
    pooled = torch.empty((G, Cc), device=yT.device, dtype=torch.float32)
    e1 = torch.empty_like(pooled) if want_e else None
    e2 = torch.empty_like(pooled) if want_e else None
    sc = sh = mu = inv = None
    if coef is not None:
        sc, sh = coef[0], coef[1]
        if want_e:
            mu, inv = coef[2], coef[3]
    call("pcaa_bn_elu_meanpool_t", _p(yT), _p(sc), _p(sh), _p(mu), _p(inv), _p(pooled), _p(e1), _p(e2), G, n, Cc, _s())
    return pooled, e1, e2


def pool_bwd_stats(dpool, e1, e2, n: int):
    _chk(dpool, torch.float32)
    G, Cc = dpool.shape
    stats2 = torch.zeros(2 * Cc, device=dpool.device, dtype=torch.float64)
    call("pcaa_pool_bwd_stats", _p(dpool), _p(e1), _p(e2), _p(stats2), G, n, Cc, _s())
    return stats2


def pool_bwd_apply_t(dpool, yT, coef, c, n: int):
    _chk(dpool, torch.float32)
    G, Cc = dpool.shape
    out = torch.empty_like(yT)
    call("pcaa_pool_bwd_apply_t", _p(dpool), _p(yT), _p(coef[0]), _p(coef[1]), _p(c[0]), _p(c[1]), _p(c[2]), _p(out), G,
         n, Cc, _s())
    return out


# ------------------------------------------------------------------------------------------------ tensor-core GEMMs
def gemm_tc_tn(a, w, mode: int, *, bias=None, stats=None, yprev=None, coef=None, out=None):
    """out[M,N] (bf16) = epilogue(a[M,K] @ w[N,K]^T); tcgen05 kernel (see include/pcaa.h for the modes)."""
    _chk(a, torch.bfloat16), _chk(w, torch.bfloat16)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    sc = sh = mu = inv = None
    if coef is not None:
        sc, sh, mu, inv = coef[0], coef[1], coef[2], coef[3]
    call("pcaa_gemm_tc_tn", _p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, mode, _p(bias),
         _p(stats), _p(yprev), _p(sc), _p(sh), _p(mu), _p(inv), _s())
    return out


def gemm_tc(a, b, mode: int, M: int, N: int, K: int, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16,
            bias=None, stats=None, yprev=None, coef=None):
    """General tcgen05 GEMM: out[M,N] = epilogue(sum_k A(m,k) B(n,k)).  `a` is stored [M,K] (or [K,M] when a_mn),
    `b` is stored [N,K] (or [K,N] when b_mn); both bf16 2-D with unit inner stride; M, N, K are the TRUE extents
    (buffers may be wider: leading dimensions come from the strides).  a_mn / b_mn may also be the pcaa_operand_layout
    codes OP_T256_K / OP_T256_MN for 3-D T256 activation tensors (see include/pcaa.h)."""
    _chk(a, torch.bfloat16, contiguous=False), _chk(b, torch.bfloat16, contiguous=False)
    if a.stride(-1) != 1 or b.stride(-1) != 1:
        raise ValueError("gemm_tc: operands need unit inner stride")
    if mode == _lib.TC_T_AFFINE_ELU_POOL:
        raise ValueError("gemm_tc: use gemm_tc_pooled for the pooled eval epilogue")
    if out is None and mode >= _lib.TC_T_BIAS_STATS:
        out = t256_empty(M, N, a.device)
    if out is None:
        if out_dtype == torch.bfloat16:
            out = torch.empty((M, (N + 7) // 8 * 8), device=a.device, dtype=torch.bfloat16)
        else:
            out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    sc = sh = mu = inv = None
    if coef is not None:
        sc, sh = coef[0], coef[1]
        if len(coef) >= 4:
            mu, inv = coef[2], coef[3]
    call("pcaa_gemm_tc", _p(a), a.stride(-2), int(a_mn), _p(b), b.stride(-2), int(b_mn), _p(out),
         out.stride(-2), _DT[out.dtype], M, N, K, mode, _p(bias), _p(stats), _p(yprev),
         0 if yprev is None else yprev.stride(-2), _p(sc), _p(sh), _p(mu), _p(inv), _s())
    return out


def gemm_tc_pooled(w, aT, M: int, P: int, K: int, n: int, *, bias=None, coef=None):
    """pooled[g, m] = mean over the n points of group g of ELU(scale[m] * (sum_k w[m,k] aT[k,p] + bias[m]) + shift[m]):
    the eval-mode shared-MLP layer with the mean pool over points fused into the GEMM epilogue (models.py:21-34 with
    running statistics + models.py:242-243, 282).  aT is T256 [tiles, K, 256]; n >= 32 and n | P.  Returns fp32 [P/n, M]."""
    _chk(w, torch.bfloat16, contiguous=False), _chk(aT, torch.bfloat16, contiguous=False)
    if n < 32 or P % n:
        raise ValueError("gemm_tc_pooled: groups of n >= 32 points with n | P")
    out = torch.empty((P // n, M), device=w.device, dtype=torch.float32)
    call("pcaa_gemm_tc", _p(w), w.stride(-2), int(_lib.OP_K), _p(aT), aT.stride(-2), int(_lib.OP_T256_MN), _p(out), n, F32, M, P, K,
         _lib.TC_T_AFFINE_ELU_POOL, _p(bias), None, None, 0, _p(coef[0]), _p(coef[1]), None, None, _s())
    return out


def colsum_ld(x, C: int, out: Optional[torch.Tensor] = None):
    """Column sums of the first C columns of a 2-D fp32 / bf16 matrix with arbitrary leading dimension."""
    R = x.shape[0]
    if out is None:
        out = torch.empty(C, device=x.device, dtype=torch.float32)
    call("pcaa_colsum_ld", _p(x), _DT[x.dtype], R, C, x.stride(0), _p(out), _s())
    return out


def gemm_tc_nt_wgrad(a, b, dW):
    """dW[N1,N2] (fp32) += a[K,N1]^T @ b[K,N2]  (bf16 operands, K = number of rows)."""
    _chk(a, torch.bfloat16), _chk(b, torch.bfloat16), _chk(dW, torch.float32)
    K, N1 = a.shape
    N2 = b.shape[1]
    call("pcaa_gemm_tc_nt_wgrad", _p(a), a.stride(0), _p(b), b.stride(0), _p(dW), dW.stride(0), N1, N2, K, _s())
    return dW


# ------------------------------------------------------------------------------------------------ Chamfer
def chamfer_fwd(preds, gts, want_idx=True):
    _chk(preds, torch.float32), _chk(gts, torch.float32)
    B, F, T, N = preds.shape
    fl = torch.empty((B, T), device=preds.device, dtype=torch.float32)
    i1 = torch.empty((B, T, N), device=preds.device, dtype=torch.int32) if want_idx else None
    i2 = torch.empty((B, T, N), device=preds.device, dtype=torch.int32) if want_idx else None
    call("pcaa_chamfer_fwd", _p(preds), _p(gts), B, F, T, N, _p(fl), _p(i1), _p(i2), _s())
    return fl, i1, i2


def chamfer_reduce(frame_loss, avg_out=True):
    B, T = frame_loss.shape
    out = torch.empty(() if avg_out else (B,), device=frame_loss.device, dtype=torch.float32)
    call("pcaa_chamfer_reduce", _p(frame_loss), B, T, 1 if avg_out else 0, _p(out), _s())
    return out


def chamfer_bwd(preds, gts, i1, i2, gout, avg_out=True):
    B, F, T, N = preds.shape
    g = torch.empty_like(preds)
    call("pcaa_chamfer_bwd", _p(preds), _p(gts), _p(i1), _p(i2), _p(gout), 1 if avg_out else 0, B, F, T, N, _p(g), _s())
    return g


def pairwise_dist(x, y):
    _chk(x, torch.float32), _chk(y, torch.float32)
    B, F, T, N = x.shape
    P = torch.empty((B, T, N, N), device=x.device, dtype=torch.float32)
    call("pcaa_pairwise_dist", _p(x), _p(y), B, F, T, N, _p(P), _s())
    return P


def ew(op: int, a: Optional[torch.Tensor], b: Optional[torch.Tensor] = None, shape=None):
    """Element-wise kernel family (see pcaa_ew); a, b fp32 contiguous."""
    ref = a if a is not None else b
    if shape is None:
        shape = ref.shape
    out = torch.empty(shape, device=ref.device, dtype=torch.float32)
    ncols = shape[-1] if len(shape) else 1
    call("pcaa_ew", op, _p(a), _p(b), _p(out), out.numel(), ncols, _s())
    return out


# ------------------------------------------------------------------------------------------------ critic
def wgangp_dstep(fv, z0, means, labels, alphas, W1, b1, W2, b2, W3, b3, gp_weight, grads):
    """grads = (gW1,gb1,gW2,gb2,gW3,gb3) pre-zeroed fp32 tensors (added to).  Returns losses[4] on device."""
    B = fv.shape[0]
    Cc = means.shape[0]
    losses = torch.empty(4, device=fv.device, dtype=torch.float32)
    call("pcaa_wgangp_dstep", _p(fv), _p(z0), _p(means), _p(labels), _p(alphas), _p(W1), _p(b1), _p(W2), _p(b2), _p(W3),
         _p(b3), float(gp_weight), _p(losses), *[_p(g) for g in grads], B, Cc, _s())
    return losses


def disc_fwd(x, labels, W1, b1, W2, b2, W3, b3, n_classes: int, want_out=True, want_dx=False, dx_scale=1.0,
             want_sum=False, out_scale=1.0):
    B = x.shape[0]
    out = torch.empty((B, 1), device=x.device, dtype=torch.float32) if want_out else None
    dx = torch.empty((B, 32), device=x.device, dtype=torch.float32) if want_dx else None
    osum = torch.empty((), device=x.device, dtype=torch.float32) if want_sum else None
    call("pcaa_disc_fwd", _p(x), _p(labels), _p(W1), _p(b1), _p(W2), _p(b2), _p(W3), _p(b3), _p(out), _p(dx),
         float(dx_scale), _p(osum), float(out_scale), B, n_classes, _s())
    if want_sum:
        return out, dx, osum
    return out, dx


# ------------------------------------------------------------------------------------------------ Adam / scoring
def adam_flat(p, g, m, v, lr, b1, b2, eps, step, grad_scale=1.0, shadow=None):
    call("pcaa_adam_flat", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, b1, b2, eps, step, grad_scale, _p(shadow), _s())


def sum_into(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst[i] += sum_k src[k, i]  (dst fp32 [n], src fp32 [nsrc, >= n] rows with a common stride)."""
    _chk(dst, torch.float32), _chk(src, torch.float32, contiguous=False)
    call("pcaa_sum_into", _p(dst), _p(src), dst.numel(), src.stride(0), src.shape[0], _s())


def sum_rows(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst[i] = sum_k src[k, i] in row order (overwrite; dst fp32 [n], src fp32 [nsrc, >= n] rows with a common stride)."""
    _chk(dst, torch.float32), _chk(src, torch.float32, contiguous=False)
    call("pcaa_sum_rows", _p(dst), _p(src), dst.numel(), src.stride(0), src.shape[0], _s())


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = src[idx[r]] along dim 0 (idx int64 on the device); rows must be multiples of 16 bytes."""
    _chk(src), _chk(idx, torch.int64)
    n = idx.numel()
    if out is None:
        out = torch.empty((n,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    row_bytes = src[0].numel() * src.element_size() if src.shape[0] else 0
    call("pcaa_gather_rows", _p(src), _p(idx), _p(out), n, row_bytes, src.shape[0], _s())
    return out


def adam_advance(step_dev, coef_dev, lr, b1, b2):
    """step_dev (int32[1]) += 1 and coef_dev (float32[2]) = Adam's bias-corrected step sizes, on the device."""
    _chk(step_dev, torch.int32), _chk(coef_dev, torch.float32)
    call("pcaa_adam_advance", _p(step_dev), _p(coef_dev), lr, b1, b2, _s())


def adam_flat_dev(p, g, m, v, b1, b2, eps, coef_dev, grad_scale=1.0, shadow=None):
    call("pcaa_adam_flat_dev", _p(p), _p(g), _p(m), _p(v), p.numel(), b1, b2, eps, _p(coef_dev), grad_scale, _p(shadow), _s())


def openset_score(emb, means):
    _chk(emb, torch.float32), _chk(means, torch.float32)
    M, D = emb.shape
    ll = torch.empty(M, device=emb.device, dtype=torch.float64)
    call("pcaa_openset_score", _p(emb), _p(means), M, means.shape[0], D, _p(ll), _s())
    return ll


def openset_vote(loglik, pred, k: int, log_thr: float, n_labels: int):
    nw = loglik.numel() // k
    out = torch.empty(nw, device=loglik.device, dtype=torch.int32)
    call("pcaa_openset_vote", _p(loglik), _p(pred), nw, k, float(log_thr), n_labels, _p(out), _s())
    return out


def sm_count() -> int:
    return _lib.load().pcaa_sm_count()
