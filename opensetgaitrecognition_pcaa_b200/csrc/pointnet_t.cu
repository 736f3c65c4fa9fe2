// Channel-major ("T256") PointNet kernels.  Activations are stored channels-first in 256-point tiles:
//     element (channel c, point p)  at  X[((p / 256) * C + c) * 256 + p % 256]          (bf16)
// i.e. [n_tiles][C][256] with the P = B*T*N points of the batch ordered (b, t, n) and the last tile ZERO padded.
// In this layout a BatchNorm channel is a row of every tile: its coefficients are per-row scalars, its statistics are
// row sums, and the tcgen05 GEMMs (gemm_tcgen05.cu, MODE_T_*) produce them in their epilogues without any cross-lane
// reduction; one GEMM tile reads / writes one contiguous [C, 256] block (DRAM-page friendly, unlike a [C, P] matrix
// whose rows are megabytes apart).
//
// Replaces, for the per-point shared MLP of the reference: Conv2d(4->512, 1x1) (models.py:21-28, 86-88),
// BatchNorm2d + ELU application (models.py:29, 33-34), AvgPool2d((1, nmax)) (models.py:242-243, 282) and their
// autograd backward.  All kernels are HBM-bound streaming passes with 16-byte accesses.
#include "common.cuh"

namespace pcaa {

__device__ __forceinline__ float elu_fast(float z) { return z > 0.f ? z : __expf(z) - 1.f; }

__device__ __forceinline__ int64_t t256(int64_t p, int c, int C) { return (((p >> 8) * C + c) << 8) + (p & 255); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return u;
}

// packed fp32x2 helpers (sm_100 FFMA2 / FADD2: two fp32 lanes per instruction)
__device__ __forceinline__ unsigned long long pk2f(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2f(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2f(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2f(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 8 consecutive points p0 .. p0+7 of feature plane f of x (B, 4, TN) fp32; out-of-range points read as 0.
// (b, tn) = (p0 / TN, p0 % TN): callers that walk p0 in fixed strides keep them incrementally (PointCursor) instead of
// paying a 64-bit division per step.
__device__ __forceinline__ void load_x8_at(const float* __restrict__ x, int64_t b, int64_t tn, int64_t p0, int64_t P,
                                           int64_t TN, int f, float (&v)[8]) {
    const int64_t off = (b * 4 + f) * TN + tn;
    if (p0 + 8 <= P && tn + 8 <= TN && (off & 3) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + off));
        const float4 c = __ldg(reinterpret_cast<const float4*>(x + off) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t p = p0 + j;
            if (p < P) {
                const int64_t bb = p / TN, t2 = p - bb * TN;
                v[j] = __ldg(x + (bb * 4 + f) * TN + t2);
            } else {
                v[j] = 0.f;
            }
        }
    }
}
__device__ __forceinline__ void load_x8(const float* __restrict__ x, int64_t p0, int64_t P, int64_t TN, int f,
                                        float (&v)[8]) {
    const int64_t b = p0 / TN;
    load_x8_at(x, b, p0 - b * TN, p0, P, TN, f, v);
}
// (sample, offset inside the sample) of a point index that advances in fixed steps
struct PointCursor {
    int64_t b, tn;
    __device__ __forceinline__ PointCursor(int64_t p0, int64_t TN) : b(p0 / TN), tn(p0 - (p0 / TN) * TN) {}
    __device__ __forceinline__ void advance(int64_t step, int64_t TN) {
        tn += step;
        while (tn >= TN) { tn -= TN; ++b; }
    }
};

// ------------------------------------------------------------------------------------------------ layer 1 forward
// grid (Cout / 64 channel groups [fast index], point super-tiles of 2048); 8 warps x 8 channels; a warp sweeps one
// 256-point tile at a time (8 points per lane = one 512-byte tile row per channel).  The channel group is the FAST block
// index: the blocks that read the same points of x are launched next to each other and share them in L2 -- with the point
// tile as the fast index x was re-read from DRAM by every channel group (ncu: 590 MB read for a 73 MB input, 0.7 % L2 hits,
// the output stream evicts it in between) and the kernel waited on those loads.
constexpr int L1_CH_PER_WARP = 8;
constexpr int L1_TILE_POINTS = 2048;

// per-channel constants of the block's 64 channels live in shared memory (broadcast reads) so that the register budget
// allows three resident CTAs per SM: the kernels are bound by the latency of their global accesses
struct L1Coef { float w[4]; float bias, scale, shift, scale_l2, shift_l2, pad0, pad1, pad2; };

__global__ void __launch_bounds__(256, 3)
pointnet_l1_fwd_t_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         __nv_bfloat16* __restrict__ yT, double* __restrict__ stats, int64_t P, int64_t TN, int Cout) {
    __shared__ L1Coef coef[8 * L1_CH_PER_WARP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 8 * L1_CH_PER_WARP) {
        const int c = min(blockIdx.x * 8 * L1_CH_PER_WARP + (int)threadIdx.x, Cout - 1);
        L1Coef k;
        const float4 t = __ldg(reinterpret_cast<const float4*>(w) + c);
        k.w[0] = t.x; k.w[1] = t.y; k.w[2] = t.z; k.w[3] = t.w;
        k.bias = bias ? __ldg(bias + c) : 0.f;
        k.scale = scale ? __ldg(scale + c) : 1.f;
        k.shift = shift ? __ldg(shift + c) : 0.f;
        k.scale_l2 = k.scale * LOG2E_F;
        k.shift_l2 = k.shift * LOG2E_F;
        k.pad0 = k.pad1 = k.pad2 = 0.f;
        coef[threadIdx.x] = k;
    }
    __syncthreads();
    const int c0 = (blockIdx.x * 8 + warp) * L1_CH_PER_WARP;
    if (c0 >= Cout) return;
    const L1Coef* ck = coef + warp * L1_CH_PER_WARP;
    const bool act = scale != nullptr;
    float s1[L1_CH_PER_WARP], s2[L1_CH_PER_WARP];
#pragma unroll
    for (int k = 0; k < L1_CH_PER_WARP; ++k) s1[k] = s2[k] = 0.f;
    const int64_t Ppad = (P + 255) & ~(int64_t)255;
    const int64_t tile0 = (int64_t)blockIdx.y * L1_TILE_POINTS;
    for (int ch = 0; ch < L1_TILE_POINTS / 256; ++ch) {
        const int64_t p0 = tile0 + ch * 256 + lane * 8;
        if (p0 >= Ppad) break;
        float xv[4][8];
#pragma unroll
        for (int f = 0; f < 4; ++f) load_x8(x, p0, P, TN, f, xv[f]);
        const int vcnt = p0 + 8 <= P ? 8 : (p0 < P ? (int)(P - p0) : 0);
#pragma unroll
        for (int k = 0; k < L1_CH_PER_WARP; ++k) {
            const float4 wk = *reinterpret_cast<const float4*>(ck[k].w);
            const float4 bk = *reinterpret_cast<const float4*>(&ck[k].bias);       // bias, scale, shift, scale_l2
            const float shl = ck[k].shift_l2;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = fmaf(wk.w, xv[3][j], bk.x);
                v = fmaf(wk.z, xv[2][j], v);
                v = fmaf(wk.y, xv[1][j], v);
                v = fmaf(wk.x, xv[0][j], v);
                v = j < vcnt ? v : 0.f;                                    // pad points are stored as zeros
                s1[k] += v;
                s2[k] = fmaf(v, v, s2[k]);
                y[j] = v;
            }
            if (act) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = j < vcnt ? elu_l2(fmaf(y[j], bk.y, bk.z), fmaf(y[j], bk.w, shl)) : 0.f;
            }
            if (c0 + k < Cout) *reinterpret_cast<uint4*>(yT + t256(p0, c0 + k, Cout)) = pack8(y);
        }
    }
    if (stats) {
#pragma unroll
        for (int k = 0; k < L1_CH_PER_WARP; ++k) {
            const float a = warp_sum(s1[k]), b = warp_sum(s2[k]);
            if (lane == 0 && c0 + k < Cout) {
                atomicAdd(&stats[c0 + k], (double)a);
                atomicAdd(&stats[Cout + c0 + k], (double)b);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ layer 1 forward, train mode
// y1 = W1 x + b1 AND a1 = ELU(scale*y1 + shift) in one pass (the BatchNorm coefficients are known up front from the input
// moments).  Two bf16 tensors are written per element computed, so this kernel is bound by instruction issue, not by HBM: all
// fp32 arithmetic runs on packed fp32x2 instructions (two points per FFMA2 / FADD2), ELU is branch-free
// (ELU(z) = max(z, min(2^(z log2 e), 1) - 1)), constants come duplicated from shared memory as ready 64-bit operands, full tiles
// take a path without per-point validity selects.  The FMA order equals pointnet_l1_fwd_t_kernel's: y1 is bit-identical.
struct L1Coef2 { float2 w[4]; float2 bias, scale, shift, scale_l2, shift_l2; };      // every constant as a (c, c) pair

__device__ __forceinline__ unsigned long long f2u(float2 v) { return pk2f(v.x, v.y); }

template <bool FULL, bool BOTH>
__device__ __forceinline__ void l1_bn_channel(const L1Coef2& c, const unsigned long long (&x)[4][4], int vcnt,
                                              __nv_bfloat16* __restrict__ yp, __nv_bfloat16* __restrict__ ap) {
    const unsigned long long w0 = f2u(c.w[0]), w1 = f2u(c.w[1]), w2 = f2u(c.w[2]), w3 = f2u(c.w[3]), b = f2u(c.bias);
    const unsigned long long sc = f2u(c.scale), sh = f2u(c.shift), scl = f2u(c.scale_l2), shl = f2u(c.shift_l2);
    const unsigned long long neg1 = pk2f(-1.f, -1.f);
    uint4 yo, ao;
    __nv_bfloat162* yh = reinterpret_cast<__nv_bfloat162*>(&yo);
    __nv_bfloat162* ah = reinterpret_cast<__nv_bfloat162*>(&ao);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned long long y2 = fma2f(w3, x[3][q], b);
        y2 = fma2f(w2, x[2][q], y2);
        y2 = fma2f(w1, x[1][q], y2);
        y2 = fma2f(w0, x[0][q], y2);
        float ylo, yhi;
        upk2f(y2, ylo, yhi);
        if (!FULL) {                                               // pad points are stored as zeros
            ylo = 2 * q < vcnt ? ylo : 0.f;
            yhi = 2 * q + 1 < vcnt ? yhi : 0.f;
            y2 = pk2f(ylo, yhi);
        }
        float zlo, zhi, llo, lhi;
        upk2f(fma2f(y2, sc, sh), zlo, zhi);
        upk2f(fma2f(y2, scl, shl), llo, lhi);
        float elo, ehi;
        upk2f(add2f(pk2f(fminf(ex2_fast(llo), 1.f), fminf(ex2_fast(lhi), 1.f)), neg1), elo, ehi);
        float alo = fmaxf(zlo, elo), ahi = fmaxf(zhi, ehi);
        if (!FULL) {
            alo = 2 * q < vcnt ? alo : 0.f;
            ahi = 2 * q + 1 < vcnt ? ahi : 0.f;
        }
        if (BOTH) yh[q] = __floats2bfloat162_rn(ylo, yhi);
        ah[q] = __floats2bfloat162_rn(alo, ahi);
    }
    if (BOTH) *reinterpret_cast<uint4*>(yp) = yo;
    *reinterpret_cast<uint4*>(ap) = ao;
}

// BOTH = true: train mode (y1 and a1); BOTH = false: eval mode (only a1 = ELU(scale*y1 + shift), running statistics)
template <bool BOTH>
__global__ void __launch_bounds__(256, 2)
pointnet_l1_fwd_bn_t_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                            const float* __restrict__ scale, const float* __restrict__ shift,
                            __nv_bfloat16* __restrict__ yT, __nv_bfloat16* __restrict__ aT, int64_t P, int64_t TN, int Cout) {
    __shared__ L1Coef2 coef[8 * L1_CH_PER_WARP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 8 * L1_CH_PER_WARP) {
        const int c = min(blockIdx.x * 8 * L1_CH_PER_WARP + (int)threadIdx.x, Cout - 1);
        L1Coef2 k;
        const float4 t = __ldg(reinterpret_cast<const float4*>(w) + c);
        k.w[0] = make_float2(t.x, t.x); k.w[1] = make_float2(t.y, t.y); k.w[2] = make_float2(t.z, t.z); k.w[3] = make_float2(t.w, t.w);
        const float b = bias ? __ldg(bias + c) : 0.f, sc = __ldg(scale + c), sh = __ldg(shift + c);
        k.bias = make_float2(b, b);
        k.scale = make_float2(sc, sc);
        k.shift = make_float2(sh, sh);
        k.scale_l2 = make_float2(sc * LOG2E_F, sc * LOG2E_F);
        k.shift_l2 = make_float2(sh * LOG2E_F, sh * LOG2E_F);
        coef[threadIdx.x] = k;
    }
    __syncthreads();
    // the block's 2048 points of x (4 planes, 32 KB) are staged in shared memory by ONE cooperative load: all 8 warps
    // (64 channels) use the same points, and the per-tile steps below then never wait on a global load (ncu: a third of
    // this kernel's stall samples sat on the first FMA after the x loads)
    __shared__ __align__(16) float xs[4][L1_TILE_POINTS];
    const int64_t tile0 = (int64_t)blockIdx.y * L1_TILE_POINTS;
    {
        const int64_t p0 = tile0 + (int64_t)threadIdx.x * 8;
        PointCursor cur(p0, TN);
        float xv[8];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            load_x8_at(x, cur.b, cur.tn, p0, P, TN, f, xv);          // points >= P read as 0
            *reinterpret_cast<float4*>(&xs[f][threadIdx.x * 8]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
            *reinterpret_cast<float4*>(&xs[f][threadIdx.x * 8 + 4]) = make_float4(xv[4], xv[5], xv[6], xv[7]);
        }
    }
    __syncthreads();
    const int c0 = (blockIdx.x * 8 + warp) * L1_CH_PER_WARP;
    if (c0 >= Cout) return;
    const L1Coef2* ck = coef + warp * L1_CH_PER_WARP;
    const int64_t Ppad = (P + 255) & ~(int64_t)255;
    for (int ch = 0; ch < L1_TILE_POINTS / 256; ++ch) {
        const int64_t p0 = tile0 + ch * 256 + lane * 8;
        if (p0 >= Ppad) break;
        unsigned long long xp[4][4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const float4 a = *reinterpret_cast<const float4*>(&xs[f][ch * 256 + lane * 8]);
            const float4 b = *reinterpret_cast<const float4*>(&xs[f][ch * 256 + lane * 8 + 4]);
            xp[f][0] = pk2f(a.x, a.y); xp[f][1] = pk2f(a.z, a.w); xp[f][2] = pk2f(b.x, b.y); xp[f][3] = pk2f(b.z, b.w);
        }
        const int vcnt = p0 + 8 <= P ? 8 : (p0 < P ? (int)(P - p0) : 0);
        if (vcnt == 8) {
#pragma unroll
            for (int k = 0; k < L1_CH_PER_WARP; ++k)
                if (c0 + k < Cout) l1_bn_channel<true, BOTH>(ck[k], xp, 8, yT + t256(p0, c0 + k, Cout), aT + t256(p0, c0 + k, Cout));
        } else {
#pragma unroll
            for (int k = 0; k < L1_CH_PER_WARP; ++k)
                if (c0 + k < Cout) l1_bn_channel<false, BOTH>(ck[k], xp, vcnt, yT + t256(p0, c0 + k, Cout), aT + t256(p0, c0 + k, Cout));
        }
    }
}

// ------------------------------------------------------------------------------------------------ layer 1 BatchNorm statistics
// y1 = W1 x + b1 is LINEAR in the 4 input features, so the batch statistics of BatchNorm 1 follow from the first and second
// moments of x alone:  mean_c = W1[c].mu + b_c,  var_c = W1[c]^T (E[x x^T] - mu mu^T) W1[c]  -- 14 numbers instead of a pass
// over the [512, P] tensor.  mom = { sum_p x_f (4), sum_p x_f x_g for f <= g (10) } in double.
__global__ void __launch_bounds__(256)
input_moments_kernel(const float* __restrict__ x, int64_t P, int64_t TN, double* __restrict__ mom) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) q[i] = 0.f;
    double ds[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) ds[i] = 0.0;
    int flush = 0;
    for (int64_t p0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8; p0 < P; p0 += (int64_t)gridDim.x * 256 * 8) {
        float xv[4][8];
#pragma unroll
        for (int f = 0; f < 4; ++f) load_x8(x, p0, P, TN, f, xv[f]);          // out-of-range points read as 0
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int i = 0;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                s[f] += xv[f][j];
#pragma unroll
                for (int g = f; g < 4; ++g) {
                    q[i] = fmaf(xv[f][j], xv[g][j], q[i]);
                    ++i;
                }
            }
        }
        if (++flush == 16) {       // fp32 partial sums of at most 128 points, then double
#pragma unroll
            for (int i = 0; i < 4; ++i) { ds[i] += (double)s[i]; s[i] = 0.f; }
#pragma unroll
            for (int i = 0; i < 10; ++i) { ds[4 + i] += (double)q[i]; q[i] = 0.f; }
            flush = 0;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) ds[i] += (double)s[i];
#pragma unroll
    for (int i = 0; i < 10; ++i) ds[4 + i] += (double)q[i];
    __shared__ double red[8][14];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        const double t = warp_sum_d(ds[i]);
        if (lane == 0) red[warp][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < 14) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(mom + threadIdx.x, t);
    }
}

// BatchNorm coefficients (scale, shift, mean, invstd as pcaa_bn_finalize gives them) + running-statistics update of a
// BatchNorm that follows a K = 4 linear layer, from the input moments
__global__ void bn_from_input_moments_kernel(const double* __restrict__ mom, double invR, double unbias, int C,
                                             const float* __restrict__ w, const float* __restrict__ bias,
                                             const float* gamma, const float* beta, float* rmean, float* rvar,
                                             float momentum, float eps, float* scale, float* shift, float* mean_o,
                                             float* invstd_o) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mu[4], cov[4][4];
    int i = 0;
    for (int f = 0; f < 4; ++f) mu[f] = mom[f] * invR;
    for (int f = 0; f < 4; ++f)
        for (int g = f; g < 4; ++g) {
            cov[f][g] = cov[g][f] = mom[4 + i] * invR - mu[f] * mu[g];
            ++i;
        }
    double wv[4];
    for (int f = 0; f < 4; ++f) wv[f] = (double)w[c * 4 + f];
    double mean = bias ? (double)bias[c] : 0.0, var = 0.0;
    for (int f = 0; f < 4; ++f) {
        mean += wv[f] * mu[f];
        for (int g = 0; g < 4; ++g) var += wv[f] * cov[f][g] * wv[g];
    }
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    if (mean_o) mean_o[c] = (float)mean;
    if (invstd_o) invstd_o[c] = invstd;
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * unbias);
}

// ------------------------------------------------------------------------------------------------ layer 1 weight gradient
// dW1[c][f] = sum_p dy[c][p] * x[f][p] with dy = c1[c]*dz + c2[c]*y + c3[c] formed on the fly (BatchNorm backward of
// layer 1 fused in: its dy is never written) or dy = dz when y is null.
constexpr int L1W_CH_PER_WARP = 4;

// two bf16 of one 32-bit word as a packed fp32x2 operand (lo = element 0, hi = element 1): two integer instructions
__device__ __forceinline__ unsigned long long bf2_to_f2(uint32_t w) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(w << 16), "r"(w & 0xffff0000u));
    return r;
}

// All arithmetic on packed fp32x2 instructions (FFMA2): per pair of points 2 FFMA2 form dy and 4 accumulate it against the
// 4 input features -- the scalar version needed ~9 issue slots per element and was issue-bound at 4.0 TB/s.  Points beyond P
// need no masking: their x reads as 0 and dz / y are the finite zero padding of the last tile.
template <bool HAS_Y>
__global__ void __launch_bounds__(256, 2)
pointnet_l1_wgrad_t_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dzT,
                           const __nv_bfloat16* __restrict__ yT, const float* __restrict__ c1,
                           const float* __restrict__ c2, const float* __restrict__ c3, float* __restrict__ dW,
                           int64_t P, int64_t TN, int Cout) {
    __shared__ float4 coef[8 * L1W_CH_PER_WARP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 8 * L1W_CH_PER_WARP) {
        const int c = min(blockIdx.x * 8 * L1W_CH_PER_WARP + (int)threadIdx.x, Cout - 1);
        coef[threadIdx.x] = HAS_Y ? make_float4(__ldg(c1 + c), __ldg(c2 + c), __ldg(c3 + c), 0.f) : make_float4(1.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int c0 = (blockIdx.x * 8 + warp) * L1W_CH_PER_WARP;
    if (c0 >= Cout) return;
    unsigned long long acc[L1W_CH_PER_WARP][4];
#pragma unroll
    for (int k = 0; k < L1W_CH_PER_WARP; ++k)
#pragma unroll
        for (int f = 0; f < 4; ++f) acc[k][f] = 0ull;
    const int64_t tile0 = (int64_t)blockIdx.y * L1_TILE_POINTS;
    PointCursor cur(tile0 + lane * 8, TN);
    for (int ch = 0; ch < L1_TILE_POINTS / 256; ++ch, cur.advance(256, TN)) {
        const int64_t p0 = tile0 + ch * 256 + lane * 8;
        if (p0 >= P) break;
        // every global load of this step is issued before the first use
        uint4 rz[L1W_CH_PER_WARP], ry[L1W_CH_PER_WARP];
#pragma unroll
        for (int k = 0; k < L1W_CH_PER_WARP; ++k) {
            const int64_t off = t256(p0, min(c0 + k, Cout - 1), Cout);
            rz[k] = __ldg(reinterpret_cast<const uint4*>(dzT + off));
            if (HAS_Y) ry[k] = __ldg(reinterpret_cast<const uint4*>(yT + off));
        }
        float xv[4][8];
#pragma unroll
        for (int f = 0; f < 4; ++f) load_x8_at(x, cur.b, cur.tn, p0, P, TN, f, xv[f]);     // out-of-range points read as 0
        unsigned long long xp[4][4];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int q = 0; q < 4; ++q) xp[f][q] = pk2f(xv[f][2 * q], xv[f][2 * q + 1]);
#pragma unroll
        for (int k = 0; k < L1W_CH_PER_WARP; ++k) {
            const float4 a = coef[warp * L1W_CH_PER_WARP + k];
            const unsigned long long a1 = pk2f(a.x, a.x), a2 = pk2f(a.y, a.y), a3 = pk2f(a.z, a.z);
            const uint32_t* wz = reinterpret_cast<const uint32_t*>(&rz[k]);
            const uint32_t* wy = reinterpret_cast<const uint32_t*>(&ry[k]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned long long d = bf2_to_f2(wz[q]);
                if (HAS_Y) d = fma2f(a1, d, fma2f(a2, bf2_to_f2(wy[q]), a3));
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[k][f] = fma2f(d, xp[f][q], acc[k][f]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < L1W_CH_PER_WARP; ++k)
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float lo, hi;
            upk2f(acc[k][f], lo, hi);
            const float t = warp_sum(lo + hi);
            if (lane == 0 && c0 + k < Cout) atomicAdd(dW + (c0 + k) * 4 + f, t);
        }
}

// ------------------------------------------------------------------------------------------------ BN + ELU apply
// The tensor is a list of n_tiles * C rows of 256 points (32 chunks of 8); row r belongs to channel r % C.  Every thread
// streams EW_UNROLL chunks (all loads issued before the first use).
constexpr int EW_UNROLL = 4;

struct ChunkPos {
    int64_t off;      // element offset of the chunk
    int64_t p0;       // first point of the chunk
    int c;            // channel
    bool ok;
};
__device__ __forceinline__ ChunkPos chunk_pos(int u, int64_t nchunks, int C) {
    ChunkPos r;
    const int64_t i = ((int64_t)blockIdx.x * EW_UNROLL + u) * 256 + threadIdx.x;
    r.ok = i < nchunks;
    const int64_t row = i >> 5;
    r.c = (int)(row % C);
    r.p0 = ((row / C) << 8) + ((i & 31) << 3);
    r.off = i << 3;
    return r;
}

// dy = c1[c]*dz + c2[c]*y + c3[c]  (BatchNorm backward), may run in place on dz
__global__ void __launch_bounds__(256)
bn_bwd_apply_t_kernel(const __nv_bfloat16* __restrict__ dzT, const __nv_bfloat16* __restrict__ yT,
                      const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
                      __nv_bfloat16* __restrict__ dyT, int64_t nchunks, int64_t P, int C) {
    uint4 rz[EW_UNROLL], ry[EW_UNROLL];
    ChunkPos cp[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        cp[u] = chunk_pos(u, nchunks, C);
        if (cp[u].ok) {
            rz[u] = __ldg(reinterpret_cast<const uint4*>(dzT + cp[u].off));
            ry[u] = __ldg(reinterpret_cast<const uint4*>(yT + cp[u].off));
        }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        if (!cp[u].ok) continue;
        const float a1 = __ldg(c1 + cp[u].c), a2 = __ldg(c2 + cp[u].c), a3 = __ldg(c3 + cp[u].c);
        float z[8], y[8];
        unpack8(rz[u], z);
        unpack8(ry[u], y);
        const bool full = cp[u].p0 + 8 <= P;
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = (full || cp[u].p0 + j < P) ? fmaf(a1, z[j], fmaf(a2, y[j], a3)) : 0.f;
        *reinterpret_cast<uint4*>(dyT + cp[u].off) = pack8(z);
    }
}

// ------------------------------------------------------------------------------------------------ mean pool over points
// pooled[g][c] = mean_{i<n} ELU(scale[c]*y(c, g*n+i) + shift[c]); when e1/e2 are given (training) also
// e1[g][c] = sum_i ELU'(z), e2[g][c] = sum_i ELU'(z)*xhat -- the group sums from which the backward's BatchNorm
// statistics follow without another pass over the activations (d pooled / d z is constant over a group).
// grid (ceil(G / 8), C / 32): warp w owns group g0 + w and walks the 32 channel rows of the block.
__global__ void __launch_bounds__(256)
bn_elu_meanpool_t_kernel(const __nv_bfloat16* __restrict__ yT, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, float* __restrict__ pooled, float* __restrict__ e1,
                         float* __restrict__ e2, int64_t G, int n, int C, int apply) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * 8 + warp;
    if (g >= G) return;
    const int cb = blockIdx.y * 32;
    const bool pairs = (n & 1) == 0;           // group starts are even -> a bf16 pair never straddles a tile boundary
    const int64_t pbeg = g * n;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        const int c = cb + i;
        if (c >= C) break;
        const float sc = apply ? __ldg(scale + c) : 1.f, sh = apply ? __ldg(shift + c) : 0.f;
        const float mu = e1 ? __ldg(mean + c) : 0.f, is = e1 ? __ldg(invstd + c) : 0.f;
        float s = 0.f, t1 = 0.f, t2 = 0.f;
        if (pairs) {
            for (int k = lane; k < n / 2; k += 32) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(yT + t256(pbeg + 2 * k, c, C)));
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float y = u ? f.y : f.x;
                    const float z = fmaf(y, sc, sh);
                    const float a = apply ? elu_fast(z) : z;
                    s += a;
                    if (e1) {
                        const float d = z > 0.f ? 1.f : a + 1.f;
                        t1 += d;
                        t2 = fmaf(d, (y - mu) * is, t2);
                    }
                }
            }
        } else {
            for (int k = lane; k < n; k += 32) {
                const float y = __bfloat162float(yT[t256(pbeg + k, c, C)]);
                const float z = fmaf(y, sc, sh);
                const float a = apply ? elu_fast(z) : z;
                s += a;
                if (e1) {
                    const float d = z > 0.f ? 1.f : a + 1.f;
                    t1 += d;
                    t2 = fmaf(d, (y - mu) * is, t2);
                }
            }
        }
        s = warp_sum(s);
        if (e1) { t1 = warp_sum(t1); t2 = warp_sum(t2); }
        if (lane == i) { r0 = s; r1 = t1; r2 = t2; }
    }
    if (cb + lane < C) {
        const int64_t o = g * C + cb + lane;
        pooled[o] = r0 / (float)n;
        if (e1) { e1[o] = r1; e2[o] = r2; }
    }
}

// stats2[c] += sum_g dpool[g][c]/n * e1[g][c]; stats2[C+c] += sum_g dpool[g][c]/n * e2[g][c]
// grid (C / 32, row splits), block (32, 8)
__global__ void __launch_bounds__(256)
pool_bwd_stats_kernel(const float* __restrict__ dpool, const float* __restrict__ e1, const float* __restrict__ e2,
                      double* __restrict__ stats2, int64_t G, int C, float inv_n) {
    __shared__ float sm[2][8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float t1 = 0.f, t2 = 0.f;
    if (c < C) {
        for (int64_t g = (int64_t)blockIdx.y * 8 + threadIdx.y; g < G; g += (int64_t)gridDim.y * 8) {
            const float d = __ldg(dpool + g * C + c) * inv_n;
            t1 = fmaf(d, __ldg(e1 + g * C + c), t1);
            t2 = fmaf(d, __ldg(e2 + g * C + c), t2);
        }
    }
    sm[0][threadIdx.y][threadIdx.x] = t1;
    sm[1][threadIdx.y][threadIdx.x] = t2;
    __syncthreads();
    if (threadIdx.y < 2 && c < C) {
        double a = 0.0;
        for (int l = 0; l < 8; ++l) a += (double)sm[threadIdx.y][l][threadIdx.x];
        atomicAdd(&stats2[(int64_t)threadIdx.y * C + c], a);
    }
}

// dy(c,p) = c1[c] * (dpool[g(p)][c]/n) * ELU'(scale*y+shift) + c2[c]*y + c3[c]   (mean-pool backward, ELU backward and
// BatchNorm backward of layer 4 in ONE pass over y4)
__global__ void __launch_bounds__(256)
pool_bwd_apply_t_kernel(const float* __restrict__ dpool, const __nv_bfloat16* __restrict__ yT,
                        const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ c1,
                        const float* __restrict__ c2, const float* __restrict__ c3, __nv_bfloat16* __restrict__ dyT,
                        int64_t nchunks, int64_t P, int n, int C, float inv_n) {
    uint4 ry[EW_UNROLL];
    ChunkPos cp[EW_UNROLL];
    float g0v[EW_UNROLL], g1v[EW_UNROLL];
    int split[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        cp[u] = chunk_pos(u, nchunks, C);
        if (cp[u].ok) {
            ry[u] = __ldg(reinterpret_cast<const uint4*>(yT + cp[u].off));
            const int64_t p0 = cp[u].p0;
            const int64_t g0 = p0 / n;
            const int64_t nxt = (g0 + 1) * n - p0;             // first index (0..8+) that belongs to the next group
            split[u] = nxt < 8 ? (int)nxt : 8;
            g0v[u] = p0 < P ? __ldg(dpool + g0 * C + cp[u].c) : 0.f;
            g1v[u] = (nxt < 8 && (g0 + 1) * n < P) ? __ldg(dpool + (g0 + 1) * C + cp[u].c) : 0.f;
        }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        if (!cp[u].ok) continue;
        const int c = cp[u].c;
        const int64_t p0 = cp[u].p0;
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float a1 = __ldg(c1 + c) * inv_n, a2 = __ldg(c2 + c), a3 = __ldg(c3 + c);
        float y[8];
        unpack8(ry[u], y);
        const bool full = p0 + 8 <= P;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z = fmaf(y[j], sc, sh);
            float gv = j < split[u] ? g0v[u] : g1v[u];
            if (n < 8) gv = (p0 + j < P) ? __ldg(dpool + ((p0 + j) / n) * C + c) : 0.f;   // tiny clouds: > 2 groups per chunk
            const float d = gv * (z > 0.f ? 1.f : __expf(z));
            y[j] = (full || p0 + j < P) ? fmaf(a1, d, fmaf(a2, y[j], a3)) : 0.f;
        }
        *reinterpret_cast<uint4*>(dyT + cp[u].off) = pack8(y);
    }
}

// ------------------------------------------------------------------------------------------------ row-streaming variants
// (groups of n >= 8 points).  A warp streams whole 512-byte tile rows (one channel x 256 points, 8 points per lane): the
// position of the group boundaries inside a tile depends on the lane only, so the per-lane group bookkeeping is set
// up once and reused for every channel row of the block.
struct LaneGroups {
    int64_t g0;        // group of the lane's first point
    int split;         // elements j < split belong to g0, the others to g0 + 1  (n >= 8: at most one boundary per chunk)
    int vcnt;          // elements j < vcnt are real points (the last tile is zero padded)
};
__device__ __forceinline__ LaneGroups lane_groups(int64_t tile, int lane, int n, int64_t P) {
    LaneGroups r;
    const int64_t p0 = (tile << 8) + (lane << 3);
    r.g0 = p0 / n;
    const int64_t nxt = (r.g0 + 1) * n - p0;
    r.split = nxt < 8 ? (int)nxt : 8;
    const int64_t left = P - p0;
    r.vcnt = left >= 8 ? 8 : (left > 0 ? (int)left : 0);
    return r;
}

constexpr int MP_ROWS_PER_WARP = 8;

// grid (n_tiles, ceil(C / 64)), 8 warps; warp w owns channels cb + w + 8*i.  pooled / e1 / e2 must be zeroed: a group
// receives one contribution per tile it touches (+ one when a lane-31 chunk straddles), added atomically.
template <bool TRAIN>
__global__ void __launch_bounds__(256)
bn_elu_meanpool_rows_kernel(const __nv_bfloat16* __restrict__ yT, const float* __restrict__ scale,
                            const float* __restrict__ shift, const float* __restrict__ mean,
                            const float* __restrict__ invstd, float* __restrict__ pooled, float* __restrict__ e1,
                            float* __restrict__ e2, int64_t P, int64_t G, int n, int C, int apply, float inv_n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int cb = blockIdx.y * (8 * MP_ROWS_PER_WARP) + warp;
    const LaneGroups lg = lane_groups(tile, lane, n, P);
    float wT[8], wA[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        wT[j] = j < lg.vcnt ? 1.f : 0.f;
        wA[j] = (j < lg.vcnt && j < lg.split) ? 1.f : 0.f;
    }
    // segmented reduction over the lanes of one group (contiguous run of lanes with the same g0)
    bool same[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int64_t other = __shfl_down_sync(0xffffffffu, lg.g0, 1 << s);
        same[s] = (lane + (1 << s) < 32) && other == lg.g0;
    }
    const int64_t prev = __shfl_up_sync(0xffffffffu, lg.g0, 1);
    const bool head = lane == 0 || prev != lg.g0;
    const bool straddle = lg.split < 8;
    const bool prev_straddle = __shfl_up_sync(0xffffffffu, straddle ? 1 : 0, 1) != 0 && lane > 0;
    const int64_t tile_off = tile * (int64_t)C * 256 + lane * 8;
    const bool full_tile = ((tile + 1) << 8) <= P;                  // block-uniform: no pad points in this tile

    uint4 raw[MP_ROWS_PER_WARP];
#pragma unroll
    for (int i = 0; i < MP_ROWS_PER_WARP; ++i) {
        const int c = cb + 8 * i;
        if (c < C) raw[i] = __ldg(reinterpret_cast<const uint4*>(yT + tile_off + (int64_t)c * 256));
    }
#pragma unroll
    for (int i = 0; i < MP_ROWS_PER_WARP; ++i) {
        const int c = cb + 8 * i;
        if (c >= C) break;
        const float sc = apply ? __ldg(scale + c) : 1.f, sh = apply ? __ldg(shift + c) : 0.f;
        float y[8];
        unpack8(raw[i], y);
        const float scl = sc * LOG2E_F, shl = sh * LOG2E_F;
        float Ts = 0.f, As = 0.f, Td = 0.f, Ad = 0.f, Tu = 0.f, Au = 0.f;
        if (full_tile) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float z = fmaf(y[j], sc, sh);
                float a = z, d = 1.f;
                if (apply) {
                    const float e = ex2_fast(fmaf(y[j], scl, shl));
                    a = z > 0.f ? z : e - 1.f;
                    d = z > 0.f ? 1.f : e;
                }
                Ts += a;
                As = fmaf(wA[j], a, As);
                if (TRAIN) {
                    const float wd = wA[j] * d;
                    Td += d;
                    Tu = fmaf(d, y[j], Tu);
                    Ad += wd;
                    Au = fmaf(wd, y[j], Au);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float z = fmaf(y[j], sc, sh);
                float a = z, d = 1.f;
                if (apply) {
                    const float e = ex2_fast(fmaf(y[j], scl, shl));
                    a = z > 0.f ? z : e - 1.f;
                    d = z > 0.f ? 1.f : e;
                }
                Ts = fmaf(wT[j], a, Ts);
                As = fmaf(wA[j], a, As);
                if (TRAIN) {
                    const float dy = d * y[j];
                    Td = fmaf(wT[j], d, Td);
                    Ad = fmaf(wA[j], d, Ad);
                    Tu = fmaf(wT[j], dy, Tu);
                    Au = fmaf(wA[j], dy, Au);
                }
            }
        }
        // the part of a straddling chunk that belongs to the next group joins the next lane's run
        float Bs = Ts - As, Bd = Td - Ad, Bu = Tu - Au;
        const float ps = __shfl_up_sync(0xffffffffu, Bs, 1);
        As += prev_straddle ? ps : 0.f;
        if (TRAIN) {
            const float pd = __shfl_up_sync(0xffffffffu, Bd, 1), pu = __shfl_up_sync(0xffffffffu, Bu, 1);
            Ad += prev_straddle ? pd : 0.f;
            Au += prev_straddle ? pu : 0.f;
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const float o = __shfl_down_sync(0xffffffffu, As, 1 << s);
            As += same[s] ? o : 0.f;
            if (TRAIN) {
                const float od = __shfl_down_sync(0xffffffffu, Ad, 1 << s), ou = __shfl_down_sync(0xffffffffu, Au, 1 << s);
                Ad += same[s] ? od : 0.f;
                Au += same[s] ? ou : 0.f;
            }
        }
        float mu = 0.f, is = 0.f;
        if (TRAIN) { mu = __ldg(mean + c); is = __ldg(invstd + c); }
        if (head && lg.g0 < G) {
            const int64_t o = lg.g0 * C + c;
            atomicAdd(pooled + o, As * inv_n);
            if (TRAIN) {
                atomicAdd(e1 + o, Ad);
                atomicAdd(e2 + o, is * (Au - mu * Ad));     // sum d*xhat = invstd * (sum d*y - mean * sum d)
            }
        }
        if (lane == 31 && straddle && lg.g0 + 1 < G) {
            const int64_t o = (lg.g0 + 1) * C + c;
            atomicAdd(pooled + o, Bs * inv_n);
            if (TRAIN) {
                atomicAdd(e1 + o, Bd);
                atomicAdd(e2 + o, is * (Bu - mu * Bd));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ staged mean pool
// Same contract as bn_elu_meanpool_rows_kernel, organised for the instruction-issue limit instead of only for the
// memory system: the CTA stages a [64 channels][256 points] block of the tile in shared memory (coalesced 16-byte
// loads, 528-byte row pitch), then LANE = CHANNEL and a warp walks 64 consecutive points of its 32 channels, so
//   * the reduction over points is a serial in-thread sum (no shuffles, no per-element group weights);
//   * group boundaries depend on the point index only => every branch on them is warp-uniform;
//   * ELU and ELU' share one exponential and need no select: with e = 2^(z*log2 e),
//       ELU'(z) = min(e, 1),  ELU(z) = max(z, 0) + min(e, 1) - 1,
//     i.e. 9 instructions per element in training (unpack, 2 FMA, EX2, MIN, MAX, 3 accumulates), 2 in eval.
// Partial sums of the (at most 255/n + 2) groups a tile touches are combined in shared memory; one global atomic per
// (group, channel, tile) as before.  grid (n_tiles, ceil(C / 64)), 256 threads, n >= 8.
constexpr int MPS_CH = 64;
constexpr int MPS_PITCH = 264;          // bf16 elements per staged row (528 B: 16-byte accesses of 8 lanes hit 32 banks)

template <bool TRAIN, bool APPLY>
__global__ void __launch_bounds__(256)
bn_elu_meanpool_staged_kernel(const __nv_bfloat16* __restrict__ yT, const float* __restrict__ scale,
                              const float* __restrict__ shift, const float* __restrict__ mean,
                              const float* __restrict__ invstd, float* __restrict__ pooled, float* __restrict__ e1,
                              float* __restrict__ e2, int64_t P, int64_t G, int n, int C, int kmax, float inv_n) {
    extern __shared__ __align__(16) unsigned char mps_smem[];
    __nv_bfloat16* rows = reinterpret_cast<__nv_bfloat16*>(mps_smem);
    constexpr int NV = TRAIN ? 3 : 1;
    float* acc = reinterpret_cast<float*>(mps_smem + (size_t)MPS_CH * MPS_PITCH * sizeof(__nv_bfloat16));   // [kmax][NV][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int cb = blockIdx.y * MPS_CH;
    const int64_t tile_off = tile * (int64_t)C * 256;
    // ---- stage: 64 rows x 32 chunks of 16 bytes, 8 per thread, copied global -> shared asynchronously (cp.async: no
    // staging registers, so more CTAs are resident per SM and one CTA's loads overlap another's arithmetic)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int id = threadIdx.x + 256 * i;
        const int r = id >> 5, ck = id & 31;
        const int c = min(cb + r, C - 1);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rows + r * MPS_PITCH + ck * 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(yT + tile_off + (int64_t)c * 256 + ck * 8) : "memory");
    }
    for (int i = threadIdx.x; i < kmax * NV * MPS_CH; i += 256) acc[i] = 0.f;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- reduce: warp = (channel half, point quarter), lane = channel
    const int chl = (warp & 1) * 32 + lane;                     // channel within the block
    const int c = min(cb + chl, C - 1);
    const int q = warp >> 1;
    float sc = 1.f, sh = 0.f;
    if (APPLY) { sc = __ldg(scale + c); sh = __ldg(shift + c); }
    const float scl = sc * LOG2E_F, shl = sh * LOG2E_F;
    const int64_t pt0 = tile << 8;
    const int64_t g_first = pt0 / n;
    const int64_t p0 = pt0 + 64 * q;
    int64_t g = p0 / n;
    int left = (int)((g + 1) * n - p0);                         // points of group g still ahead (warp-uniform)
    // packed fp32x2 arithmetic (FFMA2 / FADD2): even points accumulate in the low halves, odd points in the high halves
    unsigned long long Sa2 = 0ull, Sd2 = 0ull, Su2 = 0ull;
    const unsigned long long sc2 = pk2f(sc, sc), sh2 = pk2f(sh, sh), scl2 = pk2f(scl, scl), shl2 = pk2f(shl, shl);
    const __nv_bfloat16* row = rows + chl * MPS_PITCH + 64 * q;

    auto flush = [&](int64_t grp) {
        if (grp < G) {
            float lo, hi;
            float* a = acc + (size_t)(grp - g_first) * NV * MPS_CH + chl;
            upk2f(Sa2, lo, hi);
            atomicAdd(a, lo + hi);
            if (TRAIN) {
                upk2f(Sd2, lo, hi);
                atomicAdd(a + MPS_CH, lo + hi);
                upk2f(Su2, lo, hi);
                atomicAdd(a + 2 * MPS_CH, lo + hi);
            }
        }
        Sa2 = 0ull; Sd2 = 0ull; Su2 = 0ull;
    };
    // two points at once: w holds bf16 (point 2i | point 2i+1 << 16); mlo / mhi = 1.f or 0.f select the halves
    auto accum2 = [&](uint32_t w, bool use_lo, bool use_hi) {
        const float ylo = __uint_as_float(w << 16), yhi = __uint_as_float(w & 0xffff0000u);
        const unsigned long long y2 = pk2f(ylo, yhi);
        if (APPLY) {
            float zl, zh, al, ah;
            upk2f(fma2f(y2, sc2, sh2), zl, zh);
            upk2f(fma2f(y2, scl2, shl2), al, ah);
            float dl = fminf(ex2_fast(al), 1.f), dh = fminf(ex2_fast(ah), 1.f);
            float rl = fmaxf(zl, 0.f), rh = fmaxf(zh, 0.f);
            if (!use_lo) { dl = 0.f; rl = 0.f; }
            if (!use_hi) { dh = 0.f; rh = 0.f; }
            const unsigned long long d2 = pk2f(dl, dh);
            Sa2 = add2f(Sa2, pk2f(rl, rh));
            if (TRAIN) {
                Sd2 = add2f(Sd2, d2);
                Su2 = fma2f(d2, y2, Su2);
            } else {
                Sa2 = add2f(Sa2, d2);
            }
        } else {
            Sa2 = add2f(Sa2, pk2f(use_lo ? ylo : 0.f, use_hi ? yhi : 0.f));
        }
    };
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * k);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        if (left > 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) accum2(w[j], true, true);
            left -= 8;
        } else {
            // n >= 8: at most one group boundary per chunk (after element left - 1)
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const uint32_t wj = j == 0 ? w[0] : j == 1 ? w[1] : j == 2 ? w[2] : w[3];
                accum2(wj, true, false);
                if (--left == 0) { flush(g); ++g; left = n; }
                accum2(wj, false, true);
                if (--left == 0) { flush(g); ++g; left = n; }
            }
        }
    }
    flush(g);
    __syncthreads();
    // ---- emit: (group, channel) partial sums of this tile
    for (int i = threadIdx.x; i < kmax * MPS_CH; i += 256) {
        const int gl = i >> 6, ch = i & 63;
        const int64_t grp = g_first + gl;
        if (grp >= G || cb + ch >= C) continue;
        // points of the group inside this tile
        const int64_t lo = max(grp * (int64_t)n, pt0), hi = min((grp + 1) * (int64_t)n, pt0 + 256);
        if (hi <= lo) continue;
        const float* a = acc + (size_t)gl * NV * MPS_CH + ch;
        const int64_t o = grp * C + cb + ch;
        if (TRAIN) {
            const float sd = a[MPS_CH], su = a[2 * MPS_CH];
            atomicAdd(pooled + o, (a[0] + (sd - (float)(hi - lo))) * inv_n);
            atomicAdd(e1 + o, sd);
            atomicAdd(e2 + o, __ldg(invstd + cb + ch) * (su - __ldg(mean + cb + ch) * sd));
        } else if (APPLY) {
            atomicAdd(pooled + o, (a[0] - (float)(hi - lo)) * inv_n);
        } else {
            atomicAdd(pooled + o, a[0] * inv_n);
        }
    }
}

constexpr int EA_ROWS_PER_WARP = 4;

// outT = ELU(scale[c]*yT + shift[c]); grid (n_tiles, ceil(C / 32)): warp w streams the 512-byte rows of channels cb + w + 8*i
__global__ void __launch_bounds__(256)
bn_elu_apply_rows_kernel(const __nv_bfloat16* __restrict__ yT, const float* __restrict__ scale,
                         const float* __restrict__ shift, __nv_bfloat16* __restrict__ outT, int64_t P, int C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int cb = blockIdx.y * (8 * EA_ROWS_PER_WARP) + warp;
    const int64_t tile_off = tile * (int64_t)C * 256 + lane * 8;
    const int64_t left = P - ((tile << 8) + (lane << 3));
    const int vcnt = left >= 8 ? 8 : (left > 0 ? (int)left : 0);
    const bool full_tile = ((tile + 1) << 8) <= P;                  // block-uniform
    uint4 raw[EA_ROWS_PER_WARP];
#pragma unroll
    for (int i = 0; i < EA_ROWS_PER_WARP; ++i) {
        const int c = cb + 8 * i;
        if (c < C) raw[i] = __ldg(reinterpret_cast<const uint4*>(yT + tile_off + (int64_t)c * 256));
    }
#pragma unroll
    for (int i = 0; i < EA_ROWS_PER_WARP; ++i) {
        const int c = cb + 8 * i;
        if (c >= C) break;
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float scl = sc * LOG2E_F, shl = sh * LOG2E_F;
        float v[8];
        unpack8(raw[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = elu_l2(fmaf(v[j], sc, sh), fmaf(v[j], scl, shl));
        if (!full_tile) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = j < vcnt ? v[j] : 0.f;
        }
        *reinterpret_cast<uint4*>(outT + tile_off + (int64_t)c * 256) = pack8(v);
    }
}

constexpr int PB_ROWS_PER_WARP = 4;

// dy(c,p) = c1[c]*(dpool[g(p)][c]/n)*ELU'(scale*y+shift) + c2[c]*y + c3[c]; grid (n_tiles, ceil(C / 32)), n >= 8
__global__ void __launch_bounds__(256)
pool_bwd_apply_rows_kernel(const float* __restrict__ dpool, const __nv_bfloat16* __restrict__ yT,
                           const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ c1,
                           const float* __restrict__ c2, const float* __restrict__ c3, __nv_bfloat16* __restrict__ dyT,
                           int64_t P, int64_t G, int n, int C, float inv_n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int cb = blockIdx.y * (8 * PB_ROWS_PER_WARP) + warp;
    const LaneGroups lg = lane_groups(tile, lane, n, P);
    const int64_t tile_off = tile * (int64_t)C * 256 + lane * 8;
    const bool has0 = lg.g0 < G, has1 = lg.split < 8 && lg.g0 + 1 < G;
    const bool full_tile = ((tile + 1) << 8) <= P;                  // block-uniform
    float wsel[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wsel[j] = j < lg.split ? 1.f : 0.f;
    uint4 raw[PB_ROWS_PER_WARP];
    float g0v[PB_ROWS_PER_WARP], g1v[PB_ROWS_PER_WARP];
#pragma unroll
    for (int i = 0; i < PB_ROWS_PER_WARP; ++i) {
        const int c = cb + 8 * i;
        if (c < C) {
            raw[i] = __ldg(reinterpret_cast<const uint4*>(yT + tile_off + (int64_t)c * 256));
            g0v[i] = has0 ? __ldg(dpool + lg.g0 * C + c) : 0.f;
            g1v[i] = has1 ? __ldg(dpool + (lg.g0 + 1) * C + c) : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < PB_ROWS_PER_WARP; ++i) {
        const int c = cb + 8 * i;
        if (c >= C) break;
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        const float a1 = __ldg(c1 + c) * inv_n, a2 = __ldg(c2 + c), a3 = __ldg(c3 + c);
        const float scl = sc * LOG2E_F, shl = sh * LOG2E_F;
        const float gb = a1 * g1v[i], gd = a1 * g0v[i] - gb;          // gv_j = gb + wsel_j * gd
        // ELU'(z) = min(2^(z log2 e), 1): no select, z itself is not needed; packed fp32x2 arithmetic, two points per
        // instruction (FFMA2)
        const unsigned long long scl2 = pk2f(scl, scl), shl2 = pk2f(shl, shl), gd2 = pk2f(gd, gd), gb2 = pk2f(gb, gb);
        const unsigned long long a22 = pk2f(a2, a2), a32 = pk2f(a3, a3);
        const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
        float y[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned long long y2 = pk2f(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u));
            float al, ah;
            upk2f(fma2f(y2, scl2, shl2), al, ah);
            const unsigned long long d2 = pk2f(fminf(ex2_fast(al), 1.f), fminf(ex2_fast(ah), 1.f));
            const unsigned long long gv2 = fma2f(pk2f(wsel[2 * j], wsel[2 * j + 1]), gd2, gb2);
            upk2f(fma2f(gv2, d2, fma2f(a22, y2, a32)), y[2 * j], y[2 * j + 1]);
        }
        if (!full_tile) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = j < lg.vcnt ? y[j] : 0.f;
        }
        *reinterpret_cast<uint4*>(dyT + tile_off + (int64_t)c * 256) = pack8(y);
    }
}

}  // namespace pcaa

using namespace pcaa;
#define ST(s) ((cudaStream_t)(s))

static inline int64_t t256_chunks(int64_t P, int C) { return ((P + 255) / 256) * (int64_t)C * 32; }
static inline unsigned ew_blocks(int64_t nchunks) { return (unsigned)ceil_div(nchunks, 256 * EW_UNROLL); }

extern "C" {

int pcaa_pointnet_l1_fwd_t(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                           void* yT, double* stats, int64_t B, int64_t TN, int Cout, pcaa_stream stream) {
    if (B == 0) return PCAA_OK;
    const int64_t P = B * TN;
    PCAA_REQUIRE(((uintptr_t)yT & 15) == 0 && ((uintptr_t)w & 15) == 0 && Cout > 0, PCAA_ERR_ALIGN, "pointnet_l1_fwd_t: alignment / Cout");
    PCAA_REQUIRE((scale == nullptr) == (shift == nullptr), PCAA_ERR_SHAPE, "pointnet_l1_fwd_t: scale and shift go together");
    PCAA_REQUIRE(ceil_div(P, L1_TILE_POINTS) <= 65535, PCAA_ERR_SHAPE, "pointnet_l1: more than 134 M points in one call");
    dim3 grid((unsigned)ceil_div(Cout, 8 * L1_CH_PER_WARP), (unsigned)ceil_div(P, L1_TILE_POINTS));
    if (scale != nullptr && stats == nullptr)      // eval mode: the packed-arithmetic kernel, one output
        pointnet_l1_fwd_bn_t_kernel<false><<<grid, 256, 0, ST(stream)>>>(x, w, bias, scale, shift, (__nv_bfloat16*)yT, (__nv_bfloat16*)yT, P, TN, Cout);
    else
        pointnet_l1_fwd_t_kernel<<<grid, 256, 0, ST(stream)>>>(x, w, bias, scale, shift, (__nv_bfloat16*)yT, stats, P, TN, Cout);
    return check_launch("pointnet_l1_fwd_t");
}

int pcaa_pointnet_l1_fwd_bn_t(const float* x, const float* w, const float* bias, const float* scale, const float* shift,
                              void* yT, void* aT, int64_t B, int64_t TN, int Cout, pcaa_stream stream) {
    if (B == 0) return PCAA_OK;
    const int64_t P = B * TN;
    PCAA_REQUIRE(((uintptr_t)yT & 15) == 0 && ((uintptr_t)aT & 15) == 0 && ((uintptr_t)w & 15) == 0 && Cout > 0, PCAA_ERR_ALIGN,
                 "pointnet_l1_fwd_bn_t: alignment / Cout");
    PCAA_REQUIRE(yT && aT && scale && shift, PCAA_ERR_SHAPE, "pointnet_l1_fwd_bn_t: needs both outputs and the BatchNorm coefficients");
    PCAA_REQUIRE(ceil_div(P, L1_TILE_POINTS) <= 65535, PCAA_ERR_SHAPE, "pointnet_l1: more than 134 M points in one call");
    dim3 grid((unsigned)ceil_div(Cout, 8 * L1_CH_PER_WARP), (unsigned)ceil_div(P, L1_TILE_POINTS));
    pointnet_l1_fwd_bn_t_kernel<true><<<grid, 256, 0, ST(stream)>>>(x, w, bias, scale, shift, (__nv_bfloat16*)yT, (__nv_bfloat16*)aT, P, TN, Cout);
    return check_launch("pointnet_l1_fwd_bn_t");
}

int pcaa_input_moments(const float* x, int64_t B, int64_t TN, double* mom, pcaa_stream stream) {
    if (cudaMemsetAsync(mom, 0, sizeof(double) * 14, ST(stream)) != cudaSuccess) return check_launch("input_moments memset");
    if (B == 0) return PCAA_OK;
    const int64_t P = B * TN;
    int grid = (int)ceil_div(P, 256 * 8 * 4);
    if (grid > 148 * 4) grid = 148 * 4;
    if (grid < 1) grid = 1;
    input_moments_kernel<<<grid, 256, 0, ST(stream)>>>(x, P, TN, mom);
    return check_launch("input_moments");
}

int pcaa_bn_from_input_moments(const double* mom, int64_t R, int C, const float* w, const float* bias, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                               float* scale, float* shift, float* mean, float* invstd, pcaa_stream stream) {
    PCAA_REQUIRE(R > 0 && C > 0, PCAA_ERR_SHAPE, "bn_from_input_moments: empty");
    const double unbias = R > 1 ? (double)R / (double)(R - 1) : 1.0;
    bn_from_input_moments_kernel<<<ceil_div(C, 128), 128, 0, ST(stream)>>>(mom, 1.0 / (double)R, unbias, C, w, bias, gamma, beta,
                                                                         running_mean, running_var, momentum, eps, scale,
                                                                         shift, mean, invstd);
    return check_launch("bn_from_input_moments");
}

int pcaa_pointnet_l1_wgrad_t(const float* x, const void* dzT, const void* yT, const float* c1, const float* c2,
                             const float* c3, float* dW, int64_t B, int64_t TN, int Cout, pcaa_stream stream) {
    if (cudaMemsetAsync(dW, 0, sizeof(float) * 4 * Cout, ST(stream)) != cudaSuccess) return check_launch("pointnet_l1_wgrad_t memset");
    if (B == 0) return PCAA_OK;
    const int64_t P = B * TN;
    PCAA_REQUIRE(yT == nullptr || (c1 && c2 && c3), PCAA_ERR_SHAPE, "pointnet_l1_wgrad_t: y needs the BatchNorm-backward coefficients");
    PCAA_REQUIRE(ceil_div(P, L1_TILE_POINTS) <= 65535, PCAA_ERR_SHAPE, "pointnet_l1_wgrad_t: more than 134 M points in one call");
    dim3 grid((unsigned)ceil_div(Cout, 8 * L1W_CH_PER_WARP), (unsigned)ceil_div(P, L1_TILE_POINTS));
    if (yT != nullptr)
        pointnet_l1_wgrad_t_kernel<true><<<grid, 256, 0, ST(stream)>>>(x, (const __nv_bfloat16*)dzT, (const __nv_bfloat16*)yT, c1, c2, c3, dW, P, TN, Cout);
    else
        pointnet_l1_wgrad_t_kernel<false><<<grid, 256, 0, ST(stream)>>>(x, (const __nv_bfloat16*)dzT, nullptr, c1, c2, c3, dW, P, TN, Cout);
    return check_launch("pointnet_l1_wgrad_t");
}

int pcaa_bn_elu_apply_t(const void* yT, const float* scale, const float* shift, void* outT, int64_t P, int C,
                        pcaa_stream stream) {
    if (P == 0 || C == 0) return PCAA_OK;
    dim3 grid((unsigned)((P + 255) / 256), (unsigned)ceil_div(C, 8 * EA_ROWS_PER_WARP));
    bn_elu_apply_rows_kernel<<<grid, 256, 0, ST(stream)>>>((const __nv_bfloat16*)yT, scale, shift, (__nv_bfloat16*)outT, P, C);
    return check_launch("bn_elu_apply_t");
}

int pcaa_bn_bwd_apply_t(const void* dzT, const void* yT, const float* c1, const float* c2, const float* c3, void* dyT,
                        int64_t P, int C, pcaa_stream stream) {
    if (P == 0 || C == 0) return PCAA_OK;
    const int64_t nch = t256_chunks(P, C);
    bn_bwd_apply_t_kernel<<<ew_blocks(nch), 256, 0, ST(stream)>>>((const __nv_bfloat16*)dzT, (const __nv_bfloat16*)yT, c1, c2, c3, (__nv_bfloat16*)dyT, nch, P, C);
    return check_launch("bn_bwd_apply_t");
}

int pcaa_bn_elu_meanpool_t(const void* yT, const float* scale, const float* shift, const float* mean,
                           const float* invstd, float* pooled, float* e1, float* e2, int64_t G, int n, int C,
                           pcaa_stream stream) {
    if (G == 0 || C == 0) return PCAA_OK;
    PCAA_REQUIRE(n > 0, PCAA_ERR_SHAPE, "bn_elu_meanpool_t: bad group size");
    PCAA_REQUIRE((e1 == nullptr) == (e2 == nullptr) && (e1 == nullptr || (mean && invstd && scale)), PCAA_ERR_SHAPE,
                 "bn_elu_meanpool_t: e1/e2 need mean/invstd/scale");
    const int apply = scale != nullptr;
    if (n >= 8) {
        const int64_t P = G * n;
        const size_t bytes = sizeof(float) * (size_t)G * C;
        if (cudaMemsetAsync(pooled, 0, bytes, ST(stream)) != cudaSuccess || (e1 && cudaMemsetAsync(e1, 0, bytes, ST(stream)) != cudaSuccess) ||
            (e2 && cudaMemsetAsync(e2, 0, bytes, ST(stream)) != cudaSuccess))
            return check_launch("bn_elu_meanpool_t memset");
        const int kmax = 255 / n + 2;
        const size_t smem = (size_t)MPS_CH * MPS_PITCH * sizeof(__nv_bfloat16) + (size_t)kmax * (e1 ? 3 : 1) * MPS_CH * sizeof(float);
        dim3 grid((unsigned)((P + 255) / 256), (unsigned)ceil_div(C, MPS_CH));
        const float inv_n = 1.f / (float)n;
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(bn_elu_meanpool_staged_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(bn_elu_meanpool_staged_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(bn_elu_meanpool_staged_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            attr = true;
        }
        if (e1)
            bn_elu_meanpool_staged_kernel<true, true><<<grid, 256, smem, ST(stream)>>>((const __nv_bfloat16*)yT, scale, shift, mean, invstd, pooled, e1, e2, P, G, n, C, kmax, inv_n);
        else if (apply)
            bn_elu_meanpool_staged_kernel<false, true><<<grid, 256, smem, ST(stream)>>>((const __nv_bfloat16*)yT, scale, shift, mean, invstd, pooled, e1, e2, P, G, n, C, kmax, inv_n);
        else
            bn_elu_meanpool_staged_kernel<false, false><<<grid, 256, smem, ST(stream)>>>((const __nv_bfloat16*)yT, scale, shift, mean, invstd, pooled, e1, e2, P, G, n, C, kmax, inv_n);
        return check_launch("bn_elu_meanpool_t");
    }
    dim3 grid((unsigned)ceil_div(G, 8), (unsigned)ceil_div(C, 32));
    bn_elu_meanpool_t_kernel<<<grid, 256, 0, ST(stream)>>>((const __nv_bfloat16*)yT, scale, shift, mean, invstd, pooled, e1, e2, G, n, C, apply);
    return check_launch("bn_elu_meanpool_t");
}

int pcaa_pool_bwd_stats(const float* dpool, const float* e1, const float* e2, double* stats2, int64_t G, int n, int C,
                        pcaa_stream stream) {
    if (G == 0 || C == 0) return PCAA_OK;
    int splits = (int)(G / 64);
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    pool_bwd_stats_kernel<<<dim3((unsigned)ceil_div(C, 32), splits), dim3(32, 8), 0, ST(stream)>>>(dpool, e1, e2, stats2, G, C, 1.f / (float)n);
    return check_launch("pool_bwd_stats");
}

int pcaa_pool_bwd_apply_t(const float* dpool, const void* yT, const float* scale, const float* shift, const float* c1,
                          const float* c2, const float* c3, void* dyT, int64_t G, int n, int C, pcaa_stream stream) {
    if (G == 0 || C == 0) return PCAA_OK;
    PCAA_REQUIRE(n >= 1, PCAA_ERR_SHAPE, "pool_bwd_apply_t: bad group size");
    const int64_t P = G * n;
    if (n >= 8) {
        dim3 grid((unsigned)((P + 255) / 256), (unsigned)ceil_div(C, 8 * PB_ROWS_PER_WARP));
        pool_bwd_apply_rows_kernel<<<grid, 256, 0, ST(stream)>>>(dpool, (const __nv_bfloat16*)yT, scale, shift, c1, c2, c3, (__nv_bfloat16*)dyT, P, G, n, C, 1.f / (float)n);
        return check_launch("pool_bwd_apply_t");
    }
    const int64_t nch = t256_chunks(P, C);
    pool_bwd_apply_t_kernel<<<ew_blocks(nch), 256, 0, ST(stream)>>>(dpool, (const __nv_bfloat16*)yT, scale, shift, c1, c2, c3, (__nv_bfloat16*)dyT, nch, P, n, C, 1.f / (float)n);
    return check_launch("pool_bwd_apply_t");
}

}  // extern "C"
