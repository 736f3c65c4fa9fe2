// BatchNorm / ELU / pooling / small helper kernels of the PCAA hot path (HBM-bound, vectorised 16 B accesses).
// Rows matrices are channels-last [R, C], C % 8 == 0 for the BatchNorm family.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace pcaa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PCAA_ERR_LAUNCH;
    }
    return PCAA_OK;
}

const char* last_error_cstr() { return g_err; }

// ---- 8-wide vector access ---------------------------------------------------------------------------------
template <typename T> struct V8;
template <> struct V8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p));
        float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct V8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};

__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) { V8<float>::load(p, v); }

// geometry shared by the column-reduction kernels: 256 threads = cgs column groups x lanes row lanes
struct ColGeom {
    int cgs, lanes;
    dim3 grid, block;
};
static ColGeom col_geom(int64_t R, int C) {
    ColGeom g;
    int ncg = C / 8;
    g.cgs = ncg < 32 ? ncg : 32;
    while (256 % g.cgs) --g.cgs;  // C/8 is a power of two in practice; keep it general
    g.lanes = 256 / g.cgs;
    int gx = (ncg + g.cgs - 1) / g.cgs;
    long long want = (R + g.lanes - 1) / g.lanes;
    long long cap = (148LL * 8 + gx - 1) / gx;
    int gy = (int)(want < cap ? want : cap);
    if (gy < 1) gy = 1;
    g.grid = dim3(gx, gy);
    g.block = dim3(256);
    return g;
}

// reduce 16 per-thread partials over the row lanes of the block, then one double atomic per column
__device__ __forceinline__ void block_col_reduce(float (&s1)[8], float (&s2)[8], int cg, int lane, int cgs, int lanes,
                                                 int colgroup_global, int ncg, double* stats, int C) {
    extern __shared__ float sm[];
    float* a = sm;                       // [lanes][cgs][16]
    float* mine = a + ((size_t)lane * cgs + cg) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mine[i] = s1[i];
        mine[8 + i] = s2[i];
    }
    __syncthreads();
    // threads 0 .. cgs*16-1 finish the reduction
    for (int t = threadIdx.x; t < cgs * 16; t += blockDim.x) {
        int g = t / 16, j = t % 16;
        double acc = 0.0;
        for (int l = 0; l < lanes; ++l) acc += (double)a[((size_t)l * cgs + g) * 16 + j];
        int cgg = blockIdx.x * cgs + g;
        if (cgg < ncg) {
            int col = cgg * 8 + (j & 7);
            atomicAdd(&stats[(j < 8 ? 0 : C) + col], acc);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) colstats_kernel(const T* __restrict__ y, int64_t R, int C, double* stats,
                                                       int cgs, int lanes) {
    int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
    int ncg = C / 8;
    int cgg = blockIdx.x * cgs + cg;
    float s1[8] = {0}, s2[8] = {0};
    if (cgg < ncg) {
        for (int64_t r = (int64_t)blockIdx.y * lanes + lane; r < R; r += (int64_t)gridDim.y * lanes) {
            float v[8];
            V8<T>::load(y + r * C + cgg * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s1[i] += v[i];
                s2[i] += v[i] * v[i];
            }
        }
    }
    block_col_reduce(s1, s2, cg, lane, cgs, lanes, cgg, ncg, stats, C);
}

__global__ void bn_finalize_kernel(const double* stats, double invR, double unbias, int C, const float* gamma,
                                   const float* beta, float* rmean, float* rvar, float momentum, float eps,
                                   float* scale, float* shift, float* mean_o, float* invstd_o) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mean = stats[c] * invR;
    double var = stats[C + c] * invR - mean * mean;
    if (var < 0) var = 0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    if (mean_o) mean_o[c] = (float)mean;
    if (invstd_o) invstd_o[c] = invstd;
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * unbias);
}

__global__ void bn_eval_coeffs_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                      float eps, float* scale, float* shift, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float sc = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - rm[c] * sc;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bn_elu_apply_kernel(const TI* __restrict__ y, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, TO* __restrict__ out,
                                                           int64_t nchunks, int ncg, int hoist) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float sc[8], sh[8];
    if (hoist && tid < nchunks) {
        int cg = (int)(tid % ncg);
        load8f(scale + cg * 8, sc);
        load8f(shift + cg * 8, sh);
    }
    for (int64_t ch = tid; ch < nchunks; ch += stride) {
        if (!hoist) {
            int cg = (int)(ch % ncg);
            load8f(scale + cg * 8, sc);
            load8f(shift + cg * 8, sh);
        }
        float v[8];
        V8<TI>::load(y + ch * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = elu_f(fmaf(v[i], sc[i], sh[i]));
        V8<TO>::store(out + ch * 8, v);
    }
}

// block = (C/8 column groups [<=128]) x RL row lanes; one block per (group g, column span)
template <typename T, int RL>
__global__ void bn_elu_meanpool_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                                       const float* __restrict__ shift, float* __restrict__ pooled, int n, int C,
                                       int cgs) {
    extern __shared__ float sm[];
    int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
    int cgg = blockIdx.y * cgs + cg;
    int64_t g = blockIdx.x;
    float acc[8] = {0};
    bool active = cgg * 8 < C;
    if (active) {
        float sc[8], sh[8];
        load8f(scale + cgg * 8, sc);
        load8f(shift + cgg * 8, sh);
        const T* base = y + (g * n) * (int64_t)C + cgg * 8;
        for (int i = lane; i < n; i += RL) {
            float v[8];
            V8<T>::load(base + (int64_t)i * C, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += elu_f(fmaf(v[k], sc[k], sh[k]));
        }
    }
    float* mine = sm + ((size_t)lane * cgs + cg) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) mine[k] = acc[k];
    __syncthreads();
    if (lane == 0 && active) {
        float inv = 1.f / (float)n;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float s = 0.f;
            for (int l = 0; l < RL; ++l) s += sm[((size_t)l * cgs + cg) * 8 + k];
            o[k] = s * inv;
        }
        V8<float>::store(pooled + g * (int64_t)C + cgg * 8, o);
    }
}

template <typename TD, typename TY, typename TZ>
__global__ void __launch_bounds__(256)
elu_bwd_colstats_kernel(const TD* __restrict__ dout, const float* __restrict__ dpool, int pooled_n,
                        const TY* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd, TZ* __restrict__ dz,
                        double* stats2, int64_t R, int C, int cgs, int lanes) {
    int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
    int ncg = C / 8;
    int cgg = blockIdx.x * cgs + cg;
    float s1[8] = {0}, s2[8] = {0};
    if (cgg < ncg) {
        float sc[8], sh[8], mu[8], is[8];
        load8f(scale + cgg * 8, sc);
        load8f(shift + cgg * 8, sh);
        load8f(mean + cgg * 8, mu);
        load8f(invstd + cgg * 8, is);
        float invn = pooled_n > 0 ? 1.f / (float)pooled_n : 0.f;
        for (int64_t r = (int64_t)blockIdx.y * lanes + lane; r < R; r += (int64_t)gridDim.y * lanes) {
            float v[8], d[8];
            V8<TY>::load(y + r * C + cgg * 8, v);
            if (pooled_n > 0) {
                load8f(dpool + (r / pooled_n) * C + cgg * 8, d);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] *= invn;
            } else {
                V8<TD>::load(dout + r * C + cgg * 8, d);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float z = fmaf(v[i], sc[i], sh[i]);
                float g = d[i] * elu_grad_f(z);
                float xh = (v[i] - mu[i]) * is[i];
                d[i] = g;
                s1[i] += g;
                s2[i] += g * xh;
            }
            V8<TZ>::store(dz + r * C + cgg * 8, d);
        }
    }
    block_col_reduce(s1, s2, cg, lane, cgs, lanes, cgg, ncg, stats2, C);
}

// ------------------------------------------------------------------------------------------------ fused TCN layer steps
// The six causal-conv layers (models.py:37-79, 108-160) work on [B*30, C <= 1024] matrices: every step is a few
// microseconds, so the layer pipeline is fused to cut launches (and the round trips of intermediates), not bytes.

// forward, after the layer's GEMM (whose epilogue accumulated the column statistics): BatchNorm1d coefficients (batch
// statistics in training -- then block 0 also publishes them and updates the running statistics --, given scale / shift
// in eval) + ELU, written as the NEXT layer's bf16 im2col operand ([R, C*3], tap k of channel c at column c*3 + k holds
// the activation dil_next*(2-k) steps earlier, zero before the sequence start) and / or as the fp32 activation.
__global__ void __launch_bounds__(256)
tcn_bn_elu_next_kernel(const float* __restrict__ y, const double* __restrict__ stats, double invR, double unbias,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float* rmean, float* rvar,
                       float momentum, float eps, const float* __restrict__ scale_in, const float* __restrict__ shift_in,
                       float* __restrict__ coef_out, int64_t R, int T, int C, int dil_next,
                       __nv_bfloat16* __restrict__ col, float* __restrict__ act) {
    extern __shared__ float sm[];
    float* sc = sm;
    float* sh = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (stats != nullptr) {
            const double mean = stats[c] * invR;
            double var = stats[C + c] * invR - mean * mean;
            if (var < 0) var = 0;
            const float invstd = (float)(1.0 / sqrt(var + (double)eps));
            const float s = gamma[c] * invstd;
            sc[c] = s;
            sh[c] = beta[c] - (float)mean * s;
            if (blockIdx.x == 0) {
                coef_out[c] = s;
                coef_out[C + c] = sh[c];
                coef_out[2 * C + c] = (float)mean;
                coef_out[3 * C + c] = invstd;
                if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
                if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * unbias);
            }
        } else {
            sc[c] = scale_in[c];
            sh[c] = shift_in[c];
        }
    }
    __syncthreads();
    if (C % 8 == 0) {
        // 8 channels per thread: 32-byte reads of the (up to three) source rows, the im2col row written as one 48-byte run
        const int ncg = C / 8;
        const int64_t total = R * ncg;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int c0 = (int)(i % ncg) * 8;
            const int64_t r = i / ncg;
            const int t = (int)(r % T);
            float a[3][8];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ts = t - (2 - k) * dil_next;
                if ((k == 2 || col != nullptr) && ts >= 0) {
                    float v[8];
                    V8<float>::load(y + (r - t + ts) * C + c0, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[k][j] = elu_f(fmaf(v[j], sc[c0 + j], sh[c0 + j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[k][j] = 0.f;
                }
            }
            if (act) V8<float>::store(act + r * C + c0, a[2]);
            if (col) {
                __nv_bfloat16* dst = col + (r * C + c0) * 3;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    float w8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = g * 8 + j;              // element e of the 24: channel e / 3, tap e % 3
                        w8[j] = a[e % 3][e / 3];
                    }
                    V8<__nv_bfloat16>::store(dst + g * 8, w8);
                }
            }
        }
        return;
    }
    const int64_t total = R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const float a = elu_f(fmaf(y[i], sc[c], sh[c]));
        if (act) act[i] = a;
        if (col) {
            const int t = (int)(r % T);
            __nv_bfloat16* o = col + i * 3;
            o[2] = __float2bfloat16_rn(a);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ts = t - (2 - k) * dil_next;
                o[k] = __float2bfloat16_rn(ts >= 0 ? elu_f(fmaf(y[(r - t + ts) * C + c], sc[c], sh[c])) : 0.f);
            }
        }
    }
}

// backward, first pass of a layer: the gradient d reaching its activation is formed on the fly --
//   src_mode 0: d = src[R, C];
//   src_mode 1: col2im of the layer above, d[(b,t), c] = sum_k src[(b, t + (2-k)*dil_up), c*3 + k] for t + (2-k)*dil_up < T
//               (src = that layer's d im2col [R, C*3]);
//   src_mode 2: the mean over the T frames broadcast back, d[(b,t), c] = src[b, c] / T (src [R/T, C]);
// then dz = d * ELU'(scale*y + shift) is stored and stats2 += [sum dz, sum dz * xhat] (the BatchNorm-backward sums).
__global__ void __launch_bounds__(256)
tcn_elu_bwd_stats_kernel(const float* __restrict__ src, int src_mode, int dil_up, int T,
                         const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ dz,
                         double* stats2, int64_t R, int C, int cgs, int lanes) {
    int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
    int ncg = C / 8;
    int cgg = blockIdx.x * cgs + cg;
    float s1[8] = {0}, s2[8] = {0};
    if (cgg < ncg) {
        float sc[8], sh[8], mu[8], is[8];
        load8f(scale + cgg * 8, sc);
        load8f(shift + cgg * 8, sh);
        load8f(mean + cgg * 8, mu);
        load8f(invstd + cgg * 8, is);
        const float invT = 1.f / (float)T;
        for (int64_t r = (int64_t)blockIdx.y * lanes + lane; r < R; r += (int64_t)gridDim.y * lanes) {
            float v[8], d[8];
            V8<float>::load(y + r * C + cgg * 8, v);
            if (src_mode == 0) {
                V8<float>::load(src + r * C + cgg * 8, d);
            } else if (src_mode == 2) {
                load8f(src + (r / T) * C + cgg * 8, d);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] *= invT;
            } else {
                const int t = (int)(r % T);
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int to = t + (2 - k) * dil_up;
                    if (to < T) {
                        const float* q = src + ((r - t + to) * C + cgg * 8) * 3 + k;
#pragma unroll
                        for (int i = 0; i < 8; ++i) d[i] += __ldg(q + 3 * i);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float z = fmaf(v[i], sc[i], sh[i]);
                float g = d[i] * elu_grad_f(z);
                float xh = (v[i] - mu[i]) * is[i];
                d[i] = g;
                s1[i] += g;
                s2[i] += g * xh;
            }
            V8<float>::store(dz + r * C + cgg * 8, d);
        }
    }
    block_col_reduce(s1, s2, cg, lane, cgs, lanes, cgg, ncg, stats2, C);
}

// backward, second pass: BatchNorm-backward coefficients from the completed sums (every block derives them for all C
// channels; block 0 also writes d gamma / d beta), dy = c1*dz + c2*y + c3 stored as the bf16 operand of the two GEMMs
__global__ void __launch_bounds__(256)
tcn_bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y, const double* __restrict__ stats2,
                        double invR, const float* __restrict__ scale, const float* __restrict__ mean,
                        const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        __nv_bfloat16* __restrict__ dy, int64_t R, int C) {
    extern __shared__ float sm[];
    float* c1 = sm;
    float* c2 = sm + C;
    float* c3 = sm + 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double t1 = stats2[c], t2 = stats2[C + c];
        const double s = scale[c], is = invstd[c], mu = mean[c];
        c1[c] = (float)s;
        c2[c] = (float)(-s * is * t2 * invR);
        c3[c] = (float)(-s * t1 * invR + s * is * mu * t2 * invR);
        if (blockIdx.x == 0) {
            if (dgamma) dgamma[c] = (float)t2;
            if (dbeta) dbeta[c] = (float)t1;
        }
    }
    __syncthreads();
    const int ncg = C / 8;
    const int64_t nchunks = R * ncg;
    for (int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(ch % ncg) * 8;
        float z[8], v[8];
        V8<float>::load(dz + ch * 8, z);
        V8<float>::load(y + ch * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaf(c1[c0 + i], z[i], fmaf(c2[c0 + i], v[i], c3[c0 + i]));
        V8<__nv_bfloat16>::store(dy + ch * 8, z);
    }
}

__global__ void bn_bwd_finalize_kernel(const double* stats2, double invR, int C, const float* scale,
                                       const float* mean, const float* invstd, float* c1, float* c2, float* c3,
                                       float* dgamma, float* dbeta) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double t1 = stats2[c], t2 = stats2[C + c];
    double sc = scale[c], is = invstd[c], mu = mean[c];
    c1[c] = (float)sc;
    c2[c] = (float)(-sc * is * t2 * invR);
    c3[c] = (float)(-sc * t1 * invR + sc * is * mu * t2 * invR);
    if (dgamma) dgamma[c] = (float)t2;
    if (dbeta) dbeta[c] = (float)t1;
}

template <typename TZ, typename TY, typename TO>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const TZ* __restrict__ dz, const TY* __restrict__ y, const float* __restrict__ c1,
                    const float* __restrict__ c2, const float* __restrict__ c3, TO* __restrict__ dy, int64_t nchunks,
                    int ncg, int hoist) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float a[8], b[8], c[8];
    if (hoist && tid < nchunks) {
        int cg = (int)(tid % ncg);
        load8f(c1 + cg * 8, a);
        load8f(c2 + cg * 8, b);
        load8f(c3 + cg * 8, c);
    }
    for (int64_t ch = tid; ch < nchunks; ch += stride) {
        if (!hoist) {
            int cg = (int)(ch % ncg);
            load8f(c1 + cg * 8, a);
            load8f(c2 + cg * 8, b);
            load8f(c3 + cg * 8, c);
        }
        float z[8], v[8];
        V8<TZ>::load(dz + ch * 8, z);
        V8<TY>::load(y + ch * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaf(a[i], z[i], fmaf(b[i], v[i], c[i]));
        V8<TO>::store(dy + ch * 8, z);
    }
}

__global__ void elu_bwd_from_out_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                        float* __restrict__ dz, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float o = out[i];
        dz[i] = dout[i] * (o > 0.f ? 1.f : o + 1.f);
    }
}

// out[c] = sum_r x[r,c]; one block per 32 columns, 8 row lanes
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t R, int C, int64_t ld, float* __restrict__ out) {
    __shared__ float sm[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < C)
        for (int64_t r = threadIdx.y; r < R; r += 8) s += ld_as_float<T>(x + r * ld + c);
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
        for (int l = 0; l < 8; ++l) t += sm[l][threadIdx.x];
        out[c] = t;
    }
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) st_from_float<TO>(out + i, ld_as_float<TI>(in + i));
}

// out[r, c] = in[r, c] (or transposed) with zero padding up to ld_out; 32x32 smem tile
__global__ void pack_bf16_kernel(const float* __restrict__ in, int64_t R, int64_t C, int64_t ld_in,
                                 __nv_bfloat16* __restrict__ out, int64_t ld_out, int transpose, int64_t out_rows) {
    __shared__ float tile[32][33];
    // output coordinates
    int64_t oc0 = (int64_t)blockIdx.x * 32, or0 = (int64_t)blockIdx.y * 32;
    if (!transpose) {
        for (int i = threadIdx.y; i < 32; i += 8) {
            int64_t r = or0 + i, c = oc0 + threadIdx.x;
            if (r < out_rows && c < ld_out)
                out[r * ld_out + c] = __float2bfloat16_rn((r < R && c < C) ? in[r * ld_in + c] : 0.f);
        }
    } else {
        // out[r=c_in, c=r_in]
        for (int i = threadIdx.y; i < 32; i += 8) {
            int64_t rin = oc0 + i, cin = or0 + threadIdx.x;
            tile[i][threadIdx.x] = (rin < R && cin < C) ? in[rin * ld_in + cin] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += 8) {
            int64_t r = or0 + i, c = oc0 + threadIdx.x;
            if (r < out_rows && c < ld_out) out[r * ld_out + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
        }
    }
}

template <typename TO>
__global__ void tcn_im2col_kernel(const float* __restrict__ x, TO* __restrict__ col, int64_t BT, int T, int Cin,
                                  int dil) {
    int64_t total = BT * Cin * 3;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int k = (int)(i % 3);
        int64_t q = i / 3;
        int ci = (int)(q % Cin);
        int64_t bt = q / Cin;
        int t = (int)(bt % T);
        int ts = t - (2 - k) * dil;
        st_from_float<TO>(col + i, ts >= 0 ? x[(bt - t + ts) * Cin + ci] : 0.f);
    }
}

// bf16 im2col, 8 channels per thread (Cin % 8 == 0): three 32-byte reads of the shifted rows, one 48-byte run of the
// output row (24 consecutive bf16: channels c0 .. c0+7, taps interleaved) as three 16-byte stores
__global__ void __launch_bounds__(256)
tcn_im2col_bf16x8_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int64_t BT, int T, int Cin, int dil) {
    const int ncg = Cin / 8;
    const int64_t total = BT * ncg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cg = (int)(i % ncg);
        const int64_t bt = i / ncg;
        const int t = (int)(bt % T);
        float v[3][8];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int ts = t - (2 - k) * dil;
            if (ts >= 0) {
                V8<float>::load(x + (bt - t + ts) * Cin + cg * 8, v[k]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
            }
        }
        float o[24];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) o[j * 3 + k] = v[k][j];
        __nv_bfloat16* dst = col + (bt * Cin + cg * 8) * 3;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = o[g * 8 + j];
            V8<__nv_bfloat16>::store(dst + g * 8, w8);
        }
    }
}

__global__ void tcn_col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int64_t BT, int T, int Cin,
                                  int dil) {
    int64_t total = BT * Cin;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int ci = (int)(i % Cin);
        int64_t bt = i / Cin;
        int t = (int)(bt % T);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int to = t + (2 - k) * dil;  // output time that read this input through tap k
            if (to < T) s += dcol[((bt - t + to) * Cin + ci) * 3 + k];
        }
        dx[i] = s;
    }
}

__global__ void mean_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t G, int n, int C) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * C) return;
    int c = (int)(i % C);
    int64_t g = i / C;
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += x[(g * n + j) * C + c];
    out[i] = s / (float)n;
}

__global__ void mean_rows_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int64_t G, int n, int C) {
    int64_t total = G * n * C;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float inv = 1.f / (float)n;
    for (; i < total; i += stride) {
        int c = (int)(i % C);
        int64_t gi = i / ((int64_t)n * C);
        dx[i] = g[gi * C + c] * inv;
    }
}

// one warp per sample; C <= 1024
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ gt, float* loss,
                                  float* __restrict__ dlogits, float gscale, int32_t* __restrict__ pred, int64_t B,
                                  int C) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int lane = threadIdx.x % 32;
    if (warp >= B) return;
    const float* l = logits + (int64_t)warp * C;
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        float v = l[c];
        if (v > mx) { mx = v; am = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, mx, o);
        int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(l[c] - mx);
    se = warp_sum(se);
    float lse = mx + logf(se);
    int t = (int)gt[warp];
    // a label outside [0, C) (the reference raises on it) must not index the logits: its loss term is NaN instead, so the
    // batch loss and every gradient downstream turn NaN -- loud, without a host synchronisation
    const bool t_ok = t >= 0 && t < C;
    if (lane == 0) {
        if (pred) pred[warp] = am;
        if (loss) atomicAdd(loss, t_ok ? (lse - l[t]) / (float)B : NAN);
    }
    if (dlogits) {
        float s = gscale / (float)B;
        for (int c = lane; c < C; c += 32) dlogits[(int64_t)warp * C + c] = (expf(l[c] - lse) - (c == t ? 1.f : 0.f)) * s;
    }
}

// dst[i] += sum_k src[k * stride + i]: the local reduction of the copy-engine gradient exchange (dp.PeerExchange): the
// chunks pulled from the peers' gradient buffers are added into this rank's chunk.  16-byte accesses, grid-stride.
__global__ void __launch_bounds__(256)
sum_into_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4, int64_t stride, int nsrc, int overwrite) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 a = overwrite ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(dst)[i];
        for (int k = 0; k < nsrc; ++k) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + k * stride) + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        reinterpret_cast<float4*>(dst)[i] = a;
    }
}

// dst[r] = src[idx[r]] for rows of row_vec 16-byte words (batch assembly from an HBM-resident packed crop store);
// grid (rows, splits).  An index outside [0, n_src) yields a zero row.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ idx, uint4* __restrict__ dst,
                   int64_t row_vec, int64_t n_src) {
    const int64_t r = blockIdx.x;
    const int64_t s = idx[r];
    const bool ok = s >= 0 && s < n_src;
    const uint4* in = src + (ok ? s : 0) * row_vec;
    uint4* out = dst + r * row_vec;
    for (int64_t i = (int64_t)blockIdx.y * 256 + threadIdx.x; i < row_vec; i += (int64_t)gridDim.y * 256)
        out[i] = ok ? __ldg(in + i) : make_uint4(0u, 0u, 0u, 0u);
}

// step += 1; coef = { lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step) }: the host-side arithmetic of pcaa_adam_flat on
// a device-resident counter, so a captured CUDA graph of the train step advances the optimizer on every replay
__global__ void adam_advance_kernel(int32_t* __restrict__ step, float* __restrict__ coef, float lr, float beta1, float beta2) {
    const int s = step[0] + 1;
    step[0] = s;
    const double bc1 = 1.0 - pow((double)beta1, (double)s), bc2 = 1.0 - pow((double)beta2, (double)s);
    coef[0] = (float)((double)lr / bc1);
    coef[1] = (float)(1.0 / sqrt(bc2));
}

__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, float lr_bc1, float beta1, float beta2, float eps,
                                 float inv_sqrt_bc2, float grad_scale, __nv_bfloat16* __restrict__ shadow,
                                 const float* __restrict__ coef_dev) {
    if (coef_dev) {                      // bias corrections of a device-resident step counter (CUDA-graph replay)
        lr_bc1 = __ldg(coef_dev);
        inv_sqrt_bc2 = __ldg(coef_dev + 1);
    }
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        if (i + 3 < n) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gk = ga[k] * grad_scale;
                ma[k] = beta1 * ma[k] + (1.f - beta1) * gk;
                va[k] = beta2 * va[k] + (1.f - beta2) * gk * gk;
                float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
                pa[k] = pa[k] - lr_bc1 * (ma[k] / denom);
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (shadow) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&lo);
                u.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(shadow + i) = u;
            }
        } else {
            for (int64_t j = i; j < n; ++j) {
                float gk = g[j] * grad_scale;
                float mk = beta1 * m[j] + (1.f - beta1) * gk;
                float vk = beta2 * v[j] + (1.f - beta2) * gk * gk;
                m[j] = mk;
                v[j] = vk;
                float pk = p[j] - lr_bc1 * (mk / (sqrtf(vk) * inv_sqrt_bc2 + eps));
                p[j] = pk;
                if (shadow) shadow[j] = __float2bfloat16_rn(pk);
            }
        }
    }
}

// tiny element-wise family used by the twice-differentiable critic path (models.CGDiscriminator under
// torch.autograd.grad(create_graph=True), reference PCAA_ablation.py:955-962)
__global__ void ew_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                          int64_t n, int ncols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float x = a ? a[i] : 0.f, r;
        switch (op) {
            case PCAA_EW_MUL: r = x * b[i]; break;
            case PCAA_EW_ADD: r = x + b[i]; break;
            case PCAA_EW_ELU: r = elu_f(x); break;
            case PCAA_EW_ELU_GRAD: r = x > 0.f ? 1.f : expf(x); break;
            case PCAA_EW_ELU_GRAD2: r = x > 0.f ? 0.f : expf(x); break;
            case PCAA_EW_ADD_ROWVEC: r = x + b[i % ncols]; break;
            default: r = x;
        }
        out[i] = r;
    }
}

// ---- PointNet layer 1 -------------------------------------------------------------------------------------
// block: (Cout/8) channel groups x PL point lanes (256 threads for Cout = 512 -> 64 x 4)
template <int PL>
__global__ void __launch_bounds__(256)
pointnet_l1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       __nv_bfloat16* __restrict__ y, double* stats, int64_t P, int64_t TN, int Cout, int pts_per_block) {
    extern __shared__ float sm[];
    int ncg = Cout / 8;
    int cg = threadIdx.x % ncg, lane = threadIdx.x / ncg;
    float wr[8][4], br[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(w + (size_t)(cg * 8 + i) * 4));
        wr[i][0] = t.x; wr[i][1] = t.y; wr[i][2] = t.z; wr[i][3] = t.w;
        br[i] = bias[cg * 8 + i];
    }
    float s1[8] = {0}, s2[8] = {0};
    int64_t p0 = (int64_t)blockIdx.x * pts_per_block;
    int64_t p1 = p0 + pts_per_block < P ? p0 + pts_per_block : P;
    for (int64_t p = p0 + lane; p < p1; p += PL) {
        int64_t b = p / TN, tn = p % TN;
        const float* xb = x + b * 4 * TN + tn;
        float x0 = __ldg(xb), x1 = __ldg(xb + TN), x2 = __ldg(xb + 2 * TN), x3 = __ldg(xb + 3 * TN);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = fmaf(wr[i][0], x0, br[i]);
            a = fmaf(wr[i][1], x1, a);
            a = fmaf(wr[i][2], x2, a);
            a = fmaf(wr[i][3], x3, a);
            v[i] = a;
            s1[i] += a;
            s2[i] += a * a;
        }
        V8<__nv_bfloat16>::store(y + p * Cout + cg * 8, v);
    }
    if (stats) {
        float* mine = sm + ((size_t)lane * ncg + cg) * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) { mine[i] = s1[i]; mine[8 + i] = s2[i]; }
        __syncthreads();
        for (int t = threadIdx.x; t < ncg * 16; t += blockDim.x) {
            int g = t / 16, j = t % 16;
            double acc = 0.0;
            for (int l = 0; l < PL; ++l) acc += (double)sm[((size_t)l * ncg + g) * 16 + j];
            atomicAdd(&stats[(j < 8 ? 0 : Cout) + g * 8 + (j & 7)], acc);
        }
    }
}

template <int PL>
__global__ void __launch_bounds__(256)
pointnet_l1_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* dW, int64_t P,
                         int64_t TN, int Cout, int pts_per_block) {
    extern __shared__ float sm[];
    int ncg = Cout / 8;
    int cg = threadIdx.x % ncg, lane = threadIdx.x / ncg;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    for (int64_t blk = blockIdx.x; blk * pts_per_block < P; blk += gridDim.x) {
        int64_t p0 = blk * pts_per_block;
        int64_t p1 = p0 + pts_per_block < P ? p0 + pts_per_block : P;
        for (int64_t p = p0 + lane; p < p1; p += PL) {
            int64_t b = p / TN, tn = p % TN;
            const float* xb = x + b * 4 * TN + tn;
            float xv[4] = {__ldg(xb), __ldg(xb + TN), __ldg(xb + 2 * TN), __ldg(xb + 3 * TN)};
            float d[8];
            V8<__nv_bfloat16>::load(dy + p * Cout + cg * 8, d);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(d[i], xv[k], acc[i][k]);
        }
    }
    float* mine = sm + ((size_t)lane * ncg + cg) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) mine[i * 4 + k] = acc[i][k];
    __syncthreads();
    for (int t = threadIdx.x; t < ncg * 32; t += blockDim.x) {
        int g = t / 32, j = t % 32;
        float s = 0.f;
        for (int l = 0; l < PL; ++l) s += sm[((size_t)l * ncg + g) * 32 + j];
        atomicAdd(&dW[(size_t)(g * 8) * 4 + j], s);   // j = i*4 + k -> row g*8+i, col k
    }
}

}  // namespace pcaa

using namespace pcaa;

extern "C" {

const char* pcaa_version(void) { return "pcaa-sm100 0.1 (sm_100a; tcgen05/TMA)"; }
const char* pcaa_last_error(void) { return last_error_cstr(); }

int pcaa_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#define ST(s) ((cudaStream_t)(s))

int pcaa_colstats(const void* y, int dtype, int64_t R, int C, double* stats, pcaa_stream stream) {
    PCAA_REQUIRE(C % 8 == 0 && C > 0 && R > 0, PCAA_ERR_SHAPE, "colstats: C=%d must be a positive multiple of 8", C);
    ColGeom g = col_geom(R, C);
    size_t smem = 256 * 16 * sizeof(float);
    if (dtype == PCAA_BF16)
        colstats_kernel<__nv_bfloat16><<<g.grid, g.block, smem, ST(stream)>>>((const __nv_bfloat16*)y, R, C, stats, g.cgs, g.lanes);
    else
        colstats_kernel<float><<<g.grid, g.block, smem, ST(stream)>>>((const float*)y, R, C, stats, g.cgs, g.lanes);
    return check_launch("colstats");
}

int pcaa_bn_finalize(const double* stats, int64_t R, int C, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, float* scale, float* shift, float* mean,
                     float* invstd, pcaa_stream stream) {
    PCAA_REQUIRE(R > 0 && C > 0, PCAA_ERR_SHAPE, "bn_finalize: empty");
    double unbias = R > 1 ? (double)R / (double)(R - 1) : 1.0;
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, ST(stream)>>>(stats, 1.0 / (double)R, unbias, C, gamma, beta,
                                                               running_mean, running_var, momentum, eps, scale, shift,
                                                               mean, invstd);
    return check_launch("bn_finalize");
}

int pcaa_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, float* scale, float* shift, int C, pcaa_stream stream) {
    bn_eval_coeffs_kernel<<<ceil_div(C, 128), 128, 0, ST(stream)>>>(gamma, beta, running_mean, running_var, eps, scale,
                                                                  shift, C);
    return check_launch("bn_eval_coeffs");
}

static int ew_grid(int64_t nchunks) {
    long long b = (nchunks + 255) / 256;
    long long cap = 148LL * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int pcaa_bn_elu_apply(const void* y, int y_dtype, const float* scale, const float* shift, void* out, int out_dtype,
                      int64_t R, int C, pcaa_stream stream) {
    PCAA_REQUIRE(C % 8 == 0 && C > 0, PCAA_ERR_SHAPE, "bn_elu_apply: C=%d must be a multiple of 8", C);
    int64_t nchunks = R * (C / 8);
    int ncg = C / 8;
    int grid = ew_grid(nchunks);
    int hoist = ((int64_t)grid * 256) % ncg == 0;
#define LAUNCH(TI, TO) bn_elu_apply_kernel<TI, TO><<<grid, 256, 0, ST(stream)>>>((const TI*)y, scale, shift, (TO*)out, nchunks, ncg, hoist)
    if (y_dtype == PCAA_BF16 && out_dtype == PCAA_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else if (y_dtype == PCAA_BF16 && out_dtype == PCAA_F32) LAUNCH(__nv_bfloat16, float);
    else if (y_dtype == PCAA_F32 && out_dtype == PCAA_F32) LAUNCH(float, float);
    else LAUNCH(float, __nv_bfloat16);
#undef LAUNCH
    return check_launch("bn_elu_apply");
}

int pcaa_bn_elu_meanpool(const void* y, int y_dtype, const float* scale, const float* shift, float* pooled, int64_t G,
                         int n, int C, pcaa_stream stream) {
    PCAA_REQUIRE(C % 8 == 0 && C > 0 && n > 0, PCAA_ERR_SHAPE, "bn_elu_meanpool: bad shape C=%d n=%d", C, n);
    constexpr int RL = 4;
    int ncg = C / 8;
    int cgs = ncg < 64 ? ncg : 64;
    dim3 grid((unsigned)G, (ncg + cgs - 1) / cgs);
    size_t smem = (size_t)RL * cgs * 8 * sizeof(float);
    if (y_dtype == PCAA_BF16)
        bn_elu_meanpool_kernel<__nv_bfloat16, RL><<<grid, cgs * RL, smem, ST(stream)>>>((const __nv_bfloat16*)y, scale, shift, pooled, n, C, cgs);
    else
        bn_elu_meanpool_kernel<float, RL><<<grid, cgs * RL, smem, ST(stream)>>>((const float*)y, scale, shift, pooled, n, C, cgs);
    return check_launch("bn_elu_meanpool");
}

int pcaa_elu_bwd_colstats(const void* dout, int dout_dtype, int pooled_n, const void* y, int y_dtype,
                          const float* scale, const float* shift, const float* mean, const float* invstd, void* dz,
                          int dz_dtype, double* stats2, int64_t R, int C, pcaa_stream stream) {
    PCAA_REQUIRE(C % 8 == 0 && C > 0 && R > 0, PCAA_ERR_SHAPE, "elu_bwd_colstats: C=%d must be a multiple of 8", C);
    PCAA_REQUIRE(pooled_n == 0 || R % pooled_n == 0, PCAA_ERR_SHAPE, "elu_bwd_colstats: R %% pooled_n != 0");
    ColGeom g = col_geom(R, C);
    size_t smem = 256 * 16 * sizeof(float);
    const float* dpool = pooled_n > 0 ? (const float*)dout : nullptr;
#define LAUNCH(TD, TY, TZ) elu_bwd_colstats_kernel<TD, TY, TZ><<<g.grid, g.block, smem, ST(stream)>>>((const TD*)dout, dpool, pooled_n, (const TY*)y, scale, shift, mean, invstd, (TZ*)dz, stats2, R, C, g.cgs, g.lanes)
    if (y_dtype == PCAA_BF16 && dz_dtype == PCAA_BF16) {
        if (pooled_n > 0 || dout_dtype == PCAA_F32) LAUNCH(float, __nv_bfloat16, __nv_bfloat16);
        else LAUNCH(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
    } else if (y_dtype == PCAA_F32 && dz_dtype == PCAA_F32 && (pooled_n > 0 || dout_dtype == PCAA_F32)) {
        LAUNCH(float, float, float);
    } else {
        set_error("elu_bwd_colstats: unsupported dtype combination (%d,%d,%d)", dout_dtype, y_dtype, dz_dtype);
        return PCAA_ERR_UNSUPPORTED;
    }
#undef LAUNCH
    return check_launch("elu_bwd_colstats");
}

int pcaa_bn_bwd_finalize(const double* stats2, int64_t R, int C, const float* scale, const float* mean,
                         const float* invstd, float* c1, float* c2, float* c3, float* dgamma, float* dbeta,
                         pcaa_stream stream) {
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, ST(stream)>>>(stats2, 1.0 / (double)R, C, scale, mean, invstd, c1,
                                                                   c2, c3, dgamma, dbeta);
    return check_launch("bn_bwd_finalize");
}

int pcaa_bn_bwd_apply(const void* dz, int dz_dtype, const void* y, int y_dtype, const float* c1, const float* c2,
                      const float* c3, void* dy, int dy_dtype, int64_t R, int C, pcaa_stream stream) {
    PCAA_REQUIRE(C % 8 == 0 && C > 0, PCAA_ERR_SHAPE, "bn_bwd_apply: C=%d must be a multiple of 8", C);
    int64_t nchunks = R * (C / 8);
    int ncg = C / 8;
    int grid = ew_grid(nchunks);
    int hoist = ((int64_t)grid * 256) % ncg == 0;
    if (dz_dtype == PCAA_BF16 && y_dtype == PCAA_BF16 && dy_dtype == PCAA_BF16)
        bn_bwd_apply_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, ST(stream)>>>(
            (const __nv_bfloat16*)dz, (const __nv_bfloat16*)y, c1, c2, c3, (__nv_bfloat16*)dy, nchunks, ncg, hoist);
    else if (dz_dtype == PCAA_F32 && y_dtype == PCAA_F32 && dy_dtype == PCAA_F32)
        bn_bwd_apply_kernel<float, float, float><<<grid, 256, 0, ST(stream)>>>((const float*)dz, (const float*)y, c1, c2, c3,
                                                                            (float*)dy, nchunks, ncg, hoist);
    else if (dz_dtype == PCAA_F32 && y_dtype == PCAA_F32 && dy_dtype == PCAA_BF16)
        bn_bwd_apply_kernel<float, float, __nv_bfloat16><<<grid, 256, 0, ST(stream)>>>((const float*)dz, (const float*)y, c1, c2, c3,
                                                                                    (__nv_bfloat16*)dy, nchunks, ncg, hoist);
    else {
        set_error("bn_bwd_apply: unsupported dtype combination");
        return PCAA_ERR_UNSUPPORTED;
    }
    return check_launch("bn_bwd_apply");
}

int pcaa_elu_bwd_from_out(const float* dout, const float* out, float* dz, int64_t n, pcaa_stream stream) {
    if (n == 0) return PCAA_OK;
    elu_bwd_from_out_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(dout, out, dz, n);
    return check_launch("elu_bwd_from_out");
}

int pcaa_colsum(const float* x, int64_t R, int C, float* out, pcaa_stream stream) {
    colsum_kernel<float><<<ceil_div(C, 32), dim3(32, 8), 0, ST(stream)>>>(x, R, C, C, out);
    return check_launch("colsum");
}

int pcaa_colsum_ld(const void* x, int dtype, int64_t R, int C, int64_t ld, float* out, pcaa_stream stream) {
    if (dtype == PCAA_BF16)
        colsum_kernel<__nv_bfloat16><<<ceil_div(C, 32), dim3(32, 8), 0, ST(stream)>>>((const __nv_bfloat16*)x, R, C, ld, out);
    else
        colsum_kernel<float><<<ceil_div(C, 32), dim3(32, 8), 0, ST(stream)>>>((const float*)x, R, C, ld, out);
    return check_launch("colsum_ld");
}

int pcaa_convert(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, pcaa_stream stream) {
    if (n == 0) return PCAA_OK;
    int grid = ew_grid(n);
    if (in_dtype == PCAA_F32 && out_dtype == PCAA_BF16)
        convert_kernel<float, __nv_bfloat16><<<grid, 256, 0, ST(stream)>>>((const float*)in, (__nv_bfloat16*)out, n);
    else if (in_dtype == PCAA_BF16 && out_dtype == PCAA_F32)
        convert_kernel<__nv_bfloat16, float><<<grid, 256, 0, ST(stream)>>>((const __nv_bfloat16*)in, (float*)out, n);
    else {
        set_error("convert: unsupported dtype pair");
        return PCAA_ERR_UNSUPPORTED;
    }
    return check_launch("convert");
}

int pcaa_pack_bf16(const float* in, int64_t R, int64_t C, int64_t ld_in, void* out, int64_t ld_out, int transpose,
                   pcaa_stream stream) {
    int64_t out_rows = transpose ? C : R;
    int64_t out_cols = transpose ? R : C;
    PCAA_REQUIRE(ld_out >= out_cols, PCAA_ERR_SHAPE, "pack_bf16: ld_out too small");
    dim3 grid(ceil_div(ld_out, 32), ceil_div(out_rows, 32));
    pack_bf16_kernel<<<grid, dim3(32, 8), 0, ST(stream)>>>(in, R, C, ld_in, (__nv_bfloat16*)out, ld_out, transpose, out_rows);
    return check_launch("pack_bf16");
}

int pcaa_tcn_im2col(const float* x, void* col, int col_dtype, int64_t B, int T, int Cin, int dil, pcaa_stream stream) {
    int64_t total = B * T * Cin * 3;
    if (total == 0) return PCAA_OK;
    if (col_dtype == PCAA_BF16 && Cin % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)col & 15) == 0)
        tcn_im2col_bf16x8_kernel<<<ew_grid(B * T * (Cin / 8)), 256, 0, ST(stream)>>>(x, (__nv_bfloat16*)col, B * T, T, Cin, dil);
    else if (col_dtype == PCAA_BF16)
        tcn_im2col_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, ST(stream)>>>(x, (__nv_bfloat16*)col, B * T, T, Cin, dil);
    else
        tcn_im2col_kernel<float><<<ew_grid(total), 256, 0, ST(stream)>>>(x, (float*)col, B * T, T, Cin, dil);
    return check_launch("tcn_im2col");
}

int pcaa_tcn_col2im(const float* dcol, float* dx, int64_t B, int T, int Cin, int dil, pcaa_stream stream) {
    int64_t total = B * T * Cin;
    tcn_col2im_kernel<<<ew_grid(total), 256, 0, ST(stream)>>>(dcol, dx, B * T, T, Cin, dil);
    return check_launch("tcn_col2im");
}

int pcaa_tcn_bn_elu_next(const float* y, const double* stats, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, const float* scale,
                         const float* shift, float* coef_out, int64_t B, int T, int C, int dil_next, void* col, float* act,
                         pcaa_stream stream) {
    const int64_t R = B * T;
    if (R == 0) return PCAA_OK;
    PCAA_REQUIRE(C > 0 && C <= 4096, PCAA_ERR_SHAPE, "tcn_bn_elu_next: C=%d out of range", C);
    PCAA_REQUIRE((stats != nullptr) != (scale != nullptr && shift != nullptr), PCAA_ERR_SHAPE,
                 "tcn_bn_elu_next: give either the batch statistics (training) or scale / shift (eval)");
    PCAA_REQUIRE(stats == nullptr || (gamma && beta && coef_out), PCAA_ERR_SHAPE, "tcn_bn_elu_next: training needs gamma, beta, coef_out");
    PCAA_REQUIRE(col != nullptr || act != nullptr, PCAA_ERR_SHAPE, "tcn_bn_elu_next: no output requested");
    PCAA_REQUIRE(col == nullptr || dil_next > 0, PCAA_ERR_SHAPE, "tcn_bn_elu_next: the im2col output needs the next layer's dilation");
    const double unbias = R > 1 ? (double)R / (double)(R - 1) : 1.0;
    int grid = ew_grid(C % 8 == 0 ? R * (C / 8) : (R * C + 3) / 4);
    tcn_bn_elu_next_kernel<<<grid, 256, 2 * C * sizeof(float), ST(stream)>>>(y, stats, 1.0 / (double)R, unbias, gamma, beta,
                                                                           running_mean, running_var, momentum, eps, scale,
                                                                           shift, coef_out, R, T, C, dil_next,
                                                                           (__nv_bfloat16*)col, act);
    return check_launch("tcn_bn_elu_next");
}

int pcaa_tcn_elu_bwd_stats(const float* src, int src_mode, int dil_up, const float* y, const float* scale,
                           const float* shift, const float* mean, const float* invstd, float* dz, double* stats2,
                           int64_t B, int T, int C, pcaa_stream stream) {
    const int64_t R = B * T;
    PCAA_REQUIRE(C % 8 == 0 && C > 0 && R > 0, PCAA_ERR_SHAPE, "tcn_elu_bwd_stats: C=%d must be a multiple of 8", C);
    PCAA_REQUIRE(src_mode >= 0 && src_mode <= 2 && (src_mode != 1 || dil_up > 0), PCAA_ERR_SHAPE, "tcn_elu_bwd_stats: bad source mode");
    ColGeom g = col_geom(R, C);
    size_t smem = 256 * 16 * sizeof(float);
    tcn_elu_bwd_stats_kernel<<<g.grid, g.block, smem, ST(stream)>>>(src, src_mode, dil_up, T, y, scale, shift, mean, invstd, dz,
                                                                  stats2, R, C, g.cgs, g.lanes);
    return check_launch("tcn_elu_bwd_stats");
}

int pcaa_tcn_bn_bwd_apply(const float* dz, const float* y, const double* stats2, const float* scale, const float* mean,
                          const float* invstd, float* dgamma, float* dbeta, void* dy, int64_t R, int C,
                          pcaa_stream stream) {
    PCAA_REQUIRE(C % 8 == 0 && C > 0 && C <= 4096 && R > 0, PCAA_ERR_SHAPE, "tcn_bn_bwd_apply: C=%d must be a multiple of 8", C);
    int grid = ew_grid(R * (C / 8));
    tcn_bn_bwd_apply_kernel<<<grid, 256, 3 * C * sizeof(float), ST(stream)>>>(dz, y, stats2, 1.0 / (double)R, scale, mean, invstd,
                                                                            dgamma, dbeta, (__nv_bfloat16*)dy, R, C);
    return check_launch("tcn_bn_bwd_apply");
}

int pcaa_mean_rows(const float* x, float* out, int64_t G, int n, int C, pcaa_stream stream) {
    mean_rows_kernel<<<ceil_div(G * C, 256), 256, 0, ST(stream)>>>(x, out, G, n, C);
    return check_launch("mean_rows");
}

int pcaa_mean_rows_bwd(const float* g, float* dx, int64_t G, int n, int C, pcaa_stream stream) {
    mean_rows_bwd_kernel<<<ew_grid(G * n * C), 256, 0, ST(stream)>>>(g, dx, G, n, C);
    return check_launch("mean_rows_bwd");
}

int pcaa_softmax_ce(const float* logits, const int64_t* gt, float* loss, float* dlogits, float gscale, int32_t* pred,
                    int64_t B, int C, pcaa_stream stream) {
    if (loss) cudaMemsetAsync(loss, 0, sizeof(float), ST(stream));
    softmax_ce_kernel<<<ceil_div(B * 32, 256), 256, 0, ST(stream)>>>(logits, gt, loss, dlogits, gscale, pred, B, C);
    return check_launch("softmax_ce");
}

int pcaa_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int step, float grad_scale, void* shadow_bf16, pcaa_stream stream) {
    if (n == 0) return PCAA_OK;
    PCAA_REQUIRE(step >= 1, PCAA_ERR_SHAPE, "adam: step must be >= 1");
    PCAA_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, PCAA_ERR_ALIGN,
                 "adam: buffers must be 16-byte aligned");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    int64_t nthreads = (n + 3) / 4;
    adam_flat_kernel<<<ew_grid(nthreads), 256, 0, ST(stream)>>>(p, g, m, v, n, (float)(lr / bc1), beta1, beta2, eps,
                                                             (float)(1.0 / sqrt(bc2)), grad_scale,
                                                             (__nv_bfloat16*)shadow_bf16, nullptr);
    return check_launch("adam_flat");
}

int pcaa_sum_into(float* dst, const float* src, int64_t n, int64_t src_stride, int nsrc, pcaa_stream stream) {
    if (n == 0 || nsrc == 0) return PCAA_OK;
    PCAA_REQUIRE(n % 4 == 0 && src_stride % 4 == 0 && ((uintptr_t)dst | (uintptr_t)src) % 16 == 0 && nsrc > 0, PCAA_ERR_ALIGN,
                 "sum_into: n and the source stride must be multiples of 4 floats, buffers 16-byte aligned");
    sum_into_kernel<<<ew_grid(n / 4), 256, 0, ST(stream)>>>(dst, src, n / 4, src_stride, nsrc, 0);
    return check_launch("sum_into");
}

int pcaa_sum_rows(float* dst, const float* src, int64_t n, int64_t src_stride, int nsrc, pcaa_stream stream) {
    if (n == 0) return PCAA_OK;
    PCAA_REQUIRE(n % 4 == 0 && src_stride % 4 == 0 && ((uintptr_t)dst | (uintptr_t)src) % 16 == 0 && nsrc > 0, PCAA_ERR_ALIGN,
                 "sum_rows: n and the source stride must be multiples of 4 floats, buffers 16-byte aligned");
    sum_into_kernel<<<ew_grid(n / 4), 256, 0, ST(stream)>>>(dst, src, n / 4, src_stride, nsrc, 1);
    return check_launch("sum_rows");
}

int pcaa_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n_idx, int64_t row_bytes, int64_t n_src,
                     pcaa_stream stream) {
    if (n_idx == 0 || row_bytes == 0) return PCAA_OK;
    PCAA_REQUIRE(row_bytes % 16 == 0 && ((uintptr_t)src | (uintptr_t)dst) % 16 == 0, PCAA_ERR_ALIGN,
                 "gather_rows: rows must be multiples of 16 bytes and 16-byte aligned (row_bytes=%lld)", (long long)row_bytes);
    PCAA_REQUIRE(n_idx <= 0x7fffffffLL && n_src >= 0, PCAA_ERR_SHAPE, "gather_rows: bad row counts");
    const int64_t row_vec = row_bytes / 16;
    int splits = (int)((row_vec + 1023) / 1024);
    if (splits > 64) splits = 64;
    gather_rows_kernel<<<dim3((unsigned)n_idx, (unsigned)splits), 256, 0, ST(stream)>>>((const uint4*)src, idx, (uint4*)dst, row_vec, n_src);
    return check_launch("gather_rows");
}

int pcaa_adam_advance(int32_t* step_dev, float* coef_dev, float lr, float beta1, float beta2, pcaa_stream stream) {
    PCAA_REQUIRE(step_dev && coef_dev, PCAA_ERR_SHAPE, "adam_advance: null device pointers");
    adam_advance_kernel<<<1, 1, 0, ST(stream)>>>(step_dev, coef_dev, lr, beta1, beta2);
    return check_launch("adam_advance");
}

int pcaa_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps,
                       const float* coef_dev, float grad_scale, void* shadow_bf16, pcaa_stream stream) {
    if (n == 0) return PCAA_OK;
    PCAA_REQUIRE(coef_dev != nullptr, PCAA_ERR_SHAPE, "adam_flat_dev: null coefficient pointer");
    PCAA_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, PCAA_ERR_ALIGN,
                 "adam: buffers must be 16-byte aligned");
    int64_t nthreads = (n + 3) / 4;
    adam_flat_kernel<<<ew_grid(nthreads), 256, 0, ST(stream)>>>(p, g, m, v, n, 0.f, beta1, beta2, eps, 0.f, grad_scale,
                                                             (__nv_bfloat16*)shadow_bf16, coef_dev);
    return check_launch("adam_flat_dev");
}

int pcaa_ew(int op, const float* a, const float* b, float* out, int64_t n, int ncols, pcaa_stream stream) {
    if (n == 0) return PCAA_OK;
    PCAA_REQUIRE(op >= 0 && op <= PCAA_EW_ADD_ROWVEC, PCAA_ERR_UNSUPPORTED, "ew: unknown op %d", op);
    ew_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(op, a, b, out, n, ncols > 0 ? ncols : 1);
    return check_launch("ew");
}

int pcaa_pointnet_l1_fwd(const float* x, const float* w, const float* bias, void* y, double* stats, int64_t B,
                         int64_t TN, int Cout, pcaa_stream stream) {
    PCAA_REQUIRE(Cout % 8 == 0 && 256 % (Cout / 8) == 0, PCAA_ERR_SHAPE, "pointnet_l1_fwd: Cout=%d unsupported", Cout);
    constexpr int PL = 4;
    PCAA_REQUIRE(Cout / 8 * PL == 256, PCAA_ERR_SHAPE, "pointnet_l1_fwd: Cout must be 512");
    int64_t P = B * TN;
    int ppb = 128;
    size_t smem = 256 * 16 * sizeof(float);
    pointnet_l1_fwd_kernel<PL><<<ceil_div(P, ppb), 256, smem, ST(stream)>>>(x, w, bias, (__nv_bfloat16*)y, stats, P, TN,
                                                                         Cout, ppb);
    return check_launch("pointnet_l1_fwd");
}

int pcaa_pointnet_l1_wgrad(const float* x, const void* dy, float* dW, int64_t B, int64_t TN, int Cout,
                           pcaa_stream stream) {
    constexpr int PL = 4;
    PCAA_REQUIRE(Cout / 8 * PL == 256, PCAA_ERR_SHAPE, "pointnet_l1_wgrad: Cout must be 512");
    int64_t P = B * TN;
    int ppb = 128;
    cudaMemsetAsync(dW, 0, sizeof(float) * Cout * 4, ST(stream));
    int blocks = ceil_div(P, ppb);
    if (blocks > 148 * 4) blocks = 148 * 4;
    size_t smem = 256 * 32 * sizeof(float);
    pointnet_l1_wgrad_kernel<PL><<<blocks, 256, smem, ST(stream)>>>(x, (const __nv_bfloat16*)dy, dW, P, TN, Cout, ppb);
    return check_launch("pointnet_l1_wgrad");
}

}  // extern "C"
