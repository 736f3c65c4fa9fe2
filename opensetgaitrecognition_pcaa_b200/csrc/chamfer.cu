// SeqChamferLoss (reference utils.py:98-132): one CTA per (sample, frame).  Both clouds of a frame
// (2 x F x N fp32 = 4.8 KB at N = 150) are staged in shared memory with coalesced row loads; every thread owns
// one point of one cloud and scans the other cloud (broadcast smem reads), keeping min and arg-min in registers.
// The N x N distance matrix the reference materialises four times (xx, yy, zz, P) never exists.
#include "common.cuh"

namespace pcaa {

constexpr int CH_THREADS = 160;
constexpr int MAXF = 8;

// distance in the reference's expanded form: (|gt_i|^2 + |pred_j|^2) - 2 gt_i.pred_j   (utils.py:131)
__global__ void __launch_bounds__(CH_THREADS)
chamfer_fwd_kernel(const float* __restrict__ preds, const float* __restrict__ gts, int F, int T, int N,
                   float* __restrict__ frame_loss, int32_t* __restrict__ idx_gt_for_pred,
                   int32_t* __restrict__ idx_pred_for_gt) {
    extern __shared__ float sm[];
    float* g = sm;                 // [F][N]
    float* p = g + F * N;          // [F][N]
    float* rx = p + F * N;         // |gt_i|^2
    float* ry = rx + N;            // |pred_j|^2
    __shared__ float red[CH_THREADS / 32];
    const int bt = blockIdx.x;
    const int b = bt / T, t = bt % T;
    const int64_t base = ((int64_t)b * F * T + t) * N;      // element (b, 0, t, 0)
    const int64_t fstride = (int64_t)T * N;
    for (int i = threadIdx.x; i < F * N; i += CH_THREADS) {
        int f = i / N, n = i % N;
        g[i] = gts[base + f * fstride + n];
        p[i] = preds[base + f * fstride + n];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += CH_THREADS) {
        float a = 0.f, c = 0.f;
        for (int f = 0; f < F; ++f) {
            a = fmaf(g[f * N + i], g[f * N + i], a);
            c = fmaf(p[f * N + i], p[f * N + i], c);
        }
        rx[i] = a;
        ry[i] = c;
    }
    __syncthreads();
    float total = 0.f;
    // for every predicted point j: nearest ground-truth point  (torch.min(P, 2), utils.py:100)
    for (int j = threadIdx.x; j < N; j += CH_THREADS) {
        float pj[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) pj[f] = f < F ? p[f * N + j] : 0.f;
        float ryj = ry[j];
        float best = INFINITY;
        int bi = 0;
        for (int i = 0; i < N; ++i) {
            float zz = 0.f;
#pragma unroll
            for (int f = 0; f < MAXF; ++f)
                if (f < F) zz = fmaf(g[f * N + i], pj[f], zz);
            float d = (rx[i] + ryj) - 2.f * zz;
            if (d < best) { best = d; bi = i; }
        }
        total += best;
        if (idx_gt_for_pred) idx_gt_for_pred[(int64_t)bt * N + j] = bi;
    }
    // for every ground-truth point i: nearest predicted point  (torch.min(P, 3), utils.py:102)
    for (int i = threadIdx.x; i < N; i += CH_THREADS) {
        float gi[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) gi[f] = f < F ? g[f * N + i] : 0.f;
        float rxi = rx[i];
        float best = INFINITY;
        int bj = 0;
        for (int j = 0; j < N; ++j) {
            float zz = 0.f;
#pragma unroll
            for (int f = 0; f < MAXF; ++f)
                if (f < F) zz = fmaf(gi[f], p[f * N + j], zz);
            float d = (rxi + ry[j]) - 2.f * zz;
            if (d < best) { best = d; bj = j; }
        }
        total += best;
        if (idx_pred_for_gt) idx_pred_for_gt[(int64_t)bt * N + i] = bj;
    }
    total = warp_sum(total);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < CH_THREADS / 32; ++w) s += red[w];
        frame_loss[bt] = s;
    }
}

// ---------------------------------------------------------------------------------------------- F = 4 fast path
// The radar features are (x, y, z, doppler): F = 4 always in the reference (constants.py:29-33).  Same arithmetic as
// chamfer_fwd_kernel bit for bit (zz = fma chain over f = 0..3 from 0, d = (r_other + r_own) - 2 zz, strict `<` scan in
// index order => lowest index on ties), organised for the FP32 issue limit:
//   * packed fp32x2 FMAs (FFMA2, sm_100): two candidate points per instruction with the own point's feature as the
//     broadcast operand;
//   * a thread owns TWO points of its cloud, so one 16-byte shared-memory read of four candidates feeds 8 pairs;
//   * the running arg-min is kept per block of four candidates (3 FMNMX + compare + 2 selects per 4 pairs instead of
//     3 instructions per pair); the winning block is re-evaluated once at the end with the same arithmetic to find the
//     first index that attains the minimum;
//   * both directions run concurrently: threads [0, tp) scan the ground truth for predicted points (torch.min(P, 2),
//     utils.py:100), threads [tp, 2 tp) scan the predictions for ground-truth points (torch.min(P, 3), utils.py:102),
//     tp = ceil(N / 2); same code, the cloud pointers are per thread.
// Clouds are padded to a multiple of four candidates with r = +inf (never selected).
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

__global__ void __launch_bounds__(256)
chamfer_fwd4_kernel(const float* __restrict__ preds, const float* __restrict__ gts, int T, int N, int Np, int tp,
                    float* __restrict__ frame_loss, int32_t* __restrict__ idx_gt_for_pred,
                    int32_t* __restrict__ idx_pred_for_gt) {
    extern __shared__ __align__(16) float sm[];
    // cloud 0 = ground truth, cloud 1 = predictions: feature planes [4][Np] then squared norms [Np]
    float* const cl0 = sm;
    float* const cl1 = sm + 5 * Np;
    __shared__ float red[8];
    const int bt = blockIdx.x;
    const int b = bt / T, t = bt % T;
    const int64_t base = ((int64_t)b * 4 * T + t) * N;
    const int64_t fstride = (int64_t)T * N;
    for (int i = threadIdx.x; i < 4 * Np; i += blockDim.x) {
        const int f = i / Np, n = i - f * Np;
        const bool ok = n < N;
        cl0[i] = ok ? __ldg(gts + base + f * fstride + n) : 0.f;
        cl1[i] = ok ? __ldg(preds + base + f * fstride + n) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Np; i += blockDim.x) {
        const int c = i >= Np, n = i - c * Np;
        float* const cc = c ? cl1 : cl0;
        float a = 0.f;
#pragma unroll
        for (int f = 0; f < 4; ++f) a = fmaf(cc[f * Np + n], cc[f * Np + n], a);
        cc[4 * Np + n] = n < N ? a : INFINITY;
    }
    __syncthreads();
    float total = 0.f;
    const int tid = threadIdx.x;
    if (tid < 2 * tp) {
        const int dir = tid >= tp;                    // 0: own = prediction, other = ground truth; 1: the reverse
        const int k = tid - dir * tp;
        const float* own = dir ? cl0 : cl1;
        const float* oth = dir ? cl1 : cl0;
        const int j0 = k, j1 = min(k + tp, N - 1);
        const bool has1 = k + tp < N;
        unsigned long long w0[4], w1[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            w0[f] = pk2(own[f * Np + j0], own[f * Np + j0]);
            w1[f] = pk2(own[f * Np + j1], own[f * Np + j1]);
        }
        const float r0 = own[4 * Np + j0], r1 = own[4 * Np + j1];
        const unsigned long long rr0 = pk2(r0, r0), rr1 = pk2(r1, r1);
        const unsigned long long m2 = pk2(-2.f, -2.f), zero = pk2(0.f, 0.f);
        float best0 = INFINITY, best1 = INFINITY;
        int blk0 = 0, blk1 = 0;
        const ulonglong2* o0 = reinterpret_cast<const ulonglong2*>(oth);
        const ulonglong2* o1 = reinterpret_cast<const ulonglong2*>(oth + Np);
        const ulonglong2* o2 = reinterpret_cast<const ulonglong2*>(oth + 2 * Np);
        const ulonglong2* o3 = reinterpret_cast<const ulonglong2*>(oth + 3 * Np);
        const ulonglong2* orr = reinterpret_cast<const ulonglong2*>(oth + 4 * Np);
        const int nb = Np >> 2;
#pragma unroll 2
        for (int q = 0; q < nb; ++q) {
            const ulonglong2 a0 = o0[q], a1 = o1[q], a2 = o2[q], a3 = o3[q], ar = orr[q];
            {
                unsigned long long zl = fma2(a0.x, w0[0], zero), zh = fma2(a0.y, w0[0], zero);
                zl = fma2(a1.x, w0[1], zl); zh = fma2(a1.y, w0[1], zh);
                zl = fma2(a2.x, w0[2], zl); zh = fma2(a2.y, w0[2], zh);
                zl = fma2(a3.x, w0[3], zl); zh = fma2(a3.y, w0[3], zh);
                const unsigned long long dl = fma2(zl, m2, add2(ar.x, rr0)), dh = fma2(zh, m2, add2(ar.y, rr0));
                float d0, d1, d2, d3;
                upk2(dl, d0, d1);
                upk2(dh, d2, d3);
                const float m = fminf(fminf(d0, d1), fminf(d2, d3));
                if (m < best0) { best0 = m; blk0 = q; }
            }
            {
                unsigned long long zl = fma2(a0.x, w1[0], zero), zh = fma2(a0.y, w1[0], zero);
                zl = fma2(a1.x, w1[1], zl); zh = fma2(a1.y, w1[1], zh);
                zl = fma2(a2.x, w1[2], zl); zh = fma2(a2.y, w1[2], zh);
                zl = fma2(a3.x, w1[3], zl); zh = fma2(a3.y, w1[3], zh);
                const unsigned long long dl = fma2(zl, m2, add2(ar.x, rr1)), dh = fma2(zh, m2, add2(ar.y, rr1));
                float d0, d1, d2, d3;
                upk2(dl, d0, d1);
                upk2(dh, d2, d3);
                const float m = fminf(fminf(d0, d1), fminf(d2, d3));
                if (m < best1) { best1 = m; blk1 = q; }
            }
        }
        // first candidate of the winning block that attains the minimum (same arithmetic, scalar)
        int32_t* out_idx = dir ? idx_pred_for_gt : idx_gt_for_pred;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int j = a ? j1 : j0;
            const float rj = a ? r1 : r0, best = a ? best1 : best0;
            const int blk = a ? blk1 : blk0;
            float wj[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) wj[f] = own[f * Np + j];
            int bi = blk * 4;
#pragma unroll
            for (int e = 3; e >= 0; --e) {
                const int i = blk * 4 + e;
                float zz = 0.f;
#pragma unroll
                for (int f = 0; f < 4; ++f) zz = fmaf(oth[f * Np + i], wj[f], zz);
                const float d = fmaf(zz, -2.f, oth[4 * Np + i] + rj);
                if (d == best) bi = i;
            }
            if (a == 0 || has1) {
                total += best;
                if (out_idx) out_idx[(int64_t)bt * N + j] = bi;
            }
        }
    }
    total = warp_sum(total);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        frame_loss[bt] = s;
    }
}

// one block; avg_out: out[0] = mean over all frames, else out[b] = mean over the T frames of sample b
__global__ void chamfer_reduce_kernel(const float* __restrict__ frame_loss, int64_t B, int T, int avg_out,
                                      float* __restrict__ out) {
    __shared__ double red[8];
    if (avg_out) {
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < B * T; i += blockDim.x) s += (double)frame_loss[i];
        s = warp_sum_d(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < (int)(blockDim.x / 32); ++w) tot += red[w];
            out[0] = (float)(tot / (double)(B * T));
        }
    } else {
        for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
            double s = 0.0;
            for (int t = 0; t < T; ++t) s += (double)frame_loss[b * T + t];
            out[b] = (float)(s / (double)T);
        }
    }
}

// d loss / d pred_j = w * 2 * [ (pred_j - gt_{i*(j)}) + sum_{i : j*(i) = j} (pred_j - gt_i) ]
__global__ void __launch_bounds__(CH_THREADS)
chamfer_bwd_kernel(const float* __restrict__ preds, const float* __restrict__ gts,
                   const int32_t* __restrict__ idx_gt_for_pred, const int32_t* __restrict__ idx_pred_for_gt,
                   const float* __restrict__ gout, int avg_out, int64_t B, int F, int T, int N,
                   float* __restrict__ grad_preds) {
    extern __shared__ float sm[];
    float* g = sm;                              // [F][N]
    int* j_of_i = reinterpret_cast<int*>(g + F * N);   // [N]
    const int bt = blockIdx.x;
    const int b = bt / T, t = bt % T;
    const int64_t base = ((int64_t)b * F * T + t) * N;
    const int64_t fstride = (int64_t)T * N;
    for (int i = threadIdx.x; i < F * N; i += CH_THREADS) g[i] = gts[base + (i / N) * fstride + (i % N)];
    for (int i = threadIdx.x; i < N; i += CH_THREADS) j_of_i[i] = idx_pred_for_gt[(int64_t)bt * N + i];
    __syncthreads();
    const float w = avg_out ? gout[0] / (float)(B * T) : gout[b] / (float)T;
    for (int j = threadIdx.x; j < N; j += CH_THREADS) {
        int i1 = idx_gt_for_pred[(int64_t)bt * N + j];
        float acc[MAXF];
        float pj[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; ++f) {
            pj[f] = f < F ? preds[base + f * fstride + j] : 0.f;
            acc[f] = f < F ? pj[f] - g[f * N + i1] : 0.f;
        }
        for (int i = 0; i < N; ++i) {
            if (j_of_i[i] == j) {
#pragma unroll
                for (int f = 0; f < MAXF; ++f)
                    if (f < F) acc[f] += pj[f] - g[f * N + i];
            }
        }
#pragma unroll
        for (int f = 0; f < MAXF; ++f)
            if (f < F) grad_preds[base + f * fstride + j] = 2.f * w * acc[f];
    }
}

// P[b,t,i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  (SeqChamferLoss.batch_pairwise_dist, utils.py:109-132)
__global__ void __launch_bounds__(256)
pairwise_dist_kernel(const float* __restrict__ x, const float* __restrict__ y, int F, int T, int N,
                     float* __restrict__ P) {
    extern __shared__ float sm[];
    float* xs = sm;            // [F][N]
    float* ys = xs + F * N;    // [F][N]
    float* rx = ys + F * N;
    float* ry = rx + N;
    const int bt = blockIdx.x;
    const int b = bt / T, t = bt % T;
    const int64_t base = ((int64_t)b * F * T + t) * N;
    const int64_t fstride = (int64_t)T * N;
    for (int i = threadIdx.x; i < F * N; i += blockDim.x) {
        xs[i] = x[base + (i / N) * fstride + (i % N)];
        ys[i] = y[base + (i / N) * fstride + (i % N)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float a = 0.f, c = 0.f;
        for (int f = 0; f < F; ++f) {
            a = fmaf(xs[f * N + i], xs[f * N + i], a);
            c = fmaf(ys[f * N + i], ys[f * N + i], c);
        }
        rx[i] = a;
        ry[i] = c;
    }
    __syncthreads();
    float* out = P + (int64_t)bt * N * N;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        int i = e / N, j = e % N;
        float zz = 0.f;
        for (int f = 0; f < F; ++f) zz = fmaf(xs[f * N + i], ys[f * N + j], zz);
        out[e] = (rx[i] + ry[j]) - 2.f * zz;
    }
}

}  // namespace pcaa

using namespace pcaa;

extern "C" int pcaa_pairwise_dist(const float* x, const float* y, int64_t B, int F, int T, int N, float* P,
                                  pcaa_stream stream) {
    PCAA_REQUIRE(F >= 1 && F <= MAXF && N >= 1 && N <= 2048, PCAA_ERR_SHAPE, "pairwise_dist: unsupported F=%d N=%d", F, N);
    if (B * T == 0) return PCAA_OK;
    size_t smem = (size_t)(2 * F * N + 2 * N) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(pairwise_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr = true;
    }
    pairwise_dist_kernel<<<(unsigned)(B * T), 256, smem, (cudaStream_t)stream>>>(x, y, F, T, N, P);
    return check_launch("pairwise_dist");
}

extern "C" int pcaa_chamfer_fwd(const float* preds, const float* gts, int64_t B, int F, int T, int N, float* frame_loss,
                                int32_t* idx_gt_for_pred, int32_t* idx_pred_for_gt, pcaa_stream stream) {
    PCAA_REQUIRE(F >= 1 && F <= MAXF, PCAA_ERR_SHAPE, "chamfer: F=%d unsupported (1..%d)", F, MAXF);
    PCAA_REQUIRE(N >= 1 && N <= 2048, PCAA_ERR_SHAPE, "chamfer: N=%d unsupported (1..2048)", N);
    if (B * T == 0) return PCAA_OK;
    size_t smem = (size_t)(2 * F * N + 2 * N) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(chamfer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(chamfer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr = true;
    }
    if (F == 4 && N <= 256) {
        const int Np = (N + 3) & ~3, tp = (N + 1) / 2;
        const unsigned threads = (unsigned)((2 * tp + 31) & ~31);
        chamfer_fwd4_kernel<<<(unsigned)(B * T), threads, (size_t)10 * Np * sizeof(float), (cudaStream_t)stream>>>(
            preds, gts, T, N, Np, tp, frame_loss, idx_gt_for_pred, idx_pred_for_gt);
        return check_launch("chamfer_fwd");
    }
    chamfer_fwd_kernel<<<(unsigned)(B * T), CH_THREADS, smem, (cudaStream_t)stream>>>(preds, gts, F, T, N, frame_loss,
                                                                                   idx_gt_for_pred, idx_pred_for_gt);
    return check_launch("chamfer_fwd");
}

extern "C" int pcaa_chamfer_reduce(const float* frame_loss, int64_t B, int T, int avg_out, float* out,
                                   pcaa_stream stream) {
    if (B * T == 0) return PCAA_OK;
    chamfer_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(frame_loss, B, T, avg_out, out);
    return check_launch("chamfer_reduce");
}

extern "C" int pcaa_chamfer_bwd(const float* preds, const float* gts, const int32_t* idx_gt_for_pred,
                                const int32_t* idx_pred_for_gt, const float* gout, int avg_out, int64_t B, int F, int T,
                                int N, float* grad_preds, pcaa_stream stream) {
    PCAA_REQUIRE(F >= 1 && F <= MAXF, PCAA_ERR_SHAPE, "chamfer: F=%d unsupported (1..%d)", F, MAXF);
    PCAA_REQUIRE(N >= 1 && N <= 2048, PCAA_ERR_SHAPE, "chamfer: N=%d unsupported (1..2048)", N);
    if (B * T == 0) return PCAA_OK;
    size_t smem = (size_t)(F * N) * sizeof(float) + (size_t)N * sizeof(int);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(chamfer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr = true;
    }
    chamfer_bwd_kernel<<<(unsigned)(B * T), CH_THREADS, smem, (cudaStream_t)stream>>>(
        preds, gts, idx_gt_for_pred, idx_pred_for_gt, gout, avg_out, B, F, T, N, grad_preds);
    return check_launch("chamfer_bwd");
}
