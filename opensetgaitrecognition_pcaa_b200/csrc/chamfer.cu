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
