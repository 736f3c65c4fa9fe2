// Open-set scoring of the embeddings (reference inference_PCAA.py:129-136, 255-271).
// The reference evaluates (1/C) sum_c N(x; mu_c, I_32) with scipy in float64 and thresholds it; values are ~1e-21
// and underflow for far samples, so the kernel works in the log domain (exactly monotone restatement):
//   loglik = -D/2 ln(2 pi) - ln C + logsumexp_c( -1/2 |x - mu_c|^2 )        (float64)
// and the vote compares against ln(threshold).
#include "common.cuh"

namespace pcaa {

constexpr int SC_MAXC = 64;

// one warp per embedding: lanes split the D coordinates, prototypes are staged in shared memory
__global__ void __launch_bounds__(256)
openset_score_kernel(const float* __restrict__ emb, const float* __restrict__ means, int64_t M, int C, int D,
                     double* __restrict__ loglik) {
    extern __shared__ float mu[];   // [C][D]
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) mu[i] = means[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double cst = -0.5 * (double)D * 1.8378770664093453 /* ln(2 pi) */ - log((double)C);
    for (int64_t s = warp0; s < M; s += nwarps) {
        double e[SC_MAXC];
        double mx = -INFINITY;
        for (int c = 0; c < C; ++c) {
            double acc = 0.0;
            for (int d = lane; d < D; d += 32) {
                double df = (double)emb[s * D + d] - (double)mu[c * D + d];
                acc += df * df;
            }
            acc = warp_sum_d(acc);
            e[c] = -0.5 * acc;
            mx = fmax(mx, e[c]);
        }
        if (lane == 0) {
            double se = 0.0;
            for (int c = 0; c < C; ++c) se += exp(e[c] - mx);
            loglik[s] = cst + mx + log(se);
        }
    }
}

// one thread per window of k consecutive samples
__global__ void openset_vote_kernel(const double* __restrict__ loglik, const int32_t* __restrict__ pred,
                                    int64_t n_windows, int k, double log_thr, int n_labels, int32_t* __restrict__ out) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    int above = 0;
    for (int i = 0; i < k; ++i) above += loglik[w * k + i] > log_thr ? 1 : 0;
    int label = n_labels;
    if (2 * above > k) {
        // argmax(bincount(pred)): most frequent class, lowest class on ties (inference_PCAA.py:265-266)
        int best_cnt = 0;
        label = 0;
        for (int i = 0; i < k; ++i) {
            int ci = pred[w * k + i];
            int cnt = 0;
            for (int j = 0; j < k; ++j) cnt += pred[w * k + j] == ci ? 1 : 0;
            if (cnt > best_cnt || (cnt == best_cnt && ci < label)) { best_cnt = cnt; label = ci; }
        }
    }
    out[w] = label;
}

}  // namespace pcaa

using namespace pcaa;

extern "C" int pcaa_openset_score(const float* emb, const float* means, int64_t M, int C, int D, double* loglik,
                                  pcaa_stream stream) {
    PCAA_REQUIRE(C >= 1 && C <= SC_MAXC, PCAA_ERR_SHAPE, "openset_score: C=%d unsupported (1..%d)", C, SC_MAXC);
    PCAA_REQUIRE(D >= 1 && D <= 1024, PCAA_ERR_SHAPE, "openset_score: D=%d unsupported", D);
    if (M == 0) return PCAA_OK;
    long long blocks = (M * 32 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    openset_score_kernel<<<(unsigned)blocks, 256, (size_t)C * D * sizeof(float), (cudaStream_t)stream>>>(emb, means, M, C, D,
                                                                                                   loglik);
    return check_launch("openset_score");
}

extern "C" int pcaa_openset_vote(const double* loglik, const int32_t* pred, int64_t n_windows, int k, double log_thr,
                                 int n_labels, int32_t* out, pcaa_stream stream) {
    PCAA_REQUIRE(k >= 1 && k <= 1024, PCAA_ERR_SHAPE, "openset_vote: k=%d unsupported", k);
    if (n_windows == 0) return PCAA_OK;
    openset_vote_kernel<<<ceil_div(n_windows, 128), 128, 0, (cudaStream_t)stream>>>(loglik, pred, n_windows, k, log_thr,
                                                                                n_labels, out);
    return check_launch("openset_vote");
}
