// Generic CUDA-core GEMM with arbitrary element strides (fp32 accumulate).  Used for the small layers of the
// path (TCN convolutions as im2col GEMMs, encoder heads, decoder before the tensor-core path takes over) and as
// the on-device comparator of the tcgen05 kernels in the GPU tests.
#include "common.cuh"

namespace pcaa {

constexpr int TM = 64, TN_ = 64, TK = 16;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int64_t sam, int64_t sak, const TB* __restrict__ B, int64_t sbk, int64_t sbn,
                 TC* __restrict__ C, int64_t scm, int64_t scn, int64_t M, int64_t N, int64_t K,
                 const float* __restrict__ bias, int act, int accumulate, int64_t k_per_split) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN_ + 4];
    int tid = threadIdx.x;
    int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each 4 x 4 outputs
    int64_t m0 = (int64_t)blockIdx.x * TM, n0 = (int64_t)blockIdx.y * TN_;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_kfast = (sak == 1);
    const bool b_nfast = (sbn == 1);
    // split-K (gridDim.z > 1): this block reduces k in [kbeg, kend) and atomically adds into a zeroed fp32 C
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
    const bool split = gridDim.z > 1;
    // register-staged software pipeline: the global loads of k block i+1 are issued before the FMAs of block i (these
    // GEMMs are small and latency bound: few CTAs, nothing else to hide a load behind)
    float ra[4], rb[4];
    auto load_tiles = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m, k;
            if (a_kfast) { k = tid % TK; m = tid / TK + 16 * i; }
            else { m = tid % TM; k = tid / TM + 4 * i; }
            int64_t gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < kend) ? ld_as_float<TA>(A + gm * sam + gk * sak) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n, k;
            if (b_nfast) { n = tid % TN_; k = tid / TN_ + 4 * i; }
            else { k = tid % TK; n = tid / TK + 16 * i; }
            int64_t gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < N && gk < kend) ? ld_as_float<TB>(B + gk * sbk + gn * sbn) : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m, k;
            if (a_kfast) { k = tid % TK; m = tid / TK + 16 * i; }
            else { m = tid % TM; k = tid / TM + 4 * i; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n, k;
            if (b_nfast) { n = tid % TN_; k = tid / TN_ + 4 * i; }
            else { k = tid % TK; n = tid / TK + 16 * i; }
            Bs[k][n] = rb[i];
        }
    };
    if (kbeg < kend) load_tiles(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
        store_tiles();
        __syncthreads();
        if (k0 + TK < kend) load_tiles(k0 + TK);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            TC* c = C + gm * scm + gn * scn;
            if (split) {
                if (bias && blockIdx.z == 0) v += bias[gn];
                if constexpr (sizeof(TC) == 4) atomicAdd(reinterpret_cast<float*>(c), v);
                continue;
            }
            if (bias) v += bias[gn];
            if (act == PCAA_ACT_ELU) v = elu_f(v);
            if (accumulate) v += ld_as_float<TC>(c);
            st_from_float<TC>(c, v);
        }
    }
}

template <typename TA, typename TB, typename TC>
static int launch(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbk, int64_t sbn, void* C, int64_t scm,
                  int64_t scn, int64_t M, int64_t N, int64_t K, const float* bias, int act, int accumulate,
                  cudaStream_t st) {
    dim3 grid(ceil_div(M, TM), ceil_div(N, TN_));
    int64_t k_per_split = K > 0 ? K : 1;
    // tall-K products with few output tiles (weight gradients of the TCN / heads: K = rows) would leave most SMs idle:
    // split K over grid.z and reduce with fp32 atomics into a zeroed, contiguous fp32 C
    const int64_t tiles = (int64_t)grid.x * grid.y;
    if (sizeof(TC) == 4 && act == PCAA_ACT_NONE && !accumulate && scn == 1 && scm == N && K >= 128 && tiles < 2 * 148) {
        int splits = (int)((4 * 148 + tiles - 1) / tiles);
        int max_splits = (int)(K / 32);
        if (splits > max_splits) splits = max_splits;
        if (splits > 1) {
            k_per_split = ((K + splits - 1) / splits + TK - 1) / TK * TK;
            grid.z = (unsigned)ceil_div(K, k_per_split);
            if (cudaMemsetAsync(C, 0, sizeof(float) * M * N, st) != cudaSuccess) return check_launch("gemm_simt memset");
        }
    }
    gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, st>>>((const TA*)A, sam, sak, (const TB*)B, sbk, sbn, (TC*)C, scm, scn, M,
                                                      N, K, bias, act, accumulate, k_per_split);
    return check_launch("gemm_simt");
}

}  // namespace pcaa

using namespace pcaa;
typedef __nv_bfloat16 bf16;

extern "C" int pcaa_gemm_simt(const void* A, int a_dtype, int64_t sam, int64_t sak, const void* B, int b_dtype,
                              int64_t sbk, int64_t sbn, void* C, int c_dtype, int64_t scm, int64_t scn, int64_t M,
                              int64_t N, int64_t K, const float* bias, int act, int accumulate, pcaa_stream stream) {
    if (M == 0 || N == 0) return PCAA_OK;
    PCAA_REQUIRE(M > 0 && N > 0 && K >= 0, PCAA_ERR_SHAPE, "gemm_simt: negative dimension");
    PCAA_REQUIRE(ceil_div(N, TN_) <= 65535, PCAA_ERR_SHAPE, "gemm_simt: N=%lld too large for grid.y", (long long)N);
    cudaStream_t st = (cudaStream_t)stream;
    int key = a_dtype * 4 + b_dtype * 2 + c_dtype;
#define GO(TA, TB, TC) return launch<TA, TB, TC>(A, sam, sak, B, sbk, sbn, C, scm, scn, M, N, K, bias, act, accumulate, st)
    switch (key) {
        case 0: GO(float, float, float);
        case 1: GO(float, float, bf16);
        case 2: GO(float, bf16, float);
        case 3: GO(float, bf16, bf16);
        case 4: GO(bf16, float, float);
        case 5: GO(bf16, float, bf16);
        case 6: GO(bf16, bf16, float);
        case 7: GO(bf16, bf16, bf16);
    }
#undef GO
    set_error("gemm_simt: bad dtype");
    return PCAA_ERR_UNSUPPORTED;
}
