// Conditional critic (CGDiscriminator, reference models.py:405-421) and the fused analytic WGAN-GP critic step
// (reference PCAA_ablation.py:905-973): three critic evaluations (real / fake / interpolate), the input gradient of
// the interpolate evaluation, the gradient penalty and the gradient of the whole loss w.r.t. every critic weight
// (double backward through Linear/ELU in closed form: ELU' = e^x (x<=0) | 1, ELU'' = e^x (x<0) | 0) in ONE kernel.
#include "common.cuh"

namespace pcaa {

constexpr int H1 = 64, H2 = 32, XD = 32;   // hidden sizes, embedding size (constants.SUP_LATENT_DIM)
constexpr int NT = 128;

struct CriticSmem {
    float *W1, *b1, *W2, *b2, *w3;              // weights
    float *gW1, *gb1, *gW2, *gb2, *gw3;         // per-CTA gradient accumulators (each entry owned by one thread)
    float *u, *p1, *h1, *e1, *e1pp, *h2, *e2, *e2pp, *g2, *r, *g1, *gx, *a, *dr, *dp1, *dp2, *red;
};

__device__ __forceinline__ CriticSmem carve(float* s, int D) {
    CriticSmem m;
    m.W1 = s; s += H1 * D;
    m.b1 = s; s += H1;
    m.W2 = s; s += H2 * H1;
    m.b2 = s; s += H2;
    m.w3 = s; s += H2;
    m.gW1 = s; s += H1 * D;
    m.gb1 = s; s += H1;
    m.gW2 = s; s += H2 * H1;
    m.gb2 = s; s += H2;
    m.gw3 = s; s += H2;
    m.u = s; s += 64;
    m.p1 = s; s += H1; m.h1 = s; s += H1; m.e1 = s; s += H1; m.e1pp = s; s += H1;
    m.h2 = s; s += H2; m.e2 = s; s += H2; m.e2pp = s; s += H2; m.g2 = s; s += H2;
    m.r = s; s += H1; m.g1 = s; s += H1; m.gx = s; s += XD; m.a = s; s += XD;
    m.dr = s; s += H1; m.dp1 = s; s += H1; m.dp2 = s; s += H2; m.red = s; s += 8;
    return m;
}
static size_t critic_smem_floats(int D) { return 2 * (H1 * D + H1 + H2 * H1 + H2 + H2) + 64 + 4 * H1 + 4 * H2 + 2 * H1 + 2 * XD + 2 * H1 + H2 + 8; }

__device__ __forceinline__ void load_weights(CriticSmem& m, int D, const float* W1, const float* b1, const float* W2,
                                             const float* b2, const float* W3) {
    for (int i = threadIdx.x; i < H1 * D; i += NT) m.W1[i] = W1[i];
    for (int i = threadIdx.x; i < H2 * H1; i += NT) m.W2[i] = W2[i];
    for (int i = threadIdx.x; i < H1; i += NT) m.b1[i] = b1[i];
    for (int i = threadIdx.x; i < H2; i += NT) { m.b2[i] = b2[i]; m.w3[i] = W3[i]; }
}

// layers 1 and 2 of the critic for the input in m.u; fills p1,h1,e1,e1pp,h2,e2,e2pp.  Ends with a barrier.
__device__ __forceinline__ void critic_forward(CriticSmem& m, int D) {
    int t = threadIdx.x;
    if (t < H1) {
        float acc = m.b1[t];
        for (int d = 0; d < D; ++d) acc = fmaf(m.W1[t * D + d], m.u[d], acc);
        float ex = __expf(acc);
        m.p1[t] = acc;
        m.h1[t] = acc > 0.f ? acc : expm1f(acc);
        m.e1[t] = acc > 0.f ? 1.f : ex;
        m.e1pp[t] = acc > 0.f ? 0.f : ex;
    }
    __syncthreads();
    if (t < H2) {
        float acc = m.b2[t];
        for (int k = 0; k < H1; ++k) acc = fmaf(m.W2[t * H1 + k], m.h1[k], acc);
        float ex = __expf(acc);
        m.h2[t] = acc > 0.f ? acc : expm1f(acc);
        m.e2[t] = acc > 0.f ? 1.f : ex;
        m.e2pp[t] = acc > 0.f ? 0.f : ex;
    }
    __syncthreads();
}

// critic output for the current input (all threads get it); uses m.red
__device__ __forceinline__ float critic_output(CriticSmem& m, float b3) {
    int t = threadIdx.x;
    if (t < 32) {
        float v = m.w3[t] * m.h2[t];
        v = warp_sum(v);
        if (t == 0) m.red[0] = v + b3;
    }
    __syncthreads();
    float o = m.red[0];
    __syncthreads();
    return o;
}

// input gradient d out / d x for the current input: fills g2, r, g1, gx.  Ends with a barrier.
__device__ __forceinline__ void critic_input_grad(CriticSmem& m, int D) {
    int t = threadIdx.x;
    if (t < H2) m.g2[t] = m.w3[t] * m.e2[t];
    __syncthreads();
    if (t < H1) {
        float acc = 0.f;
        for (int q = 0; q < H2; ++q) acc = fmaf(m.W2[q * H1 + t], m.g2[q], acc);
        m.r[t] = acc;
        m.g1[t] = acc * m.e1[t];
    }
    __syncthreads();
    if (t < XD) {
        float acc = 0.f;
        for (int k = 0; k < H1; ++k) acc = fmaf(m.W1[k * D + t], m.g1[k], acc);
        m.gx[t] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT)
wgangp_dstep_kernel(const float* __restrict__ fv, const float* __restrict__ z0, const float* __restrict__ means,
                    const int64_t* __restrict__ labels, const float* __restrict__ alphas, const float* W1,
                    const float* b1, const float* W2, const float* b2, const float* W3, const float* b3,
                    float gp_weight, float* losses, float* gW1, float* gb1, float* gW2, float* gb2, float* gW3,
                    float* gb3, int B, int C, int spb) {
    extern __shared__ float smem[];
    const int D = XD + C;
    CriticSmem m = carve(smem, D);
    load_weights(m, D, W1, b1, W2, b2, W3);
    for (int i = threadIdx.x; i < H1 * D; i += NT) m.gW1[i] = 0.f;
    for (int i = threadIdx.x; i < H2 * H1; i += NT) m.gW2[i] = 0.f;
    for (int i = threadIdx.x; i < H1; i += NT) m.gb1[i] = 0.f;
    for (int i = threadIdx.x; i < H2; i += NT) { m.gb2[i] = 0.f; m.gw3[i] = 0.f; }
    __syncthreads();
    const float bias3 = b3[0];
    const float invB = 1.f / (float)B;
    const int t = threadIdx.x;
    float acc_loss = 0.f, acc_gp = 0.f, acc_fake = 0.f, acc_real = 0.f, acc_gb3 = 0.f;
    int s0 = blockIdx.x * spb, s1 = min(B, s0 + spb);
    for (int s = s0; s < s1; ++s) {
        const int lab_raw = (int)labels[s];
        // a label outside [0, C) (the reference's one_hot raises on it) must not index the prototypes: the sample's noise
        // input becomes NaN instead, so d_loss and the critic gradients turn NaN -- loud, without a host synchronisation
        const bool lab_ok = lab_raw >= 0 && lab_raw < C;
        const int lab = lab_ok ? lab_raw : 0;
        const float lab_nan = lab_ok ? 0.f : NAN;
        // ---- real (pass 0, dout = -1/B, input z) and fake (pass 1, dout = +1/B, input fv)
        for (int pass = 0; pass < 2; ++pass) {
            if (t < D) {
                float v;
                if (t < XD) v = pass == 0 ? z0[s * XD + t] + means[lab * XD + t] + lab_nan : fv[s * XD + t];
                else v = (t - XD == lab) ? 1.f : 0.f;
                m.u[t] = v;
            }
            __syncthreads();
            critic_forward(m, D);
            float out = critic_output(m, bias3);
            float dout = pass == 0 ? -invB : invB;
            if (pass == 0) acc_real += out; else acc_fake += out;
            acc_gb3 += dout;
            if (t < H2) {
                m.gw3[t] += dout * m.h2[t];
                m.dp2[t] = dout * m.w3[t] * m.e2[t];
            }
            __syncthreads();
            if (t < H1) {
                float a = 0.f;
                for (int q = 0; q < H2; ++q) a = fmaf(m.W2[q * H1 + t], m.dp2[q], a);
                m.dp1[t] = a * m.e1[t];
            }
            __syncthreads();
            for (int e = t; e < H1 * D; e += NT) m.gW1[e] += m.dp1[e / D] * m.u[e % D];
            for (int e = t; e < H2 * H1; e += NT) m.gW2[e] += m.dp2[e / H1] * m.h1[e % H1];
            if (t < H1) m.gb1[t] += m.dp1[t];
            if (t < H2) m.gb2[t] += m.dp2[t];
            __syncthreads();
        }
        // ---- gradient penalty on the interpolate
        if (t < D) {
            float v;
            if (t < XD) {
                float zz = z0[s * XD + t] + means[lab * XD + t] + lab_nan;
                v = zz + alphas[s] * (fv[s * XD + t] - zz);
            } else v = (t - XD == lab) ? 1.f : 0.f;
            m.u[t] = v;
        }
        __syncthreads();
        critic_forward(m, D);
        critic_input_grad(m, D);
        if (t < 32) {
            float q = m.gx[t] * m.gx[t];
            q = warp_sum(q);
            if (t == 0) m.red[1] = sqrtf(q + 1e-12f);
        }
        __syncthreads();
        float slope = m.red[1];
        acc_gp += (slope - 1.f) * (slope - 1.f);
        float coef = gp_weight * invB * 2.f * (slope - 1.f) / slope;
        if (t < XD) m.a[t] = coef * m.gx[t];
        __syncthreads();
        if (t < H1) {
            float dg1 = 0.f;
            for (int d = 0; d < XD; ++d) dg1 = fmaf(m.W1[t * D + d], m.a[d], dg1);
            m.dr[t] = dg1 * m.e1[t];
            m.dp1[t] = dg1 * m.r[t] * m.e1pp[t];          // through ELU' of layer 1
        }
        __syncthreads();
        if (t < H2) {
            float dg2 = 0.f;
            for (int k = 0; k < H1; ++k) dg2 = fmaf(m.W2[t * H1 + k], m.dr[k], dg2);
            m.gw3[t] += dg2 * m.e2[t];
            m.dp2[t] = dg2 * m.w3[t] * m.e2pp[t];          // through ELU' of layer 2
        }
        __syncthreads();
        if (t < H1) {
            float a = 0.f;
            for (int q = 0; q < H2; ++q) a = fmaf(m.W2[q * H1 + t], m.dp2[q], a);
            m.dp1[t] += a * m.e1[t];
        }
        __syncthreads();
        for (int e = t; e < H1 * D; e += NT) {
            int k = e / D, d = e % D;
            float v = m.dp1[k] * m.u[d];
            if (d < XD) v = fmaf(m.g1[k], m.a[d], v);
            m.gW1[e] += v;
        }
        for (int e = t; e < H2 * H1; e += NT) {
            int q = e / H1, k = e % H1;
            m.gW2[e] += fmaf(m.g2[q], m.dr[k], m.dp2[q] * m.h1[k]);
        }
        if (t < H1) m.gb1[t] += m.dp1[t];
        if (t < H2) m.gb2[t] += m.dp2[t];
        __syncthreads();
    }
    // ---- flush
    for (int e = t; e < H1 * D; e += NT) atomicAdd(&gW1[e], m.gW1[e]);
    for (int e = t; e < H2 * H1; e += NT) atomicAdd(&gW2[e], m.gW2[e]);
    if (t < H1) atomicAdd(&gb1[t], m.gb1[t]);
    if (t < H2) { atomicAdd(&gb2[t], m.gb2[t]); atomicAdd(&gW3[t], m.gw3[t]); }
    if (t == 0) {
        acc_loss = (acc_fake - acc_real) * invB + gp_weight * acc_gp * invB;
        atomicAdd(&losses[0], acc_loss);
        atomicAdd(&losses[1], acc_gp * invB);
        atomicAdd(&losses[2], acc_fake * invB);
        atomicAdd(&losses[3], acc_real * invB);
        atomicAdd(gb3, acc_gb3);
    }
}

__global__ void __launch_bounds__(NT)
disc_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, const float* W1, const float* b1,
                const float* W2, const float* b2, const float* W3, const float* b3, float* out, float* dx,
                float dx_scale, float* out_sum, float out_scale, int B, int C, int spb) {
    extern __shared__ float smem[];
    const int D = XD + C;
    CriticSmem m = carve(smem, D);
    load_weights(m, D, W1, b1, W2, b2, W3);
    __syncthreads();
    const float bias3 = b3[0];
    const int t = threadIdx.x;
    float acc = 0.f;
    int s0 = blockIdx.x * spb, s1 = min(B, s0 + spb);
    for (int s = s0; s < s1; ++s) {
        const int lab = (int)labels[s];
        if (t < D) m.u[t] = t < XD ? x[s * XD + t] : ((t - XD == lab) ? 1.f : 0.f);
        __syncthreads();
        critic_forward(m, D);
        float o = critic_output(m, bias3);
        if (out && t == 0) out[s] = o;
        acc += o;
        if (dx) {
            critic_input_grad(m, D);
            if (t < XD) dx[s * XD + t] = dx_scale * m.gx[t];
        }
        __syncthreads();
    }
    if (out_sum && t == 0) atomicAdd(out_sum, out_scale * acc);
}

}  // namespace pcaa

using namespace pcaa;

extern "C" int pcaa_wgangp_dstep(const float* fv, const float* z0, const float* means, const int64_t* labels,
                                 const float* alphas, const float* W1, const float* b1, const float* W2,
                                 const float* b2, const float* W3, const float* b3, float gp_weight, float* losses,
                                 float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3, int64_t B,
                                 int C, pcaa_stream stream) {
    PCAA_REQUIRE(B > 0 && C > 0 && C <= 32, PCAA_ERR_SHAPE, "wgangp_dstep: need 0 < C <= 32 (got %d), B > 0", C);
    cudaStream_t st = (cudaStream_t)stream;
    int D = XD + C;
    size_t smem = critic_smem_floats(D) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(wgangp_dstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(disc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_done = true;
    }
    cudaMemsetAsync(losses, 0, 4 * sizeof(float), st);
    int spb = (int)((B + 147) / 148);
    if (spb < 1) spb = 1;
    int grid = (int)((B + spb - 1) / spb);
    wgangp_dstep_kernel<<<grid, NT, smem, st>>>(fv, z0, means, labels, alphas, W1, b1, W2, b2, W3, b3, gp_weight, losses,
                                               gW1, gb1, gW2, gb2, gW3, gb3, (int)B, C, spb);
    return check_launch("wgangp_dstep");
}

extern "C" int pcaa_disc_fwd(const float* x, const int64_t* labels, const float* W1, const float* b1, const float* W2,
                             const float* b2, const float* W3, const float* b3, float* out, float* dx, float dx_scale,
                             float* out_sum, float out_scale, int64_t B, int C, pcaa_stream stream) {
    PCAA_REQUIRE(B > 0 && C > 0 && C <= 32, PCAA_ERR_SHAPE, "disc_fwd: need 0 < C <= 32 (got %d), B > 0", C);
    cudaStream_t st = (cudaStream_t)stream;
    int D = XD + C;
    size_t smem = critic_smem_floats(D) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(disc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_done = true;
    }
    int spb = (int)((B + 147) / 148);
    if (spb < 1) spb = 1;
    int grid = (int)((B + spb - 1) / spb);
    if (out_sum) cudaMemsetAsync(out_sum, 0, sizeof(float), st);
    disc_fwd_kernel<<<grid, NT, smem, st>>>(x, labels, W1, b1, W2, b2, W3, b3, out, dx, dx_scale, out_sum, out_scale,
                                           (int)B, C, spb);
    return check_launch("disc_fwd");
}
