// Shared helpers for the PCAA sm_100a kernels (error reporting, dtype access, reductions).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcaa.h"

namespace pcaa {

// thread-local last-error message, returned by pcaa_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PCAA_REQUIRE(cond, code, ...)            \
    do {                                         \
        if (!(cond)) {                           \
            ::pcaa::set_error(__VA_ARGS__);      \
            return (code);                       \
        }                                        \
    } while (0)

__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : expm1f(z); }
// derivative of ELU(alpha=1) at pre-activation z
__device__ __forceinline__ float elu_grad_f(float z) { return z > 0.f ? 1.f : __expf(z); }

// 2^x on the SFU (one MUFU.EX2, no range fix-up); ELU / ELU' from z and zl = z * log2(e) (zl comes from a second FMA with
// pre-scaled BatchNorm coefficients, so the exponential costs no extra multiply)
constexpr float LOG2E_F = 1.4426950408889634f;
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float elu_l2(float z, float zl) { return z > 0.f ? z : ex2_fast(zl) - 1.f; }
__device__ __forceinline__ float elu_grad_l2(float z, float zl) { return z > 0.f ? 1.f : ex2_fast(zl); }

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace pcaa
