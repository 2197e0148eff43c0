// nfa_math.cuh -- explicitly-rounded binary32 arithmetic shared by the device
// kernels and the host-side simulation harness (tests/host_sim).
//
// Every value that feeds a comparison in the traversal has to be rounded exactly
// like the reference CUDA build rounds it (SURVEY.md section 7 "hard parts" #1),
// so nothing here is left to the compiler's contraction heuristics: each
// operation is spelled with an explicit round-to-nearest intrinsic on the device
// and with plain IEEE ops / fmaf() on the host (host builds use
// -ffp-contract=off).
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define NFA_HD __host__ __device__ __forceinline__
#define NFA_D __device__ __forceinline__
#else
#define NFA_HD inline
#define NFA_D inline
#endif

namespace nfa {

NFA_HD float f_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
NFA_HD float f_sub(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
NFA_HD float f_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
NFA_HD float f_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
NFA_HD float f_div(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
NFA_HD float f_rcp(float a)
{
#if defined(__CUDA_ARCH__)
    return __frcp_rn(a);
#else
    return 1.0f / a;
#endif
}
NFA_HD float f_min(float a, float b) { return fminf(a, b); }
NFA_HD float f_max(float a, float b) { return fmaxf(a, b); }

NFA_HD uint32_t f_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
NFA_HD float f_from_bits(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// float -> int32 as cvt.rzi.s32.f32 does it (truncate, saturate, NaN -> 0).
NFA_HD int f_trunc_i32(float f)
{
#if defined(__CUDA_ARCH__)
    return __float2int_rz(f);
#else
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int)f;
#endif
}

NFA_HD int i_clamp(int v, int lo, int hi)
{
    int m = v < hi ? v : hi;
    return lo > m ? lo : m;
}

}  // namespace nfa
