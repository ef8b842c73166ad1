/* f32x2.h -- packed FP32 arithmetic of sm_100 (PTX add / sub / mul / fma .f32x2; SASS FADD2 / FMUL2 / FFMA2).
 *
 * One instruction does two independent FP32 operations on a pair of registers, with the same rounding as the scalar
 * forms (so results are bit-identical to two scalar instructions).  A multiplicand that is the same for both halves is
 * written as make_float2(w, w): ptxas encodes it as a broadcast operand (Rw.F32), no second register is needed.
 * The back end is bound by instruction issue, not by the FMA pipe, so halving the instruction count of its
 * multiply-add loops is what counts (tools/fp32_peak.cu measures the rates).
 */
#ifndef MP3B_F32X2_H
#define MP3B_F32X2_H

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long f2_pack(float2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
__device__ __forceinline__ float2 f2_sub(float2 a, float2 b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
/* scalar multiplicand broadcast to both halves */
__device__ __forceinline__ float2 f2_fma_s(float w, float2 b, float2 c) { return f2_fma(make_float2(w, w), b, c); }
__device__ __forceinline__ float2 f2_mul_s(float w, float2 b) { return f2_mul(make_float2(w, w), b); }
#endif

#endif
