/* fast_dct.h -- in-register fast DCT-II (unnormalised): C[n] = sum_k x[k] cos(pi n (2k+1) / (2N)).
 *
 * Even/odd recursion (N -> 2 x N/2), multiplication form:
 *   u[k] = x[k] + x[N-1-k]                                   E  = DCT_{N/2}(u)  ->  C[2m]   = E[m]
 *   v[k] = (x[k] - x[N-1-k]) * 2 cos(pi (2k+1) / (2N))       O' = DCT_{N/2}(v)  ->  O'[m]   = C[2m+1] + C[2m-1]
 *   C[1] = O'[0] / 2,  C[2m+1] = O'[m] - C[2m-1]             (because 2 cos a cos b = cos(a+b) + cos(a-b))
 * 304 operations for N = 32 instead of 1024 multiply-adds.  Every index is a compile-time constant,
 * so the arrays live in registers and the twiddles become immediates.
 * Used by the fused back end for the polyphase matrixing (a11); checked against the definition in
 * tests/test_fast_dct_cpu.py (host build) and through the PCM parity tests on the GPU.
 */
#ifndef MP3B_FAST_DCT_H
#define MP3B_FAST_DCT_H

#include "consts_gen.h"

#if defined(__CUDACC__)
#define L3_FD __device__ __forceinline__
#else
#define L3_FD inline
#endif

template <int N> struct L3Twiddle;
template <> struct L3Twiddle<2> { L3_FD static float at(int k) { return TW2[k]; } };
template <> struct L3Twiddle<4> { L3_FD static float at(int k) { return TW4[k]; } };
template <> struct L3Twiddle<8> { L3_FD static float at(int k) { return TW8[k]; } };
template <> struct L3Twiddle<16> { L3_FD static float at(int k) { return TW16[k]; } };
template <> struct L3Twiddle<32> { L3_FD static float at(int k) { return TW32[k]; } };

template <int N> struct L3Dct2 {
    L3_FD static void run(float (&x)[N])
    {
        float u[N / 2], v[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            u[k] = x[k] + x[N - 1 - k];
            v[k] = (x[k] - x[N - 1 - k]) * L3Twiddle<N>::at(k);
        }
        L3Dct2<N / 2>::run(u);
        L3Dct2<N / 2>::run(v);
        float prev = v[0] * 0.5f;
        x[0] = u[0];
        x[1] = prev;
#pragma unroll
        for (int m = 1; m < N / 2; m++) {
            x[2 * m] = u[m];
            prev = v[m] - prev;
            x[2 * m + 1] = prev;
        }
    }
};
template <> struct L3Dct2<1> {
    L3_FD static void run(float (&)[1]) {}
};

#if defined(__CUDACC__) && !defined(L3_NO_F32X2)
#include "f32x2.h"
/* The same recursion on PAIRS of independent sequences (x[k].x and x[k].y), one packed instruction (FADD2 / FMUL2,
 * sm_100) per two scalar ones: the two half-size transforms of an even/odd split are such a pair.  Operation for
 * operation the scalar algorithm, so the results are bit-identical to it. */
template <int N> struct L3Dct2P {
    L3_FD static void run(float2 (&x)[N])
    {
        float2 u[N / 2], v[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) {
            u[k] = f2_add(x[k], x[N - 1 - k]);
            v[k] = f2_mul_s(L3Twiddle<N>::at(k), f2_sub(x[k], x[N - 1 - k]));
        }
        L3Dct2P<N / 2>::run(u);
        L3Dct2P<N / 2>::run(v);
        float2 prev = f2_mul_s(0.5f, v[0]);
        x[0] = u[0];
        x[1] = prev;
#pragma unroll
        for (int m = 1; m < N / 2; m++) {
            x[2 * m] = u[m];
            prev = f2_sub(v[m], prev);
            x[2 * m + 1] = prev;
        }
    }
};
template <> struct L3Dct2P<1> {
    L3_FD static void run(float2 (&)[1]) {}
};
/* 32-point transform of one sequence: the first split in scalar form, its two 16-point halves as one packed pair */
L3_FD void l3_dct2_32_packed(float (&x)[32])
{
    float2 h[16];
#pragma unroll
    for (int k = 0; k < 16; k++)
        h[k] = make_float2(x[k] + x[31 - k], (x[k] - x[31 - k]) * L3Twiddle<32>::at(k));
    L3Dct2P<16>::run(h);
    float prev = h[0].y * 0.5f;
    x[0] = h[0].x;
    x[1] = prev;
#pragma unroll
    for (int m = 1; m < 16; m++) {
        x[2 * m] = h[m].x;
        prev = h[m].y - prev;
        x[2 * m + 1] = prev;
    }
}
#endif

#endif
