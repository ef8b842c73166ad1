/* fast_imdct.h -- in-register 18-point DCT-IV, the core of the 36-point IMDCT (a9).
 *
 * The IMDCT of the standard, x[i] = sum_k X[k] cos(pi/72 (2i+1+18)(2k+1)), i < 36, has only 18
 * distinct values: with Z[n] = sum_k X[k] cos(pi/72 (2n+1)(2k+1))  (DCT-IV, N = 18)
 *     x[i]      =  Z[9 + i]      i = 0..8          x[17 - i] = -x[i]
 *     x[18 + j] = -Z[8 - j]      j = 0..8          x[35 - j] =  x[18 + j]
 * Z is computed as  w[k] = X[k] 2cos(pi(2k+1)/72);  Y = DCT-II_18(w);  Z[0] = Y[0]/2, Z[n] = Y[n] - Z[n-1]
 * (2 cos a cos b = cos(a+b) + cos(a-b)), DCT-II_18 by one even/odd split into two 9-point DCT-IIs,
 * and those by their own symmetry (cos(pi m (17-2k)/18) = (-1)^m cos(pi m (2k+1)/18)):
 * about 170 operations instead of 324 multiply-adds, all indices compile-time constants.
 * Checked against the definition in tests/test_fast_dct_cpu.py and by the GPU parity tests.
 */
#ifndef MP3B_FAST_IMDCT_H
#define MP3B_FAST_IMDCT_H

#include "consts_gen.h"

#if defined(__CUDACC__)
#define L3_FI __device__ __forceinline__
#else
#define L3_FI inline
#endif

/* 9-point DCT-II: E[m] = sum_k u[k] cos(pi m (2k+1) / 18) */
L3_FI void l3_dct2_9(const float (&u)[9], float (&E)[9])
{
    float s[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        s[k] = u[k] + u[8 - k];
        d[k] = u[k] - u[8 - k];
    }
#pragma unroll
    for (int m = 0; m < 9; m++) {
        float acc;
        if (m & 1) {
            acc = d[0] * C9[m][0];
#pragma unroll
            for (int k = 1; k < 4; k++) acc = fmaf(d[k], C9[m][k], acc);
        } else {
            acc = u[4] * C9[m][4]; /* cos(pi m / 2) = +-1 */
#pragma unroll
            for (int k = 0; k < 4; k++) acc = fmaf(s[k], C9[m][k], acc);
        }
        E[m] = acc;
    }
}

/* Z[n] = sum_k X[k] cos(pi/72 (2n+1)(2k+1)), n < 18 */
L3_FI void l3_dct4_18(const float (&X)[18], float (&Z)[18])
{
    float u[9], v[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const float a = X[k] * TW_IV18[k], b = X[17 - k] * TW_IV18[17 - k];
        u[k] = a + b;
        v[k] = (a - b) * TW_II18[k];
    }
    float E[9], O[9];
    l3_dct2_9(u, E);
    l3_dct2_9(v, O);
    /* Y[2m] = E[m]; Y[1] = O[0]/2, Y[2m+1] = O[m] - Y[2m-1]; then Z[0] = Y[0]/2, Z[n] = Y[n] - Z[n-1] */
    float yodd = O[0] * 0.5f;
    float z = E[0] * 0.5f;
    Z[0] = z;
    z = yodd - z;
    Z[1] = z;
#pragma unroll
    for (int m = 1; m < 9; m++) {
        z = E[m] - z;
        Z[2 * m] = z;
        yodd = O[m] - yodd;
        z = yodd - z;
        Z[2 * m + 1] = z;
    }
}

#endif
