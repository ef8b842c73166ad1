// k_hybrid.cu -- K3a: IMDCT-36 / 3x IMDCT-12 with windowing (a9); K3b: overlap-add and
// frequency inversion (a10).
//
// K3a: one warp per unit, lane = subband.  Each lane keeps its 18 spectral lines in registers
// and evaluates the 36-point IMDCT through its two symmetries (x[17-i] = -x[i], x[35-i] = x[18+i]
// for i < 9), i.e. 18 distinct outputs x 18 inputs = 324 FMAs with compile-time coefficient
// addresses in constant memory (FFMA with a constant-bank operand, no loads in the loop).
// Short blocks evaluate three 12-point transforms the same way.  Output is slot-major
// [36][32] so that the lanes of a warp write one 128-byte row per store.
// K3b: elementwise, float4: out[t][sb] = cur[t][sb] + prev[18 + t][sb], odd-odd negation.
// Restates oracle/l3_oracle.c::hybrid in float32 (coefficients rounded from double).
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include <math.h>

#include "kernels.h"

namespace {

__constant__ float c_cosA[9][18];  // out[i],      i = 0..8
__constant__ float c_cosB[9][18];  // out[18 + i], i = 0..8
__constant__ float c_cos12[12][6];
__constant__ float c_win[4][36];

constexpr int K3_WARPS = 8;

// 36 windowed IMDCT outputs of one subband, for one lane.  bt = window type of this subband.
__device__ __forceinline__ void imdct_lane(const float (&X)[18], int bt, float (&o)[36])
{
    if (bt != 2) {
        float a[9], b[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int k = 0; k < 18; k++) {
                sa = fmaf(X[k], c_cosA[i][k], sa);
                sb = fmaf(X[k], c_cosB[i][k], sb);
            }
            a[i] = sa;
            b[i] = sb;
        }
        const float *w = c_win[bt];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            o[i] = a[i] * w[i];
            o[17 - i] = -a[i] * w[17 - i];
            o[18 + i] = b[i] * w[18 + i];
            o[35 - i] = b[i] * w[35 - i];
        }
    } else {
        float y[3][12];
#pragma unroll
        for (int wdw = 0; wdw < 3; wdw++)
#pragma unroll
            for (int i = 0; i < 12; i++) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 6; k++) s = fmaf(X[3 * k + wdw], c_cos12[i][k], s);
                y[wdw][i] = s * c_win[2][i];
            }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            o[i] = 0.f;
            o[6 + i] = y[0][i];
            o[12 + i] = y[0][6 + i] + y[1][i];
            o[18 + i] = y[1][6 + i] + y[2][i];
            o[24 + i] = y[2][6 + i];
            o[30 + i] = 0.f;
        }
    }
}

__global__ void __launch_bounds__(K3_WARPS * 32)
k_imdct(const L3UnitDesc *__restrict__ units, uint32_t u_lo, uint32_t nunits, const float *__restrict__ xr,
        float *__restrict__ imd)
{
    __shared__ float s_x[K3_WARPS][32 * 19];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x * K3_WARPS + warp >= nunits) return;
    const uint32_t u = u_lo + blockIdx.x * K3_WARPS + warp;
    const uint8_t flags = units[u].flags;
    const float *src = xr + (size_t)u * 576;
    // coalesced load, stored with a row stride of 19 so that lane-strided reads are conflict-free
    for (int i = lane; i < 576; i += 32) s_x[warp][(i / 18) * 19 + (i % 18)] = src[i];
    __syncwarp();
    float X[18], o[36];
#pragma unroll
    for (int k = 0; k < 18; k++) X[k] = s_x[warp][lane * 19 + k];
    int bt = flags & L3F_BT_MASK;
    if (bt == 2 && (flags & L3F_MIXED) && lane < 2) bt = 0;
    imdct_lane(X, bt, o);
    float *dst = imd + (size_t)u * (36 * 32);
#pragma unroll
    for (int i = 0; i < 36; i++) dst[i * 32 + lane] = o[i];
}

__global__ void __launch_bounds__(256)
k_overlap(const L3UnitDesc *__restrict__ units, uint32_t u_lo, uint32_t nunits, const float4 *__restrict__ imd,
          float4 *__restrict__ sb)
{
    // one thread per float4 of the [18][32] output block: 144 per unit
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx / 144 >= nunits) return;
    const uint32_t u = u_lo + (uint32_t)(idx / 144);
    const uint32_t e = (uint32_t)(idx % 144), t = e >> 3; // slot 0..17; 8 float4 per slot
    const L3UnitDesc d = units[u];
    const uint32_t nch = (d.hdr & L3H_STEREO) ? 2 : 1;
    float4 v = imd[(size_t)u * 288 + e];
    if (!(d.pos & L3P_FIRST)) {
        const float4 p = imd[(size_t)(u - nch) * 288 + 144 + e];
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (t & 1) { v.y = -v.y; v.w = -v.w; } // odd slot: negate odd subbands
    sb[(size_t)u * 144 + e] = v;
}

} // namespace

void l3_hybrid_init(void)
{
    static float A[9][18], B[9][18], C12[12][6], W[4][36];
    for (int i = 0; i < 9; i++)
        for (int k = 0; k < 18; k++) {
            A[i][k] = (float)cos(M_PI / 72.0 * (2 * i + 1 + 18) * (2 * k + 1));
            B[i][k] = (float)cos(M_PI / 72.0 * (2 * (18 + i) + 1 + 18) * (2 * k + 1));
        }
    for (int i = 0; i < 12; i++)
        for (int k = 0; k < 6; k++) C12[i][k] = (float)cos(M_PI / 24.0 * (2 * i + 1 + 6) * (2 * k + 1));
    for (int i = 0; i < 36; i++) {
        W[0][i] = (float)sin(M_PI / 36.0 * (i + 0.5));
        W[1][i] = i < 18 ? (float)sin(M_PI / 36.0 * (i + 0.5))
                         : (i < 24 ? 1.f : (i < 30 ? (float)sin(M_PI / 12.0 * (i - 18 + 0.5)) : 0.f));
        W[2][i] = i < 12 ? (float)sin(M_PI / 12.0 * (i + 0.5)) : 0.f;
        W[3][i] = i < 6 ? 0.f : (i < 12 ? (float)sin(M_PI / 12.0 * (i - 6 + 0.5))
                                        : (i < 18 ? 1.f : (float)sin(M_PI / 36.0 * (i + 0.5))));
    }
    cudaMemcpyToSymbol(c_cosA, A, sizeof A);
    cudaMemcpyToSymbol(c_cosB, B, sizeof B);
    cudaMemcpyToSymbol(c_cos12, C12, sizeof C12);
    cudaMemcpyToSymbol(c_win, W, sizeof W);
}

void l3_launch_imdct_range(const L3UnitDesc *units, uint32_t u_lo, uint32_t nunits, const float *xr, float *imd,
                           cudaStream_t st)
{
    if (!nunits) return;
    k_imdct<<<(nunits + K3_WARPS - 1) / K3_WARPS, K3_WARPS * 32, 0, st>>>(units, u_lo, nunits, xr, imd);
}

void l3_launch_overlap_range(const L3UnitDesc *units, uint32_t u_lo, uint32_t nunits, const float *imd, float *sb,
                             cudaStream_t st)
{
    if (!nunits) return;
    const uint64_t n = (uint64_t)nunits * 144;
    k_overlap<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(units, u_lo, nunits,
                                                            reinterpret_cast<const float4 *>(imd),
                                                            reinterpret_cast<float4 *>(sb));
}
