// k_resample.cu -- sample-rate conversion of the decoded batch (the output step
// just after the hot path: a player's audio device runs at one rate, its files at several).
//
// Rational-ratio polyphase FIR: out_rate / in_rate = L / M in lowest terms.  By definition
//     y[n] = sum_i x[i] * h[n M + D - i L],        D = half * L (the filter's centre)
// i.e. zero-stuff by L, filter with h (a Kaiser-windowed sinc of 2 half L + 1 taps, cut-off just
// below the lower of the two Nyquist frequencies, DC gain L), keep every M-th sample.  With
// u = n M + D, q = u div L, p = u mod L this is T = 2 half + 1 multiply-adds per output sample:
//     y[n] = sum_{j < T} hp[p][j] * x[q - j],      hp[p][j] = h[p + j L]
// One thread per output frame (all channels), FP32 accumulation.  The fast kernel keeps the whole
// polyphase table and, per tile of 256 output frames, the input span they touch (converted to float,
// zero outside the stream) in shared memory: row stride T is odd, so lanes (whose phases differ) read
// their taps without bank conflicts, and neighbouring lanes share input samples.  Rate pairs whose
// table does not fit shared memory use the general kernel (taps through L1).
// No reference code exists for this step (/root/reference/README.md:1-84); tests compare it with the
// textbook definition evaluated in float64 by numpy on the filter returned by mp3b_resample_filter.
#include <math.h>

#include <vector>

#include "kernels.h"
#include "mp3b.h"

namespace {

__device__ __forceinline__ int16_t rs_to_s16(float v)
{
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(v * 32768.f));
    return (int16_t)r;
}

template <typename T> __device__ __forceinline__ float rs_load(const T *p);
template <> __device__ __forceinline__ float rs_load<int16_t>(const int16_t *p) { return (float)*p * (1.f / 32768.f); }
template <> __device__ __forceinline__ float rs_load<float>(const float *p) { return *p; }
__device__ __forceinline__ void rs_store(int16_t *p, float v) { *p = rs_to_s16(v); }
__device__ __forceinline__ void rs_store(float *p, float v) { *p = v; }

template <typename T>
__global__ void __launch_bounds__(256)
k_resample(const T *__restrict__ in, T *__restrict__ out, const L3ResampleJob *__restrict__ jobs,
           const float *__restrict__ hp, int L, int M, int taps, int half)
{
    const L3ResampleJob jb = jobs[blockIdx.y];
    const int nch = jb.channels;
    const T *x = in + jb.in_off;
    T *y = out + jb.out_off;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < jb.out_n; n += (long long)gridDim.x * blockDim.x) {
        const long long u = n * M + (long long)half * L;
        const long long q = u / L;
        const int p = (int)(u - q * L);
        const float *h = hp + (size_t)p * taps;
        float a0 = 0.f, a1 = 0.f;
        // taps whose input sample exists: 0 <= q - j < in_n
        const int j_lo = q >= jb.in_n ? (int)(q - jb.in_n + 1) : 0;
        const int j_hi = q < taps - 1 ? (int)q : taps - 1;
        if (nch == 2) {
            for (int j = j_lo; j <= j_hi; j++) {
                const float c = __ldg(h + j);
                const T *s = x + (q - j) * 2;
                a0 = fmaf(c, rs_load<T>(s), a0);
                a1 = fmaf(c, rs_load<T>(s + 1), a1);
            }
            rs_store(y + n * 2, a0);
            rs_store(y + n * 2 + 1, a1);
        } else {
            for (int j = j_lo; j <= j_hi; j++) a0 = fmaf(__ldg(h + j), rs_load<T>(x + (q - j)), a0);
            rs_store(y + n, a0);
        }
    }
}

// Tiled kernel: dynamic shared memory = [L * taps floats, padded to 16 bytes][span * nch floats].
template <typename T, int NCH>
__global__ void __launch_bounds__(256)
k_resample_tiled(const T *__restrict__ in, T *__restrict__ out, const L3ResampleJob *__restrict__ jobs,
                 const float *__restrict__ hp, int L, int M, int taps, int half, int span)
{
    extern __shared__ __align__(16) float rs_smem[];
    float *s_h = rs_smem, *s_x = rs_smem + (((size_t)L * taps + 3) & ~(size_t)3); // float2 reads need alignment
    const L3ResampleJob jb = jobs[blockIdx.y];
    if (jb.channels != NCH) return;
    const T *x = in + jb.in_off;
    T *y = out + jb.out_off;
    for (int i = threadIdx.x; i < L * taps; i += 256) s_h[i] = hp[i];
    const long long D = (long long)half * L;
    for (long long n0 = (long long)blockIdx.x * 256; n0 < jb.out_n; n0 += (long long)gridDim.x * 256) {
        // input frames the tile touches: [q_first - (taps - 1), q_last]
        const long long q_first = (n0 * M + D) / L, base = q_first - (taps - 1);
        __syncthreads(); // the previous tile's readers are done with s_x (and s_h is loaded)
        for (int i = threadIdx.x; i < span * NCH; i += 256) {
            const long long fr = base + i / NCH;
            s_x[i] = (fr >= 0 && fr < jb.in_n) ? rs_load<T>(x + fr * NCH + i % NCH) : 0.f;
        }
        __syncthreads();
        const long long n = n0 + threadIdx.x;
        if (n < jb.out_n) {
            const long long u = n * M + D, q = u / L;
            const int p = (int)(u - q * L);
            const float *h = s_h + (size_t)p * taps;
            const float *xs = s_x + (size_t)(q - base) * NCH; // frame q; frame q - j is NCH * j floats below
            float a0 = 0.f, a1 = 0.f;
            if (NCH == 2) {
#pragma unroll 4
                for (int j = 0; j < taps; j++) {
                    const float2 v = *reinterpret_cast<const float2 *>(xs - 2 * j);
                    a0 = fmaf(h[j], v.x, a0);
                    a1 = fmaf(h[j], v.y, a1);
                }
                rs_store(y + n * 2, a0);
                rs_store(y + n * 2 + 1, a1);
            } else {
#pragma unroll 4
                for (int j = 0; j < taps; j++) a0 = fmaf(h[j], xs[-j], a0);
                rs_store(y + n, a0);
            }
        }
    }
}

// Register-coefficient kernel for the common tap counts (65: every upsampling pair; 71, 97, 129: 48 -> 44.1 kHz,
// 3 : 2 and 2 : 1 downsampling).  The tiled kernel above is bound by shared-
// memory bandwidth: per tap and output frame it loads one coefficient (4 bytes) and one input frame (8
// bytes stereo) = 3 LSU cycles per warp for 2 FMAs.  Here a thread owns output t of every block of S =
// blockDim outputs; S is a multiple of L, so the thread's phase -- and therefore its coefficient row -- is
// the same in every block and lives in registers for the thread's whole life: no coefficient loads, 2 LSU
// cycles per tap.  (Two neighbouring outputs per thread, sharing the input loads, was tried: 168 registers,
// two CTAs per SM, input loads with 2-way bank conflicts -- 15.7 ms against 12.1 for the tiled kernel.)
constexpr int RU_NB = 16; // blocks staged per chunk

template <typename T, int NCH, int RU_TAPS>
__global__ void __launch_bounds__(320, RU_TAPS <= 71 ? 2 : 1)
k_resample_reg(const T *__restrict__ in, T *__restrict__ out, const L3ResampleJob *__restrict__ jobs,
              const float *__restrict__ hp, int L, int M, int half, int span)
{
    extern __shared__ __align__(16) float rs_smem[];
    float *s_x = rs_smem;
    const L3ResampleJob jb = jobs[blockIdx.y];
    if (jb.channels != NCH) return;
    const T *x = in + jb.in_off;
    T *y = out + jb.out_off;
    const int t = threadIdx.x, S = blockDim.x;
    const long long D = (long long)half * L;
    const long long step = S / L * M; // input frames per block (S is a multiple of L)
    // this thread's output inside a block: its last input frame (relative to the block) and its phase
    const long long u0 = (long long)t * M + D;
    const int qt = (int)(u0 / L), p0 = (int)(u0 - (long long)qt * L);
    float c[RU_TAPS];
#pragma unroll
    for (int j = 0; j < RU_TAPS; j++) c[j] = __ldg(hp + (size_t)p0 * RU_TAPS + j);
    const int q_lo = (int)(D / L); // thread 0's qt
    const long long chunk_out = (long long)RU_NB * S;
    for (long long n0 = (long long)blockIdx.x * chunk_out; n0 < jb.out_n; n0 += (long long)gridDim.x * chunk_out) {
        const long long qc = n0 / L * M; // exact: n0 is a multiple of S
        const long long base = qc + q_lo - (RU_TAPS - 1);
        __syncthreads();
        for (int i = t; i < span * NCH; i += S) {
            const long long fr = base + i / NCH;
            s_x[i] = (fr >= 0 && fr < jb.in_n) ? rs_load<T>(x + fr * NCH + i % NCH) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < RU_NB; k++) {
            const long long n = n0 + (long long)k * S + t;
            if (n >= jb.out_n) break;
            const float *xs = s_x + (size_t)(qc + (long long)k * step + qt - base) * NCH; // frame q of this output
            float a0 = 0.f, a1 = 0.f;
            if (NCH == 2) {
#pragma unroll
                for (int j = 0; j < RU_TAPS; j++) {
                    const float2 v = *reinterpret_cast<const float2 *>(xs - 2 * j);
                    a0 = fmaf(c[j], v.x, a0);
                    a1 = fmaf(c[j], v.y, a1);
                }
                rs_store(y + n * 2, a0);
                rs_store(y + n * 2 + 1, a1);
            } else {
#pragma unroll
                for (int j = 0; j < RU_TAPS; j++) a0 = fmaf(c[j], xs[-j], a0);
                rs_store(y + n, a0);
            }
        }
    }
}

double bessel_i0(double x)
{
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 64; k++) {
        t *= (x / (2.0 * k)) * (x / (2.0 * k));
        s += t;
        if (t < 1e-18 * s) break;
    }
    return s;
}

int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

} // namespace

// The FIR for in_rate -> out_rate, in polyphase order hp[p][j] = h[p + j L].  Returns the number of
// coefficients (L * taps), 0 for unusable rates.
size_t l3_resample_design(int in_rate, int out_rate, std::vector<float> *hp, int *pL, int *pM, int *ptaps, int *phalf)
{
    if (in_rate <= 0 || out_rate <= 0 || in_rate > 768000 || out_rate > 768000) return 0;
    const int g = gcd_i(in_rate, out_rate);
    const int L = out_rate / g, M = in_rate / g;
    if (L > 4096 || M > 4096) return 0;
    // 32 zero crossings on each side at the lower of the two rates
    const int half = L >= M ? 32 : (32 * M + L - 1) / L;
    const int taps = 2 * half + 1;
    const long long D = (long long)half * L, N = 2 * D + 1;
    // cut-off centred so that the Kaiser transition band (beta 9, ~ -90 dB) ends at the lower Nyquist frequency
    const double fc = 0.5 / (double)(L > M ? L : M) * 0.91; // cycles per (upsampled) sample
    const double beta = 9.0, i0b = bessel_i0(beta);
    hp->assign((size_t)L * taps, 0.f);
    for (long long k = 0; k < N; k++) {
        const double t = (double)(k - D);
        const double s = t == 0.0 ? 1.0 : sin(2.0 * M_PI * fc * t) / (2.0 * M_PI * fc * t);
        const double r = t / (double)D;
        const double w = bessel_i0(beta * sqrt(r * r < 1.0 ? 1.0 - r * r : 0.0)) / i0b;
        const double v = 2.0 * fc * L * s * w;
        (*hp)[(size_t)(k % L) * taps + (size_t)(k / L)] = (float)v;
    }
    *pL = L;
    *pM = M;
    *ptaps = taps;
    *phalf = half;
    return hp->size();
}

void l3_launch_resample(const void *in, void *out, int pcm_format, const L3ResampleJob *jobs, int njobs,
                        long long max_out_n, const float *hp, int L, int M, int taps, int half, cudaStream_t st,
                        int channel_mask)
{
    if (njobs <= 0 || max_out_n <= 0 || !(channel_mask & 3)) return;
    // the register-coefficient kernel, when the tap count is one it is built for and a block size exists that
    // is a multiple of L
    if (taps == 65 || taps == 71 || taps == 97 || taps == 129) {
        const int T = L * ((128 + L - 1) / L); // threads per CTA = outputs per block: a multiple of L
        const long long S = T;
        const int span = (int)((RU_NB * S / L) * M) + taps + 4;
        const size_t smem = (size_t)span * 2 * sizeof(float);
        if (T <= 320 && smem <= 48 * 1024) { // (the default dynamic shared-memory limit: odd rate pairs can exceed it)
            const long long chunks = (max_out_n + RU_NB * S - 1) / (RU_NB * S);
            long long want = (2048 + njobs - 1) / njobs;
            const unsigned gx = (unsigned)(want < 1 ? 1 : (want > chunks ? chunks : want));
            for (int j0 = 0; j0 < njobs; j0 += 65535) {
                const int nj = njobs - j0 < 65535 ? njobs - j0 : 65535;
                dim3 grid(gx, (unsigned)nj);
#define RS_REG(TT, TAPS)                                                                                                 \
    do {                                                                                                                 \
        if (channel_mask & 2)                                                                                            \
            k_resample_reg<TT, 2, TAPS><<<grid, T, smem, st>>>(static_cast<const TT *>(in), static_cast<TT *>(out),      \
                                                               jobs + j0, hp, L, M, half, span);                         \
        if (channel_mask & 1)                                                                                            \
            k_resample_reg<TT, 1, TAPS><<<grid, T, smem, st>>>(static_cast<const TT *>(in), static_cast<TT *>(out),      \
                                                               jobs + j0, hp, L, M, half, span);                         \
    } while (0)
#define RS_REG_T(TT)                                                                                                     \
    do {                                                                                                                 \
        if (taps == 65) RS_REG(TT, 65);                                                                                  \
        else if (taps == 71) RS_REG(TT, 71);                                                                             \
        else if (taps == 97) RS_REG(TT, 97);                                                                             \
        else RS_REG(TT, 129);                                                                                            \
    } while (0)
                if (pcm_format == MP3B_PCM_S16) RS_REG_T(int16_t);
                else RS_REG_T(float);
#undef RS_REG_T
#undef RS_REG
            }
            return;
        }
    }
    // tiled kernel when the table and a tile's input span fit shared memory
    const int span = (int)((255ll * M) / L) + taps + 2;
    const size_t smem = ((((size_t)L * taps + 3) & ~(size_t)3) + (size_t)span * 2) * sizeof(float);
    if (smem <= 160 * 1024) {
        static std::atomic<unsigned long long> configured{0};
        if (l3_device_needs_setup(configured)) {
            cudaFuncSetAttribute(k_resample_tiled<int16_t, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            cudaFuncSetAttribute(k_resample_tiled<int16_t, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            cudaFuncSetAttribute(k_resample_tiled<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            cudaFuncSetAttribute(k_resample_tiled<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            l3_device_setup_done(configured);
        }
        // a few CTAs per stream, each walking many tiles, so that the table is loaded once per CTA
        const long long tiles = (max_out_n + 255) / 256;
        long long want = (4096 + njobs - 1) / njobs;
        unsigned gxt = (unsigned)(want < 1 ? 1 : (want > tiles ? tiles : want));
        for (int j0 = 0; j0 < njobs; j0 += 65535) {
            const int nj = njobs - j0 < 65535 ? njobs - j0 : 65535;
            dim3 grid(gxt, (unsigned)nj);
            // mono and stereo streams of the same rate share the job list: each instantiation skips the others
            if (pcm_format == MP3B_PCM_S16) {
                k_resample_tiled<int16_t, 2><<<grid, 256, smem, st>>>(static_cast<const int16_t *>(in), static_cast<int16_t *>(out), jobs + j0, hp, L, M, taps, half, span);
                k_resample_tiled<int16_t, 1><<<grid, 256, smem, st>>>(static_cast<const int16_t *>(in), static_cast<int16_t *>(out), jobs + j0, hp, L, M, taps, half, span);
            } else {
                k_resample_tiled<float, 2><<<grid, 256, smem, st>>>(static_cast<const float *>(in), static_cast<float *>(out), jobs + j0, hp, L, M, taps, half, span);
                k_resample_tiled<float, 1><<<grid, 256, smem, st>>>(static_cast<const float *>(in), static_cast<float *>(out), jobs + j0, hp, L, M, taps, half, span);
            }
        }
        return;
    }
    unsigned gx = (unsigned)((max_out_n + 255) / 256);
    if (gx > 4096) gx = 4096; // grid-stride beyond that
    for (int j0 = 0; j0 < njobs; j0 += 65535) {
        const int nj = njobs - j0 < 65535 ? njobs - j0 : 65535;
        dim3 grid(gx, (unsigned)nj);
        if (pcm_format == MP3B_PCM_S16)
            k_resample<int16_t><<<grid, 256, 0, st>>>(static_cast<const int16_t *>(in), static_cast<int16_t *>(out),
                                                      jobs + j0, hp, L, M, taps, half);
        else
            k_resample<float><<<grid, 256, 0, st>>>(static_cast<const float *>(in), static_cast<float *>(out), jobs + j0,
                                                    hp, L, M, taps, half);
    }
}
