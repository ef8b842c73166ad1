// k_stretch.cu -- time-scale modification of the decoded batch without pitch change (WSOLA): the one
// playback feature the reference describes ("slow-speed listening, repeat each sentence",
// /root/reference/README.md:46).  The reference has no code for it.
//
// Waveform-similarity overlap-add, defined so that it is exactly reproducible:
//   speed = num / den (input seconds per output second), Hs = synthesis hop (512 / 256 / 128 samples
//   by sample rate), frame N = 2 Hs, search radius R = Hs / 2, output length floor(in_n den / num).
//   c[n]  = the int8 alignment signal: (left + right) >> 9 for stereo, sample >> 8 for mono (s16
//           domain; float PCM is first rounded to s16), clamped to +-127, zero outside the stream.
//   frame 0 starts at p_0 = 0.  Frame m >= 1 nominally starts at a_m = floor(m Hs num / den); it is
//   moved by the d in [-R, R] that best continues the previous frame, found coarse-to-fine:
//     score(d)  = sum_{k<N}        c[a_m + d + k] c[p_{m-1} + Hs + k]      (full)
//     score2(d) = sum_{k<N, k even} c[a_m + d + k] c[p_{m-1} + Hs + k]      (every second sample)
//   d0 = the first maximum of score2 over d = -R, -R + 4, ..., R; then d = the first maximum of score
//   over [d0 - 3, d0 + 3] (clipped to [-R, R]).  7 x fewer multiply-adds than scoring every d in full.
//   Output segment m (Hs samples): x[p_0 + k] for m = 0, else
//       (1 - w[k]) x[p_{m-1} + Hs + k] + w[k] x[p_m + k],   w[k] = 0.5 - 0.5 cos(pi k / Hs)
//   (the two halves of a periodic Hann window of length N; samples outside the stream are zero).
// The search is integer arithmetic (__dp4a on packed int8: exact, so the chosen offsets equal the
// numpy oracle's bit for bit); the overlap-add is FP32.
// Mapping: the chain over frames is serial per stream, so one CTA per stream walks it; per frame the
// template and the search region (and their every-second-sample copies) are staged in shared memory
// once, one thread scores each coarse candidate (dp4a on packed bytes), the CTA reduces to the argmax,
// one warp scores each fine candidate (unaligned bytes realigned by a funnel shift), and all threads
// write the output segment.
#include <limits.h>
#include <math.h>

#include "kernels.h"
#include "mp3b.h"

namespace {

constexpr int TS_THREADS = 256;
constexpr int TS_MAX_HS = 512;

template <typename T> __device__ __forceinline__ int ts_s16(const T *p);
template <> __device__ __forceinline__ int ts_s16<int16_t>(const int16_t *p) { return *p; }
template <> __device__ __forceinline__ int ts_s16<float>(const float *p)
{
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(*p * 32768.f));
    return (int)(short)r;
}
template <typename T> __device__ __forceinline__ float ts_f(const T *p);
template <> __device__ __forceinline__ float ts_f<int16_t>(const int16_t *p) { return (float)*p * (1.f / 32768.f); }
template <> __device__ __forceinline__ float ts_f<float>(const float *p) { return *p; }
__device__ __forceinline__ void ts_store(int16_t *p, float v)
{
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(v * 32768.f));
    *p = (int16_t)r;
}
__device__ __forceinline__ void ts_store(float *p, float v) { *p = v; }

// alignment signal at input frame n (zero outside [0, in_n))
template <typename T>
__device__ __forceinline__ int ts_align(const T *x, long long n, long long in_n, int nch)
{
    if (n < 0 || n >= in_n) return 0;
    int v = nch == 2 ? (ts_s16<T>(x + n * 2) + ts_s16<T>(x + n * 2 + 1)) >> 9 : ts_s16<T>(x + n) >> 8;
    return max(-127, min(127, v));
}

template <typename T>
__global__ void __launch_bounds__(TS_THREADS)
k_stretch(const T *__restrict__ in, T *__restrict__ out, const L3StretchJob *__restrict__ jobs, int num, int den,
          int *__restrict__ offsets_out /* optional: chosen d per frame, [job][max_frames] */, int max_frames)
{
    __shared__ __align__(16) signed char s_t[2 * TS_MAX_HS + 16];      // template: c[p_prev + Hs + k], k < N
    __shared__ __align__(16) signed char s_r[3 * TS_MAX_HS + 16 + 16]; // region: c[a - R + i], i < N + 2 R (+ pad)
    __shared__ __align__(16) signed char s_t2[TS_MAX_HS + 16];         // template, every second sample
    __shared__ __align__(16) signed char s_r2[3 * TS_MAX_HS / 2 + 32]; // region, every second sample
    __shared__ int s_best[TS_THREADS / 32], s_bestd[TS_THREADS / 32];
    __shared__ long long s_p;
    const L3StretchJob jb = jobs[blockIdx.x];
    const int nch = jb.channels, Hs = jb.hop, N = 2 * Hs, R = Hs / 2;
    const T *x = in + jb.in_off;
    T *y = out + jb.out_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nseg = (jb.out_n + Hs - 1) / Hs;
    const float wstep = 3.14159265358979323846f / (float)Hs;
    long long p_prev = 0;
    for (long long m = 0; m < nseg; m++) {
        long long p = 0;
        if (m > 0) {
            const long long a = (m * Hs * (long long)num) / den, tpos = p_prev + Hs;
            for (int i = tid; i < N; i += TS_THREADS) {
                const signed char v = (signed char)ts_align<T>(x, tpos + i, jb.in_n, nch);
                s_t[i] = v;
                if (!(i & 1)) s_t2[i >> 1] = v;
            }
            for (int i = tid; i < N + 2 * R + 16; i += TS_THREADS) {
                const signed char v = (signed char)(i < N + 2 * R ? ts_align<T>(x, a - R + i, jb.in_n, nch) : 0);
                s_r[i] = v;
                if (!(i & 1)) s_r2[i >> 1] = v;
            }
            __syncthreads();
            // ---- coarse: candidates d = -R + 4 j, every second sample; region byte 4 j = decimated byte 2 j
            int best = INT_MIN, bestd = 0;
            {
                const int *tw = reinterpret_cast<const int *>(s_t2);
                for (int j = tid; 4 * j <= 2 * R; j += TS_THREADS) {
                    const unsigned *rw = reinterpret_cast<const unsigned *>(s_r2) + (j >> 1);
                    const unsigned sh = (unsigned)(j & 1) * 16u;
                    int acc = 0;
                    unsigned lo = rw[0];
                    for (int k = 0; k < N / 8; k++) {
                        const unsigned hi = rw[k + 1];
                        acc = __dp4a((int)__funnelshift_r(lo, hi, sh), tw[k], acc);
                        lo = hi;
                    }
                    if (acc > best) { best = acc; bestd = j; }
                }
            }
            auto cta_argmax = [&]() { // larger score, then smaller candidate index; result in s_best[0] / s_bestd[0]
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int ob = __shfl_xor_sync(0xffffffffu, best, o), od = __shfl_xor_sync(0xffffffffu, bestd, o);
                    if (ob > best || (ob == best && od < bestd)) { best = ob; bestd = od; }
                }
                if (lane == 0) { s_best[warp] = best; s_bestd[warp] = bestd; }
                __syncthreads();
                if (tid == 0) {
                    int bb = s_best[0], dd = s_bestd[0];
                    for (int w = 1; w < TS_THREADS / 32; w++)
                        if (s_best[w] > bb || (s_best[w] == bb && s_bestd[w] < dd)) { bb = s_best[w]; dd = s_bestd[w]; }
                    s_best[0] = bb;
                    s_bestd[0] = dd;
                }
                __syncthreads();
            };
            cta_argmax();
            const int c0 = 4 * s_bestd[0]; // coarse winner as a region byte offset (d0 = c0 - R)
            __syncthreads();
            // ---- fine: candidates c0 - 3 .. c0 + 3 (clipped to 0 .. 2 R) in full, one warp each
            best = INT_MIN;
            bestd = 0;
            if (warp < 7) {
                const int c = c0 - 3 + warp;
                if (c >= 0 && c <= 2 * R) {
                    const int *tw = reinterpret_cast<const int *>(s_t);
                    const unsigned *rw = reinterpret_cast<const unsigned *>(s_r) + (c >> 2);
                    const unsigned sh = (unsigned)(c & 3) * 8u;
                    int acc = 0;
                    for (int k = lane; k < N / 4; k += 32)
                        acc = __dp4a((int)__funnelshift_r(rw[k], rw[k + 1], sh), tw[k], acc);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    best = acc;
                    bestd = c;
                }
            }
            cta_argmax();
            if (tid == 0) {
                const int d = s_bestd[0] - R;
                s_p = a + d;
                if (offsets_out && m < max_frames) offsets_out[(size_t)blockIdx.x * max_frames + m] = d;
            }
            __syncthreads();
            p = s_p;
        } else if (tid == 0 && offsets_out && max_frames > 0)
            offsets_out[(size_t)blockIdx.x * max_frames] = 0;
        // output segment m
        const long long o0 = m * Hs;
        for (int i = tid; i < Hs * nch; i += TS_THREADS) {
            const int k = i / nch, c = i - k * nch;
            if (o0 + k >= jb.out_n) continue;
            const long long ia = p + k, ib = p_prev + Hs + k;
            const float va = (ia >= 0 && ia < jb.in_n) ? ts_f<T>(x + ia * nch + c) : 0.f;
            float v = va;
            if (m > 0) {
                const float vb = (ib >= 0 && ib < jb.in_n) ? ts_f<T>(x + ib * nch + c) : 0.f;
                const float w = 0.5f - 0.5f * cosf(wstep * (float)k);
                v = (1.f - w) * vb + w * va;
            }
            ts_store(y + (o0 + k) * nch + c, v);
        }
        p_prev = p;
        __syncthreads(); // s_t / s_r are rewritten by the next frame
    }
}

} // namespace

int l3_stretch_hop(int sample_rate) { return sample_rate >= 32000 ? 512 : (sample_rate >= 16000 ? 256 : 128); }

void l3_launch_stretch(const void *in, void *out, int pcm_format, const L3StretchJob *jobs, int njobs, int num, int den,
                       int *offsets_out, int max_frames, cudaStream_t st)
{
    if (njobs <= 0) return;
    if (pcm_format == MP3B_PCM_S16)
        k_stretch<int16_t><<<njobs, TS_THREADS, 0, st>>>(static_cast<const int16_t *>(in), static_cast<int16_t *>(out), jobs,
                                                         num, den, offsets_out, max_frames);
    else
        k_stretch<float><<<njobs, TS_THREADS, 0, st>>>(static_cast<const float *>(in), static_cast<float *>(out), jobs, num,
                                                       den, offsets_out, max_frames);
}
