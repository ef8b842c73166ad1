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
// one frame (1 or 2 channels) as floats; stereo frames are element-pair aligned (pcm offsets of stereo streams are even)
template <typename T> __device__ __forceinline__ void ts_load_frame(const T *p, int nch, float &l, float &r);
template <> __device__ __forceinline__ void ts_load_frame<int16_t>(const int16_t *p, int nch, float &l, float &r)
{
    if (nch == 2) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
        l = (float)(short)(w & 0xffffu); // (s16 units: the cross-fade is linear, the scale cancels)
        r = (float)((int)w >> 16);
    } else
        l = (float)*p;
}
template <> __device__ __forceinline__ void ts_load_frame<float>(const float *p, int nch, float &l, float &r)
{
    if (nch == 2) {
        const float2 w = *reinterpret_cast<const float2 *>(p);
        l = w.x;
        r = w.y;
    } else
        l = *p;
}
__device__ __forceinline__ void ts_store_frame(int16_t *p, int nch, float l, float r)
{
    int a, b;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(a) : "f"(l));
    if (nch == 2) {
        asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(b) : "f"(r));
        *reinterpret_cast<uint32_t *>(p) = ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16);
    } else
        *p = (int16_t)a;
}
__device__ __forceinline__ void ts_store_frame(float *p, int nch, float l, float r)
{
    if (nch == 2) *reinterpret_cast<float2 *>(p) = make_float2(l, r);
    else *p = l;
}

// L + R of one stereo frame in the s16 domain (the alignment signal is (L + R) >> 9)
template <typename T> __device__ __forceinline__ int ts_pair_sum(const T *p) { return ts_s16<T>(p) + ts_s16<T>(p + 1); }
template <> __device__ __forceinline__ int ts_pair_sum<int16_t>(const int16_t *p) // stereo frames are 4-byte aligned
{
    return __dp2a_lo((int)*reinterpret_cast<const uint32_t *>(p), 0x0101, 0);
}

template <typename T>
__global__ void __launch_bounds__(TS_THREADS)
k_stretch(const T *__restrict__ in, T *__restrict__ out, const L3StretchJob *__restrict__ jobs, int num, int den,
          int *__restrict__ offsets_out /* optional: chosen d per frame, [job][max_frames] */, int max_frames)
{
    __shared__ __align__(16) signed char s_t[2 * TS_MAX_HS + 16];      // template: c[p_prev + Hs + k], k < N
    __shared__ __align__(16) signed char s_r[4 * TS_MAX_HS + 16 + 16]; // c[a - R + i]: the region (i < N + 2 R) and, behind it,
                                                                       // what the template of a slowed-down stream reaches
    __shared__ __align__(16) signed char s_t2[TS_MAX_HS + 16];         // template, every second sample
    __shared__ __align__(16) signed char s_r2[3 * TS_MAX_HS / 2 + 32]; // region, every second sample
    __shared__ unsigned long long s_key[2]; // arg max of the coarse / fine search: score (order-preserving) << 32 | ~candidate
    __shared__ long long s_p;
    __shared__ float s_w[TS_MAX_HS]; // the cross-fade window, once per CTA
    const L3StretchJob jb = jobs[blockIdx.x];
    const int nch = jb.channels, Hs = jb.hop, N = 2 * Hs, R = Hs / 2;
    const T *x = in + jb.in_off;
    T *y = out + jb.out_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nseg = (jb.out_n + Hs - 1) / Hs;
    const float wstep = 3.14159265358979323846f / (float)Hs;
    for (int k = tid; k < Hs; k += TS_THREADS) s_w[k] = 0.5f - 0.5f * cosf(wstep * (float)k);
    __syncthreads();
    long long p_prev = 0;
    // nominal start a_m = floor(m Hs num / den), advanced exactly without a 64-bit division per segment
    const int astep = Hs * num, aq = astep / den, ar = astep % den;
    long long a = 0;
    int arem = 0;
    // The search region of a segment overlaps the previous one's (by 83 % at half speed), and the template -- the
    // natural continuation of the previous frame -- starts (Hs - advance) + d + R bytes into it: the alignment signal
    // is therefore kept in shared memory from segment to segment as ONE run c[ra .. ra + ULEN) that covers both, moved
    // by the advance of `a`, and only its new tail is converted from the PCM: a few hundred samples per segment instead
    // of 2,576.  (Speeds above 1 put the template in front of the run: it is then converted on its own, as before.)
    const int ULEN = 4 * Hs + 16;    // region N + 2 R = 3 Hs, template reach up to Hs more, 16 bytes for the word reads
    long long ra = 0;                // s_r holds c[ra .. ra + ULEN)
    bool have_run = false;
    for (long long m = 0; m < nseg; m++) {
        long long p = 0;
        if (m > 0) {
            a += aq;
            arem += ar;
            if (arem >= den) { arem -= den; a += 1; }
            const long long tpos = p_prev + Hs, ra_new = a - R;
            // alignment signal c[start + i], i in [i0, i1), into dst[i] (zero outside the stream)
            auto fill = [&](signed char *dst, long long start, int i0, int i1) {
                const int lo = (int)max((long long)i0, min(-start, (long long)i1));
                const int hi = (int)max((long long)i0, min(jb.in_n - start, (long long)i1));
                for (int i = i0 + tid; i < i1; i += TS_THREADS) {
                    int v = 0;
                    if (i >= lo && i < hi) {
                        const T *q = x + (start + i) * nch;
                        v = nch == 2 ? ts_pair_sum<T>(q) >> 9 : ts_s16<T>(q) >> 8;
                        v = max(-127, min(127, v));
                    }
                    dst[i] = (signed char)v;
                }
            };
            unsigned *uw = reinterpret_cast<unsigned *>(s_r);
            if (tid == 0) s_key[0] = s_key[1] = 0ull; // (read last behind the previous segment's searches; next written after two barriers)
            // ---- (A) move the run by delta = ra_new - ra bytes (through registers: it overlaps itself), (B) convert the
            // new tail
            const long long delta = ra_new - ra;
            const bool shift = have_run && delta >= 0 && delta < ULEN;
            const int keepw = shift ? (int)((ULEN - delta) >> 2) : 0; // whole words that survive
            constexpr int UW = (4 * TS_MAX_HS + 16) / 4 / TS_THREADS + 1; // words per thread
            unsigned kept[UW];
            {
                const int w0 = (int)(delta >> 2);
                const unsigned sh = (unsigned)(delta & 3) * 8u;
#pragma unroll
                for (int q = 0; q < UW; q++) {
                    const int i = tid + q * TS_THREADS;
                    kept[q] = i < keepw ? __funnelshift_r(uw[w0 + i], uw[w0 + i + 1], sh) : 0u;
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < UW; q++) {
                const int i = tid + q * TS_THREADS;
                if (i < keepw) uw[i] = kept[q];
            }
            fill(s_r, ra_new, keepw * 4, ULEN);
            ra = ra_new;
            have_run = true;
            __syncthreads();
            // ---- (C) the template out of the run (or from the PCM when it is not inside), every-second-sample copies
            const long long toff = tpos - ra_new;
            const bool t_in_run = toff >= 0 && toff + N <= ULEN - 4;
            if (t_in_run) {
                const int w0 = (int)(toff >> 2);
                const unsigned sh = (unsigned)(toff & 3) * 8u;
                for (int w = tid; w < N / 4; w += TS_THREADS)
                    reinterpret_cast<unsigned *>(s_t)[w] = __funnelshift_r(uw[w0 + w], uw[w0 + w + 1], sh);
                for (int w = tid; w < N / 8; w += TS_THREADS) {
                    const unsigned t0 = __funnelshift_r(uw[w0 + 2 * w], uw[w0 + 2 * w + 1], sh);
                    const unsigned t1 = __funnelshift_r(uw[w0 + 2 * w + 1], uw[w0 + 2 * w + 2], sh);
                    reinterpret_cast<unsigned *>(s_t2)[w] = __byte_perm(t0, t1, 0x6420);
                }
            } else {
                fill(s_t, tpos, 0, N);
                __syncthreads();
                const unsigned *tw = reinterpret_cast<const unsigned *>(s_t);
                for (int w = tid; w < N / 8; w += TS_THREADS) reinterpret_cast<unsigned *>(s_t2)[w] = __byte_perm(tw[2 * w], tw[2 * w + 1], 0x6420);
            }
            for (int w = tid; w < (N + 2 * R + 16) / 8; w += TS_THREADS) reinterpret_cast<unsigned *>(s_r2)[w] = __byte_perm(uw[2 * w], uw[2 * w + 1], 0x6420);
            __syncthreads();
            // ---- coarse: candidates d = -R + 4 j, every second sample; region byte 4 j = decimated byte 2 j.  A thread
            // scores the candidates 2 i and 2 i + 1: they read the same words of the decimated region (byte offsets 4 i and
            // 4 i + 2), so every load serves two dot products.
            int best = INT_MIN, bestd = 0;
            const int ncand = R / 2 + 1; // 4 j <= 2 R
            const int i = tid >> 1, h = tid & 1;
            {
                // (the dot products are the serial part of a segment: each pair of candidates is split over two adjacent
                // lanes -- the halves of the template -- and every lane keeps two accumulators per candidate)
                const int *tw = reinterpret_cast<const int *>(s_t2);
                const int kh = N / 16; // words per half
                int a0 = 0, a1 = 0, b0 = 0, b1 = 0;
                if (2 * i < ncand) {
                    const unsigned *rw = reinterpret_cast<const unsigned *>(s_r2) + i + h * kh;
                    const int *tq = tw + h * kh;
                    unsigned lo = rw[0];
#pragma unroll 4
                    for (int k = 0; k < kh; k += 2) { // kh is 16, 32 or 64
                        const unsigned mid = rw[k + 1], hi = rw[k + 2];
                        const int t0 = tq[k], t1 = tq[k + 1];
                        a0 = __dp4a((int)lo, t0, a0);
                        a1 = __dp4a((int)__funnelshift_r(lo, mid, 16u), t0, a1);
                        b0 = __dp4a((int)mid, t1, b0);
                        b1 = __dp4a((int)__funnelshift_r(mid, hi, 16u), t1, b1);
                        lo = hi;
                    }
                }
                int acc0 = a0 + b0, acc1 = a1 + b1;
                acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); // (all lanes: the pair's two halves)
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
                if (h == 0 && 2 * i < ncand) {
                    if (acc0 > best) { best = acc0; bestd = 2 * i; }               // (first maximum: the smaller index wins ties)
                    if (2 * i + 1 < ncand && acc1 > best) { best = acc1; bestd = 2 * i + 1; }
                }
            }
            // arg max over the CTA: larger score, then smaller candidate index, as one unsigned 64-bit maximum
            {
                const bool has = best != INT_MIN || (h == 0 && 2 * i < ncand);
                const unsigned us = has ? (unsigned)best ^ 0x80000000u : 0u;
                if (warp * 32 < 2 * ((ncand + 1) / 2)) { // (warp-uniform: the warps that hold candidates)
                    const unsigned mx = __reduce_max_sync(0xffffffffu, us);
                    const unsigned mi = __reduce_min_sync(0xffffffffu, has && us == mx ? (unsigned)bestd : 0xffffffffu);
                    if (lane == 0) atomicMax(&s_key[0], ((unsigned long long)mx << 32) | (0xffffffffu - mi));
                }
            }
            __syncthreads();
            const int c0 = 4 * (int)(0xffffffffu - (unsigned)(s_key[0] & 0xffffffffull)); // coarse winner as a region byte offset (d0 = c0 - R)
            // ---- fine: candidates c0 - 3 .. c0 + 3 (clipped to 0 .. 2 R) in full, one warp each
            if (warp < 7) {
                const int c = c0 - 3 + warp;
                if (c >= 0 && c <= 2 * R) {
                    const int *tw = reinterpret_cast<const int *>(s_t);
                    const unsigned *rw = reinterpret_cast<const unsigned *>(s_r) + (c >> 2);
                    const unsigned sh = (unsigned)(c & 3) * 8u;
                    int acc = 0;
                    for (int k = lane; k < N / 4; k += 32)
                        acc = __dp4a((int)__funnelshift_r(rw[k], rw[k + 1], sh), tw[k], acc);
                    acc = __reduce_add_sync(0xffffffffu, acc);
                    if (lane == 0)
                        atomicMax(&s_key[1], ((unsigned long long)((unsigned)acc ^ 0x80000000u) << 32) | (0xffffffffu - (unsigned)c));
                }
            }
            __syncthreads();
            if (tid == 0) {
                const int d = (int)(0xffffffffu - (unsigned)(s_key[1] & 0xffffffffull)) - R;
                s_p = a + d;
                if (offsets_out && m < max_frames) offsets_out[(size_t)blockIdx.x * max_frames + m] = d;
            }
            __syncthreads();
            p = s_p;
        } else if (tid == 0 && offsets_out && max_frames > 0)
            offsets_out[(size_t)blockIdx.x * max_frames] = 0;
        // output segment m
        const long long o0 = m * Hs;
        {
            // a thread per output frame (both channels); valid input ranges as 32-bit offsets, once per segment
            const int kmax = (int)min((long long)Hs, jb.out_n - o0);
            const long long sa = p, sb = p_prev + Hs;
            const int alo = (int)max(0ll, min(-sa, (long long)Hs)), ahi = (int)max(0ll, min(jb.in_n - sa, (long long)Hs));
            const int blo = (int)max(0ll, min(-sb, (long long)Hs)), bhi = (int)max(0ll, min(jb.in_n - sb, (long long)Hs));
            for (int k = tid; k < kmax; k += TS_THREADS) {
                float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
                if (k >= alo && k < ahi) ts_load_frame<T>(x + (sa + k) * nch, nch, a0, a1);
                float v0 = a0, v1 = a1;
                if (m > 0) {
                    if (k >= blo && k < bhi) ts_load_frame<T>(x + (sb + k) * nch, nch, b0, b1);
                    const float w = s_w[k];
                    v0 = fmaf(w, a0 - b0, b0); // (1 - w) b + w a
                    v1 = fmaf(w, a1 - b1, b1);
                }
                ts_store_frame(y + (o0 + k) * nch, nch, v0, v1);
            }
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
