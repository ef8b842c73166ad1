// k_segments.cu -- sentence boundaries of the decoded batch: the "repeat each sentence" half of the one
// playback feature the reference describes (/root/reference/README.md:46).  The
// reference has no code for it, so the definition is ours, chosen to be exactly reproducible (integers only):
//   x[n]  = the stream's PCM as s16 (float PCM is rounded to s16 first), m[n] = (L + R) >> 1 or the mono sample
//   W     = sample_rate / 100 samples (10 ms), window k = samples [k W, min((k + 1) W, N)), cnt_k its length
//   E[k]  = sum of m[n]^2 over window k                                     (64-bit, exact)
//   window k is silent  iff  E[k] <= thr^2 cnt_k      (thr = RMS threshold in s16 units)
//   a pause is a run of >= G consecutive silent windows; a sentence is a maximal run of windows that starts
//   and ends with a voiced window and holds no pause; sentences shorter than S windows are dropped.
//   Sentence [a, b) in windows is reported as samples [a W, min(b W, N)).
// Two kernels: k_window_energy reads the arena once (HBM-bound: a warp per window, aligned 16-byte loads,
// L + R by one dp2a, 64-bit multiply-add, shuffle reduction); k_find_sentences is one warp per stream over the small energy array (ballot
// per 32 windows, then the run-length state machine over the runs of the bit mask).
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "mp3b.h"

namespace {

constexpr int SG_WARPS = 8;
constexpr int SG_NW = 1; // windows a warp works on at a time (2, interleaved, measured slower: 0.40 vs 0.39 ms)

template <typename T> __device__ __forceinline__ int sg_s16(T v);
template <> __device__ __forceinline__ int sg_s16<int16_t>(int16_t v) { return v; }
template <> __device__ __forceinline__ int sg_s16<float>(float v)
{
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(v * 32768.f));
    return (int)(short)r;
}

// One frame's mono value m (s16 domain) from its elements.
template <typename T, int NCH> __device__ __forceinline__ int sg_frame(const T *p)
{
    if (NCH == 2) return (sg_s16<T>(p[0]) + sg_s16<T>(p[1])) >> 1;
    return sg_s16<T>(p[0]);
}

// Sum of m^2 over the frames of one aligned 16-byte vector.
template <typename T, int NCH> __device__ __forceinline__ void sg_vec(const uint4 v, long long &acc);
template <> __device__ __forceinline__ void sg_vec<int16_t, 2>(const uint4 v, long long &acc)
{
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = __dp2a_lo((int)w[i], 0x0101, 0) >> 1; // L + R in one instruction
        acc += (long long)m * m;
    }
}
template <> __device__ __forceinline__ void sg_vec<int16_t, 1>(const uint4 v, long long &acc)
{
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a = (int)(short)(w[i] & 0xffffu), b = (int)w[i] >> 16;
        acc += (long long)a * a;
        acc += (long long)b * b;
    }
}
template <> __device__ __forceinline__ void sg_vec<float, 2>(const uint4 v, long long &acc)
{
    const int a = (sg_s16<float>(__uint_as_float(v.x)) + sg_s16<float>(__uint_as_float(v.y))) >> 1;
    const int b = (sg_s16<float>(__uint_as_float(v.z)) + sg_s16<float>(__uint_as_float(v.w))) >> 1;
    acc += (long long)a * a;
    acc += (long long)b * b;
}
template <> __device__ __forceinline__ void sg_vec<float, 1>(const uint4 v, long long &acc)
{
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a = sg_s16<float>(__uint_as_float(w[i]));
        acc += (long long)a * a;
    }
}

// A warp per window (SG_NW = 1).  A window's frames are read as aligned 16-byte vectors (a window starts
// wherever k W frames lands, so up to FPV - 1 frames before the first and after the last whole vector are
// read one by one by the first lanes).  The energy is a 64-bit integer sum, so the order of the additions
// does not matter and the result is exact.
template <typename T, int NCH>
__device__ __forceinline__ void sg_windows(const T *__restrict__ x, const L3SegJob &jb, unsigned long long *__restrict__ energy)
{
    constexpr int FB = (int)sizeof(T) * NCH, FPV = 16 / FB, NW = SG_NW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long W = jb.window;
    for (unsigned k = (blockIdx.x * SG_WARPS + warp) * NW; k < jb.nwin; k += gridDim.x * SG_WARPS * NW) {
        const T *p[NW];
        const uint4 *pv[NW];
        int nh[NW], nv[NW], nt[NW];
        long long acc[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const long long n0 = min((long long)(k + w) * W, jb.samples), n1 = min(n0 + W, jb.samples);
            const int len = (int)(n1 - n0); // 0 for the pair's second window past the stream's end
            p[w] = x + n0 * NCH;
            nh[w] = min(len, (int)((16u - (unsigned)(reinterpret_cast<uintptr_t>(p[w]) & 15u)) & 15u) / FB);
            nv[w] = (len - nh[w]) / FPV;
            nt[w] = len - nh[w] - nv[w] * FPV;
            pv[w] = reinterpret_cast<const uint4 *>(p[w] + nh[w] * NCH);
            acc[w] = 0;
        }
        int nvmax = nv[0];
#pragma unroll
        for (int w = 1; w < NW; w++) nvmax = max(nvmax, nv[w]);
        // the trip count is the same on every lane (a per-lane bound would split the warp between the
        // unrolled loop and its remainder: measured 15.8 active threads per instruction)
#pragma unroll 4
        for (int v0 = 0; v0 < nvmax; v0 += 32) {
            const int v = v0 + lane;
#pragma unroll
            for (int w = 0; w < NW; w++)
                if (v < nv[w]) sg_vec<T, NCH>(__ldg(pv[w] + v), acc[w]);
        }
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if (lane < nh[w] + nt[w]) {
                const int f = lane < nh[w] ? lane : nh[w] + nv[w] * FPV + (lane - nh[w]);
                const int m = sg_frame<T, NCH>(p[w] + f * NCH);
                acc[w] += (long long)m * m;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int w = 0; w < NW; w++) acc[w] += __shfl_xor_sync(0xffffffffu, acc[w], o);
        }
        if (lane < NW && k + lane < jb.nwin) energy[jb.win_base + k + lane] = (unsigned long long)(lane ? acc[NW - 1] : acc[0]);
    }
}

template <typename T>
__global__ void __launch_bounds__(SG_WARPS * 32)
k_window_energy(const T *__restrict__ pcm, const L3SegJob *__restrict__ jobs, int njobs, unsigned long long *__restrict__ energy)
{
    // grid: x = window tile (fastest: CTAs launched together read neighbouring memory), y / z = stream
    const unsigned j = blockIdx.z * 65535u + blockIdx.y;
    if (j >= (unsigned)njobs) return;
    const L3SegJob jb = jobs[j];
    if (jb.channels == 2) sg_windows<T, 2>(pcm + jb.off, jb, energy); // pcm offsets of stereo streams are even
    else sg_windows<T, 1>(pcm + jb.off, jb, energy);
}

__global__ void __launch_bounds__(128)
k_find_sentences(const L3SegJob *__restrict__ jobs, int njobs, const unsigned long long *__restrict__ energy,
                 unsigned long long thr2, int G, int S, long long *__restrict__ seg, int *__restrict__ nseg)
{
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= njobs) return;
    const L3SegJob jb = jobs[j];
    const unsigned long long *E = energy + jb.win_base;
    const long long W = jb.window;
    bool in_seg = false;
    unsigned start = 0, last_voiced = 0, count = 0;
    int silent_run = 0;
    auto emit = [&]() {
        if ((int)(last_voiced - start) >= S) {
            if (lane == 0 && count < jb.seg_cap) {
                seg[2 * (size_t)(jb.seg_base + count)] = (long long)start * W;
                seg[2 * (size_t)(jb.seg_base + count) + 1] = min((long long)last_voiced * W, jb.samples);
            }
            count++;
        }
        in_seg = false;
    };
    // 128 windows per trip (four loads in flight per lane), one ballot per 32; the state machine then walks
    // the RUNS of each mask (count-trailing-zeros), not its bits: the same on every lane
    for (unsigned k0 = 0; k0 < jb.nwin; k0 += 128) {
        unsigned masks[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned k = k0 + 32u * q + lane;
            bool voiced = false;
            if (k < jb.nwin) {
                const long long cnt = min(W, jb.samples - (long long)k * W);
                voiced = __ldg(E + k) > thr2 * (unsigned long long)cnt;
            }
            masks[q] = __ballot_sync(0xffffffffu, voiced);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned kq = k0 + 32u * q;
            if (kq >= jb.nwin) break;
            const int nb = (int)min(32u, jb.nwin - kq);
            int pos = 0;
            while (pos < nb) {
                const unsigned rem = masks[q] >> pos;
                if (rem & 1u) { // a run of voiced windows
                    const int run = min(nb - pos, __ffs((int)~rem) ? __ffs((int)~rem) - 1 : 32);
                    if (!in_seg) { in_seg = true; start = kq + pos; }
                    last_voiced = kq + pos + run;
                    silent_run = 0;
                    pos += run;
                } else { // a run of silent windows
                    const int run = min(nb - pos, rem ? __ffs((int)rem) - 1 : 32);
                    if (in_seg) {
                        silent_run += run;
                        if (silent_run >= G) emit();
                    }
                    pos += run;
                }
            }
        }
    }
    if (in_seg) emit();
    if (lane == 0) nseg[j] = (int)min(count, jb.seg_cap);
}

} // namespace

void l3_launch_segments(const void *pcm, int pcm_format, const L3SegJob *jobs, int njobs, unsigned max_nwin,
                        unsigned long long *energy, unsigned long long thr2, int G, int S, long long *seg, int *nseg,
                        cudaStream_t st)
{
    if (njobs <= 0) return;
    static const int wpw = [] { // window groups per warp (tuning override, read once, thread-safe)
        const char *e = getenv("MP3B_SG_WPW");
        return e ? std::max(1, atoi(e)) : 8;
    }();
    const unsigned per_cta = (unsigned)(SG_WARPS * SG_NW * wpw);
    const unsigned nx = std::min(65535u, std::max(1u, (max_nwin + per_cta - 1) / per_cta));
    const dim3 grid(nx, (unsigned)std::min(njobs, 65535), (unsigned)((njobs + 65534) / 65535));
    if (pcm_format == MP3B_PCM_S16)
        k_window_energy<int16_t><<<grid, SG_WARPS * 32, 0, st>>>(static_cast<const int16_t *>(pcm), jobs, njobs, energy);
    else
        k_window_energy<float><<<grid, SG_WARPS * 32, 0, st>>>(static_cast<const float *>(pcm), jobs, njobs, energy);
    k_find_sentences<<<(njobs + 3) / 4, 128, 0, st>>>(jobs, njobs, energy, thr2, G, S, seg, nseg);
}
