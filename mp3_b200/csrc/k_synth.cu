// k_synth.cu -- K4: polyphase synthesis filterbank (a11).
//
// Per slot of 32 subband samples S[k] the standard computes V[i] = sum_k cos((16+i)(2k+1)pi/64) S[k]
// (64x32), shifts V into a 1024-entry FIFO and forms 32 PCM samples as a 512-tap dot product with
// the window D.  All 64 V values are signed copies of the 32-point transform
//     C[n] = sum_k cos(n (2k+1) pi / 64) S[k],  n = 0..31
// (V[i] = C[16+i] for i < 16, V[16] = 0, V[i] = -C[48-i] for 17 <= i <= 47, V[48] = -C[0],
//  V[i] = -C[i-48] for i >= 49), so the matrixing is a batched [slots x 32] x [32 x 32] product and
// the FIFO is never materialised: PCM of slot T is
//     pcm[j] = sum_{l<16} W[l][j] * C_{T-l}[src_{l&1}[j]]
// with the signs folded into the rearranged window W.  A tile (one stream, G granules, both
// channels) keeps its C vectors in shared memory; the 15-slot history comes from the previous
// granule's subband samples (re-transformed: a 10 % halo at G = 8), or is zero at stream start.
// The 32-point transform is the in-register fast DCT-II of fast_dct.h (304 operations per slot instead
// of 1024 multiply-adds), one thread per (channel, slot) row; rows have an odd stride so that both the
// row-per-lane transform and the column-per-lane window are free of bank conflicts.  This kernel is the
// product path of Layer II (k_layer2.cu) and the last stage of the staged Layer III pipeline.
// FP32 FMA throughout; the window is a sliding accumulation (a warp per granule and channel: each of the 33
// rows it touches is loaded once), the subband samples come in by 16-byte loads.
// Restates oracle/l3_oracle.c::synth_slot in float32.
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include <math.h>

#include "fast_dct.h"
#include "iso_tables.h"
#include "kernels.h"
#include "mp3b.h"

namespace {

constexpr int K4_G = 8;                 // granules per tile
constexpr int K4_SLOTS = 15 + K4_G * 18; // with history
constexpr int K4_THREADS = 256;
constexpr int K4_FS = 33;               // row stride

__device__ float g_synwin[16][32]; // W[l][j]

__device__ __forceinline__ int16_t to_s16(float v)
{
    float s = rintf(v * 32768.f);
    s = fminf(fmaxf(s, -32768.f), 32767.f);
    return (int16_t)s;
}

template <int FMT>
__global__ void __launch_bounds__(K4_THREADS)
k_synth(const uint2 *__restrict__ tiles, uint32_t ntiles, const uint32_t *__restrict__ gran_unit0,
        const float *__restrict__ sb_all, const uint32_t *__restrict__ tile_shift, void *__restrict__ pcm)
{
    extern __shared__ float s_c[]; // [nch][K4_SLOTS][K4_FS]
    const uint32_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    const uint2 tl = tiles[tile];
    const uint32_t g0 = tl.x, ng = tl.y & 0xffu;
    const uint32_t gu0 = gran_unit0[g0];
    const bool first = (gu0 & L3G_FIRST) != 0;
    const int nch = (gu0 & L3G_STEREO) ? 2 : 1;
    const uint32_t u0 = gu0 & L3G_UNIT_MASK;
    // Layer I / II: the subband samples live in a dense buffer of those streams only (unit - shift)
    const float *sb = sb_all - (tile_shift ? (size_t)tile_shift[tile] * 576 : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = K4_THREADS / 32;
    const int nslots = 15 + (int)ng * 18;

    // ---- load subband samples: history (last 15 slots of the previous granule) + ng granules.
    // units of this stream are contiguous: unit(g, c) = u0 + (g - g0) * nch + c
    // (16-byte loads: a unit's 576 samples are contiguous and 16-byte aligned; the padded rows take scalar stores)
    for (int c = 0; c < nch; c++) {
        float *dst = s_c + c * (K4_SLOTS * K4_FS);
        for (int i = threadIdx.x; i < 15 * 8; i += K4_THREADS) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!first) v = __ldg(reinterpret_cast<const float4 *>(sb + (size_t)(u0 - nch + c) * 576 + 3 * 32) + i);
            float *o = dst + (i >> 3) * K4_FS + (i & 7) * 4;
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
        for (int i = threadIdx.x; i < (int)ng * 144; i += K4_THREADS) {
            const int gi = i / 144, r = i - gi * 144;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(sb + (size_t)(u0 + gi * nch + c) * 576) + r);
            float *o = dst + (15 + gi * 18 + (r >> 3)) * K4_FS + (r & 7) * 4;
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
    }
    __syncthreads();

    // ---- matrixing: C[n] = sum_k S[k] cos(n(2k+1)pi/64), in place, one thread per (channel, slot) row
    for (int it = threadIdx.x; it < nch * nslots; it += K4_THREADS) {
        float *row = s_c + (it / nslots) * (K4_SLOTS * K4_FS) + (it % nslots) * K4_FS;
        float x[32];
#pragma unroll
        for (int k = 0; k < 32; k++) x[k] = row[k];
        L3Dct2<32>::run(x);
#pragma unroll
        for (int k = 0; k < 32; k++) row[k] = x[k];
    }
    __syncthreads();

    // ---- windowing: a warp per (granule, channel); sliding accumulation over the 33 rows a granule's 18
    // output slots touch: each row is loaded once (2 LDS) and feeds the up to 16 outputs it contributes to,
    // the 16-entry accumulator ring and the window taps live in registers (everything is unrolled)
    {
        float w[16];
#pragma unroll
        for (int l = 0; l < 16; l++) w[l] = g_synwin[l][lane];
        const int src_e = lane <= 15 ? 16 + lane : (lane == 16 ? 0 : 48 - lane); // lane 16: weight is 0
        const int src_o = lane <= 16 ? 16 - lane : lane - 16;
        for (int task = warp; task < (int)ng * nch; task += nwarps) {
            const int gi = task / nch, c = task - gi * nch;
            const float *Fc = s_c + c * (K4_SLOTS * K4_FS) + gi * 18 * K4_FS; // rows T0 - 15 .. T0 + 17 of this granule
            // PCM element index of (granule gi, slot 0, sample lane, channel c)
            const size_t e0 = (size_t)(u0 + gi * nch) * 576 + (size_t)lane * nch + c;
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = 0.f;
#pragma unroll
            for (int q = 0; q < 33; q++) {
                const float e = Fc[q * K4_FS + src_e], o = Fc[q * K4_FS + src_o];
#pragma unroll
                for (int l = 0; l < 16; l++) {
                    const int T = q + l - 15;
                    if (T >= 0 && T < 18) acc[T & 15] = fmaf(w[l], (l & 1) ? o : e, acc[T & 15]);
                }
                if (q >= 15) {
                    const float val = acc[(q - 15) & 15];
                    acc[(q - 15) & 15] = 0.f;
                    const size_t e1 = e0 + (size_t)(q - 15) * 32 * nch;
                    if (FMT == MP3B_PCM_S16) reinterpret_cast<int16_t *>(pcm)[e1] = to_s16(val);
                    else reinterpret_cast<float *>(pcm)[e1] = val;
                }
            }
        }
    }
}

} // namespace

int l3_synth_tile_granules(void) { return K4_G; }

void l3_synth_init(void)
{
    static float win[16][32];
    for (int l = 0; l < 16; l++)
        for (int j = 0; j < 32; j++) {
            const int i = l >> 1;
            double v;
            if (!(l & 1)) v = l3_dwin(64 * i + j) * (j <= 15 ? 1.0 : (j == 16 ? 0.0 : -1.0));
            else v = -l3_dwin(64 * i + 32 + j);
            win[l][j] = (float)v;
        }
    cudaMemcpyToSymbol(g_synwin, win, sizeof win);
}

void l3_launch_synth(const uint2 *tiles, uint32_t ntiles, const uint32_t *gran_unit0, const float *sb,
                     const uint32_t *tile_shift, void *pcm, int pcm_format, cudaStream_t st)
{
    if (!ntiles) return;
    const size_t smem = (size_t)2 * K4_SLOTS * K4_FS * sizeof(float);
    static std::atomic<unsigned long long> attr{0};
    if (l3_device_needs_setup(attr)) {
        cudaFuncSetAttribute(k_synth<MP3B_PCM_S16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_synth<MP3B_PCM_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        l3_device_setup_done(attr);
    }
    if (pcm_format == MP3B_PCM_S16)
        k_synth<MP3B_PCM_S16><<<ntiles, K4_THREADS, smem, st>>>(tiles, ntiles, gran_unit0, sb, tile_shift, pcm);
    else
        k_synth<MP3B_PCM_F32><<<ntiles, K4_THREADS, smem, st>>>(tiles, ntiles, gran_unit0, sb, tile_shift, pcm);
}
