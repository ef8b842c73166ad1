// k_fused.cu -- KF: the fused back end (a6-a11): requantise + stereo + reorder + alias reduction +
// IMDCT/window + overlap-add/frequency inversion + polyphase synthesis in ONE kernel.
//
// Why: as separate kernels these stages move 21 kB per unit through HBM (17.5x the algorithmic
// minimum); fused, a unit costs its 1152-byte int16 spectrum + 40 scalefactor
// bytes in and 1152 bytes of s16 PCM out, and every intermediate lives in shared memory.
//
// Mapping.  A CTA (256 threads) owns a tile = up to `G` consecutive granules of one stream, both
// channels, and walks it in batches of 4 granules.  The only state carried from batch to batch is
// what the signal flow really needs: the second IMDCT half of the last granule (overlap-add) and
// the last 15 transformed slots (the synthesis FIR reaches 15 slots back).  A tile that does not
// start at the head of its stream first re-derives that state from the two preceding granules
// (warm-up: decoded, not output), so tiles are independent and a long stream is time-parallel.
//   S1  4 groups of 64 threads, one granule each: gains, requantise, intensity/MS decisions,
//       reorder, alias butterflies                                  -> X  (padded rows of 19)
//   S2  8 warps = 4 granules x 2 channels, lane = subband: 36-point IMDCT by its two symmetries
//       (324 FMA, coefficients as constant-bank operands), window, frequency inversion
//                                                                   -> F (first halves), H (second)
//   S3  warp per (channel, slot), lane = output index n: S = F + H(previous granule);
//       C[n] = sum_k S[k] cos(n(2k+1)pi/64) in place (the 64 "V" values are signed copies of C)
//   S4  warp per slot, lane = sample j, both channels: 16-tap dot product with the rearranged
//       window; (L, R) pairs stored interleaved, one full 128-byte row per store
// Results must equal the staged pipeline's (tests/test_gpu_parity.py runs both).
// No reference code exists for these stages (/root/reference/README.md:1-84).
#include <math.h>

#include "iso_tables.h"
#include "kernels.h"
#include "mp3b.h"

namespace {

constexpr int KF_B = 4;         // granules per batch
constexpr int KF_THREADS = 256;
constexpr int KF_ROWS = 15 + KF_B * 18;
constexpr int XROW = 19;        // padded subband row of X
constexpr int XSZ = 32 * XROW;  // 608 floats per channel spectrum

__constant__ float f_pow2q[4];
__constant__ float f_is_kl[7], f_is_kr[7];
__constant__ float f_lsf_pow[2][16];
__constant__ float f_cs[8], f_ca[8];
__constant__ uint8_t f_pretab[22];
__constant__ float f_cosA[9][18], f_cosB[9][18], f_cos12[12][6], f_win[4][36];
__device__ float f_dct32[32][32];  // [k][n]
__device__ float f_synwin[16][32];

struct GroupScratch {
    float gain[2][40];
    float kl[40], kr[40];
    int nz[40];
    uint8_t mode[40];
};

struct FusedShared {
    float X[KF_B][2][XSZ];          // spectra of the batch, padded rows
    float F[2][KF_ROWS][32];        // rows 0..14: C history; rows 15..: first IMDCT halves -> S -> C
    float H[2][KF_B + 1][18][32];   // second IMDCT halves; [0] = last granule of the previous batch
    GroupScratch gs[KF_B];
};

__device__ __forceinline__ int xpad(int i) { return i + i / 18; }

__device__ __forceinline__ void group_sync(int group)
{
    asm volatile("bar.sync %0, 64;" ::"r"(group + 1));
}

__device__ __forceinline__ int16_t to_s16(float v)
{
    float s = rintf(v * 32768.f);
    s = fminf(fmaxf(s, -32768.f), 32767.f);
    return (int16_t)s;
}

// ---- S1: one granule, executed by one 64-thread group -------------------------------------------
__device__ __forceinline__ void stage_requant(FusedShared &S, int grp, int t64, uint32_t u0, int nch,
                                              const L3UnitDesc *__restrict__ units, const int16_t *__restrict__ is_in,
                                              const uint8_t *__restrict__ sf_in, const L3BandTables *__restrict__ bands,
                                              const float *__restrict__ pow43, float *tmp /* [2][576] scratch */)
{
    GroupScratch &G = S.gs[grp];
    const L3UnitDesc d0 = units[u0], d1 = units[u0 + (nch - 1)];
    const int row = (d0.hdr >> L3H_SR_SHIFT) & 7;
    const int lay0 = (d0.flags & L3F_BT_MASK) == 2 ? ((d0.flags & L3F_MIXED) ? 2 : 1) : 0;
    const int lay1 = (d1.flags & L3F_BT_MASK) == 2 ? ((d1.flags & L3F_MIXED) ? 2 : 1) : 0;

    for (int t = t64; t < 80; t += 64) {
        const int c = t / 40, b = t % 40;
        if (c < nch) {
            const L3UnitDesc &dd = c ? d1 : d0;
            const int lay = c ? lay1 : lay0;
            float gn = 0.f;
            if (b < bands->nbands[row][lay]) {
                const int s = sf_in[(size_t)(u0 + c) * 40 + b] & 0x7f;
                const int win = bands->win[row][lay][b];
                const int sh = (dd.flags & L3F_SFSCALE) ? 4 : 2;
                int q = (int)dd.global_gain - 210;
                if (win < 0) q -= sh * (s + ((dd.flags & L3F_PREFLAG) ? f_pretab[bands->sfb[row][lay][b]] : 0));
                else q -= 8 * dd.sbg[win] + sh * s;
                gn = ldexpf(f_pow2q[q & 3], q >> 2);
            }
            G.gain[c][b] = gn;
        }
    }
    if (t64 < 40) { G.nz[t64] = 0; G.mode[t64] = 0; }
    group_sync(grp);

    for (int c = 0; c < nch; c++) {
        const int16_t *is = is_in + (size_t)(u0 + c) * 576;
        const uint8_t *l2b = bands->line2band[row][c ? lay1 : lay0];
        for (int i = t64; i < 576; i += 64) {
            const int v = is[i], b = l2b[i];
            const float a = __ldg(pow43 + (v < 0 ? -v : v)) * G.gain[c][b];
            tmp[c * 576 + i] = v < 0 ? -a : a;
            if (c == 1 && v != 0) G.nz[b] = 1;
        }
    }
    group_sync(grp);

    const bool ms = (d0.hdr & L3H_MS) != 0, ist = (d0.hdr & L3H_IS) != 0;
    const bool ok = (d0.flags & L3F_VALID) != 0;
    if (nch == 2 && ist && ok && t64 == 0) {
        const int nb = bands->nbands[row][lay1];
        const bool lsf = (d1.hdr & L3H_LSF) != 0;
        const uint8_t *sf1 = sf_in + (size_t)(u0 + 1) * 40;
        int found[3] = {0, 0, 0}, found_long = 0;
        bool first_long = true;
        for (int b = nb - 1; b >= 0; b--) {
            const int w = bands->win[row][lay1][b];
            int *fnd;
            if (w >= 0) fnd = &found[w];
            else {
                if (first_long) { found_long = found[0] | found[1] | found[2]; first_long = false; }
                fnd = &found_long;
            }
            if (*fnd) continue;
            if (G.nz[b]) { *fnd = 1; continue; }
            const int sfb = bands->sfb[row][lay1][b];
            int bsf = b;
            if (w >= 0 && sfb == 12) bsf = b - 3;
            if (w < 0 && sfb == 21) bsf = b - 1;
            const int p = sf1[bsf];
            if (!lsf) {
                if (p < 7) { G.mode[b] = 1; G.kl[b] = f_is_kl[p]; G.kr[b] = f_is_kr[p]; }
            } else if (!(p & 0x80)) {
                const int j = d1.sfc & 1;
                G.mode[b] = 1;
                G.kl[b] = (p & 1) ? f_lsf_pow[j][(p + 1) >> 1] : 1.f;
                G.kr[b] = (p & 1) ? 1.f : f_lsf_pow[j][p >> 1];
            }
        }
    }
    group_sync(grp);

    const float isq2 = 0.70710678118654752440f;
    const bool joint = nch == 2 && ok && (ms || ist);
    float *X0 = S.X[grp][0], *X1 = S.X[grp][1];
    for (int i = t64; i < 576; i += 64) {
        float l = tmp[i], r = nch == 2 ? tmp[576 + i] : 0.f;
        if (joint) {
            const int b = bands->line2band[row][lay1][i];
            if (ist && G.mode[b]) { const float a = l; l = a * G.kl[b]; r = a * G.kr[b]; }
            else if (ms) { const float a = l, c = r; l = (a + c) * isq2; r = (a - c) * isq2; }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (c >= nch) break;
            const int lay = c ? lay1 : lay0;
            int dst = i;
            if (lay != 0) {
                const int b = bands->line2band[row][lay][i];
                const int w = bands->win[row][lay][b];
                if (w >= 0) {
                    const int wd = bands->width[row][lay][b], s = bands->start[row][lay][b];
                    dst = (s - w * wd) + 3 * (i - s) + w;
                }
            }
            (c ? X1 : X0)[xpad(dst)] = c ? r : l;
        }
    }
    group_sync(grp);

    for (int t = t64; t < 248 * nch; t += 64) {
        const int c = t / 248, k = t % 248, sb = 1 + (k >> 3), i = k & 7;
        const uint8_t fl = c ? d1.flags : d0.flags;
        const int nb = (fl & L3F_BT_MASK) == 2 ? ((fl & L3F_MIXED) ? 1 : 0) : 31;
        if (sb <= nb) {
            float *Xc = c ? X1 : X0;
            const int ilo = (sb - 1) * XROW + 17 - i, ihi = sb * XROW + i;
            const float lo = Xc[ilo], hi = Xc[ihi];
            Xc[ilo] = lo * f_cs[i] - hi * f_ca[i];
            Xc[ihi] = hi * f_cs[i] + lo * f_ca[i];
        }
    }
}

// ---- S2: IMDCT of one (granule, channel) by one warp, lane = subband ----------------------------
__device__ __forceinline__ void stage_imdct(const float *__restrict__ X, int lane, uint8_t flags,
                                            float *__restrict__ Fdst /* [18][32] */, float *__restrict__ Hdst /* [18][32] */)
{
    float x[18];
#pragma unroll
    for (int k = 0; k < 18; k++) x[k] = X[lane * XROW + k];
    int bt = flags & L3F_BT_MASK;
    if (bt == 2 && (flags & L3F_MIXED) && lane < 2) bt = 0;
    const float sgn = (lane & 1) ? -1.f : 1.f; // frequency inversion: odd subband, odd slot
    if (bt != 2) {
        const float *w = f_win[bt];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int k = 0; k < 18; k++) {
                sa = fmaf(x[k], f_cosA[i][k], sa);
                sb = fmaf(x[k], f_cosB[i][k], sb);
            }
            // out[i] = sa, out[17-i] = -sa, out[18+i] = sb, out[35-i] = sb
            const float s_i = (i & 1) ? sgn : 1.f, s_m = ((17 - i) & 1) ? sgn : 1.f;
            Fdst[i * 32 + lane] = sa * w[i] * s_i;
            Fdst[(17 - i) * 32 + lane] = -sa * w[17 - i] * s_m;
            Hdst[i * 32 + lane] = sb * w[18 + i] * s_i;
            Hdst[(17 - i) * 32 + lane] = sb * w[35 - i] * s_m;
        }
    } else {
        float y[3][12];
#pragma unroll
        for (int wdw = 0; wdw < 3; wdw++)
#pragma unroll
            for (int i = 0; i < 12; i++) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 6; k++) s = fmaf(x[3 * k + wdw], f_cos12[i][k], s);
                y[wdw][i] = s * f_win[2][i];
            }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const float s_i = (i & 1) ? sgn : 1.f; // 6 and 12 are even: parity of i everywhere
            Fdst[i * 32 + lane] = 0.f;
            Fdst[(6 + i) * 32 + lane] = y[0][i] * s_i;
            Fdst[(12 + i) * 32 + lane] = (y[0][6 + i] + y[1][i]) * s_i;
            Hdst[i * 32 + lane] = (y[1][6 + i] + y[2][i]) * s_i;
            Hdst[(6 + i) * 32 + lane] = y[2][6 + i] * s_i;
            Hdst[(12 + i) * 32 + lane] = 0.f;
        }
    }
}

template <int FMT>
__global__ void __launch_bounds__(KF_THREADS)
k_backend(const uint4 *__restrict__ tiles, uint32_t ntiles, const uint32_t *__restrict__ gran_unit0,
          const L3UnitDesc *__restrict__ units, const int16_t *__restrict__ is_in, const uint8_t *__restrict__ sf_in,
          const L3BandTables *__restrict__ bands, const float *__restrict__ pow43, void *__restrict__ pcm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedShared &S = *reinterpret_cast<FusedShared *>(smem_raw);
    if (blockIdx.x >= ntiles) return;
    const uint4 tl = tiles[blockIdx.x];
    const int warm = (int)tl.z, ng = (int)tl.y;    // granules before g0 to re-derive state from
    const uint32_t gstart = tl.x - (uint32_t)warm;  // first granule processed
    const int total = ng + warm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t gu_first = gran_unit0[gstart];
    const int nch = (gu_first & L3G_STEREO) ? 2 : 1;
    const uint32_t ubase = gu_first & L3G_UNIT_MASK; // units of a stream are contiguous

    // history starts at zero (stream head, or about to be re-derived by the warm-up granules)
    for (int i = tid; i < 2 * 15 * 32; i += KF_THREADS) S.F[i / 480][(i % 480) / 32][i % 32] = 0.f;
    for (int i = tid; i < 2 * 576; i += KF_THREADS) S.H[i / 576][0][(i % 576) / 32][i % 32] = 0.f;
    __syncthreads();

    float cn[32];
#pragma unroll
    for (int k = 0; k < 32; k++) cn[k] = f_dct32[k][lane];
    float wn[16];
#pragma unroll
    for (int l = 0; l < 16; l++) wn[l] = f_synwin[l][lane];
    const int src_e = lane <= 15 ? 16 + lane : (lane == 16 ? 0 : 48 - lane);
    const int src_o = lane <= 16 ? 16 - lane : lane - 16;

    for (int b0 = 0; b0 < total; b0 += KF_B) {
        const int nb = min(KF_B, total - b0);
        // ---- S1
        {
            const int grp = tid >> 6, t64 = tid & 63;
            if (grp < nb) {
                const uint32_t u0 = ubase + (uint32_t)(b0 + grp) * nch;
                float *tmp = &S.F[grp >> 1][15][0] + (grp & 1) * 1152; // rows 15.. of F are free until S2
                stage_requant(S, grp, t64, u0, nch, units, is_in, sf_in, bands, pow43, tmp);
            }
        }
        __syncthreads();
        // ---- S2
        {
            const int gi = warp >> 1, c = warp & 1;
            if (gi < nb && c < nch) {
                const uint8_t fl = units[ubase + (uint32_t)(b0 + gi) * nch + c].flags;
                stage_imdct(S.X[gi][c], lane, fl, &S.F[c][15 + gi * 18][0], &S.H[c][gi + 1][0][0]);
            }
        }
        __syncthreads();
        // ---- S3
        for (int it = warp; it < nch * nb * 18; it += KF_THREADS / 32) {
            const int c = it / (nb * 18), s = it % (nb * 18), gi = s / 18, t = s % 18;
            float *row = &S.F[c][15 + s][0];
            const float *hrow = &S.H[c][gi][t][0];
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                const float4 a = *reinterpret_cast<const float4 *>(row + k);
                const float4 h = *reinterpret_cast<const float4 *>(hrow + k);
                acc = fmaf(a.x + h.x, cn[k], acc);
                acc = fmaf(a.y + h.y, cn[k + 1], acc);
                acc = fmaf(a.z + h.z, cn[k + 2], acc);
                acc = fmaf(a.w + h.w, cn[k + 3], acc);
            }
            __syncwarp();
            row[lane] = acc;
        }
        __syncthreads();
        // ---- S4
        for (int s = warp; s < nb * 18; s += KF_THREADS / 32) {
            const int gi = s / 18, t = s % 18;
            if (b0 + gi < warm) continue; // warm-up granule: state only
            float out[2] = {0.f, 0.f};
            for (int c = 0; c < nch; c++) {
                const float *base = &S.F[c][15 + s][0];
                float acc = 0.f;
#pragma unroll
                for (int l = 0; l < 16; l += 2) {
                    acc = fmaf(wn[l], base[-l * 32 + src_e], acc);
                    acc = fmaf(wn[l + 1], base[-(l + 1) * 32 + src_o], acc);
                }
                out[c] = acc;
            }
            const size_t e0 = (size_t)(ubase + (uint32_t)(b0 + gi) * nch) * 576 + (size_t)(t * 32 + lane) * nch;
            if (FMT == MP3B_PCM_S16) {
                int16_t *p = reinterpret_cast<int16_t *>(pcm);
                if (nch == 2)
                    *reinterpret_cast<uint32_t *>(p + e0) =
                        (uint16_t)to_s16(out[0]) | ((uint32_t)(uint16_t)to_s16(out[1]) << 16);
                else
                    p[e0] = to_s16(out[0]);
            } else {
                float *p = reinterpret_cast<float *>(pcm);
                if (nch == 2) *reinterpret_cast<float2 *>(p + e0) = make_float2(out[0], out[1]);
                else p[e0] = out[0];
            }
        }
        __syncthreads();
        // ---- S5: carry the state to the next batch
        for (int i = tid; i < 2 * 15 * 32; i += KF_THREADS) {
            const int c = i / 480, r = (i % 480) / 32, k = i % 32;
            S.F[c][r][k] = S.F[c][nb * 18 + r][k];
        }
        for (int i = tid; i < 2 * 576; i += KF_THREADS) {
            const int c = i / 576, r = (i % 576) / 32, k = i % 32;
            S.H[c][0][r][k] = S.H[c][nb][r][k];
        }
        __syncthreads();
    }
}

} // namespace

void l3_fused_init(void)
{
    float p2[4], kl[7], kr[7], lp[2][16], cs[8], ca[8];
    for (int k = 0; k < 4; k++) p2[k] = (float)pow(2.0, k / 4.0);
    for (int p = 0; p < 7; p++) {
        if (p == 6) { kl[p] = 1.f; kr[p] = 0.f; }
        else {
            double t = tan(p * M_PI / 12.0);
            kl[p] = (float)(t / (1.0 + t));
            kr[p] = (float)(1.0 / (1.0 + t));
        }
    }
    for (int j = 0; j < 2; j++)
        for (int n = 0; n < 16; n++) lp[j][n] = (float)pow(2.0, -(j + 1) * n / 4.0);
    for (int i = 0; i < 8; i++) {
        double ci = l3_alias_ci[i];
        cs[i] = (float)(1.0 / sqrt(1.0 + ci * ci));
        ca[i] = (float)(ci / sqrt(1.0 + ci * ci));
    }
    cudaMemcpyToSymbol(f_pow2q, p2, sizeof p2);
    cudaMemcpyToSymbol(f_is_kl, kl, sizeof kl);
    cudaMemcpyToSymbol(f_is_kr, kr, sizeof kr);
    cudaMemcpyToSymbol(f_lsf_pow, lp, sizeof lp);
    cudaMemcpyToSymbol(f_cs, cs, sizeof cs);
    cudaMemcpyToSymbol(f_ca, ca, sizeof ca);
    cudaMemcpyToSymbol(f_pretab, l3_pretab, sizeof l3_pretab);

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
    cudaMemcpyToSymbol(f_cosA, A, sizeof A);
    cudaMemcpyToSymbol(f_cosB, B, sizeof B);
    cudaMemcpyToSymbol(f_cos12, C12, sizeof C12);
    cudaMemcpyToSymbol(f_win, W, sizeof W);

    static float dct[32][32], win[16][32];
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) dct[k][n] = (float)cos(n * (2 * k + 1) * M_PI / 64.0);
    for (int l = 0; l < 16; l++)
        for (int j = 0; j < 32; j++) {
            const int i = l >> 1;
            double v;
            if (!(l & 1)) v = l3_dwin(64 * i + j) * (j <= 15 ? 1.0 : (j == 16 ? 0.0 : -1.0));
            else v = -l3_dwin(64 * i + 32 + j);
            win[l][j] = (float)v;
        }
    cudaMemcpyToSymbol(f_dct32, dct, sizeof dct);
    cudaMemcpyToSymbol(f_synwin, win, sizeof win);
    const int smem = (int)sizeof(FusedShared);
    cudaFuncSetAttribute(k_backend<MP3B_PCM_S16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_backend<MP3B_PCM_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

void l3_launch_backend(const uint4 *tiles, uint32_t ntiles, const uint32_t *gran_unit0, const L3UnitDesc *units,
                       const int16_t *is_in, const uint8_t *sf_in, const L3DevTables &T, void *pcm, int pcm_format,
                       cudaStream_t st)
{
    if (!ntiles) return;
    const size_t smem = sizeof(FusedShared);
    if (pcm_format == MP3B_PCM_S16)
        k_backend<MP3B_PCM_S16><<<ntiles, KF_THREADS, smem, st>>>(tiles, ntiles, gran_unit0, units, is_in, sf_in,
                                                                  T.bands, T.pow43, pcm);
    else
        k_backend<MP3B_PCM_F32><<<ntiles, KF_THREADS, smem, st>>>(tiles, ntiles, gran_unit0, units, is_in, sf_in,
                                                                  T.bands, T.pow43, pcm);
}
