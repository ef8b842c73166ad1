// k_requant.cu -- K2: requantise (a6) + MS / intensity stereo (a7) + short-block reorder and alias
// reduction (a8), fused: one CTA per granule (both channels, because joint stereo couples them).
//
// HBM-bound by design: reads 576 int16 + 40 scalefactor bytes per unit, writes 576 float.  All
// intermediate state (per-band gains, intensity decisions, the reorder permutation and the alias
// butterflies) lives in shared memory; global accesses are contiguous rows.
// Restates oracle/l3_oracle.c::{requantise, stereo, reorder, alias_reduce} in float32.
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include <math.h>

#include "iso_tables.h"
#include "kernels.h"

namespace {

__constant__ float c_pow2q[4];        // 2^(k/4)
__constant__ float c_is_kl[7], c_is_kr[7];
__constant__ float c_lsf_pow[2][16];  // 2^(-(j+1) n / 4)
__constant__ float c_cs[8], c_ca[8];
__constant__ uint8_t c_pretab[22];

constexpr int K2_THREADS = 192;

struct GranShared {
    float x[2][576];
    float y[2][576];
    float gain[2][40];
    float kl[40], kr[40];
    uint8_t mode[40];  // 1 = intensity
    int nz[40];        // right channel band has a nonzero line
};

__global__ void __launch_bounds__(K2_THREADS)
k_requant(const L3UnitDesc *__restrict__ units, const uint32_t *__restrict__ gran_unit0, uint32_t g_lo,
          uint32_t ngranules,
          const int16_t *__restrict__ is_in, const uint8_t *__restrict__ sf_in,
          const L3BandTables *__restrict__ bands, const float *__restrict__ pow43, float *__restrict__ xr_out)
{
    __shared__ GranShared S;
    if (blockIdx.x >= ngranules) return;
    const uint32_t g = g_lo + blockIdx.x;
    const uint32_t gu = gran_unit0[g], u0 = gu & L3G_UNIT_MASK;
    const int nch = (gu & L3G_STEREO) ? 2 : 1;
    const int tid = threadIdx.x;

    L3UnitDesc d[2];
    d[0] = units[u0];
    d[1] = units[u0 + (nch - 1)];
    const int row = (d[0].hdr >> L3H_SR_SHIFT) & L3H_SR_MASK;
    int lay[2];
#pragma unroll
    for (int c = 0; c < 2; c++) lay[c] = (d[c].flags & L3F_BT_MASK) == 2 ? ((d[c].flags & L3F_MIXED) ? 2 : 1) : 0;

    // ---- per-band gains
    if (tid < 80) {
        const int c = tid / 40, b = tid % 40;
        if (c < nch) {
            const L3UnitDesc &dd = d[c];
            float gn = 0.f;
            if (b < bands->nbands[row][lay[c]]) {
                const int s = sf_in[(size_t)(u0 + c) * 40 + b] & 0x7f;
                const int win = bands->win[row][lay[c]][b];
                const int sh = (dd.flags & L3F_SFSCALE) ? 4 : 2;
                int q = (int)dd.global_gain - 210;
                if (win < 0) q -= sh * (s + ((dd.flags & L3F_PREFLAG) ? c_pretab[bands->sfb[row][lay[c]][b]] : 0));
                else q -= 8 * dd.sbg[win] + sh * s;
                gn = ldexpf(c_pow2q[q & 3], q >> 2);
            }
            S.gain[c][b] = gn;
        }
    }
    if (tid < 40) { S.nz[tid] = 0; S.mode[tid] = 0; }
    __syncthreads();

    // ---- requantise
    for (int c = 0; c < nch; c++) {
        const int16_t *is = is_in + (size_t)(u0 + c) * 576;
        const uint8_t *l2b = bands->line2band[row][lay[c]];
        for (int i = tid; i < 576; i += K2_THREADS) {
            const int v = is[i], b = l2b[i];
            const float a = __ldg(pow43 + (v < 0 ? -v : v)) * S.gain[c][b];
            S.x[c][i] = v < 0 ? -a : a;
            if (c == 1 && v != 0) S.nz[b] = 1;
        }
    }
    __syncthreads();

    // ---- stereo decisions (right channel's band layout governs)
    const bool ms = (d[0].hdr & L3H_MS) != 0, ist = (d[0].hdr & L3H_IS) != 0;
    const bool ok = (d[0].flags & L3F_VALID) != 0;
    if (nch == 2 && ist && ok && tid == 0) {
        const int nb = bands->nbands[row][lay[1]];
        const bool lsf = (d[1].hdr & L3H_LSF) != 0;
        const uint8_t *sf1 = sf_in + (size_t)(u0 + 1) * 40;
        int found[3] = {0, 0, 0}, found_long = 0;
        bool first_long = true;
        for (int b = nb - 1; b >= 0; b--) {
            const int w = bands->win[row][lay[1]][b];
            int *fnd;
            if (w >= 0) fnd = &found[w];
            else {
                if (first_long) { found_long = found[0] | found[1] | found[2]; first_long = false; }
                fnd = &found_long;
            }
            if (*fnd) continue;
            if (S.nz[b]) { *fnd = 1; continue; }
            const int sfb = bands->sfb[row][lay[1]][b];
            int bsf = b;
            if (w >= 0 && sfb == 12) bsf = b - 3;
            if (w < 0 && sfb == 21) bsf = b - 1;
            const int p = sf1[bsf];
            if (!lsf) {
                if (p < 7) { S.mode[b] = 1; S.kl[b] = c_is_kl[p]; S.kr[b] = c_is_kr[p]; }
            } else if (!(p & 0x80)) {
                const int j = d[1].sfc & 1;
                S.mode[b] = 1;
                S.kl[b] = (p & 1) ? c_lsf_pow[j][(p + 1) >> 1] : 1.f;
                S.kr[b] = (p & 1) ? 1.f : c_lsf_pow[j][p >> 1];
            }
        }
    }
    __syncthreads();

    // ---- apply stereo, write in reordered position
    const float isq2 = 0.70710678118654752440f;
    const bool joint = nch == 2 && ok && (ms || ist);
    for (int i = tid; i < 576; i += K2_THREADS) {
        float l = S.x[0][i], r = nch == 2 ? S.x[1][i] : 0.f;
        if (joint) {
            const int b = bands->line2band[row][lay[1]][i];
            if (ist && S.mode[b]) { const float a = l; l = a * S.kl[b]; r = a * S.kr[b]; }
            else if (ms) { const float a = l, c = r; l = (a + c) * isq2; r = (a - c) * isq2; }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (c >= nch) break;
            int dst = i;
            if (lay[c] != 0) {
                const int b = bands->line2band[row][lay[c]][i];
                const int w = bands->win[row][lay[c]][b];
                if (w >= 0) {
                    const int wd = bands->width[row][lay[c]][b], s = bands->start[row][lay[c]][b];
                    dst = (s - w * wd) + 3 * (i - s) + w;
                }
            }
            S.y[c][dst] = c == 0 ? l : r;
        }
    }
    __syncthreads();

    // ---- alias reduction
    for (int t = tid; t < 248 * nch; t += K2_THREADS) {
        const int c = t / 248, k = t % 248, sb = 1 + (k >> 3), i = k & 7;
        const int nb = (d[c].flags & L3F_BT_MASK) == 2 ? ((d[c].flags & L3F_MIXED) ? 1 : 0) : 31;
        if (sb <= nb) {
            const float lo = S.y[c][sb * 18 - 1 - i], hi = S.y[c][sb * 18 + i];
            S.y[c][sb * 18 - 1 - i] = lo * c_cs[i] - hi * c_ca[i];
            S.y[c][sb * 18 + i] = hi * c_cs[i] + lo * c_ca[i];
        }
    }
    __syncthreads();

    for (int c = 0; c < nch; c++) {
        float *o = xr_out + (size_t)(u0 + c) * 576;
        for (int i = tid; i < 576; i += K2_THREADS) o[i] = S.y[c][i];
    }
}

} // namespace

void l3_requant_init(void)
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
    cudaMemcpyToSymbol(c_pow2q, p2, sizeof p2);
    cudaMemcpyToSymbol(c_is_kl, kl, sizeof kl);
    cudaMemcpyToSymbol(c_is_kr, kr, sizeof kr);
    cudaMemcpyToSymbol(c_lsf_pow, lp, sizeof lp);
    cudaMemcpyToSymbol(c_cs, cs, sizeof cs);
    cudaMemcpyToSymbol(c_ca, ca, sizeof ca);
    cudaMemcpyToSymbol(c_pretab, l3_pretab, sizeof l3_pretab);
}

void l3_launch_requant_range(const L3UnitDesc *units, const uint32_t *gran_unit0, uint32_t g_lo, uint32_t ngranules,
                             const int16_t *is_in, const uint8_t *sf_in, const L3DevTables &T, float *xr_out,
                             cudaStream_t st)
{
    if (!ngranules) return;
    k_requant<<<ngranules, K2_THREADS, 0, st>>>(units, gran_unit0, g_lo, ngranules, is_in, sf_in, T.bands, T.pow43,
                                                xr_out);
}
