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
//   S1  all threads, one (granule, line) pair each: per-band gains, requantise both channels,
//       MS / intensity stereo, short-block reorder                  -> X  (padded rows of 19)
//       (intensity decisions need the right channel's zero bands first: a short pre-pass)
//   S2  8 warps = 4 granules x 2 channels, lane = subband: alias butterflies by shuffle, 36-point
//       IMDCT as an in-register fast 18-point DCT-IV (fast_imdct.h, ~170 operations instead of 648
//       multiply-adds), window, frequency inversion; first halves -> F, second halves stay in
//       registers and are added to the next granule's rows (overlap-add)
//   S3  thread per (channel, slot): C[n] = sum_k S[k] cos(n(2k+1)pi/64) as an in-register fast DCT-II
//       (304 operations, fast_dct.h), in place (the 64 "V" values of the standard are signed copies
//       of C); rows have an odd stride so this lane = row pattern is bank-conflict free too
//   S4  warp per (granule, channel), lane = sample j: the 16-tap window as a fully unrolled sliding
//       accumulation over the 33 rows a granule touches (2 LDS per row instead of 16 per output);
//       PCM is staged in shared memory and leaves the CTA as one TMA bulk store per batch
//   Input staging: the next batch's spectra are fetched by TMA bulk copies (cp.async.bulk + mbarrier,
//   only the vectors that hold data) while the current batch computes.
// Results must equal the staged pipeline's (tests/test_gpu_parity.py runs both).
// No reference code exists for these stages (/root/reference/README.md:1-84).
#include <math.h>

#include <type_traits>

#include "consts_gen.h"
#include "fast_dct.h"
#include "fast_imdct.h"
#include "iso_tables.h"
#include "kernels.h"
#include "mp3b.h"

namespace {

constexpr int KF_B = 4;         // granules per batch
constexpr int KF_THREADS = 256;
constexpr int KF_ROWS = 15 + KF_B * 18;
constexpr int XROW = 19;        // padded subband row of X
constexpr int XSZ = 32 * XROW;  // 608 floats per channel spectrum
constexpr int KF_POW_LUT = 512;
constexpr int FS = 33;         // row stride of F: odd, so rows are conflict-free both by lane = column and lane = row

__constant__ float f_pow2q[4];
__constant__ float f_is_kl[7], f_is_kr[7];
__constant__ float f_lsf_pow[2][16];
__constant__ float f_cs[8], f_ca[8];
__constant__ uint8_t f_pretab[22];
__constant__ float f_win[4][36];
__device__ float f_synwin[16][32];

struct GranMeta {
    L3UnitDesc d[2];
    int row, lay[2];
    int joint;      // 0 = independent channels, 1 = MS and/or intensity processing applies
    int ms, ist;
};

// The Huffman output of the next batch (9216 bytes) is fetched while the current batch is in S3 / S4,
// when X is dead except for the PCM staging area at its start.  With s16 output that area is 9216 bytes
// too, so the fetch buffer lives INSIDE X (bytes 9216 .. 18431) and the CTA needs no separate buffer --
// which is what brings shared memory under 56 KB and a fourth CTA onto the SM.  f32 output stages twice
// as much and keeps its own buffer (three CTAs per SM).
template <int FMT> struct IsOwn { __align__(16) int16_t buf[KF_B * 2][576]; };
template <> struct IsOwn<MP3B_PCM_S16> {};

template <int FMT>
struct FusedSharedT {
    __align__(16) float X[KF_B][2][XSZ]; // spectra of the batch, padded rows; PCM staging in S4; (s16) next batch's is
    float F[2][KF_ROWS][FS];        // rows 0..14: C history; rows 15..: first IMDCT halves -> S -> C
    __align__(16) float Hc[2][18][32]; // second IMDCT half of the last granule of the previous batch
    float pow43[KF_POW_LUT];        // |is|^(4/3) for the common small values
    float win[16][32];              // per-lane window taps (copied from global once)
    float gain[KF_B][2][40];
    float kl[KF_B][40], kr[KF_B][40];
    uint8_t nz[KF_B][40];           // right-channel band has a non-zero line (intensity bound)
    uint8_t mode[KF_B][40];         // 1 = intensity-coded band
    GranMeta gm[KF_B];
    int any_ist;
    int any_short;                  // some unit of the batch has short / mixed blocks
    __align__(16) uint8_t sf_buf[KF_B * 2][40]; // next batch's scalefactors (cp.async)
    __align__(8) uint64_t bar;  // mbarrier: completion of the bulk copies into is_buf
    IsOwn<FMT> own;
    // next batch's Huffman output, fetched by TMA while this batch computes
    __device__ __forceinline__ int16_t (*is_buf())[576]
    {
        if constexpr (FMT == MP3B_PCM_S16)
            return reinterpret_cast<int16_t(*)[576]>(reinterpret_cast<unsigned char *>(&X[0][0][0]) + KF_B * 2 * 576 * 2);
        else
            return own.buf;
    }
};
static_assert(KF_B * 2 * 576 * 2 * 2 <= (int)sizeof(float) * KF_B * 2 * XSZ, "is buffer must fit behind the s16 staging area");

// f_pow2q[q & 3] * 2^(q >> 2), exactly (q >> 2 stays within the normal exponent range: -82 .. 11)
__device__ __forceinline__ float gain_of(int q) { return __int_as_float((127 + (q >> 2)) << 23) * f_pow2q[q & 3]; }

// s16: four CTAs per SM, f32: three -- n x (shared memory + 1 KB reserved per CTA) must fit 228 KB
static_assert(sizeof(FusedSharedT<MP3B_PCM_S16>) <= 56 * 1024, "s16 back end: too large for 4 CTAs per SM");
static_assert(sizeof(FusedSharedT<MP3B_PCM_F32>) <= 75 * 1024, "f32 back end: too large for 3 CTAs per SM");

__device__ __forceinline__ int xpad(int i) { return i + ((i * 3641) >> 16); } // i + i / 18 for i < 608

__device__ __forceinline__ int16_t to_s16(float v)
{
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(v * 32768.f));
    return (int16_t)r;
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) with an mbarrier for completion -----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Start fetching the Huffman output (1152 B / unit) and scalefactors (40 B / unit) of `n` consecutive
// units into shared memory; the caller waits (cp_async_wait_all + barrier) before reading them.
// Only the vectors that hold data are fetched (nzv_in[u] of the 72 per unit); the all-zero tail of
// each spectrum is never written by the Huffman kernel nor read here: it is zeroed in place.
template <class SH>
__device__ __forceinline__ void prefetch_units(SH &S, int tid, uint32_t u_first, int n,
                                               const int16_t *__restrict__ is_in, const uint8_t *__restrict__ sf_in,
                                               const uint8_t *__restrict__ nzv_in)
{
    // spectra: one TMA bulk copy per unit, issued by the lanes of the last warp (which has no S3 work);
    // every warp zeroes the tail of "its" unit
    if ((tid >> 5) == KF_THREADS / 32 - 1) {
        const int k = tid & 31;
        const uint32_t b = k < n ? 16u * nzv_in[u_first + k] : 0u;
        uint32_t tot = b;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (k == 0) {
            fence_proxy_async(); // earlier generic-proxy reads of is_buf are ordered before the async writes
            mbar_expect_tx(&S.bar, tot);
        }
        __syncwarp();
        if (b) bulk_g2s(&S.is_buf()[k][0], is_in + (size_t)(u_first + k) * 576, b, &S.bar);
    }
    const int unit = tid >> 5, lane = tid & 31;
    if (unit < n || (unit == n && (n & 1))) { // an odd mono batch leaves one slot of the last pair empty: all zero
        char *si = reinterpret_cast<char *>(&S.is_buf()[unit][0]);
        const int nv = unit < n ? nzv_in[u_first + unit] : 0;
#pragma unroll
        for (int v = lane; v < 72; v += 32)
            if (v >= nv) *reinterpret_cast<uint4 *>(si + v * 16) = make_uint4(0, 0, 0, 0);
    }
    const char *gs = reinterpret_cast<const char *>(sf_in + (size_t)u_first * 40);
    char *ss = reinterpret_cast<char *>(&S.sf_buf[0][0]);
    for (int i = tid; i < n * 5; i += KF_THREADS) cp_async8(ss + i * 8, gs + i * 8);
    cp_async_commit();
}

// ---- per-batch metadata (descriptors, layouts), loaded one batch ahead ---------------------------
// A batch is held as pairs of unit slots: the two channels of a stereo granule, or two consecutive
// granules of a mono stream (which doubles the granules per batch, so that mono tiles keep all eight
// warps busy).  `nun` = units in the batch; an odd mono batch repeats its last descriptor.
template <class SH>
__device__ __forceinline__ void load_meta(SH &S, int tid, uint32_t u_first, int nun,
                                          const L3UnitDesc *__restrict__ units)
{
    if (tid < ((nun + 1) & ~1)) S.gm[tid >> 1].d[tid & 1] = units[u_first + (uint32_t)min(tid, nun - 1)];
    for (int i = tid; i < KF_B * 40; i += KF_THREADS) {
        (&S.nz[0][0])[i] = 0;
        (&S.mode[0][0])[i] = 0;
    }
    if (tid == 0) { S.any_ist = 0; S.any_short = 0; }
}

template <class SH>
__device__ __forceinline__ void finish_meta(SH &S, int tid, int nb, int nch)
{
    if (tid < nb) {
        GranMeta &m = S.gm[tid];
        m.row = (m.d[0].hdr >> L3H_SR_SHIFT) & L3H_SR_MASK;
        for (int c = 0; c < 2; c++)
            m.lay[c] = (m.d[c].flags & L3F_BT_MASK) == 2 ? ((m.d[c].flags & L3F_MIXED) ? 2 : 1) : 0;
        const bool ok = (m.d[0].flags & L3F_VALID) != 0;
        m.ms = (nch == 2 && ok && (m.d[0].hdr & L3H_MS)) ? 1 : 0;
        m.ist = (nch == 2 && ok && (m.d[0].hdr & L3H_IS)) ? 1 : 0;
        m.joint = m.ms | m.ist;
        if (m.ist) S.any_ist = 1;
        if (m.lay[0] | m.lay[1]) S.any_short = 1;
    }
}

// ---- S1a: per-band gains, and the right channel's non-zero bands for intensity granules ---------
template <class SH>
__device__ __forceinline__ void stage_gains(SH &S, int tid, int nb, int nch,
                                            const L3BandTables *__restrict__ bands)
{
    // 64 threads per granule: lanes 0..39 of each of its two warps take one band of one channel
    {
        const int gi = tid >> 6, c = (tid >> 5) & 1, b = tid & 31;
        if (gi < nb && c < nch && S.gm[gi].lay[c] == 0) {
            // long blocks (22 bands, the band index is the sfb): no table lookups
            const L3UnitDesc &dd = S.gm[gi].d[c];
            if (b < 22) {
                const int s = S.sf_buf[gi * nch + c][b] & 0x7f;
                const int sh = (dd.flags & L3F_SFSCALE) ? 4 : 2;
                const int q = (int)dd.global_gain - 210 - sh * (s + ((dd.flags & L3F_PREFLAG) ? f_pretab[b] : 0));
                // the table-free requantise path (both channels long, no intensity) takes the MS factor
                // 1 / sqrt 2 from the gains; the general path applies it per line
                const GranMeta &m = S.gm[gi];
                const bool fold = m.ms && (m.lay[0] | m.lay[1]) == 0 && !m.ist;
                S.gain[gi][c][b] = gain_of(q) * (fold ? 0.70710678118654752440f : 1.f);
            }
        }
    }
    if (S.any_short) {
        // short / mixed layouts: up to 39 bands, lane and lane + 32 of the same (granule, channel) warp
        const int gi = tid >> 6, c = (tid >> 5) & 1, lane = tid & 31;
        if (gi < nb && c < nch && S.gm[gi].lay[c] != 0) {
            const GranMeta &m = S.gm[gi];
            const L3UnitDesc &dd = m.d[c];
            const int lay = m.lay[c], nbands = bands->nbands[m.row][lay];
            const int sh = (dd.flags & L3F_SFSCALE) ? 4 : 2;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int b = lane + 32 * r;
                if (b >= 40) break;
                float gn = 0.f;
                if (b < nbands) {
                    const int s = S.sf_buf[gi * nch + c][b] & 0x7f;
                    const int win = bands->win[m.row][lay][b];
                    int q = (int)dd.global_gain - 210;
                    if (win < 0) q -= sh * (s + ((dd.flags & L3F_PREFLAG) ? f_pretab[bands->sfb[m.row][lay][b]] : 0));
                    else q -= 8 * dd.sbg[win] + sh * s;
                    gn = gain_of(q);
                }
                S.gain[gi][c][b] = gn;
            }
        }
    }
    if (S.any_ist) {
        // which bands of the right channel hold a non-zero line: one 16-byte vector (8 lines) per item
        for (int it = tid; it < nb * 72; it += KF_THREADS) {
            const int gi = it / 72, v = it - gi * 72;
            const GranMeta &m = S.gm[gi];
            if (!m.ist) continue;
            const uint4 q = reinterpret_cast<const uint4 *>(&S.is_buf()[gi * nch + 1][0])[v];
            if ((q.x | q.y | q.z | q.w) == 0u) continue;
            const uint32_t *lm = bands->lmap[m.row][m.lay[1]] + v * 8;
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (w[k] & 0xffffu) S.nz[gi][__ldg(lm + 2 * k) & 0xffu] = 1;
                if (w[k] >> 16) S.nz[gi][__ldg(lm + 2 * k + 1) & 0xffu] = 1;
            }
        }
    }
}

// ---- S1b: intensity decisions, one warp per intensity granule, lane = band (and band + 32) ------
// A band of the right channel is intensity coded when no band of its own class (short window 0 / 1 / 2,
// or long) at or above it holds a non-zero line -- and, for the long bands of a mixed block, no short
// band does either (11172-3 2.4.3.4: intensity applies above the last non-zero band).  Scanned serially
// from the top this is a chain; as "highest non-zero band per class" it is four ballots.
template <class SH>
__device__ __forceinline__ void stage_intensity(SH &S, int gi, int u1 /* batch-local unit */, int lane,
                                                const L3BandTables *__restrict__ bands)
{
    const GranMeta &m = S.gm[gi];
    const int row = m.row, lay1 = m.lay[1];
    const int nbands = bands->nbands[row][lay1];
    const bool lsf = (m.d[1].hdr & L3H_LSF) != 0;
    const uint8_t *sf1 = S.sf_buf[u1];
    int cls[2], top[4] = {-1, -1, -1, -1}; // class of this lane's two bands; highest non-zero band per class
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int b = lane + 32 * r;
        cls[r] = b < nbands ? (bands->win[row][lay1][b] < 0 ? 3 : bands->win[row][lay1][b]) : -1;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t lo = __ballot_sync(0xffffffffu, cls[0] == c && S.nz[gi][lane]);
        const uint32_t hi = __ballot_sync(0xffffffffu, cls[1] == c && lane + 32 < 40 && S.nz[gi][lane + 32 < 40 ? lane + 32 : 0]);
        top[c] = hi ? 63 - __clz(hi) : (lo ? 31 - __clz(lo) : -1);
    }
    const bool any_short_nz = top[0] >= 0 || top[1] >= 0 || top[2] >= 0;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int b = lane + 32 * r;
        if (cls[r] < 0) continue;
        const int c = cls[r];
        const bool blocked = b <= top[c] || (c == 3 && any_short_nz);
        if (blocked) continue;
        const int sfb = bands->sfb[row][lay1][b];
        int bsf = b;
        if (c != 3 && sfb == 12) bsf = b - 3;
        if (c == 3 && sfb == 21) bsf = b - 1;
        const int p = sf1[bsf];
        if (!lsf) {
            if (p < 7) { S.mode[gi][b] = 1; S.kl[gi][b] = f_is_kl[p]; S.kr[gi][b] = f_is_kr[p]; }
        } else if (!(p & 0x80)) {
            const int j = m.d[1].sfc & 1;
            S.mode[gi][b] = 1;
            S.kl[gi][b] = (p & 1) ? f_lsf_pow[j][(p + 1) >> 1] : 1.f;
            S.kr[gi][b] = (p & 1) ? 1.f : f_lsf_pow[j][p >> 1];
        }
    }
}

template <class SH>
__device__ __forceinline__ float requant1(const SH &S, int v, float gain, const float *__restrict__ pow43)
{
    const int m = v < 0 ? -v : v;
    const float p = m < KF_POW_LUT ? S.pow43[m] : __ldg(pow43 + m);
    const float a = p * gain;
    return v < 0 ? -a : a;
}

// ---- S1c: requantise + stereo + reorder, one (granule, line) per item -----------------------------
// `bq` holds this thread's nine long-block band indices (lines t64 + 64 q), one byte each: they
// depend only on the stream's sample rate, so the common case (both channels long blocks, no
// intensity stereo) needs no table lookups, no reorder and no branches per line.
// First half: every thread pulls its nine line pairs of the Huffman output into registers.  With s16
// output that buffer lives inside X, which the second half overwrites: the caller puts a barrier between
// the two.
template <class SH>
__device__ __forceinline__ void requant_load(SH &S, int tid, int nb, uint32_t (&v)[9])
{
    const int gi = tid >> 6, t64 = tid & 63;
    if (gi >= nb) return;
    const uint16_t *s0 = reinterpret_cast<const uint16_t *>(S.is_buf()[gi * 2]);
    const uint16_t *s1 = reinterpret_cast<const uint16_t *>(S.is_buf()[gi * 2 + 1]);
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = (uint32_t)s0[t64 + 64 * q] | ((uint32_t)s1[t64 + 64 * q] << 16); // packed pair
}

template <class SH>
__device__ __forceinline__ void stage_requant(SH &S, int tid, int nb, const L3BandTables *__restrict__ bands,
                                              const float *__restrict__ pow43, const uint32_t (&bq)[3],
                                              const uint32_t (&v)[9])
{
    constexpr int ITEMS = 576 / 64; // 9 lines per thread: 64 threads per granule
    const float isq2 = 0.70710678118654752440f;
    const int gi = tid >> 6, t64 = tid & 63;
    if (gi >= nb) return;
    const GranMeta &m = S.gm[gi];
    const int lay0 = m.lay[0], lay1 = m.lay[1];
    const float *g0 = S.gain[gi][0], *g1 = S.gain[gi][1];
    if ((lay0 | lay1) == 0 && !m.ist) {
        float *X0 = S.X[gi][0], *X1 = S.X[gi][1];
        // MS: (M +- S) / sqrt 2; stage_gains has folded the factor into this granule's band gains
        auto lines = [&](auto ms_tag) {
#pragma unroll
            for (int q = 0; q < ITEMS; q++) {
                const int b = (bq[q >> 2] >> (8 * (q & 3))) & 0xff;
                const int w0 = (int)(short)(v[q] & 0xffffu), w1 = (int)v[q] >> 16;
                const int m0 = abs(w0), m1 = abs(w1);
                const float p0 = m0 < KF_POW_LUT ? S.pow43[m0] : __ldg(pow43 + m0);
                const float p1 = m1 < KF_POW_LUT ? S.pow43[m1] : __ldg(pow43 + m1);
                const float a = __int_as_float(__float_as_int(p0 * g0[b]) | (w0 & 0x80000000));
                const float c = __int_as_float(__float_as_int(p1 * g1[b]) | (w1 & 0x80000000));
                const int xp = xpad(t64 + 64 * q);
                X0[xp] = decltype(ms_tag)::value ? a + c : a;
                X1[xp] = decltype(ms_tag)::value ? a - c : c;
            }
        };
        if (m.ms) lines(std::true_type{});
        else lines(std::false_type{});
        return;
    }
    // general path (short / mixed blocks, intensity stereo): band and reordered padded position of every
    // line from the per-rate line map (global memory, L1-resident)
    const uint32_t *lm0 = bands->lmap[m.row][lay0], *lm1 = bands->lmap[m.row][lay1];
    const bool ist = m.ist != 0, ms = m.ms != 0;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
        const int i = t64 + 64 * q;
        const uint32_t e0 = __ldg(lm0 + i), e1 = __ldg(lm1 + i);
        const int b1 = (int)(e1 & 0xffu);
        float l = requant1(S, (int)(short)(v[q] & 0xffffu), g0[e0 & 0xffu], pow43);
        float r = requant1(S, (int)v[q] >> 16, g1[b1], pow43);
        if (ist && S.mode[gi][b1]) { const float a = l; l = a * S.kl[gi][b1]; r = a * S.kr[gi][b1]; }
        else if (ms) { const float a = l, c = r; l = (a + c) * isq2; r = (a - c) * isq2; }
        S.X[gi][1][e1 >> 8] = r;
        S.X[gi][0][e0 >> 8] = l;
    }
}

// ---- S2: alias reduction + IMDCT of one (granule, channel) by one warp, lane = subband ----------
// First halves go to Fdst (plus `carry`, the previous batch's last second half, for the batch's first
// granule); the second half stays in registers (h) and is added to the next granule's rows after a
// barrier, so no second-half buffer is needed.
__device__ __forceinline__ void stage_imdct(const float *__restrict__ X, int lane, uint8_t flags,
                                            float *__restrict__ Fdst /* [18][32] */,
                                            const float *__restrict__ carry /* [18][32] or null */, float (&h)[18])
{
    float x[18];
#pragma unroll
    for (int k = 0; k < 18; k++) x[k] = X[lane * XROW + k];
    int bt = flags & L3F_BT_MASK;
    const bool mixed = bt == 2 && (flags & L3F_MIXED);
    // alias butterflies between subband `lane - 1` (its lines 17-i) and `lane` (its lines i), i < 8
    {
        const int nbnd = bt == 2 ? (mixed ? 1 : 0) : 31; // boundaries 1..nbnd are processed
        const bool lo_side = lane + 1 <= nbnd;           // this lane is the lower subband of a boundary
        const bool hi_side = lane >= 1 && lane <= nbnd;  // this lane is the upper subband of a boundary
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float from_next = __shfl_down_sync(0xffffffffu, x[i], 1);    // next subband's line i
            const float from_prev = __shfl_up_sync(0xffffffffu, x[17 - i], 1); // previous subband's line 17-i
            const float hi = x[i], lo = x[17 - i];
            if (hi_side) x[i] = hi * f_cs[i] + from_prev * f_ca[i];
            if (lo_side) x[17 - i] = lo * f_cs[i] - from_next * f_ca[i];
        }
    }
    if (mixed && lane < 2) bt = 0;
    const float sgn = (lane & 1) ? -1.f : 1.f; // frequency inversion: odd subband, odd slot
    if (bt != 2) {
        const float *w = f_win[bt];
        float Z[18];
        l3_dct4_18(x, Z); // the 18 distinct IMDCT values (fast_imdct.h)
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const float sa = Z[9 + i], sb = -Z[8 - i];
            // out[i] = sa, out[17-i] = -sa, out[18+i] = sb, out[35-i] = sb; slots i and 17 - i have
            // opposite parity, so exactly one of each pair takes the inversion sign
            const float sa_s = sa * sgn, sb_s = sb * sgn;
            float f0 = ((i & 1) ? sa_s : sa) * w[i], f1 = -((i & 1) ? sa : sa_s) * w[17 - i];
            if (carry) { f0 += carry[i * 32 + lane]; f1 += carry[(17 - i) * 32 + lane]; }
            Fdst[i * FS + lane] = f0;
            Fdst[(17 - i) * FS + lane] = f1;
            h[i] = ((i & 1) ? sb_s : sb) * w[18 + i];
            h[17 - i] = ((i & 1) ? sb : sb_s) * w[35 - i];
        }
    } else {
        // 12-point IMDCT per window.  Only six of the twelve values are distinct (y[5-i] = -y[i],
        // y[11-i] = y[6+i]: the kernel is a 6-point DCT-IV), so six dot products per window, not twelve.
        float y[3][12];
#pragma unroll
        for (int wdw = 0; wdw < 3; wdw++)
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    a = fmaf(x[3 * k + wdw], K12[i][k], a);
                    b = fmaf(x[3 * k + wdw], K12[6 + i][k], b);
                }
                y[wdw][i] = a * f_win[2][i];
                y[wdw][5 - i] = -a * f_win[2][5 - i];
                y[wdw][6 + i] = b * f_win[2][6 + i];
                y[wdw][11 - i] = b * f_win[2][11 - i];
            }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const float s_i = (i & 1) ? sgn : 1.f; // 6 and 12 are even: parity of i everywhere
            float f0 = 0.f, f1 = y[0][i] * s_i, f2 = (y[0][6 + i] + y[1][i]) * s_i;
            if (carry) {
                f0 += carry[i * 32 + lane];
                f1 += carry[(6 + i) * 32 + lane];
                f2 += carry[(12 + i) * 32 + lane];
            }
            Fdst[i * FS + lane] = f0;
            Fdst[(6 + i) * FS + lane] = f1;
            Fdst[(12 + i) * FS + lane] = f2;
            h[i] = (y[1][6 + i] + y[2][i]) * s_i;
            h[6 + i] = y[2][6 + i] * s_i;
            h[12 + i] = 0.f;
        }
    }
}

// ---- S4: synthesis window for one (granule, channel), lane = sample j ---------------------------
// rows r0 .. r0+32 of Fc are the transformed slots T0-15 .. T0+17.  Output slot T (0..17) is
//   sum_{l<16} W[l][j] * C_{T-l}[l even ? src_e : src_o]
// evaluated as a sliding accumulation over the rows: each row is loaded once (2 LDS) and feeds the
// up to 16 outputs it contributes to.  Everything is unrolled, so the 16-entry accumulator ring and
// the window taps live in registers.
template <typename Emit>
__device__ __forceinline__ void stage_window(const float *__restrict__ Fc, int r0, int src_e, int src_o,
                                             const float (&wn)[16], Emit emit)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0.f;
#pragma unroll
    for (int q = 0; q < 33; q++) {
        const float e = Fc[(r0 + q) * FS + src_e], o = Fc[(r0 + q) * FS + src_o];
#pragma unroll
        for (int l = 0; l < 16; l++) {
            const int T = q + l - 15;
            if (T >= 0 && T < 18) acc[T & 15] = fmaf(wn[l], (l & 1) ? o : e, acc[T & 15]);
        }
        if (q >= 15) {
            emit(q - 15, acc[(q - 15) & 15]);
            acc[(q - 15) & 15] = 0.f;
        }
    }
}

// One tile.  MONO is a compile-time switch so that the stereo path carries none of the mono index
// arithmetic: pairs of consecutive mono granules take the two slots a stereo granule's channels would,
// eight granules per batch, and the rows of F form ONE time sequence (F[0] and F[1] are contiguous).
template <int FMT, bool MONO>
__device__ __forceinline__ void backend_tile(FusedSharedT<FMT> &S, const int warm, const int total, const uint32_t ubase,
                                             const L3UnitDesc *__restrict__ units, const int16_t *__restrict__ is_in,
                                             const uint8_t *__restrict__ sf_in, const uint8_t *__restrict__ nzv_in,
                                             const L3BandTables *__restrict__ bands, const float *__restrict__ pow43,
                                             void *__restrict__ pcm)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int nch = MONO ? 1 : 2;
    typedef typename std::conditional<FMT == MP3B_PCM_S16, int16_t, float>::type pcm_t;
    constexpr bool mono = MONO;
    constexpr int KFG = mono ? 2 * KF_B : KF_B;
    float *const Fm = &S.F[0][0][0];

    if (tid == 0) mbar_init(&S.bar, 1);
    __syncthreads();
    // history starts at zero (stream head, or about to be re-derived by the warm-up granules)
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < 2 * 15 * FS; i += KF_THREADS) S.F[i / (15 * FS)][0][i % (15 * FS)] = 0.f;
        for (int i = tid; i < 2 * 144; i += KF_THREADS) reinterpret_cast<float4 *>(&S.Hc[i / 144][0][0])[i % 144] = z;
        for (int i = tid; i < KF_POW_LUT; i += KF_THREADS) S.pow43[i] = pow43[i];
        for (int i = tid; i < 512; i += KF_THREADS) (&S.win[0][0])[i] = (&f_synwin[0][0])[i];
        load_meta(S, tid, ubase, min(KFG, total) * nch, units);
        prefetch_units(S, tid, ubase, min(KFG, total) * nch, is_in, sf_in, nzv_in);
    }
    __syncthreads();

    const int src_e = lane <= 15 ? 16 + lane : (lane == 16 ? 0 : 48 - lane);
    const int src_o = lane <= 16 ? 16 - lane : lane - 16;
    pcm_t *stage = reinterpret_cast<pcm_t *>(&S.X[0][0][0]);
    uint32_t bq[3] = {0, 0, 0}; // long-block band index of this thread's nine lines (see stage_requant)
    {
        const int row0 = (units[ubase].hdr >> L3H_SR_SHIFT) & L3H_SR_MASK;
#pragma unroll
        for (int q = 0; q < 9; q++)
            bq[q >> 2] |= (uint32_t)bands->line2band[row0][0][(tid & 63) + 64 * q] << (8 * (q & 3));
    }

    for (int b0 = 0; b0 < total; b0 += KFG) {
        const int nb = min(KFG, total - b0);           // granules of this batch
        const int np = mono ? (nb + 1) >> 1 : nb;      // slot pairs of this batch
        const uint32_t u_first = ubase + (uint32_t)b0 * nch;
        // ---- S1
        finish_meta(S, tid, np, nch);
        cp_async_wait_all();                        // scalefactors (LDGSTS)
        mbar_wait(&S.bar, (uint32_t)(b0 / KFG) & 1u); // spectra (TMA); one phase per batch
        if (tid == KF_THREADS - 32) bulk_wait_read_all(); // the previous batch's PCM has left the staging buffer
        __syncthreads();
        stage_gains(S, tid, np, 2, bands);
        __syncthreads();
        if (S.any_ist) { // stereo only
            if ((tid & 32) == 0 && (tid >> 6) < np && S.gm[tid >> 6].ist) // first warp of each granule's 64 threads
                stage_intensity(S, tid >> 6, (tid >> 6) * 2 + 1, lane, bands);
            __syncthreads();
        }
        {
            uint32_t v[9];
            requant_load(S, tid, np, v);
            if (FMT == MP3B_PCM_S16) __syncthreads(); // the Huffman output sits inside X, which is written next
            stage_requant(S, tid, np, bands, pow43, bq, v);
        }
        __syncthreads();
        // ---- S2: alias + IMDCT; second halves travel in registers to the next granule's rows
        {
            // warp -> slot (gi, c); j = its granule inside the batch, seq = the row sequence it belongs to
            const int gi = warp >> 1, c = warp & 1;
            const int j = mono ? warp : gi;
            const bool act = j < nb;
            float *const seq = mono ? Fm : &S.F[c][0][0];
            float *const hc = &S.Hc[mono ? 0 : c][0][0];
            float h[18];
            if (act)
                stage_imdct(S.X[gi][c], lane, S.gm[gi].d[c].flags, seq + (15 + j * 18) * FS, j == 0 ? hc : nullptr, h);
            __syncthreads();
            if (act) {
                float *dst = j + 1 < nb ? seq + (15 + (j + 1) * 18) * FS : hc;
                if (j + 1 < nb) {
#pragma unroll
                    for (int t = 0; t < 18; t++) dst[t * FS + lane] += h[t];
                } else {
#pragma unroll
                    for (int t = 0; t < 18; t++) dst[t * 32 + lane] = h[t];
                }
            }
        }
        __syncthreads();
        // next batch: descriptors (gm is dead until the next S1) and, behind S3..S5, spectra + scalefactors
        if (b0 + KFG < total) {
            load_meta(S, tid, u_first + (uint32_t)KFG * nch, min(KFG, total - b0 - KFG) * nch, units);
            prefetch_units(S, tid, u_first + (uint32_t)KFG * nch, min(KFG, total - b0 - KFG) * nch, is_in, sf_in,
                           nzv_in);
        }
        // ---- S3: 32-point transform of every slot, in place; one thread per (channel, slot) row,
        // the whole transform in registers (fast_dct.h: 304 operations instead of 1024 FMAs)
        {
            const int rows_c = nb * 18; // rows per sequence: stereo has two sequences, mono one
            if (tid < nch * rows_c) {
                const int c = tid >= rows_c ? 1 : 0;
                float *row = &S.F[c][0][0] + (15 + tid - c * rows_c) * FS;
                float x[32];
#pragma unroll
                for (int k = 0; k < 32; k++) x[k] = row[k];
                L3Dct2<32>::run(x);
#pragma unroll
                for (int k = 0; k < 32; k++) row[k] = x[k];
            }
        }
        __syncthreads();
        // ---- S4: window -> PCM staging (X is free now)
        {
            const int c = mono ? 0 : (warp & 1), j = mono ? warp : (warp >> 1);
            if (j < nb && b0 + j >= warm) {
                float wn[16];
#pragma unroll
                for (int l = 0; l < 16; l++) wn[l] = S.win[l][lane];
                pcm_t *dst = stage + (size_t)j * 576 * nch + c;
                stage_window(&S.F[c][0][0], j * 18, src_e, src_o, wn, [&](int t, float val) {
                    if (FMT == MP3B_PCM_S16) dst[(t * 32 + lane) * nch] = (pcm_t)to_s16(val);
                    else dst[(t * 32 + lane) * nch] = (pcm_t)val;
                });
                fence_proxy_async(); // make the staged PCM visible to the TMA store issued after the barrier
            }
        }
        __syncthreads();
        // ---- S5: PCM out (TMA bulk store), carry the state to the next batch
        {
            const int first_out = max(0, warm - b0); // warm-up granules of this batch produce no PCM
            if (first_out < nb && tid == KF_THREADS - 32) { // one TMA bulk store for the whole batch
                const size_t e0 = (size_t)(u_first + (uint32_t)first_out * nch) * 576;
                const uint32_t bytes = (uint32_t)((nb - first_out) * 576 * nch) * (uint32_t)sizeof(pcm_t);
                bulk_s2g(reinterpret_cast<pcm_t *>(pcm) + e0, stage + (size_t)first_out * 576 * nch, bytes);
            }
            for (int k = tid; k < 15 * FS; k += KF_THREADS) {
                Fm[k] = Fm[nb * 18 * FS + k];
                if (!mono) S.F[1][0][k] = S.F[1][nb * 18][k];
            }
        }
        __syncthreads();
    }
    if (tid == KF_THREADS - 32) bulk_wait_read_all(); // the last PCM store must have read the staging buffer
}

template <int FMT>
__global__ void __launch_bounds__(KF_THREADS, FMT == MP3B_PCM_S16 ? 4 : 3)
k_backend(const uint4 *__restrict__ tiles, uint32_t ntiles, const uint32_t *__restrict__ gran_unit0,
          const L3UnitDesc *__restrict__ units, const int16_t *__restrict__ is_in, const uint8_t *__restrict__ sf_in,
          const uint8_t *__restrict__ nzv_in, const L3BandTables *__restrict__ bands,
          const float *__restrict__ pow43, void *__restrict__ pcm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSharedT<FMT> &S = *reinterpret_cast<FusedSharedT<FMT> *>(smem_raw);
    if (blockIdx.x >= ntiles) return;
    const uint4 tl = tiles[blockIdx.x];
    const int warm = (int)tl.z, ng = (int)tl.y;    // granules before g0 to re-derive state from
    const uint32_t gstart = tl.x - (uint32_t)warm;  // first granule processed
    const uint32_t gu_first = gran_unit0[gstart];
    const uint32_t ubase = gu_first & L3G_UNIT_MASK; // units of a stream are contiguous
    if (gu_first & L3G_STEREO)
        backend_tile<FMT, false>(S, warm, ng + warm, ubase, units, is_in, sf_in, nzv_in, bands, pow43, pcm);
    else
        backend_tile<FMT, true>(S, warm, ng + warm, ubase, units, is_in, sf_in, nzv_in, bands, pow43, pcm);
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

    static float W[4][36];
    for (int i = 0; i < 36; i++) {
        W[0][i] = (float)sin(M_PI / 36.0 * (i + 0.5));
        W[1][i] = i < 18 ? (float)sin(M_PI / 36.0 * (i + 0.5))
                         : (i < 24 ? 1.f : (i < 30 ? (float)sin(M_PI / 12.0 * (i - 18 + 0.5)) : 0.f));
        W[2][i] = i < 12 ? (float)sin(M_PI / 12.0 * (i + 0.5)) : 0.f;
        W[3][i] = i < 6 ? 0.f : (i < 12 ? (float)sin(M_PI / 12.0 * (i - 6 + 0.5))
                                        : (i < 18 ? 1.f : (float)sin(M_PI / 36.0 * (i + 0.5))));
    }
    cudaMemcpyToSymbol(f_win, W, sizeof W);

    static float win[16][32];
    for (int l = 0; l < 16; l++)
        for (int j = 0; j < 32; j++) {
            const int i = l >> 1;
            double v;
            if (!(l & 1)) v = l3_dwin(64 * i + j) * (j <= 15 ? 1.0 : (j == 16 ? 0.0 : -1.0));
            else v = -l3_dwin(64 * i + 32 + j);
            win[l][j] = (float)v;
        }
    cudaMemcpyToSymbol(f_synwin, win, sizeof win);
    cudaFuncSetAttribute(k_backend<MP3B_PCM_S16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(FusedSharedT<MP3B_PCM_S16>));
    cudaFuncSetAttribute(k_backend<MP3B_PCM_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(FusedSharedT<MP3B_PCM_F32>));
}

void l3_launch_backend(const uint4 *tiles, uint32_t ntiles, const uint32_t *gran_unit0, const L3UnitDesc *units,
                       const int16_t *is_in, const uint8_t *sf_in, const uint8_t *nzv_in, const L3DevTables &T,
                       void *pcm, int pcm_format, cudaStream_t st)
{
    if (!ntiles) return;
    if (pcm_format == MP3B_PCM_S16)
        k_backend<MP3B_PCM_S16><<<ntiles, KF_THREADS, sizeof(FusedSharedT<MP3B_PCM_S16>), st>>>(tiles, ntiles, gran_unit0, units, is_in, sf_in,
                                                                  nzv_in, T.bands, T.pow43, pcm);
    else
        k_backend<MP3B_PCM_F32><<<ntiles, KF_THREADS, sizeof(FusedSharedT<MP3B_PCM_F32>), st>>>(tiles, ntiles, gran_unit0, units, is_in, sf_in,
                                                                  nzv_in, T.bands, T.pow43, pcm);
}
