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
#include "f32x2.h"
#include "fast_dct.h"
#include "fast_imdct.h"
#include "iso_tables.h"
#include "kernels.h"
#include "mp3b.h"

#ifndef KF_CTAS_S16
#define KF_CTAS_S16 4    // resident CTAs per SM the s16 kernel is compiled for (4: 64 registers; 3: 80)
#endif
#ifndef KF_MS_MERGE
#define KF_MS_MERGE 1    // S1: one copy of the long-block requantiser for MS and non-MS granules (smaller code:
                         // cfg3 -2.3 %, cfg4 -1.3 %, cfg2 unchanged)
#endif
#ifndef KF_GEN_ROLL
#define KF_GEN_ROLL 0    // S1 general path: the two lines of a pair as a rolled loop (smaller code)
#endif
#ifndef KF_WIN_F32X2
#define KF_WIN_F32X2 1   // S4: the synthesis window as packed FFMA2 (two output slots per instruction)
#endif
#ifndef KF_DCT_F32X2
#define KF_DCT_F32X2 1   // S3: the two 16-point halves of the 32-point transform as one packed pair
#endif

namespace {

constexpr int KF_B = 4;         // granules per batch
constexpr int KF_THREADS = 256;
constexpr int KF_ROWS = 15 + KF_B * 18;
constexpr int XROW = 18;        // subband row of X: unpadded -- S1 stores line pairs (8 bytes), S2 reads a row as nine
                                // 8-byte words, and at a stride of 9 such words the 16 lanes of a half warp hit 16 banks pairs
constexpr int XSZ = 32 * XROW;  // 576 floats per channel spectrum
constexpr int KF_POW_LUT = 1024; // sign(x) |x|^(4/3) for x = -512 .. 511, indexed by x & 1023
constexpr int FS = 36;         // row stride of F: 16-byte aligned rows, so S3 moves a row as eight 16-byte words (rows r .. r+7
                               // of a quarter warp start 4 banks apart: conflict-free); by lane = column any stride is

__constant__ float f_pow2q[4];
__constant__ float f_is_kl[7], f_is_kr[7];
__constant__ float f_lsf_pow[2][16];
__constant__ float f_cs[8], f_ca[8];
__constant__ uint8_t f_pretab[22];
__constant__ float f_win[4][36];
__device__ float f_synwin[2][16][32]; // [0] for float PCM; [1] scaled by 32768 for s16 PCM (exact: a power of two)

struct GranMeta {
    L3UnitDesc d[2];
    int row, lay[2];
    int slow;       // bit 16 c + 2q + h: the 64 lines from 128 q + 64 h on of unit c may hold |is| > 511 (S1)
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
    __align__(16) float X[KF_B][2][XSZ]; // spectra of the batch; PCM staging in S4; (s16) next batch's is
    __align__(16) float F[2][KF_ROWS][FS]; // rows 0..14: C history; rows 15..: first IMDCT halves -> S -> C
    __align__(16) float Hc[2][18][32]; // second IMDCT half of the last granule of the previous batch
    float pow43[KF_POW_LUT];        // sign(x) |x|^(4/3), x = -512 .. 511 at index x & 1023
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
// Huffman books whose escapes (linbits >= 9) can produce |is| > 511: table_select 22, 23, 29, 30, 31
constexpr uint32_t KF_BIG_TABLES = (1u << 22) | (1u << 23) | (1u << 29) | (1u << 30) | (1u << 31);

// f_pow2q[q & 3] * 2^(q >> 2), exactly (q >> 2 stays within the normal exponent range: -82 .. 11)
__device__ __forceinline__ float gain_of(int q) { return __int_as_float((127 + (q >> 2)) << 23) * f_pow2q[q & 3]; }

// s16: four CTAs per SM, f32: three -- n x (shared memory + 1 KB reserved per CTA) must fit 228 KB
static_assert(sizeof(FusedSharedT<MP3B_PCM_S16>) <= 56 * 1024, "s16 back end: too large for 4 CTAs per SM");
static_assert(sizeof(FusedSharedT<MP3B_PCM_F32>) <= 75 * 1024, "f32 back end: too large for 3 CTAs per SM");

__device__ __forceinline__ int16_t to_s16(float v) // v already in s16 units (the s16 window set is scaled)
{
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(v));
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
// The vector counts the fetch needs: nzv of unit `lane` (for the last warp's copies) and of unit `warp` (for the tail
// this warp zeroes), packed as lane's | warp's << 8.  Loaded ahead of prefetch_units so that the load's latency does
// not sit between a barrier and the copies.
__device__ __forceinline__ uint32_t prefetch_counts(int tid, uint32_t u_first, int n, const uint8_t *__restrict__ nzv_in)
{
    const int unit = tid >> 5, k = tid & 31;
    const uint32_t a = (unit == KF_THREADS / 32 - 1 && k < n) ? nzv_in[u_first + k] : 0u;
    const uint32_t b = unit < n ? nzv_in[u_first + unit] : 0u;
    return a | (b << 8);
}

template <class SH>
__device__ __forceinline__ void prefetch_units(SH &S, int tid, uint32_t u_first, int n,
                                               const int16_t *__restrict__ is_in, const uint8_t *__restrict__ sf_in,
                                               uint32_t counts)
{
    // spectra: one TMA bulk copy per unit, issued by the lanes of the last warp (which has no S3 work);
    // every warp zeroes the tail of "its" unit
    if ((tid >> 5) == KF_THREADS / 32 - 1) {
        const int k = tid & 31;
        const uint32_t b = 16u * (counts & 0xffu);
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
        const int nv = (int)(counts >> 8);
#pragma unroll
        for (int v = lane; v < 72; v += 32)
            if (v >= nv) *reinterpret_cast<uint4 *>(si + v * 16) = make_uint4(0, 0, 0, 0);
    }
    const char *gs = reinterpret_cast<const char *>(sf_in + (size_t)u_first * 40);
    char *ss = reinterpret_cast<char *>(&S.sf_buf[0][0]);
    for (int i = tid; i < n * 5; i += KF_THREADS) cp_async8(ss + i * 8, gs + i * 8);
    // the empty slot of an odd mono batch: scalefactors 0, so that its (unused) gains stay finite -- the merged
    // requantiser multiplies them by zero, and stale bytes can decode to an infinite gain
    if ((n & 1) && tid < 5) *reinterpret_cast<uint2 *>(ss + (n * 5 + tid) * 8) = make_uint2(0u, 0u);
    cp_async_commit();
}

// ---- per-batch metadata (descriptors, layouts), loaded one batch ahead ---------------------------
// A batch is held as pairs of unit slots: the two channels of a stereo granule, or two consecutive
// granules of a mono stream (which doubles the granules per batch, so that mono tiles keep all eight
// warps busy).  `nun` = units in the batch; an odd mono batch repeats its last descriptor.
template <class SH>
__device__ __forceinline__ void load_meta(SH &S, int tid, uint32_t u_first, int nun, int nch,
                                          const L3UnitDesc *__restrict__ units)
{
    for (int i = tid; i < KF_B * 40; i += KF_THREADS) {
        (&S.nz[0][0])[i] = 0;
        (&S.mode[0][0])[i] = 0;
    }
    // Everything else is the last warp's (it has no S3 work, so this is off the critical path): lane k takes
    // unit slot k, the two slots of a pair combine by shuffle.
    if ((tid >> 5) != KF_THREADS / 32 - 1) return;
    const int k = tid & 31, npairs = (nun + 1) >> 1;
    const bool mine = k < 2 * npairs;
    L3UnitDesc d;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(units + u_first + (uint32_t)min(mine ? k : 0, nun - 1));
        uint4 *dst = reinterpret_cast<uint4 *>(&d);
        dst[0] = __ldg(src);
        dst[1] = __ldg(src + 1);
    }
    const int lay = (d.flags & L3F_BT_MASK) == 2 ? ((d.flags & L3F_MIXED) ? 2 : 1) : 0;
    // Which 64-line spans (one warp's share of one S1 trip) can hold a value outside the small signed
    // |is|^(4/3) table: only the big_values regions coded with a book whose escapes reach that far.
    // Everything else -- other books (|is| <= 270), the count1 region (+-1), the zero tail -- cannot.
    int slow = 0;
    {
        const int lim[4] = {0, d.r1, d.r2, 2 * (int)d.big_values};
#pragma unroll
        for (int r = 0; r < 3; r++)
            if (((KF_BIG_TABLES >> d.tsel[r]) & 1u) && lim[r + 1] > lim[r]) {
                const int lo = lim[r] >> 6, hi = (lim[r + 1] - 1) >> 6;
                slow |= (2 << hi) - (1 << lo); // bits lo .. hi
            }
    }
    const int lay_o = __shfl_xor_sync(0xffffffffu, lay, 1), slow_o = __shfl_xor_sync(0xffffffffu, slow, 1);
    const bool ok = (d.flags & L3F_VALID) != 0;
    const int ms = (nch == 2 && ok && (d.hdr & L3H_MS)) ? 1 : 0, ist = (nch == 2 && ok && (d.hdr & L3H_IS)) ? 1 : 0;
    const bool head = mine && !(k & 1); // slot 0 of a pair writes the pair's record
    const uint32_t any_i = __ballot_sync(0xffffffffu, head && ist), any_s = __ballot_sync(0xffffffffu, mine && lay != 0);
    if (mine) S.gm[k >> 1].d[k & 1] = d;
    if (head) {
        GranMeta &m = S.gm[k >> 1];
        m.row = (d.hdr >> L3H_SR_SHIFT) & L3H_SR_MASK;
        m.lay[0] = lay;
        m.lay[1] = lay_o;
        m.ms = ms;
        m.ist = ist;
        m.slow = slow | (slow_o << 16);
    }
    if (k == 0) { S.any_ist = any_i != 0; S.any_short = any_s != 0; }
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
#if KF_MS_MERGE
                // (merged copy: the right channel of an MS granule is requantised with the opposite sign, see stage_requant)
                S.gain[gi][c][b] = gain_of(q) * (fold ? (c ? -0.70710678118654752440f : 0.70710678118654752440f) : 1.f);
#else
                S.gain[gi][c][b] = gain_of(q) * (fold ? 0.70710678118654752440f : 1.f);
#endif
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

// One spectral value of a packed pair (low / high half of a 32-bit word of the Huffman output) times its band gain.
// Fast: the signed table, for values known to lie in -512 .. 511 (index = the value's low ten bits, scaled to bytes).
template <class SH> __device__ __forceinline__ float rq_lo(const SH &S, uint32_t w, float g)
{
    return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(S.pow43) + ((w << 2) & 0xffcu)) * g;
}
template <class SH> __device__ __forceinline__ float rq_hi(const SH &S, uint32_t w, float g)
{
    return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(S.pow43) + ((w >> 14) & 0xffcu)) * g;
}
// Any value (escapes up to 8206): the full table in global memory beyond the shared one.  Branch-free: the shared
// lookup is always made (index clamped), the global one is a predicated load that only lanes with |v| > 511 perform.
template <class SH>
__device__ __forceinline__ float rq_any(const SH &S, int v, float gain, const float *__restrict__ pow43)
{
    const int m = v < 0 ? -v : v;
    float p = S.pow43[min(m, 511)];
    asm("{\n\t.reg .pred q;\n\tsetp.gt.s32 q, %1, 511;\n\t@q ld.global.nc.f32 %0, [%2];\n\t}" : "+f"(p) : "r"(m), "l"(pow43 + m));
    return __int_as_float(__float_as_int(p * gain) ^ (v & 0x80000000)); // (xor: the gain may carry a sign of its own)
}

// ---- S1c: requantise + stereo + reorder, one pair of adjacent lines of both channels per item -----------
// 64 threads per granule: thread t64 takes the line pairs t64 + 64 q, q < 5 (288 pairs; the second warp
// has no fifth trip).  A pair never straddles a scalefactor band (all band edges are even) nor a subband
// row (18 is even), so it shares one gain per channel and leaves as one 8-byte store.
// `bq` holds the long-block band index of this thread's five pairs, one byte each: they depend only on the
// stream's sample rate, so the common case (both channels long blocks, no intensity stereo) needs no table
// lookups, no reorder and no branches per line.
// First half: every thread pulls its five word pairs of the Huffman output into registers.  With s16
// output that buffer lives inside X, which the second half overwrites: the caller puts a barrier between
// the two.
template <class SH>
__device__ __forceinline__ void requant_load(SH &S, int tid, int nb, uint32_t (&v)[10])
{
    const int gi = tid >> 6, t64 = tid & 63;
    if (gi >= nb) return;
    const uint32_t *s0 = reinterpret_cast<const uint32_t *>(S.is_buf()[gi * 2]);
    const uint32_t *s1 = reinterpret_cast<const uint32_t *>(S.is_buf()[gi * 2 + 1]);
#pragma unroll
    for (int q = 0; q < 5; q++)
        if (q < 4 || t64 < 32) { // (warp-uniform)
            v[2 * q] = s0[t64 + 64 * q];
            v[2 * q + 1] = s1[t64 + 64 * q];
        }
}

template <class SH>
__device__ __forceinline__ void stage_requant(SH &S, int tid, int nb, const L3BandTables *__restrict__ bands,
                                              const float *__restrict__ pow43, const uint32_t (&bq)[2],
                                              const uint32_t (&v)[10])
{
    const float isq2 = 0.70710678118654752440f;
    const int gi = tid >> 6, t64 = tid & 63;
    if (gi >= nb) return;
    const GranMeta &m = S.gm[gi];
    const int lay0 = m.lay[0], lay1 = m.lay[1];
    const float *g0 = S.gain[gi][0], *g1 = S.gain[gi][1];
    // bit 2q (unit 0) / 16 + 2q (unit 1): this warp's span of trip q may hold values beyond the shared table
    const uint32_t slow = (uint32_t)m.slow >> ((tid >> 5) & 1);
    if ((lay0 | lay1) == 0 && !m.ist) {
        float2 *X0 = reinterpret_cast<float2 *>(S.X[gi][0]), *X1 = reinterpret_cast<float2 *>(S.X[gi][1]);
        // MS: (M +- S) / sqrt 2; stage_gains has folded the factor into this granule's band gains
        const float mf = m.ms ? 1.f : 0.f;
        (void)mf;
        auto lines = [&](auto ms_tag) {
#pragma unroll
            for (int q = 0; q < 5; q++) {
                if (q == 4 && t64 >= 32) break;
                uint32_t bw = bq[q >> 2];
                asm volatile("" : "+r"(bw)); // opaque: the gain addresses are cheap to form, costly to keep (spills)
                const int b = (bw >> (8 * (q & 3))) & 0xff;
                const float ga = g0[b], gb = g1[b];
                const uint32_t wa = v[2 * q], wb = v[2 * q + 1];
                // (the two lines of a pair times their band gain as ONE packed multiply)
                float2 A, C;
                if (slow & (1u << (2 * q))) // (warp-uniform)
                    A = make_float2(rq_any(S, (int)(short)(wa & 0xffffu), ga, pow43), rq_any(S, (int)wa >> 16, ga, pow43));
                else
                    A = f2_mul_s(ga, make_float2(rq_lo(S, wa, 1.f), rq_hi(S, wa, 1.f)));
                if (slow & (0x10000u << (2 * q)))
                    C = make_float2(rq_any(S, (int)(short)(wb & 0xffffu), gb, pow43), rq_any(S, (int)wb >> 16, gb, pow43));
                else
                    C = f2_mul_s(gb, make_float2(rq_lo(S, wb, 1.f), rq_hi(S, wb, 1.f)));
                const int p = t64 + 64 * q;
                const float a0 = A.x, a1 = A.y, c0 = C.x, c1 = C.y;
                (void)a0; (void)a1; (void)c0; (void)c1;
#if KF_MS_MERGE
                // c = -S for an MS granule (sign folded into the gains): M + S = a - mf c, M - S = mf a + c; mf = 0: (a, c)
                X0[p] = f2_fma_s(-mf, C, A);
                X1[p] = f2_fma_s(mf, A, C);
#else
                X0[p] = decltype(ms_tag)::value ? make_float2(a0 + c0, a1 + c1) : make_float2(a0, a1);
                X1[p] = decltype(ms_tag)::value ? make_float2(a0 - c0, a1 - c1) : make_float2(c0, c1);
#endif
            }
        };
#if KF_MS_MERGE
        lines(std::false_type{});
#else
        if (m.ms) lines(std::true_type{});
        else lines(std::false_type{});
#endif
        return;
    }
    // general path (short / mixed blocks, intensity stereo): band and reordered position of every line from the
    // per-rate line map (global memory, L1-resident)
    const uint2 *lm0 = reinterpret_cast<const uint2 *>(bands->lmap[m.row][lay0]);
    const uint2 *lm1 = reinterpret_cast<const uint2 *>(bands->lmap[m.row][lay1]);
    const bool ist = m.ist != 0, ms = m.ms != 0;
#pragma unroll
    for (int q = 0; q < 5; q++) {
        if (q == 4 && t64 >= 32) break;
        const int p = t64 + 64 * q;
        const uint2 e0 = __ldg(lm0 + p), e1 = __ldg(lm1 + p);
        const uint32_t wa = v[2 * q], wb = v[2 * q + 1];
        const bool sla = (slow & (1u << (2 * q))) != 0, slb = (slow & (0x10000u << (2 * q))) != 0;
#if KF_GEN_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int h = 0; h < 2; h++) {
            const uint32_t f0 = h ? e0.y : e0.x, f1 = h ? e1.y : e1.x;
            const int b1 = (int)(f1 & 0xffu);
            const float ga = g0[f0 & 0xffu], gb = g1[b1];
            float l, r;
            if (sla) l = rq_any(S, h ? (int)wa >> 16 : (int)(short)(wa & 0xffffu), ga, pow43);
            else l = h ? rq_hi(S, wa, ga) : rq_lo(S, wa, ga);
            if (slb) r = rq_any(S, h ? (int)wb >> 16 : (int)(short)(wb & 0xffffu), gb, pow43);
            else r = h ? rq_hi(S, wb, gb) : rq_lo(S, wb, gb);
            if (ist && S.mode[gi][b1]) { const float a = l; l = a * S.kl[gi][b1]; r = a * S.kr[gi][b1]; }
            else if (ms) { const float a = l, c = r; l = (a + c) * isq2; r = (a - c) * isq2; }
            S.X[gi][1][f1 >> 8] = r;
            S.X[gi][0][f0 >> 8] = l;
        }
    }
}

// ---- S2: alias reduction + IMDCT of one (granule, channel) by one warp, lane = subband ----------
// First halves go to the granule's rows of F.  Second halves (overlap-add with the NEXT granule) go to
// [slot t][32 subbands] with the 16-byte groups of row t XORed by t & 7: the granule's own, now dead, spectrum X
// -- S3 adds them to the next granule's rows while it loads those, 16 bytes at a time, and the swizzle keeps the
// eight rows a quarter warp reads together in different banks -- or, for the last granule of a batch, the carry
// buffer Hc (plain layout), whose previous content the warp of the batch's FIRST granule adds to its rows here.
// Those two warps touch Hc in the same phase: the reader announces on a named barrier (role 1) that its loads
// are done, the writer waits there (role 2) before storing; a batch of one granule is both (role 3: program order).
// (barrier 0 is __syncthreads; 1 and 2 serve the two channel sequences; immediates, so that no more are reserved)
__device__ __forceinline__ void named_arrive(int id)
{
    if (id == 1) asm volatile("barrier.arrive 1, 64;" ::: "memory");
    else asm volatile("barrier.arrive 2, 64;" ::: "memory");
}
__device__ __forceinline__ void named_sync(int id)
{
    if (id == 1) asm volatile("barrier.sync 1, 64;" ::: "memory");
    else asm volatile("barrier.sync 2, 64;" ::: "memory");
}

__device__ __forceinline__ void stage_imdct(float *__restrict__ X, int lane, uint8_t flags,
                                            float *__restrict__ Fdst /* [18][FS] */, float *__restrict__ Hc,
                                            int role, int bar_id)
{
    const bool reads_carry = (role & 1) != 0, writes_carry = (role & 2) != 0;
    float h[18];
    float x[18];
    {
        const float2 *Xr = reinterpret_cast<const float2 *>(X + lane * XROW); // 72-byte rows: 8-byte aligned
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const float2 t = Xr[k];
            x[2 * k] = t.x;
            x[2 * k + 1] = t.y;
        }
        __syncwarp(); // the warp's second halves will overwrite this spectrum
    }
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
        float Z[18];
        l3_dct4_18(x, Z); // the 18 distinct IMDCT values (fast_imdct.h)
        // window + frequency inversion; the window of each block type as immediates (a warp-uniform switch)
        auto emit = [&](auto bt_tag) {
            constexpr int BT = decltype(bt_tag)::value;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const float sa = Z[9 + i], sb = -Z[8 - i];
                // out[i] = sa, out[17-i] = -sa, out[18+i] = sb, out[35-i] = sb; slots i and 17 - i have
                // opposite parity, so exactly one of each pair takes the inversion sign
                const float sa_s = sa * sgn, sb_s = sb * sgn;
                Fdst[i * FS + lane] = ((i & 1) ? sa_s : sa) * WIN36[BT][i];
                Fdst[(17 - i) * FS + lane] = ((i & 1) ? sa : sa_s) * -WIN36[BT][17 - i];
                h[i] = ((i & 1) ? sb_s : sb) * WIN36[BT][18 + i];
                h[17 - i] = ((i & 1) ? sb : sb_s) * WIN36[BT][35 - i];
            }
        };
        if (bt == 0) emit(std::integral_constant<int, 0>{});
        else if (bt == 1) emit(std::integral_constant<int, 1>{});
        else emit(std::integral_constant<int, 3>{});
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
                y[wdw][i] = a * WIN36[2][i];
                y[wdw][5 - i] = -a * WIN36[2][5 - i];
                y[wdw][6 + i] = b * WIN36[2][6 + i];
                y[wdw][11 - i] = b * WIN36[2][11 - i];
            }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const float s_i = (i & 1) ? sgn : 1.f; // 6 and 12 are even: parity of i everywhere
            const float f0 = 0.f, f1 = y[0][i] * s_i, f2 = (y[0][6 + i] + y[1][i]) * s_i;
            Fdst[i * FS + lane] = f0;
            Fdst[(6 + i) * FS + lane] = f1;
            Fdst[(12 + i) * FS + lane] = f2;
            h[i] = (y[1][6 + i] + y[2][i]) * s_i;
            h[6 + i] = y[2][6 + i] * s_i;
            h[12 + i] = 0.f;
        }
    }
    if (reads_carry) { // the batch's first granule: + the previous batch's last second half (this lane's own column);
                       // a pass of its own: specialising the loops above for it costs more (code size) than it saves
#pragma unroll
        for (int t = 0; t < 18; t++) Fdst[t * FS + lane] += Hc[t * 32 + lane];
    }
    if (role == 1) named_arrive(bar_id); // (the sums are stored: the loads of Hc are complete)
    if (role == 2) named_sync(bar_id);
    if (writes_carry) {
#pragma unroll
        for (int t = 0; t < 18; t++) Hc[t * 32 + lane] = h[t];
    } else {
#pragma unroll
        for (int t = 0; t < 18; t++) X[t * 32 + (lane ^ ((t & 7) << 2))] = h[t];
    }
}

// ---- S4: synthesis window for one (granule, channel), lane = sample j ---------------------------
// rows r0 .. r0+32 of Fc are the transformed slots T0-15 .. T0+17.  Output slot T (0..17) is
//   sum_{l<16} W[l][j] * C_{T-l}[l even ? src_e : src_o]
// evaluated as a sliding accumulation over the rows: each row is loaded once (2 LDS) and feeds the
// up to 16 outputs it contributes to.  Everything is unrolled, so the 16-entry accumulator ring and
// the window taps live in registers.
#if KF_WIN_F32X2
// Packed form (FFMA2, sm_100): two consecutive output slots T = 2k, 2k + 1 share every weight,
//   (out[2k], out[2k+1]) = sum_i W[2i] (e[2a+1], e[2a+2]) + W[2i+1] (o[2a], o[2a+1]),  a = k + 7 - i,
// so an instruction takes a scalar weight (broadcast operand) times a PAIR of rows: 144 FFMA2 per granule and
// channel instead of 288 FFMA.  e / o: the row's even- / odd-tap column of this lane; rows are relative to r0.
template <typename Emit>
__device__ __forceinline__ void stage_window(const float *__restrict__ Fc, int r0, int src_e, int src_o,
                                             const float (&wn)[16], Emit emit)
{
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 16; a++) {
        const float *row = Fc + (r0 + 2 * a) * FS;
        const float2 ep = make_float2(row[FS + src_e], row[2 * FS + src_e]); // rows 2a + 1, 2a + 2
        const float2 op = make_float2(row[src_o], row[FS + src_o]);         // rows 2a, 2a + 1
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = a - 7 + i;
            if (k >= 0 && k <= 8) {
                acc[k & 7] = f2_fma_s(wn[2 * i], ep, acc[k & 7]);
                acc[k & 7] = f2_fma_s(wn[2 * i + 1], op, acc[k & 7]);
            }
        }
        if (a >= 7) {
            const int k = a - 7;
            emit(2 * k, acc[k & 7].x);
            emit(2 * k + 1, acc[k & 7].y);
            acc[k & 7] = make_float2(0.f, 0.f);
        }
    }
}
#else
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
#endif

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

    // (the kernel body has initialised the barrier, zeroed the history and built the |x|^(4/3) table)
    load_meta(S, tid, ubase, min(KFG, total) * nch, nch, units);
    prefetch_units(S, tid, ubase, min(KFG, total) * nch, is_in, sf_in,
                   prefetch_counts(tid, ubase, min(KFG, total) * nch, nzv_in));
    __syncthreads();

    const int src_e = lane <= 15 ? 16 + lane : (lane == 16 ? 0 : 48 - lane);
    const int src_o = lane <= 16 ? 16 - lane : lane - 16;
    pcm_t *stage = reinterpret_cast<pcm_t *>(&S.X[0][0][0]);
    uint32_t bq[2] = {0, 0}; // long-block band index of this thread's five line pairs (see stage_requant)
    {
        const int row0 = (units[ubase].hdr >> L3H_SR_SHIFT) & L3H_SR_MASK;
#pragma unroll
        for (int q = 0; q < 5; q++)
            if (q < 4 || (tid & 63) < 32)
                bq[q >> 2] |= (uint32_t)bands->line2band[row0][0][2 * ((tid & 63) + 64 * q)] << (8 * (q & 3));
    }

    for (int b0 = 0; b0 < total; b0 += KFG) {
        const int nb = min(KFG, total - b0);           // granules of this batch
        const int np = mono ? (nb + 1) >> 1 : nb;      // slot pairs of this batch
        const uint32_t u_first = ubase + (uint32_t)b0 * nch;
        // ---- S1
        cp_async_wait_all();                        // scalefactors (LDGSTS)
        mbar_wait(&S.bar, (uint32_t)(b0 / KFG) & 1u); // spectra (TMA); one phase per batch
        if (tid == KF_THREADS - 32) bulk_wait_read_all(); // the previous batch's PCM has left the staging buffer
        __syncthreads();
        {
            // the Huffman output into registers first: with s16 output it sits inside X, which stage_requant
            // overwrites -- the barrier behind the gains then covers that hazard too
            uint32_t v[10];
            requant_load(S, tid, np, v);
            stage_gains(S, tid, np, 2, bands);
            __syncthreads();
            if (S.any_ist) { // stereo only
                if ((tid & 32) == 0 && (tid >> 6) < np && S.gm[tid >> 6].ist) // first warp of each granule's 64 threads
                    stage_intensity(S, tid >> 6, (tid >> 6) * 2 + 1, lane, bands);
                __syncthreads();
            }
            stage_requant(S, tid, np, bands, pow43, bq, v);
        }
        __syncthreads();
        // ---- S2: alias + IMDCT; first halves -> F rows, second halves -> the granule's dead spectrum / Hc
        {
            // warp -> slot (gi, c); j = its granule inside the batch, seq = the row sequence it belongs to
            const int gi = warp >> 1, c = warp & 1;
            const int j = mono ? warp : gi;
            if (j < nb) {
                const int role = (j == 0 ? 1 : 0) | (j == nb - 1 ? 2 : 0);
                stage_imdct(S.X[gi][c], lane, S.gm[gi].d[c].flags, (mono ? Fm : &S.F[c][0][0]) + (15 + j * 18) * FS,
                            &S.Hc[mono ? 0 : c][0][0], role, 1 + (mono ? 0 : c));
            }
        }
        __syncthreads();
        // next batch: descriptors (gm is dead until the next S1), and the vector counts of its spectra
        uint32_t counts = 0;
        if (b0 + KFG < total) {
            load_meta(S, tid, u_first + (uint32_t)KFG * nch, min(KFG, total - b0 - KFG) * nch, nch, units);
            counts = prefetch_counts(tid, u_first + (uint32_t)KFG * nch, min(KFG, total - b0 - KFG) * nch, nzv_in);
        }
        // ---- S3: overlap-add + 32-point transform of every slot, in place; one thread per (channel, slot) row,
        // the whole transform in registers (fast_dct.h: 304 operations instead of 1024 FMAs).  The row's other
        // summand, the previous granule's second half, comes from that granule's X (rotated rows, see S2); the
        // batch's first granule took its carry in S2.
        {
            const int rows_c = nb * 18; // rows per sequence: stereo has two sequences, mono one
            if (tid < nch * rows_c) {
                const int c = tid >= rows_c ? 1 : 0, r = tid - c * rows_c;
                const int j = (r * 57) >> 10, t = r - 18 * j; // r / 18 for r < 160
                float4 *row = reinterpret_cast<float4 *>(&S.F[c][0][0] + (15 + r) * FS);
                float x[32];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float4 q = row[k];
                    x[4 * k] = q.x; x[4 * k + 1] = q.y; x[4 * k + 2] = q.z; x[4 * k + 3] = q.w;
                }
                if (j > 0) {
                    const int jp = j - 1; // slot of the previous granule: stereo (jp, c), mono (jp >> 1, jp & 1)
                    const float4 *hrow = reinterpret_cast<const float4 *>(
                        (mono ? &S.X[jp >> 1][jp & 1][0] : &S.X[jp][c][0]) + t * 32);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float4 q = hrow[k ^ (t & 7)];
                        x[4 * k] += q.x; x[4 * k + 1] += q.y; x[4 * k + 2] += q.z; x[4 * k + 3] += q.w;
                    }
                }
#if KF_DCT_F32X2
                l3_dct2_32_packed(x);
#else
                L3Dct2<32>::run(x);
#endif
#pragma unroll
                for (int k = 0; k < 8; k++) row[k] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
            }
        }
        __syncthreads();
        // next batch's spectra + scalefactors, behind S4 / S5 (not earlier: with s16 output the fetch buffer lies
        // inside X, whose second halves S3 has just consumed)
        if (b0 + KFG < total)
            prefetch_units(S, tid, u_first + (uint32_t)KFG * nch, min(KFG, total - b0 - KFG) * nch, is_in, sf_in,
                           counts);
        // ---- S4: window -> PCM staging (X is free now)
        {
            const int c = mono ? 0 : (warp & 1), j = mono ? warp : (warp >> 1);
            if (j < nb && b0 + j >= warm) {
                float wn[16]; // this lane's 16 taps (L1-resident; the s16 set carries the factor 32768)
#pragma unroll
                for (int l = 0; l < 16; l++) wn[l] = __ldg(&f_synwin[FMT == MP3B_PCM_S16 ? 1 : 0][l][lane]);
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
            {
                // the last 15 transformed slots become the history rows (16-byte words; stereo: two sequences)
                constexpr int NV = 15 * FS / 4;
                for (int i = tid; i < (mono ? NV : 2 * NV); i += KF_THREADS) {
                    float4 *seq = reinterpret_cast<float4 *>(i < NV ? Fm : &S.F[1][0][0]);
                    const int k = i < NV ? i : i - NV;
                    seq[k] = seq[nb * 18 * (FS / 4) + k];
                }
            }
        }
        // (no barrier here: nothing before the next batch's first barrier touches what S5 reads or writes)
    }
    if (tid == KF_THREADS - 32) bulk_wait_read_all(); // the last PCM store must have read the staging buffer
}

template <int FMT>
__global__ void __launch_bounds__(KF_THREADS, FMT == MP3B_PCM_S16 ? KF_CTAS_S16 : 3)
k_backend(const uint4 *__restrict__ tiles, uint32_t ntiles, const uint32_t *__restrict__ gran_unit0,
          const L3UnitDesc *__restrict__ units, const int16_t *__restrict__ is_in, const uint8_t *__restrict__ sf_in,
          const uint8_t *__restrict__ nzv_in, const L3BandTables *__restrict__ bands,
          const float *__restrict__ pow43, void *__restrict__ pcm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSharedT<FMT> &S = *reinterpret_cast<FusedSharedT<FMT> *>(smem_raw);
    pdl_launch_dependents();
    if (blockIdx.x >= ntiles) return;
    // ---- set-up that reads constant data only: it may run while the Huffman kernel is still draining (programmatic
    // dependent launch, kernels.h).  History starts at zero (stream head, or about to be re-derived by the warm-up
    // granules).
    {
        const int tid = threadIdx.x;
        if (tid == 0) mbar_init(&S.bar, 1);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < 2 * 15 * FS; i += KF_THREADS) S.F[i / (15 * FS)][0][i % (15 * FS)] = 0.f;
        static_assert(FS % 4 == 0, "rows move as 16-byte words");
        for (int i = tid; i < 2 * 144; i += KF_THREADS) reinterpret_cast<float4 *>(&S.Hc[i / 144][0][0])[i % 144] = z;
        for (int i = tid; i < KF_POW_LUT; i += KF_THREADS) { // signed: index = the value's low ten bits
            const int x = i < KF_POW_LUT / 2 ? i : i - KF_POW_LUT;
            S.pow43[i] = x < 0 ? -pow43[-x] : pow43[x];
        }
    }
    pdl_wait(); // everything below reads what the indexing and Huffman kernels wrote
    __syncthreads();
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

    static float win[2][16][32];
    for (int l = 0; l < 16; l++)
        for (int j = 0; j < 32; j++) {
            const int i = l >> 1;
            double v;
            if (!(l & 1)) v = l3_dwin(64 * i + j) * (j <= 15 ? 1.0 : (j == 16 ? 0.0 : -1.0));
            else v = -l3_dwin(64 * i + 32 + j);
            win[0][l][j] = (float)v;
            win[1][l][j] = (float)v * 32768.f;
        }
    cudaMemcpyToSymbol(f_synwin, win, sizeof win);
    cudaFuncSetAttribute(k_backend<MP3B_PCM_S16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(FusedSharedT<MP3B_PCM_S16>));
    cudaFuncSetAttribute(k_backend<MP3B_PCM_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(FusedSharedT<MP3B_PCM_F32>));
}

void l3_launch_backend(const uint4 *tiles, uint32_t ntiles, const uint32_t *gran_unit0, const L3UnitDesc *units,
                       const int16_t *is_in, const uint8_t *sf_in, const uint8_t *nzv_in, const L3DevTables &T,
                       void *pcm, int pcm_format, cudaStream_t st, bool pdl)
{
    if (!ntiles) return;
    if (pcm_format == MP3B_PCM_S16)
        l3_launch_k(k_backend<MP3B_PCM_S16>, dim3(ntiles), dim3(KF_THREADS), sizeof(FusedSharedT<MP3B_PCM_S16>), st, pdl, tiles,
                    ntiles, gran_unit0, units, is_in, sf_in, nzv_in, T.bands, T.pow43, pcm);
    else
        l3_launch_k(k_backend<MP3B_PCM_F32>, dim3(ntiles), dim3(KF_THREADS), sizeof(FusedSharedT<MP3B_PCM_F32>), st, pdl, tiles,
                    ntiles, gran_unit0, units, is_in, sf_in, nzv_in, T.bands, T.pow43, pcm);
}
