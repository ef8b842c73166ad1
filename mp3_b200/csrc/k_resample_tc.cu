// k_resample_tc.cu -- sample-rate conversion of stereo s16 PCM on the tensor cores (tcgen05, sm_100a).
//
// The FP32 resampler (k_resample.cu) is bound by shared-memory loads: 8 bytes per tap and output frame, 7 ms for the
// 1,024 x 10 s batch at 44.1 -> 48 kHz, whatever the FMA pipe could do.  The same arithmetic as a product the tensor
// core reads straight out of shared memory:
//     y[n] = sum_j hp[p(n)][j] x[q(n) - j],     u = n M + D,  q = u div L,  p = u mod L            (k_resample.cu)
// For a tile of 128 consecutive outputs n0 .. n0 + 127 all inputs lie in [base, base + K), base = q(n0) - (T - 1),
//     y[n0 + r] = sum_c A[r][c] x[base + c],    A[r][c] = hp[p_r][(q_r - q_0) + (T - 1) - c]  (0 outside the band),
// and A depends on n0 only through (n0 M + D) mod L: tiles t = n0 / 128 of the same "kind" t mod NK share it
// (NK = 5 for 44.1 -> 48 kHz).  So:  D[128 x N] = A_k[128 x K] * X[K x N],  one column of X per (tile, channel).
// Exact split-precision operands in fp16 (kind::f16, FP32 accumulation in TMEM):
//   * a sample x (s16) = 256 hi + lo with hi in -128 .. 127, lo in 0 .. 255: both halves are exact fp16 numbers (the
//     factor 256 is applied to the hi accumulator when it is read out);
//   * a coefficient c (float) = c1 + c2 + O(2^-22 c) with c1 = fp16(c), c2 = fp16(c - c1);
//   four products per K step: (c1 + c2)(256 hi + lo).  What is dropped is 2^-22 of a coefficient times a 16-bit
//   sample: below the FP32 kernel's own rounding.  Results equal the FP32 kernel's within its 1-LSB bound
//   (tests/test_resample.py runs both against scipy in float64).
// One persistent CTA per SM: A_k stays in shared memory (96 KB: hi, lo), X is built by all threads from the PCM arena
// (two buffers: the next 64 columns are staged while the tensor core works on these), one thread issues the MMAs
// (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 64, K = 16; completion by tcgen05.commit -> mbarrier), four warps
// read the 128 x 64 accumulator back (tcgen05.ld 32x32b) and store s16 frames, 128 contiguous bytes per warp and tile.
// Descriptor layouts follow CUTLASS cute/arch/mma_sm100_desc.hpp (as tools/tc_matrixing.cu, whose tf32 variant was
// measured for the synthesis matrixing and lost there: that stage's operands are produced in registers; these sit in
// memory anyway).  No reference code exists for this step (/root/reference/README.md:1-84).
#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "kernels.h"
#include "mp3b.h"

namespace {

constexpr int RT_ROWS = 128;     // outputs per tile = MMA M
constexpr int RT_COLS = 64;      // signal columns per group = 32 (tile, channel-pair) entries
constexpr int RT_PAIRS = RT_COLS / 2;
constexpr int RT_N = 2 * RT_COLS; // MMA N: the hi halves of the 64 columns, then their lo halves -- one MMA per coefficient
                                  // piece and K step yields A x hi and A x lo side by side; the read-out adds them
constexpr int RT_LBO_B = RT_N / 8 * 128 + 16; // K-chunk stride of the signal operand: 16 bytes of padding, so that the
                                              // threads that build it (one K chunk each) store to different banks
constexpr int RT_KMAX = 192;     // largest padded input window per tile (multiple of 16)
#ifndef RT_PWARPS_N
#define RT_PWARPS_N 8 /* (16 was measured with four read-out warps: no faster; with eight it does not fit the register file) */
#endif
constexpr int RT_PWARPS = RT_PWARPS_N;                 // producer warps (8 or 16)
constexpr int RT_EWARPS = 8;                           // read-out warps: warp e owns TMEM lanes 32 (e % 4) .., columns 32 (e / 4) ..
constexpr int RT_THREADS = 32 * (RT_EWARPS + RT_PWARPS + 2); // read-out warps, producer warps, 1 MMA-issue warp, 1 metadata warp

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle, 2-byte elements: 8-row x 16-byte (8 elements of K) core matrices.  Element (row r, k) of a
// [rows x K] operand at ((k / 8) * (rows / 8) + r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2 bytes;
// LBO (next 8 elements of K) = rows / 8 * 128, SBO (next 8 rows) = 128.
__host__ __device__ __forceinline__ uint32_t canon_off16(int rows, int r, int k)
{
    return (uint32_t)(((k >> 3) * (rows >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}
// SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48), SWIZZLE_NONE = 0 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// InstrDescriptor: c_format F32 = 1 [4,6), a_format / b_format F16 = 0 [7,10) / [10,13), A and B K-major (0) [15], [16],
// N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t RT_IDESC = (1u << 4) | ((uint32_t)(RT_N >> 3) << 17) | ((uint32_t)(RT_ROWS >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(RT_IDESC), "r"(accumulate) : "memory");
}

struct RtMeta {          // one (tile, channel pair) of the staged group
    long long in_off;    // element offset of the stream's PCM in the arena
    long long base;      // first input frame of the tile's window (may be negative, may run past in_n)
    long long out_off;   // element offset of the tile's first output frame
    int in_n;            // the stream's frames
    int rows;            // output frames of the tile that exist (0: padding entry)
};

static_assert(RT_THREADS <= 576, "more threads than registers: the read-out warps hold 64 accumulator values each");
// (measurement switches, all off: -DRT_DBG_NO_STORE / RT_DBG_NO_LOAD / RT_DBG_NO_MMA take the PCM stores, the PCM loads or
// all but one K step out of the kernel -- the numbers in DESIGN.md that show it is bound by neither)
constexpr int RT_META = 16; // ring of per-group metadata: written up to four groups ahead of the producers, read by the
                            // read-out warps up to four groups behind them
struct RtShared {
    RtMeta meta[RT_META][RT_PAIRS];
    __align__(8) uint64_t full[2];      // signal operand b staged (one producer thread arrives)
    __align__(8) uint64_t mma_done[2];  // the MMAs on operand / accumulator b have completed (tcgen05.commit)
    __align__(8) uint64_t tmem_free[2]; // accumulator b has been read out and stored (one lane per epilogue warp)
    uint32_t tmem_base;
    int job_cursor; // job of the last group's first entry: entries come in increasing order, the next search starts here
    volatile int meta_done; // groups whose metadata the metadata warp has published
    volatile int prod_done; // groups the producers have staged (throttles the metadata warp: the ring has RT_META slots)
    __align__(16) uint32_t scratch[RT_PWARPS][RT_KMAX]; // per producer warp: one input window on its way from lane = frame to lane = K chunk
};

__device__ __forceinline__ void rt_bar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rt_bar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_bar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok, spins = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap(); // never hang the device on a bad descriptor
    } while (!ok);
}

// (hi, lo) halves of eight stereo frames -> four 16-byte K chunks: left hi, left lo, right hi, right lo
__device__ __forceinline__ void split8(const uint32_t (&w)[8], uint4 &lhi, uint4 &llo, uint4 &rhi, uint4 &rlo)
{
    uint32_t hi[8], lo[8]; // per frame: (left, right) as half2
    const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f), k1152 = __floats2half2_rn(1152.f, 1152.f);
#pragma unroll
    for (int f = 0; f < 8; f++) {
        // 0x6400 | b is the fp16 number 1024 + b for b in 0 .. 1023
        uint32_t l = (w[f] & 0x00ff00ffu) | 0x64006400u;
        uint32_t h = ((w[f] >> 8) & 0x00ff00ffu) ^ 0x64806480u; // 1024 + (signed high byte + 128)
        __half2 lh = __hsub2(*reinterpret_cast<__half2 *>(&l), k1024);
        __half2 hh = __hsub2(*reinterpret_cast<__half2 *>(&h), k1152); // (the factor 256 is applied at the read-out)
        lo[f] = *reinterpret_cast<uint32_t *>(&lh);
        hi[f] = *reinterpret_cast<uint32_t *>(&hh);
    }
    lhi = make_uint4(__byte_perm(hi[0], hi[1], 0x5410), __byte_perm(hi[2], hi[3], 0x5410), __byte_perm(hi[4], hi[5], 0x5410),
                     __byte_perm(hi[6], hi[7], 0x5410));
    rhi = make_uint4(__byte_perm(hi[0], hi[1], 0x7632), __byte_perm(hi[2], hi[3], 0x7632), __byte_perm(hi[4], hi[5], 0x7632),
                     __byte_perm(hi[6], hi[7], 0x7632));
    llo = make_uint4(__byte_perm(lo[0], lo[1], 0x5410), __byte_perm(lo[2], lo[3], 0x5410), __byte_perm(lo[4], lo[5], 0x5410),
                     __byte_perm(lo[6], lo[7], 0x5410));
    rlo = make_uint4(__byte_perm(lo[0], lo[1], 0x7632), __byte_perm(lo[2], lo[3], 0x7632), __byte_perm(lo[4], lo[5], 0x7632),
                     __byte_perm(lo[6], lo[7], 0x7632));
}

// Roles: warps 0-3 read accumulators out and store PCM (warp w owns TMEM lanes 32 w ..), warps 4-11 build the signal
// operand (raw frames of the group after next are already in registers: two register sets, so that a full group's
// time hides the latency of the loads), one warp (one lane) issues the MMAs, one warp works out a few groups ahead which
// stream and tile every entry of a group is (dependent loads of the entry table: off everybody's critical path).  They meet only through mbarriers: full[b] -> MMA -> mma_done[b] ->
// epilogue -> tmem_free[b]; mma_done[b] also tells the producers that operand buffer b may be overwritten.
constexpr int RT_WIN = RT_PAIRS / RT_PWARPS; // input windows per producer warp and group
constexpr int RT_LD = RT_KMAX / 32;       // coalesced 32-frame loads per window

__global__ void __launch_bounds__(RT_THREADS, 1)
k_resample_tc(const int16_t *__restrict__ in, int16_t *__restrict__ out, const L3ResampleJob *__restrict__ jobs, int njobs,
              const uint32_t *__restrict__ prefix /* [NK][njobs + 1] */, const uint16_t *__restrict__ A_all, int L, int M,
              int taps, int half, int NK, int Kpad)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ RtShared S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kind = blockIdx.y;
    const uint32_t a_bytes = (uint32_t)RT_ROWS * Kpad * 2, b_bytes = (uint32_t)(Kpad >> 3) * RT_LBO_B;
    unsigned char *Ahi = dyn, *Alo = dyn + a_bytes;
    unsigned char *Bbuf = dyn + 2 * a_bytes; // [buffer][b_bytes]: columns 0 .. 63 hi halves, 64 .. 127 lo halves
    const uint32_t *pfx = prefix + (size_t)kind * (njobs + 1);
    const uint32_t npairs = pfx[njobs];
    const uint32_t ngroups = (npairs + RT_PAIRS - 1) / RT_PAIRS;
    if (blockIdx.x >= ngroups) return;
    const int nit = (int)((ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x); // groups of this CTA

    // ---- A_k (hi, lo) into shared memory, TMEM, barriers
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(A_all + (size_t)kind * 2 * RT_ROWS * Kpad);
        uint4 *dst = reinterpret_cast<uint4 *>(dyn);
        for (uint32_t i = tid; i < 2 * a_bytes / 16; i += RT_THREADS) dst[i] = __ldg(src + i);
    }
    if (warp == 0) { // 256 TMEM columns: two 128 x 128 FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        S.job_cursor = 0;
        S.meta_done = 0;
        S.prod_done = 0;
        for (int b = 0; b < 2; b++) {
            rt_bar_init(&S.full[b], 1);
            rt_bar_init(&S.mma_done[b], 1);
            rt_bar_init(&S.tmem_free[b], RT_EWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // A: generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const long long D = (long long)half * L;

    if (warp >= RT_EWARPS && warp < RT_EWARPS + RT_PWARPS) {
        // ================= producers: metadata and raw frames two groups ahead =================
        const int ptid = tid - 32 * RT_EWARPS;
        const int kchunks = Kpad >> 3;
        // A window = Kpad consecutive stereo frames (4 bytes each) of one stream.  Loads: lane = frame, 128 contiguous
        // bytes per instruction, the words of the group after next parked in registers.  Conversion needs 8 consecutive
        // frames per thread (one 16-byte K chunk per operand piece): the window passes through a 768-byte scratch row
        // of the warp (transpose), then lane kc builds chunk kc.
        const int pw = warp - RT_EWARPS;
        uint32_t *scratch = reinterpret_cast<uint32_t *>(S.scratch[pw]);
        auto load_raw = [&](int it, uint32_t (&w)[RT_WIN][RT_LD]) {
            for (uint32_t spins = 0; S.meta_done <= it;) // (the metadata warp runs a few groups ahead: normally no wait)
                if (++spins > (1u << 26)) __trap();
            __threadfence_block();
            const RtMeta *mt = S.meta[it & (RT_META - 1)];
#pragma unroll
            for (int q = 0; q < RT_WIN; q++) {
                const RtMeta &m = mt[pw + RT_PWARPS * q];
                const uint32_t *src = reinterpret_cast<const uint32_t *>(in + m.in_off);
                const long long base = m.base, in_n = m.rows > 0 ? m.in_n : 0;
#pragma unroll
                for (int f = 0; f < RT_LD; f++) {
                    const int x = lane + 32 * f;
                    const long long fi = base + x;
#ifdef RT_DBG_NO_LOAD
                    w[q][f] = (x < Kpad && fi >= 0 && fi < in_n) ? (uint32_t)fi : 0u;
#else
                    w[q][f] = (x < Kpad && fi >= 0 && fi < in_n) ? __ldg(src + fi) : 0u;
#endif
                }
            }
        };
        uint32_t w0[RT_WIN][RT_LD], w1[RT_WIN][RT_LD]; // raw frames of the groups `it` (even / odd)
        // one group: convert the register set into operand buffer b, publish it, refill the set for group it + 2
        auto step = [&](int it, uint32_t (&w)[RT_WIN][RT_LD]) {
            const int b = it & 1, k = it >> 1;
            if (k > 0) rt_bar_wait(&S.mma_done[b], (uint32_t)(k - 1) & 1u); // the MMAs that read this buffer are done
            unsigned char *Bb = Bbuf + (size_t)b * b_bytes;
#pragma unroll
            for (int q = 0; q < RT_WIN; q++) {
#pragma unroll
                for (int f = 0; f < RT_LD; f++) scratch[lane + 32 * f] = w[q][f];
                __syncwarp();
                if (lane < kchunks) {
                    const uint4 a0 = reinterpret_cast<const uint4 *>(scratch)[2 * lane], a1 = reinterpret_cast<const uint4 *>(scratch)[2 * lane + 1];
                    const uint32_t ww[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    uint4 lhi, llo, rhi, rlo;
                    split8(ww, lhi, llo, rhi, rlo);
                    // column c of the group: hi half at MMA column c, lo half at 64 + c
                    const int pi = pw + RT_PWARPS * q;
                    unsigned char *chunk = Bb + (size_t)lane * RT_LBO_B;
                    const int cl = 2 * pi, cr = 2 * pi + 1;
                    *reinterpret_cast<uint4 *>(chunk + (cl >> 3) * 128 + (cl & 7) * 16) = lhi;
                    *reinterpret_cast<uint4 *>(chunk + (cr >> 3) * 128 + (cr & 7) * 16) = rhi;
                    *reinterpret_cast<uint4 *>(chunk + ((cl + RT_COLS) >> 3) * 128 + (cl & 7) * 16) = llo;
                    *reinterpret_cast<uint4 *>(chunk + ((cr + RT_COLS) >> 3) * 128 + (cr & 7) * 16) = rlo;
                }
                __syncwarp();
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy stores -> visible to the tensor core
            asm volatile("barrier.sync 1, %0;" ::"n"(32 * RT_PWARPS) : "memory");
            if (ptid == 0) {
                rt_bar_arrive(&S.full[b]);
                S.prod_done = it + 1;
            }
            if (it + 2 < nit) load_raw(it + 2, w);
        };
        load_raw(0, w0);
        if (nit > 1) load_raw(1, w1);
        for (int it = 0; it < nit; it += 2) {
            step(it, w0);
            if (it + 1 < nit) step(it + 1, w1);
        }
    } else if (warp == RT_EWARPS + RT_PWARPS) {
        // ================= MMA issue: one thread =================
        if (lane == 0) {
            const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo);
            const uint32_t lbo_a = RT_ROWS / 8 * 128;
            for (int it = 0; it < nit; it++) {
                const int b = it & 1, k = it >> 1;
                rt_bar_wait(&S.full[b], (uint32_t)k & 1u);
                if (k > 0) rt_bar_wait(&S.tmem_free[b], (uint32_t)(k - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t bb = smem_u32(Bbuf + (size_t)b * b_bytes);
                const uint32_t d = tmem + (uint32_t)b * RT_N;
#ifdef RT_DBG_NO_MMA
                for (int ks = 0; ks < 1; ks++) {
#else
                for (int ks = 0; ks < (Kpad >> 4); ks++) { // K step = 16 elements = two 16-byte chunks
#endif
                    const uint64_t dah = make_desc(a_hi + 2 * ks * lbo_a, lbo_a, 128);
                    const uint64_t dal = make_desc(a_lo + 2 * ks * lbo_a, lbo_a, 128);
                    const uint64_t db = make_desc(bb + 2 * ks * RT_LBO_B, RT_LBO_B, 128);
                    mma_f16(d, dah, db, ks > 0);
                    mma_f16(d, dal, db, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"(smem_u32(&S.mma_done[b])) : "memory");
            }
        }
    } else if (warp == RT_EWARPS + RT_PWARPS + 1) {
        // ================= metadata: which stream and tile each of the group's 32 entries is, a few groups ahead =================
        // q0 of tile t = kind + NK j is q0(kind) + j * (NK 128 M / L): no division per entry (NK 128 M is a multiple of L)
        const long long q0k = ((long long)kind * RT_ROWS * M + D) / L, qstep = (long long)NK * RT_ROWS * M / L;
        for (int it = 0; it < nit; it++) {
            for (uint32_t spins = 0; S.prod_done < it - 3;) // slot it & 15: read last by the read-out of group it - 16, long finished
                if (++spins > (1u << 26)) __trap();
            const uint32_t e = (blockIdx.x + (uint32_t)it * gridDim.x) * RT_PAIRS + lane;
            RtMeta m;
            m.rows = 0;
            m.in_n = 0;
            m.in_off = m.base = m.out_off = 0;
            int lo = S.job_cursor; // last job with pfx[job] <= e: a short walk from where the previous group started
            __syncwarp();
            if (e < npairs) {
                while (lo + 1 < njobs && pfx[lo + 1] <= e) lo++;
                const L3ResampleJob jb = jobs[lo];
                const long long j = (long long)(e - pfx[lo]), t = kind + (long long)NK * j;
                m.in_off = jb.in_off;
                m.in_n = (int)jb.in_n;
                m.base = q0k + j * qstep - (taps - 1);
                m.out_off = jb.out_off + t * RT_ROWS * 2;
                const long long left = jb.out_n - t * RT_ROWS;
                m.rows = (int)(left < RT_ROWS ? left : RT_ROWS);
            }
            S.meta[it & (RT_META - 1)][lane] = m;
            if (lane == 0) S.job_cursor = lo;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) S.meta_done = it + 1;
        }
    } else {
        // ================= epilogue: row r = 32 warp + lane of the accumulator; frame i of the group = columns 2i, 2i + 1
        // (signal hi part) plus 64 + 2i, 64 + 2i + 1 (lo part) =================
        const int quarter = warp & 3, hcol = warp >> 2; // TMEM lanes 32 quarter .., signal columns 32 hcol .. (16 frames)
        const int r = quarter * 32 + lane;
        for (int it = 0; it < nit; it++) {
            const int b = it & 1, k = it >> 1;
            rt_bar_wait(&S.mma_done[b], (uint32_t)k & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const RtMeta *mt = S.meta[it & (RT_META - 1)] + hcol * 16;
            uint32_t v[32], u[32];
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * RT_N + hcol * 32);
#define RT_LD32(dst, addr)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
                 : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]), "=r"(dst[4]), "=r"(dst[5]), "=r"(dst[6]), "=r"(dst[7]), \
                   "=r"(dst[8]), "=r"(dst[9]), "=r"(dst[10]), "=r"(dst[11]), "=r"(dst[12]), "=r"(dst[13]), "=r"(dst[14]),      \
                   "=r"(dst[15]), "=r"(dst[16]), "=r"(dst[17]), "=r"(dst[18]), "=r"(dst[19]), "=r"(dst[20]), "=r"(dst[21]),    \
                   "=r"(dst[22]), "=r"(dst[23]), "=r"(dst[24]), "=r"(dst[25]), "=r"(dst[26]), "=r"(dst[27]), "=r"(dst[28]),    \
                   "=r"(dst[29]), "=r"(dst[30]), "=r"(dst[31])                                                                 \
                 : "r"(addr) : "memory")
            RT_LD32(v, taddr);
            RT_LD32(u, taddr + RT_COLS);
#undef RT_LD32
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // the accumulator is in registers: the MMAs of the group after next may overwrite it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) rt_bar_arrive(&S.tmem_free[b]);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const RtMeta &m = mt[i];
                int l16, r16;
                asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(l16) : "f"(fmaf(__uint_as_float(v[2 * i]), 256.f, __uint_as_float(u[2 * i]))));
                asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r16) : "f"(fmaf(__uint_as_float(v[2 * i + 1]), 256.f, __uint_as_float(u[2 * i + 1]))));
#ifdef RT_DBG_NO_STORE
                if (r < m.rows && l16 == 123456789)
#else
                if (r < m.rows)
#endif
                    *reinterpret_cast<uint32_t *>(out + m.out_off + 2 * r) = ((uint32_t)l16 & 0xffffu) | ((uint32_t)r16 << 16);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

} // namespace

// Plan for a rate pair: the tile kinds and their coefficient matrices (hi, lo) in the MMA's shared-memory layout.
// Returns false when the pair is not served by this path (window too long for the shared-memory budget, too many kinds).
bool l3_resample_tc_plan(const float *hp, int L, int M, int taps, int half, L3RsTcPlan *plan)
{
    if (L <= 0 || M <= 0 || taps <= 0) return false;
    const long long D = (long long)half * L;
    const long long step = ((long long)RT_ROWS * M) % L;
    int NK = 1;
    if (step) {
        long long a = L, b = step;
        while (b) { const long long t = a % b; a = b; b = t; }
        NK = (int)(L / a);
    }
    if (NK > 160) return false;
    int span_max = 0;
    for (int k = 0; k < NK; k++) {
        const long long n0 = (long long)k * RT_ROWS;
        const long long q0 = (n0 * M + D) / L, q1 = ((n0 + RT_ROWS - 1) * M + D) / L;
        span_max = std::max(span_max, (int)(q1 - q0));
    }
    const int Kpad = (span_max + taps + 15) / 16 * 16;
    if (Kpad > RT_KMAX) return false;
    plan->L = L;
    plan->M = M;
    plan->taps = taps;
    plan->half = half;
    plan->NK = NK;
    plan->Kpad = Kpad;
    plan->A.assign((size_t)NK * 2 * RT_ROWS * Kpad, 0);
    for (int k = 0; k < NK; k++) {
        uint16_t *hi = plan->A.data() + (size_t)k * 2 * RT_ROWS * Kpad, *lo = hi + (size_t)RT_ROWS * Kpad;
        const long long n0 = (long long)k * RT_ROWS;
        const long long q0 = (n0 * M + D) / L;
        for (int r = 0; r < RT_ROWS; r++) {
            const long long u = (n0 + r) * M + D, q = u / L;
            const int p = (int)(u - q * L);
            for (int j = 0; j < taps; j++) {
                const int c = (int)(q - q0) + (taps - 1) - j;
                const float v = hp[(size_t)p * taps + j];
                const __half h1 = __float2half_rn(v);
                const __half h2 = __float2half_rn(v - __half2float(h1));
                const uint32_t off = canon_off16(RT_ROWS, r, c) / 2;
                hi[off] = *reinterpret_cast<const uint16_t *>(&h1);
                lo[off] = *reinterpret_cast<const uint16_t *>(&h2);
            }
        }
    }
    return true;
}

// (tile, channel pair) entries per kind, cumulative over the jobs: prefix[k * (njobs + 1) + j]; mono jobs count zero.
// Returns the total number of entries.
unsigned long long l3_resample_tc_prefix(const L3RsTcPlan &plan, const L3ResampleJob *jobs, int njobs, std::vector<uint32_t> *prefix)
{
    prefix->assign((size_t)plan.NK * (njobs + 1), 0u);
    uint64_t total = 0;
    for (int k = 0; k < plan.NK; k++) {
        uint32_t *p = prefix->data() + (size_t)k * (njobs + 1);
        uint64_t run = 0;
        for (int j = 0; j < njobs; j++) {
            p[j] = (uint32_t)run;
            if (jobs[j].channels == 2 && (jobs[j].in_n > 0x7fffffffll || jobs[j].out_n > 0x7fffffffll)) return 0; // (32-bit frame counts in the kernel)
            if (jobs[j].channels == 2 && jobs[j].out_n > 0) {
                const long long ntile = (jobs[j].out_n + RT_ROWS - 1) / RT_ROWS;
                if (ntile > k) run += (uint64_t)((ntile - k + plan.NK - 1) / plan.NK);
            }
        }
        if (run > 0xfffffff0ull) return 0; // (does not fit the 32-bit table: the caller falls back)
        p[njobs] = (uint32_t)run;
        total += run;
    }
    return total;
}

void l3_launch_resample_tc(const void *in, void *out, const L3ResampleJob *jobs_dev, int njobs, const uint32_t *prefix_dev,
                           uint32_t max_entries_per_kind, const uint16_t *A_dev, const L3RsTcPlan &plan, int sm_count,
                           cudaStream_t st)
{
    const size_t smem = (size_t)2 * RT_ROWS * plan.Kpad * 2 + (size_t)2 * (plan.Kpad / 8) * RT_LBO_B;
    static std::atomic<unsigned long long> configured{0};
    if (l3_device_needs_setup(configured)) {
        cudaFuncSetAttribute(k_resample_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
        l3_device_setup_done(configured);
    }
    const uint32_t groups = (max_entries_per_kind + RT_PAIRS - 1) / RT_PAIRS;
    uint32_t gx = (uint32_t)std::max(1, sm_count / plan.NK); // one resident CTA per SM (196 KB of shared memory), one wave
    gx = std::min(gx, std::max(1u, groups));
    k_resample_tc<<<dim3(gx, (unsigned)plan.NK), RT_THREADS, smem, st>>>(
        static_cast<const int16_t *>(in), static_cast<int16_t *>(out), jobs_dev, njobs, prefix_dev, A_dev, plan.L, plan.M,
        plan.taps, plan.half, plan.NK, plan.Kpad);
}
