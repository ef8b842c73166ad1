// k_huffman.cu -- K1: scalefactor decode (a4) and Huffman big_values / count1 decode (a5).
//
// One thread per unit (granule-channel).  Decoding a unit is a serial bit walk (every codeword's
// position depends on the previous one), so the parallelism is across units: 1.57 M of them in
// the 1024-stream workload.  What the GPU mapping has to get right is
//   * the bits: the main data of a CTA's chunk of units (<= 512 consecutive units) is one
//     contiguous range of the arena, so it is staged into shared memory by ONE TMA bulk copy
//     (cp.async.bulk + mbarrier) and byte-swapped once; after that a peek is three LDS and two
//     funnel shifts, a skip is one add, and no global-load latency sits on the decode chain
//     (a chunk whose range does not fit falls back to a register-window reader over global memory
//     whose refill is predicated, not branched);
//   * table lookups: all 15 code books as one two-level LUT (9-bit root, 15.5 KB) in shared
//     memory, so a codeword costs two LDS, not a tree walk;
//   * convergence: one flat loop over pairs for all three regions (table, code length, escapes
//     and signs are data -- selects -- not control flow), then one loop over count1 quadruples
//     whose (symbol, sign bits) -> four lines mapping is a 256-entry table;
//   * balance: trip counts differ between units, so a chunk is sorted (big_values for the pair
//     loop, remaining bits for the count1 loop) into groups of 32 that the warps pull longest
//     first from a shared counter: lanes of a group run similar loops and warps finish together;
//   * stores: big_values pairs are packed 8 lines per 16-byte store (a shift register of four
//     words; all lanes push in lock step); count1 quadruples leave as two packed words.
// Output is bit-exact by construction (integer work only) and is compared word for word with
// oracle/l3_oracle.c::huffman_decode / read_scalefactors in tests/test_gpu_parity.py.
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include <algorithm>
#include <stdlib.h>

#include "kernels.h"

namespace {

__constant__ uint8_t c_lsf_nsfb[6][3][4] = {
    {{6, 5, 5, 5}, {9, 9, 9, 9}, {6, 9, 9, 9}},
    {{6, 5, 7, 3}, {9, 9, 12, 6}, {6, 9, 12, 6}},
    {{11, 10, 0, 0}, {18, 18, 0, 0}, {15, 18, 0, 0}},
    {{7, 7, 7, 0}, {12, 12, 12, 0}, {6, 15, 12, 0}},
    {{6, 6, 6, 3}, {12, 9, 9, 6}, {6, 12, 9, 6}},
    {{8, 8, 5, 0}, {15, 12, 9, 0}, {6, 18, 9, 0}},
};

constexpr int K1_THREADS = 256;
constexpr int K1_CHUNK = 512;       // most units per CTA (16 groups of 32 for 8 warps)
constexpr uint64_t K1_MAX_DYN_SMEM = 200 * 1024; // dynamic shared memory the kernel may be launched with
constexpr int K1_TAIL_PAD = 160;    // bytes staged beyond the last unit's part2_3 end (look-ahead, overruns)

__device__ __forceinline__ uint32_t bits_at(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit, int n);

// ---- bit reader over the staged (shared memory, already in bit order) main data -----------------
struct SmemReader {
    const uint32_t *w;
    uint32_t pos;       // bit position relative to the stage start
    uint64_t base_bit;  // arena bit offset of the stage start
    uint32_t span_bits; // staged bits

    __device__ __forceinline__ void init(const uint32_t *stage, uint32_t bit, uint64_t stage_bit0, uint32_t nbits)
    {
        w = stage;
        pos = bit;
        base_bit = stage_bit0;
        span_bits = nbits;
    }
    // n (0..5) bits at an absolute arena position (scfsi reuse): from the stage when it covers them
    __device__ __forceinline__ uint32_t bits_abs(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit, int n) const
    {
        const uint64_t rel = bit - base_bit;
        if (bit >= base_bit && rel + 64 <= span_bits) {
            const uint32_t r = (uint32_t)rel;
            const uint32_t *q = w + (r >> 5);
            return __funnelshift_l(__funnelshift_l(q[1], q[0], r), 0u, n);
        }
        return bits_at(arena, arena_bytes, bit, n);
    }
    __device__ __forceinline__ uint32_t bitpos() const { return pos; }
    __device__ __forceinline__ uint32_t peek32() const
    {
        const uint32_t *q = w + (pos >> 5);
        return __funnelshift_l(q[1], q[0], pos);
    }
    __device__ __forceinline__ void peek64(uint32_t &hi, uint32_t &lo) const
    {
        const uint32_t *q = w + (pos >> 5);
        const uint32_t a = q[0], b = q[1], c = q[2];
        hi = __funnelshift_l(b, a, pos); // the shift count wraps at 32
        lo = __funnelshift_l(c, b, pos);
    }
    __device__ __forceinline__ void skip(int k) { pos += (uint32_t)k; }
    __device__ __forceinline__ void skip_long(int k) { pos += (uint32_t)k; }
    __device__ __forceinline__ uint32_t get(int k) // k in 0..16 (0 reads nothing), branch-free
    {
        const uint32_t v = (peek32() >> 1) >> (31 - k);
        pos += (uint32_t)k;
        return v;
    }
};

// ---- the same stage through a three-word register window -----------------------------------------
// A peek costs no load at all (two funnel shifts out of w0..w2); a skip moves the window by a predicated load of ONE
// word, and only in the lanes that crossed a word boundary.  The kernel is bound by shared-memory wavefronts (every lane
// reads its own unit's bits, at random banks): this reader issues one load per 32 bits consumed instead of three per
// code word, at the price of a few more integer instructions per skip.
struct SmemWinReader {
    const uint32_t *w;      // stage base (random access for scfsi reuse)
    const uint32_t *p;      // word `wi` of the stage
    uint32_t w0, w1, w2;    // words wi, wi + 1, wi + 2
    uint32_t wi, pos;       // pos: bit position relative to the stage start
    uint64_t base_bit;
    uint32_t span_bits;

    __device__ __forceinline__ void init(const uint32_t *stage, uint32_t bit, uint64_t stage_bit0, uint32_t nbits)
    {
        w = stage;
        pos = bit;
        wi = bit >> 5;
        p = stage + wi;
        w0 = p[0];
        w1 = p[1];
        w2 = p[2];
        base_bit = stage_bit0;
        span_bits = nbits;
    }
    __device__ __forceinline__ uint32_t bits_abs(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit, int n) const
    {
        const uint64_t rel = bit - base_bit;
        if (bit >= base_bit && rel + 64 <= span_bits) {
            const uint32_t r = (uint32_t)rel;
            const uint32_t *q = w + (r >> 5);
            return __funnelshift_l(__funnelshift_l(q[1], q[0], r), 0u, n);
        }
        return bits_at(arena, arena_bytes, bit, n);
    }
    __device__ __forceinline__ uint32_t bitpos() const { return pos; }
    __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_l(w1, w0, pos); }
    __device__ __forceinline__ void peek64(uint32_t &hi, uint32_t &lo) const
    {
        hi = __funnelshift_l(w1, w0, pos); // the shift count wraps at 32
        lo = __funnelshift_l(w2, w1, pos);
    }
    __device__ __forceinline__ void advance()
    {
        const uint32_t cross = (pos >> 5) != wi ? 1u : 0u;
        uint32_t nw = w2;
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.u32 %0, [%1+12];\n\t}"
                     : "+r"(nw) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(cross));
        w0 = cross ? w1 : w0;
        w1 = cross ? w2 : w1;
        w2 = nw;
        p += cross;
        wi += cross;
    }
    __device__ __forceinline__ void skip(int k) // k in 0..32: at most one word boundary
    {
        pos += (uint32_t)k;
        advance();
    }
    __device__ __forceinline__ void skip_long(int k) // k in 0..64: a second boundary is rare (long escapes)
    {
        pos += (uint32_t)k;
        advance();
        if ((pos >> 5) != wi) advance();
    }
    __device__ __forceinline__ uint32_t get(int k) // k in 0..16 (0 reads nothing), branch-free
    {
        const uint32_t v = (peek32() >> 1) >> (31 - k);
        skip(k);
        return v;
    }
};
#ifndef K1_WINDOW
#define K1_WINDOW 0
#endif
#if K1_WINDOW
typedef SmemWinReader StageReader;
#else
typedef SmemReader StageReader;
#endif

// ---- fallback reader over the big-endian arena in global memory ---------------------------------
// Three consecutive words are cached in registers (the third one raw, loaded a word ahead).  The
// refill has no branch: the load is predicated and the window moves by selects, so lanes that cross
// a word boundary and lanes that do not stay on one instruction stream.  Reads past the arena
// return 0.  Positions are 32-bit and relative to the word the unit starts in.
struct GlobalReader {
    const uint32_t *words;
    uint32_t wlimit;
    uint32_t w0, w1, w2;
    uint32_t wi;
    uint32_t pos;

    __device__ __forceinline__ uint32_t load_raw(uint32_t w) const { return w < wlimit ? __ldg(words + w) : 0u; }
    __device__ __forceinline__ void init(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit_off)
    {
        const uint64_t w = bit_off >> 5, total = arena_bytes >> 2;
        words = reinterpret_cast<const uint32_t *>(arena) + w;
        const uint64_t left = total > w ? total - w : 0;
        wlimit = left > 0x7fffffffull ? 0x7fffffffu : (uint32_t)left;
        wi = 0;
        w0 = __byte_perm(load_raw(0), 0, 0x0123);
        w1 = __byte_perm(load_raw(1), 0, 0x0123);
        w2 = load_raw(2);
        pos = (uint32_t)(bit_off & 31);
    }
    __device__ __forceinline__ uint32_t bits_abs(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit, int n) const
    {
        return bits_at(arena, arena_bytes, bit, n);
    }
    __device__ __forceinline__ uint32_t bitpos() const { return pos; }
    __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_l(w1, w0, pos & 31); }
    __device__ __forceinline__ void peek64(uint32_t &hi, uint32_t &lo) const
    {
        const uint32_t sh = pos & 31;
        hi = __funnelshift_l(w1, w0, sh);
        lo = __funnelshift_l(__byte_perm(w2, 0, 0x0123), w1, sh);
    }
    __device__ __forceinline__ void advance()
    {
        const bool cross = (pos >> 5) != wi;
        const uint32_t nxt = wi + 3;
        const uint32_t ok = (cross && nxt < wlimit) ? 1u : 0u;
        uint32_t nw = 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}"
                     : "+r"(nw) : "l"(words + nxt), "r"(ok));
        w0 = cross ? w1 : w0;
        w1 = cross ? __byte_perm(w2, 0, 0x0123) : w1;
        w2 = cross ? nw : w2;
        wi += cross ? 1u : 0u;
    }
    __device__ __forceinline__ void skip(int k) // k in 0..32
    {
        pos += (uint32_t)k;
        advance();
    }
    __device__ __forceinline__ void skip_long(int k) // k in 0..64: may cross two words (rare)
    {
        pos += (uint32_t)k;
        advance();
        if ((pos >> 5) != wi) advance();
    }
    __device__ __forceinline__ uint32_t get(int k)
    {
        const uint32_t v = (peek32() >> 1) >> (31 - k);
        skip(k);
        return v;
    }
};

// read `n` (0..5) bits at an absolute bit position (used for scfsi reuse from granule 0)
__device__ __forceinline__ uint32_t bits_at(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit, int n)
{
    if (n == 0) return 0;
    const uint32_t *words = reinterpret_cast<const uint32_t *>(arena);
    uint32_t wend = (uint32_t)(arena_bytes >> 2), w = (uint32_t)(bit >> 5);
    uint32_t a = w < wend ? __byte_perm(__ldg(words + w), 0, 0x0123) : 0u;
    uint32_t b = w + 1 < wend ? __byte_perm(__ldg(words + w + 1), 0, 0x0123) : 0u;
    unsigned long long v = ((unsigned long long)a << 32) | b;
    return (uint32_t)((v << (bit & 31)) >> (64 - n));
}

// ---- TMA bulk copy global -> shared with an mbarrier for completion ------------------------------
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
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// exclusive scan of 160 histogram buckets by the first warp: 5 per lane
__device__ __forceinline__ void cta_scan160(uint32_t *hist)
{
    if (threadIdx.x < 32) {
        uint32_t v[5], sum = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) { v[k] = hist[threadIdx.x * 5 + k]; sum += v[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += t;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < 5; k++) { hist[threadIdx.x * 5 + k] = run; run += v[k]; }
    }
}

template <int CH>
struct K1SharedT {
    uint32_t info[32];        // per table_select: base | root << 16 | linbits << 24
    uint8_t quad[128];        // count1 book A (6-bit peek), then book B (4-bit code = 15 - value)
    uint2 c1[256];            // (symbol vwxy << 4 | next 4 bits) -> the four signed lines, packed
    uint32_t hist[160];
    uint16_t perm[CH];  // sorted order -> slot (unit of the chunk)
    uint32_t state[CH]; // per slot after the pair loop: bits consumed (13) | line index (10) << 13 | dead << 31
    uint8_t est[CH];    // per slot: estimated count1 iterations (sort key)
    unsigned long long lo_bit, hi_bit; // bit range of the chunk's main data in the arena
    uint32_t next;            // group counter of the running phase
    __align__(8) uint64_t bar;
};
typedef K1SharedT<K1_CHUNK> K1Shared;

// ---------------------------------------------------------------- part 2: scalefactors (a4)
template <class R>
__device__ __forceinline__ void read_scalefactors(R &br, const L3UnitDesc &d, uint32_t u, bool valid,
                                                  const L3UnitDesc *__restrict__ units,
                                                  const uint8_t *__restrict__ arena, uint64_t arena_bytes,
                                                  uint32_t (&sfw)[10])
{
    const int bt = d.flags & L3F_BT_MASK;
    const bool mixed = (d.flags & L3F_MIXED) != 0;
    const bool lsf = (d.hdr & L3H_LSF) != 0;
#pragma unroll
    for (int i = 0; i < 10; i++) sfw[i] = 0;
    if (valid) {
        if (!lsf) {
            // slen1 / slen2 nibble tables indexed by scalefac_compress
            const unsigned long long SL1 = 0x4433322211130000ull, SL2 = 0x3232132132103210ull;
            const int s1 = (int)((SL1 >> (4 * d.sfc)) & 15), s2 = (int)((SL2 >> (4 * d.sfc)) & 15);
            if (bt == 2) {
                const int n1 = mixed ? 17 : 18, ntot = mixed ? 35 : 36;
                // six fields (<= 24 bits) per peek
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const uint32_t v = br.peek32();
                    int sh = 0;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const int b = 6 * c + k;
                        const int w = b < n1 ? s1 : (b < ntot ? s2 : 0);
                        sfw[b >> 2] |= (((v << sh) >> 1) >> (31 - w)) << (8 * (b & 3));
                        sh += w;
                    }
                    br.skip(sh);
                }
            } else {
                const bool gr1 = (d.pos & L3P_GR) != 0;
                const uint32_t scfsi = gr1 ? (d.pos >> L3P_SCFSI_SHIFT) & 15u : 0u;
                // granule 0 of the same channel sits nch units earlier
                const L3UnitDesc *d0p = units + (u - ((d.hdr & L3H_STEREO) ? 2u : 1u));
                uint64_t g0_bit = 0;
                int g0_s1 = 0, g0_s2 = 0, g0_n1 = 11, g0_ntot = 21;
                if (scfsi) {
                    const L3UnitDesc d0 = *d0p;
                    g0_bit = d0.bit_off;
                    g0_s1 = (int)((SL1 >> (4 * d0.sfc)) & 15);
                    g0_s2 = (int)((SL2 >> (4 * d0.sfc)) & 15);
                    if ((d0.flags & L3F_BT_MASK) == 2) {
                        g0_n1 = (d0.flags & L3F_MIXED) ? 17 : 18;
                        g0_ntot = (d0.flags & L3F_MIXED) ? 35 : 36;
                    }
                }
                // the four scfsi groups (bands 0-5, 6-10, 11-15, 16-20): a group that is transmitted is <= 6 fields of
                // one width (<= 24 bits) and comes out of ONE peek -- 4 dependent reads per unit instead of 21
#pragma unroll
                for (int grp = 0; grp < 4; grp++) {
                    const int b0 = grp == 0 ? 0 : (grp == 1 ? 6 : (grp == 2 ? 11 : 16)), nb = grp == 0 ? 6 : 5;
                    if ((scfsi >> grp) & 1u) {
#pragma unroll
                        for (int k = 0; k < nb; k++) { // b-th transmitted scalefactor of granule 0
                            const int b = b0 + k;
                            const int w0 = b < g0_n1 ? g0_s1 : (b < g0_ntot ? g0_s2 : 0);
                            const uint64_t off = b < g0_n1 ? (uint64_t)(b * g0_s1)
                                                           : (uint64_t)(g0_n1 * g0_s1 + (b - g0_n1) * g0_s2);
                            sfw[b >> 2] |= br.bits_abs(arena, arena_bytes, g0_bit + off, w0) << (8 * (b & 3));
                        }
                    } else {
                        const int w = grp < 2 ? s1 : s2;
                        const uint32_t v = br.peek32();
#pragma unroll
                        for (int k = 0; k < nb; k++) {
                            const int b = b0 + k;
                            sfw[b >> 2] |= (((v << (k * w)) >> 1) >> (31 - w)) << (8 * (b & 3)); // (w = 0 reads as 0)
                        }
                        br.skip(nb * w);
                    }
                }
            }
        } else {
            int sfc = d.sfc, sl0, sl1, sl2, sl3, tbl;
            const bool ist = (d.hdr & L3H_IS) && (d.pos & L3P_CH);
            if (!ist) {
                if (sfc < 400) { sl0 = (sfc >> 4) / 5; sl1 = (sfc >> 4) % 5; sl2 = (sfc & 15) >> 2; sl3 = sfc & 3; tbl = 0; }
                else if (sfc < 500) { sfc -= 400; sl0 = (sfc >> 2) / 5; sl1 = (sfc >> 2) % 5; sl2 = sfc & 3; sl3 = 0; tbl = 1; }
                else { sfc -= 500; sl0 = sfc / 3; sl1 = sfc % 3; sl2 = 0; sl3 = 0; tbl = 2; }
            } else {
                sfc >>= 1;
                if (sfc < 180) { sl0 = sfc / 36; sl1 = (sfc % 36) / 6; sl2 = sfc % 6; sl3 = 0; tbl = 3; }
                else if (sfc < 244) { sfc -= 180; sl0 = (sfc & 63) >> 4; sl1 = (sfc & 15) >> 2; sl2 = sfc & 3; sl3 = 0; tbl = 4; }
                else { sfc -= 244; sl0 = sfc / 3; sl1 = sfc % 3; sl2 = 0; sl3 = 0; tbl = 5; }
            }
            const int lay = bt == 2 ? (mixed ? 2 : 1) : 0;
            const int e0 = c_lsf_nsfb[tbl][lay][0], e1 = e0 + c_lsf_nsfb[tbl][lay][1];
            const int e2 = e1 + c_lsf_nsfb[tbl][lay][2], e3 = e2 + c_lsf_nsfb[tbl][lay][3];
            // five fields (<= 25 bits) per peek
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint32_t pk = br.peek32();
                int sh = 0;
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int b = 5 * c + k;
                    const int w = b < e0 ? sl0 : (b < e1 ? sl1 : (b < e2 ? sl2 : (b < e3 ? sl3 : 0)));
                    uint32_t v = ((pk << sh) >> 1) >> (31 - w);
                    if (ist && w && v == (1u << w) - 1u) v |= 0x80u; // illegal intensity position
                    sfw[b >> 2] |= v << (8 * (b & 3));
                    sh += w;
                }
                br.skip(sh);
            }
        }
    }
}

// ---------------------------------------------------------------- part 3a: big_values pairs (a5)
// Returns the state word of the unit: bits consumed | line index << 13.
// The loop is bound by the integer (ALU) pipe, so it is written to keep that pipe's instruction
// count down: the table of the current region lives in registers and changes in a rarely taken
// branch; "the unit ran out of bits" is simply `position >= limit` (positions never move back, so
// no sticky flag: a pair that starts past the limit decodes to (0, 0) and consumes nothing, exactly
// what stopping at the first such pair does); the empty book is a real table whose two leaves are
// (0, 0) of length 0; bit fields come out of the stream with single funnel shifts; four pairs are
// decoded per trip, so the 16-byte store needs no shift register.
template <class R, class SH>
__device__ __forceinline__ uint32_t decode_pairs(R &br, uint32_t start, const L3UnitDesc &d, bool valid,
                                                 const uint16_t *__restrict__ s_lut, const SH &S,
                                                 uint32_t *__restrict__ out32)
{
    uint4 *out = reinterpret_cast<uint4 *>(out32);
    const int bv2 = valid ? d.big_values * 2 : 0;
    const int r1 = d.r1, r2 = d.r2;
    const uint32_t limit = start + d.p23len;
    int i = 0, nst = 0;
    // current region: table info unpacked, and the line at which the next region starts
    int reg = -1, nextb = 0;
    uint32_t base = 0;
    int root = 1, lin = 0;
    auto one = [&]() -> uint32_t {
        if (i >= nextb) { // region change: at most three times per unit
            uint32_t info;
            do {
                reg++;
                info = S.info[reg == 0 ? d.tsel[0] : (reg == 1 ? d.tsel[1] : d.tsel[2])];
                nextb = reg == 0 ? r1 : (reg == 1 ? r2 : 0x7fffffff);
            } while (i >= nextb);
            base = info & 0xffffu;
            root = (int)((info >> 16) & 0xffu);
            lin = (int)(info >> 24);
        }
        const bool act = br.bitpos() < limit;
        // 64 bits of look-ahead: code (<= 19) + escapes and signs (<= 28)
        uint32_t hi, lo;
        br.peek64(hi, lo);
        const uint32_t e1 = s_lut[base + __funnelshift_l(hi, 0u, root)];
        // codes longer than the root: one second-level lookup covers the rest
        // (always issued, as selects: a branch here would split the warp on almost every pair)
        const bool lng = (e1 & 0x8000u) != 0;
        const uint32_t w2 = (e1 >> 11) & 15u;
        const uint32_t sub = lng ? (e1 & 0x7ffu) + __funnelshift_l(hi << root, 0u, w2) : 0u;
        const uint32_t e2 = s_lut[base + sub];
        const uint32_t e = lng ? e2 : e1;
        const int len = (int)((e >> 8) & 15u) + (lng ? root : 0);
        int x = (e >> 4) & 15, y = e & 15;
        uint32_t rest = __funnelshift_l(lo, hi, len);
        const int lx = x == 15 ? lin : 0;
        x += (int)__funnelshift_l(rest, 0u, lx);
        rest <<= lx;
        const int sx = x != 0;
        {
            const int m = (int)rest >> 31; // sign bit (consumed only when x != 0; negating 0 gives 0 anyway)
            x = (x ^ m) - m;
        }
        rest <<= sx;
        const int ly = y == 15 ? lin : 0;
        y += (int)__funnelshift_l(rest, 0u, ly);
        rest <<= ly;
        const int sy = y != 0;
        {
            const int m = (int)rest >> 31;
            y = (y ^ m) - m;
        }
        const int n = len + lx + sx + ly + sy;
        br.skip_long(act ? n : 0);
        i += 2;
        return act ? __byte_perm((uint32_t)x, (uint32_t)y, 0x5410) : 0u;
    };
    while (i + 8 <= bv2) { // whole 16-byte vectors: the same trips for every lane of a sorted group
        const uint32_t a = one(), b = one(), c = one(), e = one();
        out[nst++] = make_uint4(a, b, c, e);
    }
    if (i < bv2) { // up to 3 pairs more, as single words (the count1 loop stores words as well)
        uint32_t *o = out32 + nst * 4;
        o[0] = one();
        if (i < bv2) {
            o[1] = one();
            if (i < bv2) o[2] = one();
        }
    }
    return min(br.bitpos() - start, 8191u) | ((uint32_t)i << 13);
}

// ---------------------------------------------------------------- part 3b: count1 quadruples (a5)
// code (<= 6 bits) + up to 4 sign bits per iteration; returns the line index after the last one.
// Real music has long count1 regions (the +-1 lines of the upper bands: 300 lines per granule at 128 kbit/s, where the
// generator's random spectra have 45), and with a thread per unit every store instruction touches 32 different units:
// two 4-byte stores per quadruple made this loop store-bound (135 SM cycles per warp and quadruple).  The quadruples are
// therefore collected in a 16-byte register window and leave as one 16-byte store per two of them: a quadruple's two
// words go to word position (i >> 1) & 3 of the window -- by selects, the position differs between lanes --, the window
// is stored when its last word has been written, and it starts out with the words the pair loop left in that vector.
// On return the vector that holds line i has been written completely (zeros behind the data).
template <class R, class SH>
__device__ __forceinline__ int decode_count1(R &br, uint32_t limit, int i, uint32_t qoff, const SH &S,
                                             uint32_t *__restrict__ out32)
{
    uint4 *out = reinterpret_cast<uint4 *>(out32);
    uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    {
        const int wp = (i >> 1) & 3; // words of the current vector that the pair loop has written already
        if (wp) {
            const uint4 cur = out[i >> 3];
            v0 = cur.x;
            v1 = wp > 1 ? cur.y : 0u;
            v2 = wp > 2 ? cur.z : 0u;
        }
    }
    // Two quadruples per trip while both are certain to be taken (<= 20 bits: one 32-bit peek holds them): the second
    // look-up does not wait for a new peek, and their four words complete exactly one vector.  The last quadruples of
    // a unit -- where the end of part2_3_length or of the spectrum has to be checked between them -- take the
    // single-step loop below, which has the exact rules.
    while (i <= 568 && br.bitpos() < limit) {
        const uint32_t bits = br.peek32();
        const uint32_t e1 = S.quad[qoff + (bits >> 26)];
        const uint32_t sym1 = e1 & 15u, len1 = e1 >> 4, n1 = len1 + __popc(sym1);
        const uint32_t bits2 = bits << n1;
        const uint32_t e2 = S.quad[qoff + (bits2 >> 26)];
        const uint32_t sym2 = e2 & 15u, len2 = e2 >> 4, n2 = len2 + __popc(sym2);
        const uint32_t p1 = br.bitpos() + n1;
        if (p1 >= limit || p1 + n2 > limit) break;
        br.skip((int)(n1 + n2));
        const uint2 a = S.c1[(sym1 << 4) | ((bits << len1) >> 28)], b = S.c1[(sym2 << 4) | ((bits2 << len2) >> 28)];
        const int wp = (i >> 1) & 3; // (the same in every trip of this loop)
        out[i >> 3] = make_uint4(wp == 0 ? a.x : v0, wp == 0 ? a.y : (wp == 1 ? a.x : v1),
                                 wp == 0 ? b.x : (wp == 1 ? a.y : (wp == 2 ? a.x : v2)),
                                 wp == 0 ? b.y : (wp == 1 ? b.x : (wp == 2 ? a.y : a.x)));
        v0 = wp == 1 ? b.y : (wp == 2 ? b.x : (wp == 3 ? a.y : 0u));
        v1 = wp == 2 ? b.y : (wp == 3 ? b.x : 0u);
        v2 = wp == 3 ? b.y : 0u;
        i += 8;
    }
    while (i <= 572 && br.bitpos() < limit) {
        const uint32_t bits = br.peek32();
        const uint32_t e = S.quad[qoff + (bits >> 26)];
        const uint32_t sym = e & 15u, len = e >> 4;
        const uint32_t s4 = (bits << len) >> 28;
        br.skip((int)(len + __popc(sym)));
        if (br.bitpos() > limit) break; // overran part2_3_length: discard this quadruple
        const uint2 vw = S.c1[(sym << 4) | s4];
        const int wp = (i >> 1) & 3;
        v0 = wp == 0 ? vw.x : v0;
        v1 = wp == 0 ? vw.y : (wp == 1 ? vw.x : v1);
        v2 = wp == 1 ? vw.y : (wp == 2 ? vw.x : v2);
        v3 = wp == 2 ? vw.y : (wp == 3 ? vw.x : v3);
        if (wp >= 2) { // the vector is complete
            out[i >> 3] = make_uint4(v0, v1, v2, v3);
            v0 = wp == 3 ? vw.y : 0u; // (the quadruple's second word opens the next vector)
            v1 = v2 = v3 = 0u;
        }
        i += 4;
    }
    if ((i >> 1) & 3) out[i >> 3] = make_uint4(v0, v1, v2, v3); // the last, partly filled vector: zeros behind the data
    return i;
}

// NT threads per CTA (256: four CTAs per SM; 512: two, with chunks of up to 1,024 units -- more groups per warp to
// balance, see the launcher), chunks of up to 2 NT units
template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)
k_huffman(const uint8_t *__restrict__ arena, uint64_t arena_bytes, const L3UnitDesc *__restrict__ units,
          uint32_t u_lo, uint32_t nunits, const uint16_t *__restrict__ g_lut, uint32_t lut_len,
          uint32_t stage_bytes, uint32_t chunk, const L3HuffInfo *__restrict__ g_info, const uint8_t *__restrict__ g_quad,
          int16_t *__restrict__ is_out, uint8_t *__restrict__ sf_out, uint8_t *__restrict__ nzv_out, int zero_fill)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ K1SharedT<2 * NT> S;
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem_raw);                // stage_bytes
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(smem_raw + stage_bytes);  // lut_len entries
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t chunk_first = blockIdx.x * chunk;
    const int n = (int)min(chunk, nunits - chunk_first);
    const L3UnitDesc *cu = units + u_lo + chunk_first;

    // ---- code tables into shared memory: constant data, so this part may run while the kernel before this one in the
    // stream is still draining (programmatic dependent launch, kernels.h)
    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(&S.bar, 1);
        S.lo_bit = ~0ull;
        S.hi_bit = 0ull;
        S.next = 0;
    }
    for (int k = tid; k < 160; k += NT) S.hist[k] = 0;
    {   // 16 bytes per load (the table's allocation and its shared copy are both padded to a multiple of 16 bytes)
        const uint4 *src = reinterpret_cast<const uint4 *>(g_lut);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t k = tid; k < (lut_len * 2 + 15) / 16; k += NT) dst[k] = __ldg(src + k);
    }
    if (tid < 32)
        S.info[tid] = g_info->base[tid] | ((uint32_t)g_info->root[tid] << 16) | ((uint32_t)g_info->linbits[tid] << 24);
    if (tid < 64) S.quad[tid] = g_quad[tid];
    else if (tid < 128) S.quad[tid] = (uint8_t)((4u << 4) | (15u - ((tid - 64u) >> 2)));
    if (tid < 256) {
        const uint32_t sym = tid >> 4, s4 = tid & 15u;
        uint32_t val[4], k = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            val[c] = 0;
            if (sym & (8u >> c)) { val[c] = ((s4 >> (3 - k)) & 1u) ? 0xffffu : 1u; k++; }
        }
        S.c1[tid] = make_uint2(val[0] | (val[1] << 16), val[2] | (val[3] << 16));
    }
    pdl_wait(); // unit descriptors and the main-data arena: written by the indexing kernels
    __syncthreads();
    // ---- histogram, the chunk's bit range
    uint32_t bucket[2];
    {
        unsigned long long lo = ~0ull, hi = 0ull;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int slot = tid + r * NT;
            bucket[r] = 0;
            if (slot < n) {
                const L3UnitDesc dd = cu[slot];
                if (dd.flags & L3F_VALID) {
                    bucket[r] = (uint32_t)min((int)dd.big_values, 288) >> 1;
                    lo = min(lo, (unsigned long long)dd.bit_off);
                    hi = max(hi, (unsigned long long)dd.bit_off + dd.p23len);
                }
                atomicAdd(&S.hist[bucket[r]], 1u);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0 && hi) {
            atomicMin(&S.lo_bit, lo);
            atomicMax(&S.hi_bit, hi);
        }
    }
    __syncthreads();
    // stage the range [a0, a1) if it fits: one TMA bulk copy, issued now, awaited after the sort
    const uint64_t a0 = S.hi_bit ? ((S.lo_bit >> 3) & ~15ull) : 0ull;
    const uint64_t a1 = S.hi_bit ? ((((S.hi_bit + 7) >> 3) + K1_TAIL_PAD + 15) & ~15ull) : 0ull;
    const bool staged = a1 - a0 <= stage_bytes;
    const uint32_t span = (uint32_t)(a1 - a0);
    const uint32_t ncopy = staged ? (uint32_t)(min(a1, arena_bytes) > a0 ? min(a1, arena_bytes) - a0 : 0ull) : 0u;
    if (tid == 0 && ncopy) {
        mbar_expect_tx(&S.bar, ncopy);
        bulk_g2s(stage, arena + a0, ncopy, &S.bar);
    }
    // sort the chunk by big_values, longest first
    cta_scan160(S.hist);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int slot = tid + r * NT;
        if (slot < n) S.perm[n - 1 - (int)atomicAdd(&S.hist[bucket[r]], 1u)] = (uint16_t)slot;
    }
    if (ncopy) {
        mbar_wait(&S.bar, 0);
        for (uint32_t k = tid; k < (ncopy >> 2); k += NT) stage[k] = __byte_perm(stage[k], 0, 0x0123);
    }
    if (staged)
        for (uint32_t k = (ncopy >> 2) + tid; k < (span >> 2); k += NT) stage[k] = 0u; // past the arena
    __syncthreads();
    for (int k = tid; k < 160; k += NT) S.hist[k] = 0; // reused by the count1 sort
    __syncthreads();

    const int ngroups = (n + 31) >> 5;
    // ---- phase 1: scalefactors + big_values pairs; warps pull groups of 32 units, longest first
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(&S.next, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if ((int)g >= ngroups) break;
        const int idx = (int)g * 32 + lane;
        if (idx >= n) continue;
        const uint32_t slot = S.perm[idx];
        const uint32_t u = u_lo + chunk_first + slot;
        const L3UnitDesc d = units[u];
        const bool valid = (d.flags & L3F_VALID) != 0;
        uint32_t sfw[10];
        uint32_t *out32 = reinterpret_cast<uint32_t *>(is_out + (size_t)u * 576);
        uint32_t stw;
        if (staged) {
            StageReader br;
            br.init(stage, valid ? (uint32_t)(d.bit_off - a0 * 8) : 0u, a0 * 8, span * 8);
            const uint32_t start = br.bitpos();
            read_scalefactors(br, d, u, valid, units, arena, arena_bytes, sfw);
            stw = decode_pairs(br, start, d, valid, s_lut, S, out32);
        } else {
            GlobalReader br;
            br.init(arena, arena_bytes, d.bit_off);
            const uint32_t start = br.bitpos();
            read_scalefactors(br, d, u, valid, units, arena, arena_bytes, sfw);
            stw = decode_pairs(br, start, d, valid, s_lut, S, out32);
        }
        {
            uint2 *o = reinterpret_cast<uint2 *>(sf_out + (size_t)u * 40);
#pragma unroll
            for (int k = 0; k < 5; k++) o[k] = make_uint2(sfw[2 * k], sfw[2 * k + 1]);
        }
        // count1 work left, estimated from the bits left (the sort key of phase 2)
        uint32_t est = 0;
        if (!(stw & 0x80000000u)) {
            const int used = (int)(stw & 0x1fffu), i = (int)((stw >> 13) & 0x3ffu);
            const int by_bits = ((int)d.p23len - used + 3) >> 2, by_lines = (576 - i) >> 2;
            est = (uint32_t)max(0, min(min(by_bits, by_lines) + 1, 159));
        }
        S.state[slot] = stw;
        S.est[slot] = (uint8_t)est;
        atomicAdd(&S.hist[est], 1u);
    }
    __syncthreads();
    if (tid == 0) S.next = 0;
    cta_scan160(S.hist);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int slot = tid + r * NT;
        if (slot < n) S.perm[n - 1 - (int)atomicAdd(&S.hist[S.est[slot]], 1u)] = (uint16_t)slot;
    }
    __syncthreads();

    // ---- phase 2: count1 quadruples, zero padding of the last vector, vector counts
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(&S.next, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if ((int)g >= ngroups) break;
        const int idx = (int)g * 32 + lane;
        if (idx >= n) continue;
        const uint32_t slot = S.perm[idx];
        const uint32_t u = u_lo + chunk_first + slot;
        const L3UnitDesc d = units[u];
        const uint32_t stw = S.state[slot];
        const uint32_t used = stw & 0x1fffu, p23 = d.p23len;
        int i = (int)((stw >> 13) & 0x3ffu);
        uint32_t *out32 = reinterpret_cast<uint32_t *>(is_out + (size_t)u * 576);
        bool c1 = false; // decode_count1 ran: it completes the last vector itself
        if (!(stw & 0x80000000u) && used < p23) {
            const uint32_t qoff = (d.flags & L3F_C1TAB) ? 64u : 0u;
            c1 = true;
            if (staged) {
                StageReader br;
                br.init(stage, (uint32_t)(d.bit_off - a0 * 8) + used, a0 * 8, span * 8);
                i = decode_count1(br, br.bitpos() + (p23 - used), i, qoff, S, out32);
            } else {
                GlobalReader br;
                br.init(arena, arena_bytes, d.bit_off + used);
                i = decode_count1(br, br.bitpos() + (p23 - used), i, qoff, S, out32);
            }
        }
        // zero the rest of the last 16-byte vector.  The all-zero tail of the spectrum is not written:
        // nzv_out[u] tells the consumer how many 16-byte vectors (8 lines each) hold data, and it treats
        // the rest as zero.  zero_fill (staged pipeline, parity dumps) writes the tail anyway.
        const int nst = (i + 7) >> 3;
        if (!c1)
            for (int k = i >> 1; k < nst * 4; k++) out32[k] = 0u;
        nzv_out[u] = (uint8_t)nst;
        if (zero_fill) {
            uint4 *out = reinterpret_cast<uint4 *>(out32);
            for (int k = nst; k < 72; k++) out[k] = make_uint4(0, 0, 0, 0);
        }
    }
}


// =====================================================================================================================
// The sorted variant (MP3B_K1_MODE=sorted; measured, NOT the default -- numbers in DESIGN.md): the same decode with
// the balance across warps the chunked kernel cannot have.
//
// In k_huffman a CTA sorts ITS chunk, so its eight warps get groups of very different length (a chunk holds 9 groups,
// longest first: ncu shows 17.6 % of all warp time waiting at the barrier behind the pair phase, and 28 of 32 lanes
// active inside it).  Here blocks of 2,048 consecutive units are counting-sorted by length (k_hsort_local) and ONE
// persistent CTA per SM runs up to 32 independent warps: a warp pulls the next 32 units of the sorted order from a
// global counter, copies their bits into its private part of shared memory (the lanes together, contiguous bytes per
// unit, byte-swapped on the way) and decodes its 32 units, which now have the same trip count (31.8 of 32 lanes
// active).  No CTA barrier after the tables are loaded; tables are loaded once per SM, not once per chunk.  Two passes,
// each with its own order: A = scalefactors + big_values pairs, sorted by (big_values, block type); B = count1
// quadruples + the zero tail, sorted by the bits pass A left.  A group whose bits do not fit the warp's part of shared
// memory (very high bitrates) reads global memory through the register window instead.
// Why blocks and not the whole wave: with a wave-wide order the 32 lanes of a store instruction hit 32 different
// 2-MB pages and the load / store unit replays it page by page (1.66 ms); inside a block they share one or two.
// Why it still loses (1.18 ms against 0.875): every group starts with a chain of dependent global loads (counter, order,
// descriptor, bits: 27 % of pass A's warp time is long-scoreboard stall, 55 % of pass B's), which the chunked kernel
// pays once per 288 units with one TMA copy.
constexpr int HS_KEYS = 512;           // pass A: (min(big_values, 288) >> 1) * 2 + (block_type == 2); pass B: see hs_key_b
constexpr int HS_SORT_THREADS = 256;
constexpr int HS_SORT_ITEMS = 8;       // units per thread of the sort kernel
constexpr int HS_BLOCK = HS_SORT_THREADS * HS_SORT_ITEMS; // units sorted together
constexpr int HS_CTL_WORDS = 4;        // control block (device words): the group counters of the two passes

__device__ __forceinline__ uint32_t hs_key_a(const L3UnitDesc &d)
{
    if (!(d.flags & L3F_VALID)) return 0u;
    return ((uint32_t)min((int)d.big_values, 288) >> 1) * 2u + ((d.flags & L3F_BT_MASK) == 2 ? 1u : 0u);
}
// count1 work left behind the pairs: by the bits (a quadruple takes about five) and by the lines that are left
__device__ __forceinline__ uint32_t hs_key_b(const L3UnitDesc &d, uint32_t stw)
{
    const uint32_t used = stw & 0x1fffu, i = (stw >> 13) & 0x3ffu;
    if (!(d.flags & L3F_VALID) || used >= d.p23len || i > 572u) return 0u;
    const uint32_t by_bits = ((uint32_t)d.p23len - used) >> 2, by_lines = (576u - i) >> 1;
    return (1u + min(min(by_bits, by_lines), 254u)) * 2u + ((d.flags & L3F_C1TAB) ? 1u : 0u);
}

// Counting sort of each block of HS_BLOCK consecutive units by descending key, in shared memory: perm[position] = unit
// (both relative to u_lo; a block keeps its place, so a group of 32 consecutive positions holds units that lie within
// HS_BLOCK units of each other).  Order inside a key is arbitrary (atomics); the decoded output does not depend on it.
// PASS 0 computes the key from the descriptor, PASS 1 reads the key pass A of the decode left in `keys`.
template <int PASS>
__global__ void __launch_bounds__(HS_SORT_THREADS)
k_hsort_local(const L3UnitDesc *__restrict__ units, uint32_t u_lo, uint32_t nunits, const uint16_t *__restrict__ keys,
              uint32_t *__restrict__ perm)
{
    __shared__ uint32_t cnt[HS_KEYS];   // count per key, then the first position of the key inside the block
    pdl_launch_dependents();
    for (int k = threadIdx.x; k < HS_KEYS; k += HS_SORT_THREADS) cnt[k] = 0;
    pdl_wait();
    __syncthreads();
    const uint32_t base = blockIdx.x * HS_BLOCK;
    uint32_t key[HS_SORT_ITEMS], rank[HS_SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < HS_SORT_ITEMS; r++) {
        const uint32_t i = base + r * HS_SORT_THREADS + threadIdx.x;
        key[r] = i < nunits ? (PASS == 0 ? hs_key_a(units[u_lo + i]) : (uint32_t)keys[i]) : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < HS_SORT_ITEMS; r++) rank[r] = key[r] != 0xffffffffu ? atomicAdd(&cnt[key[r]], 1u) : 0u;
    __syncthreads();
    if (threadIdx.x < 32) { // descending exclusive scan by one warp: units with a larger key come first
        uint32_t carry = 0;
        for (int b = 0; b < HS_KEYS / 32; b++) {
            const int c = HS_KEYS - 1 - (b * 32 + (int)threadIdx.x);
            const uint32_t v = cnt[c];
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += t;
            }
            cnt[c] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < HS_SORT_ITEMS; r++)
        if (key[r] != 0xffffffffu) perm[base + cnt[key[r]] + rank[r]] = base + r * HS_SORT_THREADS + threadIdx.x;
}

// The bits of the warp's 32 units into its region: unit k's words [w0_k, w0_k + nw_k) to region[off_k ..), byte-swapped.
// Eight units at a time, two words per lane and unit (a unit of up to 256 bytes), all sixteen loads in flight before
// the first store; longer units finish in a loop.
__device__ __forceinline__ void hs_stage_group(uint32_t *__restrict__ region, const uint32_t *__restrict__ words,
                                               uint32_t wlimit, uint32_t w0, uint32_t nw, uint32_t off, int lane)
{
#pragma unroll 1
    for (int k0 = 0; k0 < 32; k0 += 8) {
        uint32_t v[8][2], n_k[8], o_k[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            n_k[q] = __shfl_sync(0xffffffffu, nw, k0 + q);
            o_k[q] = __shfl_sync(0xffffffffu, off, k0 + q);
            const uint32_t w = __shfl_sync(0xffffffffu, w0, k0 + q) + lane;
            v[q][0] = (lane < n_k[q] && w < wlimit) ? __ldg(words + w) : 0u;
            v[q][1] = (lane + 32 < n_k[q] && w + 32 < wlimit) ? __ldg(words + w + 32) : 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (lane < n_k[q]) region[o_k[q] + lane] = __byte_perm(v[q][0], 0, 0x0123);
            if (lane + 32 < n_k[q]) region[o_k[q] + lane + 32] = __byte_perm(v[q][1], 0, 0x0123);
        }
#pragma unroll 1
        for (int q = 0; q < 8; q++) {
            const uint32_t n = __shfl_sync(0xffffffffu, nw, k0 + q);
            if (n <= 64) continue; // (warp-uniform)
            const uint32_t o = __shfl_sync(0xffffffffu, off, k0 + q), wb = __shfl_sync(0xffffffffu, w0, k0 + q);
            for (uint32_t j = 64 + lane; j < n; j += 32)
                region[o + j] = wb + j < wlimit ? __byte_perm(__ldg(words + wb + j), 0, 0x0123) : 0u;
        }
    }
}

// PASS 0: scalefactors + pairs (state word and pass-B key out); PASS 1: count1 + zero tail + vector count.
template <int PASS>
__global__ void __launch_bounds__(1024, 1)
k_huffman_sorted(const uint8_t *__restrict__ arena, uint64_t arena_bytes, const L3UnitDesc *__restrict__ units,
                 uint32_t u_lo, uint32_t nunits, const uint32_t *__restrict__ perm, uint32_t *__restrict__ group_counter,
                 uint32_t *__restrict__ state, uint16_t *__restrict__ keys_b,
                 const uint16_t *__restrict__ g_lut, uint32_t lut_len, uint32_t lut_bytes, uint32_t region_words,
                 const L3HuffInfo *__restrict__ g_info, const uint8_t *__restrict__ g_quad, int16_t *__restrict__ is_out,
                 uint8_t *__restrict__ sf_out, uint8_t *__restrict__ nzv_out, int zero_fill)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ K1Shared S; // (info, quad, c1 are used; the chunk bookkeeping is not)
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *region = reinterpret_cast<uint32_t *>(smem_raw + lut_bytes) + (size_t)warp * region_words;

    pdl_launch_dependents();
    if (PASS == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_lut);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t k = tid; k < (lut_len * 2 + 15) / 16; k += blockDim.x) dst[k] = __ldg(src + k);
    }
    for (int t = tid; t < 256; t += blockDim.x) {
        if (t < 32) S.info[t] = g_info->base[t] | ((uint32_t)g_info->root[t] << 16) | ((uint32_t)g_info->linbits[t] << 24);
        if (t < 64) S.quad[t] = g_quad[t];
        else if (t < 128) S.quad[t] = (uint8_t)((4u << 4) | (15u - ((t - 64u) >> 2)));
        const uint32_t sym = t >> 4, s4 = t & 15u;
        uint32_t val[4], k = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            val[c] = 0;
            if (sym & (8u >> c)) { val[c] = ((s4 >> (3 - k)) & 1u) ? 0xffffu : 1u; k++; }
        }
        S.c1[t] = make_uint2(val[0] | (val[1] << 16), val[2] | (val[3] << 16));
    }
    pdl_wait(); // descriptors, main-data arena, the sorted order: written by the kernels before this one
    __syncthreads();

    const uint32_t ngroups = (nunits + 31) >> 5;
    const uint32_t wlimit = (uint32_t)min((unsigned long long)(arena_bytes >> 2), 0xffffffffull);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(arena);
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(group_counter, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= ngroups) break;
        const uint32_t idx = g * 32 + lane;
        const bool mine = idx < nunits;
        const uint32_t rel = perm[mine ? idx : g * 32], u = u_lo + rel;
        const L3UnitDesc d = units[u];
        const bool valid = (d.flags & L3F_VALID) != 0;
        uint32_t stw = 0, used = 0;
        bool work = mine && valid; // this lane reads bits
        if (PASS == 1) {
            stw = state[rel];
            used = stw & 0x1fffu;
            work = work && used < d.p23len;
        }
        // Words this lane's reader may touch.  Pass A: scalefactors (<= 200 bits even when part2_3_length says less: a
        // damaged side info), one code word past the end (<= 47 bits), the 64-bit look-ahead of a peek.  Pass B: the
        // bits behind the pairs, one quadruple past the end (<= 10 bits), the look-ahead.
        const uint64_t bit0 = d.bit_off + used;
        const uint32_t w0 = (uint32_t)(bit0 >> 5);
        const uint32_t span = PASS == 0 ? max((uint32_t)d.p23len, 200u) + 47u : ((uint32_t)d.p23len - used) + 10u;
        const uint32_t nw = work ? (((uint32_t)bit0 & 31u) + span + 64u + 31u) / 32u + 1u : 0u;
        uint32_t off = nw; // exclusive prefix over the lanes = the unit's place in the warp's region
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, off, o);
            if (lane >= o) off += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, off, 31);
        off -= nw;
        const bool staged = total <= region_words; // (warp-uniform)
        if (staged) {
            __syncwarp(); // the previous group's readers are done with the region
            hs_stage_group(region, words, wlimit, w0, nw, off, lane);
            __syncwarp();
        }
        if (mine) {
            uint32_t *out32 = reinterpret_cast<uint32_t *>(is_out + (size_t)u * 576);
            auto pass = [&](auto &br) {
                if (PASS == 0) {
                    uint32_t sfw[10];
                    const uint32_t start = br.bitpos();
                    read_scalefactors(br, d, u, valid, units, arena, arena_bytes, sfw);
                    const uint32_t st = decode_pairs(br, start, d, valid, s_lut, S, out32);
                    uint2 *o = reinterpret_cast<uint2 *>(sf_out + (size_t)u * 40);
#pragma unroll
                    for (int k = 0; k < 5; k++) o[k] = make_uint2(sfw[2 * k], sfw[2 * k + 1]);
                    const uint32_t kb = hs_key_b(d, st);
                    state[rel] = st;
                    keys_b[rel] = (uint16_t)kb;
                } else {
                    int i = (int)((stw >> 13) & 0x3ffu);
                    if (work) i = decode_count1(br, br.bitpos() + ((uint32_t)d.p23len - used), i, (d.flags & L3F_C1TAB) ? 64u : 0u, S, out32);
                    // zero the rest of the last 16-byte vector (decode_count1 does it itself); the all-zero tail of the
                    // spectrum is not written: nzv_out[u] tells the consumer how many vectors hold data (zero_fill:
                    // staged pipeline, parity dumps)
                    const int nst = (i + 7) >> 3;
                    if (!work)
                        for (int k = i >> 1; k < nst * 4; k++) out32[k] = 0u;
                    nzv_out[u] = (uint8_t)nst;
                    if (zero_fill) {
                        uint4 *out = reinterpret_cast<uint4 *>(out32);
                        for (int k = nst; k < 72; k++) out[k] = make_uint4(0, 0, 0, 0);
                    }
                }
            };
            if (staged) {
                SmemReader br;
                br.init(region + off, work ? ((uint32_t)bit0 & 31u) : 0u, (uint64_t)w0 * 32u, nw * 32u);
                pass(br);
            } else {
                GlobalReader br;
                br.init(arena, arena_bytes, bit0);
                pass(br);
            }
        }
    }
}


// =====================================================================================================================
// The warp-per-unit variant (MP3B_K1_MODE=warp; measured, NOT the default -- numbers in DESIGN.md): what BASELINE.json's
// north_star sketches ("a warp-per-granule Huffman / count1 decode").  A code word's position depends on every code
// word before it, so the 32 lanes cannot share one unit's walk; what they can do is SPECULATE: lane l decodes the code
// word that would start at bit `pos + l`, for all 32 start positions at once (the look-ahead words are the same for
// all lanes: broadcast loads), and the true chain is then followed through the lanes' lengths by shuffles -- lane 0 is
// real, the lane at its end is real, and so on until the chain leaves the 32-bit window, the region or the unit.  A
// round yields window / (average code length) pairs (four to six) for one look-up per lane.  Scalefactors: a lane per
// band, positions by a warp prefix sum over the field widths.  Persistent CTAs; warps pull units from a counter.
struct WarpBits {
    const uint32_t *words; // arena as words
    uint32_t wlimit;
    // 64 bits at absolute bit position p (zeros beyond the arena)
    __device__ __forceinline__ void peek64(uint64_t p, uint32_t &hi, uint32_t &lo) const
    {
        const uint64_t w = p >> 5;
        const uint32_t a = w < wlimit ? __byte_perm(__ldg(words + w), 0, 0x0123) : 0u;
        const uint32_t b = w + 1 < wlimit ? __byte_perm(__ldg(words + w + 1), 0, 0x0123) : 0u;
        const uint32_t c = w + 2 < wlimit ? __byte_perm(__ldg(words + w + 2), 0, 0x0123) : 0u;
        const uint32_t sh = (uint32_t)p & 31u;
        hi = __funnelshift_l(b, a, sh);
        lo = __funnelshift_l(c, b, sh);
    }
    __device__ __forceinline__ uint32_t bits(uint64_t p, int n) const // n in 0..16
    {
        uint32_t hi, lo;
        peek64(p, hi, lo);
        return (hi >> 1) >> (31 - n);
    }
};

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t &total)
{
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

// Part 2 of one unit by the whole warp: lane b takes bands b and b + 32.  Same rules as read_scalefactors.
// Returns the number of bits part 2 occupies in this unit.
__device__ __forceinline__ uint32_t warp_scalefactors(const WarpBits &wb, const L3UnitDesc &d, uint32_t u, bool valid,
                                                      const L3UnitDesc *__restrict__ units, uint8_t *__restrict__ sf_out,
                                                      int lane)
{
    const int bt = d.flags & L3F_BT_MASK;
    const bool mixed = (d.flags & L3F_MIXED) != 0, lsf = (d.hdr & L3H_LSF) != 0;
    int w[2] = {0, 0};          // field width of bands lane, lane + 32 in this unit
    uint64_t reuse[2] = {0, 0}; // scfsi: absolute bit position of the band's value in granule 0
    int rw[2] = {0, 0};         // ... and its width there
    bool use_g0[2] = {false, false};
    bool ist = false;
    if (valid) {
        if (!lsf) {
            const unsigned long long SL1 = 0x4433322211130000ull, SL2 = 0x3232132132103210ull;
            const int s1 = (int)((SL1 >> (4 * d.sfc)) & 15), s2 = (int)((SL2 >> (4 * d.sfc)) & 15);
            if (bt == 2) {
                const int n1 = mixed ? 17 : 18, ntot = mixed ? 35 : 36;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int b = lane + 32 * r;
                    w[r] = b < n1 ? s1 : (b < ntot ? s2 : 0);
                }
            } else {
                const uint32_t scfsi = (d.pos & L3P_GR) ? (d.pos >> L3P_SCFSI_SHIFT) & 15u : 0u;
                int g0_s1 = 0, g0_s2 = 0, g0_n1 = 11, g0_ntot = 21;
                uint64_t g0_bit = 0;
                if (scfsi) {
                    const L3UnitDesc d0 = units[u - ((d.hdr & L3H_STEREO) ? 2u : 1u)];
                    g0_bit = d0.bit_off;
                    g0_s1 = (int)((SL1 >> (4 * d0.sfc)) & 15);
                    g0_s2 = (int)((SL2 >> (4 * d0.sfc)) & 15);
                    if ((d0.flags & L3F_BT_MASK) == 2) {
                        g0_n1 = (d0.flags & L3F_MIXED) ? 17 : 18;
                        g0_ntot = (d0.flags & L3F_MIXED) ? 35 : 36;
                    }
                }
                const int b = lane;
                if (b < 21) {
                    const int grp = b < 6 ? 0 : (b < 11 ? 1 : (b < 16 ? 2 : 3));
                    if ((scfsi >> grp) & 1u) {
                        use_g0[0] = true;
                        rw[0] = b < g0_n1 ? g0_s1 : (b < g0_ntot ? g0_s2 : 0);
                        reuse[0] = g0_bit + (b < g0_n1 ? (uint64_t)(b * g0_s1) : (uint64_t)(g0_n1 * g0_s1 + (b - g0_n1) * g0_s2));
                    } else
                        w[0] = b < 11 ? s1 : s2;
                }
            }
        } else {
            int sfc = d.sfc, sl0, sl1, sl2, sl3, tbl;
            ist = (d.hdr & L3H_IS) && (d.pos & L3P_CH);
            if (!ist) {
                if (sfc < 400) { sl0 = (sfc >> 4) / 5; sl1 = (sfc >> 4) % 5; sl2 = (sfc & 15) >> 2; sl3 = sfc & 3; tbl = 0; }
                else if (sfc < 500) { sfc -= 400; sl0 = (sfc >> 2) / 5; sl1 = (sfc >> 2) % 5; sl2 = sfc & 3; sl3 = 0; tbl = 1; }
                else { sfc -= 500; sl0 = sfc / 3; sl1 = sfc % 3; sl2 = 0; sl3 = 0; tbl = 2; }
            } else {
                sfc >>= 1;
                if (sfc < 180) { sl0 = sfc / 36; sl1 = (sfc % 36) / 6; sl2 = sfc % 6; sl3 = 0; tbl = 3; }
                else if (sfc < 244) { sfc -= 180; sl0 = (sfc & 63) >> 4; sl1 = (sfc & 15) >> 2; sl2 = sfc & 3; sl3 = 0; tbl = 4; }
                else { sfc -= 244; sl0 = sfc / 3; sl1 = sfc % 3; sl2 = 0; sl3 = 0; tbl = 5; }
            }
            const int lay = bt == 2 ? (mixed ? 2 : 1) : 0;
            const int e0 = c_lsf_nsfb[tbl][lay][0], e1 = e0 + c_lsf_nsfb[tbl][lay][1];
            const int e2 = e1 + c_lsf_nsfb[tbl][lay][2], e3 = e2 + c_lsf_nsfb[tbl][lay][3];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int b = lane + 32 * r;
                w[r] = b >= 40 ? 0 : (b < e0 ? sl0 : (b < e1 ? sl1 : (b < e2 ? sl2 : (b < e3 ? sl3 : 0))));
            }
        }
    }
    uint32_t tot0, tot1;
    const uint32_t o0 = warp_excl_scan((uint32_t)w[0], lane, tot0);
    const uint32_t o1 = tot0 + warp_excl_scan((uint32_t)w[1], lane, tot1);
    uint32_t v0 = use_g0[0] ? wb.bits(reuse[0], rw[0]) : wb.bits(d.bit_off + o0, w[0]);
    uint32_t v1 = wb.bits(d.bit_off + o1, w[1]);
    if (ist && w[0] && v0 == (1u << w[0]) - 1u) v0 |= 0x80u; // illegal intensity position
    if (ist && w[1] && v1 == (1u << w[1]) - 1u) v1 |= 0x80u;
    sf_out[(size_t)u * 40 + lane] = (uint8_t)v0;
    if (lane < 8) sf_out[(size_t)u * 40 + 32 + lane] = (uint8_t)v1;
    return tot0 + tot1;
}

constexpr int KW_THREADS = 256;
__global__ void __launch_bounds__(KW_THREADS)
k_huffman_warp(const uint8_t *__restrict__ arena, uint64_t arena_bytes, const L3UnitDesc *__restrict__ units, uint32_t u_lo,
               uint32_t nunits, uint32_t *__restrict__ unit_counter, const uint16_t *__restrict__ g_lut, uint32_t lut_len,
               const L3HuffInfo *__restrict__ g_info, const uint8_t *__restrict__ g_quad, int16_t *__restrict__ is_out,
               uint8_t *__restrict__ sf_out, uint8_t *__restrict__ nzv_out, int zero_fill)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_info[32];
    __shared__ uint8_t s_quad[128];
    __shared__ uint2 s_c1[256];
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    pdl_launch_dependents();
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_lut);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (uint32_t k = tid; k < (lut_len * 2 + 15) / 16; k += KW_THREADS) dst[k] = __ldg(src + k);
    }
    if (tid < 32) s_info[tid] = g_info->base[tid] | ((uint32_t)g_info->root[tid] << 16) | ((uint32_t)g_info->linbits[tid] << 24);
    if (tid < 64) s_quad[tid] = g_quad[tid];
    else if (tid < 128) s_quad[tid] = (uint8_t)((4u << 4) | (15u - ((tid - 64u) >> 2)));
    {
        const uint32_t sym = tid >> 4, s4 = tid & 15u;
        uint32_t val[4], k = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            val[c] = 0;
            if (sym & (8u >> c)) { val[c] = ((s4 >> (3 - k)) & 1u) ? 0xffffu : 1u; k++; }
        }
        s_c1[tid] = make_uint2(val[0] | (val[1] << 16), val[2] | (val[3] << 16));
    }
    pdl_wait();
    __syncthreads();

    WarpBits wb;
    wb.words = reinterpret_cast<const uint32_t *>(arena);
    wb.wlimit = (uint32_t)min((unsigned long long)(arena_bytes >> 2), 0xffffffffull);
    for (;;) {
        uint32_t ui = 0;
        if (lane == 0) ui = atomicAdd(unit_counter, 1u);
        ui = __shfl_sync(0xffffffffu, ui, 0);
        if (ui >= nunits) break;
        const uint32_t u = u_lo + ui;
        const L3UnitDesc d = units[u];
        const bool valid = (d.flags & L3F_VALID) != 0;
        uint32_t *out32 = reinterpret_cast<uint32_t *>(is_out + (size_t)u * 576);
        const uint64_t start = d.bit_off, limit = start + d.p23len;
        uint64_t pos = start + warp_scalefactors(wb, d, u, valid, units, sf_out, lane);
        // ---- big_values pairs
        const int bv2 = valid ? d.big_values * 2 : 0;
        int i = 0;
        while (i < bv2) {
            const int reg = i < d.r1 ? 0 : (i < d.r2 ? 1 : 2);
            const int nextb = min(bv2, reg == 0 ? (int)d.r1 : (reg == 1 ? (int)d.r2 : bv2));
            const uint32_t info = s_info[d.tsel[reg]];
            const uint32_t base = info & 0xffffu;
            const int root = (int)((info >> 16) & 0xffu), lin = (int)(info >> 24);
            // this lane's guess: the pair that starts at pos + lane
            uint32_t hi, lo;
            wb.peek64(pos + lane, hi, lo);
            const uint32_t e1 = s_lut[base + __funnelshift_l(hi, 0u, root)];
            const bool lng = (e1 & 0x8000u) != 0;
            const uint32_t w2 = (e1 >> 11) & 15u;
            const uint32_t e = lng ? s_lut[base + (e1 & 0x7ffu) + __funnelshift_l(hi << root, 0u, w2)] : e1;
            const int len = (int)((e >> 8) & 15u) + (lng ? root : 0);
            int x = (e >> 4) & 15, y = e & 15;
            uint32_t rest = __funnelshift_l(lo, hi, len);
            const int lx = x == 15 ? lin : 0;
            x += (int)__funnelshift_l(rest, 0u, lx);
            rest <<= lx;
            const int sx = x != 0;
            { const int m = (int)rest >> 31; x = (x ^ m) - m; }
            rest <<= sx;
            const int ly = y == 15 ? lin : 0;
            y += (int)__funnelshift_l(rest, 0u, ly);
            rest <<= ly;
            const int sy = y != 0;
            { const int m = (int)rest >> 31; y = (y ^ m) - m; }
            const int n = len + lx + sx + ly + sy;
            const uint32_t word = __byte_perm((uint32_t)x, (uint32_t)y, 0x5410);
            // follow the true chain through the lanes
            const int maxk = (nextb - i) >> 1;
            int cur = 0, k = 0, myk = -1;
            bool dead = false, empty = false;
            while (cur < 32 && k < maxk) {
                if (pos + cur >= limit) { dead = true; break; } // out of bits: this and all later pairs are (0, 0)
                const int n_cur = __shfl_sync(0xffffffffu, n, cur);
                if (n_cur == 0) { empty = true; break; }        // the empty book: the rest of the region is (0, 0)
                if (lane == cur) myk = k;
                cur += n_cur;
                k++;
            }
            if (myk >= 0) out32[(i >> 1) + myk] = word;
            pos += (uint32_t)cur;
            i += 2 * k;
            if (dead || empty) {
                const int stop = dead ? bv2 : nextb;
                for (int q = (i >> 1) + lane; q < (stop >> 1); q += 32) out32[q] = 0u;
                i = stop;
            }
        }
        // ---- count1 quadruples
        if (valid) {
            const uint32_t qoff = (d.flags & L3F_C1TAB) ? 64u : 0u;
            bool stop = false;
            while (!stop && i <= 572 && pos < limit) {
                uint32_t hi, lo;
                wb.peek64(pos + lane, hi, lo);
                const uint32_t e = s_quad[qoff + (hi >> 26)];
                const uint32_t sym = e & 15u, len = e >> 4;
                const uint32_t s4 = (hi << len) >> 28;
                const int n = (int)(len + __popc(sym));
                const uint2 vw = s_c1[(sym << 4) | s4];
                int cur = 0, k = 0, myk = -1;
                while (cur < 32 && i + 4 * k <= 572 && pos + cur < limit) {
                    const int n_cur = __shfl_sync(0xffffffffu, n, cur);
                    if (pos + cur + n_cur > limit) { stop = true; break; } // overran part2_3_length: discard, and end
                    if (lane == cur) myk = k;
                    cur += n_cur;
                    k++;
                }
                if (myk >= 0) {
                    out32[(i >> 1) + 2 * myk] = vw.x;
                    out32[(i >> 1) + 2 * myk + 1] = vw.y;
                }
                pos += (uint32_t)cur;
                i += 4 * k;
            }
        }
        const int nst = (i + 7) >> 3;
        for (int q = (i >> 1) + lane; q < nst * 4; q += 32) out32[q] = 0u;
        if (lane == 0) nzv_out[u] = (uint8_t)nst;
        if (zero_fill) {
            uint4 *out = reinterpret_cast<uint4 *>(out32);
            for (int q = nst + lane; q < 72; q += 32) out[q] = make_uint4(0, 0, 0, 0);
        }
    }
}

} // namespace

void l3_launch_huffman_range(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                             uint32_t nunits, uint32_t avg_unit_bytes, const L3DevTables &T, int16_t *is_out,
                             uint8_t *sf_out, uint8_t *nzv_out, int zero_fill, cudaStream_t st, bool pdl)
{
    if (!nunits) return;
    // stage size: the average chunk plus slack, bounded so that at least one CTA fits an SM; chunks
    // that do not fit (high bitrates, VBR peaks) take the global-memory reader
    static std::atomic<unsigned long long> configured{0};
    if (l3_device_needs_setup(configured)) {
        cudaFuncSetAttribute(k_huffman<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_MAX_DYN_SMEM);
        cudaFuncSetAttribute(k_huffman<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_MAX_DYN_SMEM);
        l3_device_setup_done(configured);
    }
    const uint64_t lut_bytes = ((uint64_t)T.huff_lut_len * sizeof(uint16_t) + 15) & ~15ull;
    // Chunk length and stage size: as many CTAs per SM as still leave every warp a few groups to pull.
    // The stage holds the chunk's main data (average bytes per unit plus slack); a chunk whose range
    // does not fit (VBR peaks) takes the global-memory reader.
    // tuning overrides, read once (function-local statics: initialisation is thread-safe, contexts on several
    // host threads may get here together)
    static const long f_stage = [] { const char *e = getenv("MP3B_K1_STAGE"); return e ? atol(e) : 0l; }();
    static const long f_chunk = [] { const char *e = getenv("MP3B_K1_CHUNK"); return e ? atol(e) : 0l; }();
    static const long f_threads = [] { const char *e = getenv("MP3B_K1_THREADS"); return e ? atol(e) : 0l; }();
    const int nt = f_threads == 512 ? 512 : 256;
    const int max_chunk = 2 * nt, max_ctas = 1024 / nt;
    const uint64_t per_unit = (uint64_t)avg_unit_bytes + avg_unit_bytes / 24 + 1;
    const uint64_t fixed = lut_bytes + (nt == 512 ? 10240 : 6656) + 1024; // LUT, static shared memory, per-CTA reservation
    uint64_t want = 0;
    uint32_t chunk = 0;
    // prefer 4 CTAs per SM if they get >= 8 groups each (measured: 0.88 ms against 0.91 with 3 x 480 units;
    // the code tables through L1 instead of a shared copy, to make room for a fifth CTA: 1.05 - 1.16 ms)
    for (int k = max_ctas; k >= 1 && chunk < (uint32_t)nt; k--) {
        // never more dynamic shared memory than the kernel was given (200 KB): large units -- 160 kbit/s mono at
        // 8 kHz is 1,440 bytes per unit -- would otherwise ask for more at one CTA per SM and fail to launch
        const uint64_t budget = std::min<uint64_t>(227ull * 1024 / k - fixed, K1_MAX_DYN_SMEM - lut_bytes) & ~15ull;
        uint64_t c = (budget - 1280) / per_unit / 32 * 32;
        c = c > (uint64_t)max_chunk ? (uint64_t)max_chunk : c;
        if (c >= 32) { chunk = (uint32_t)c; want = (c * per_unit + 1280 + 15) & ~15ull; }
    }
    if (!chunk) { chunk = 32; want = 32 * 1024; }
    if (f_chunk > 0) { chunk = (uint32_t)(f_chunk > max_chunk ? max_chunk : f_chunk) / 32 * 32; want = (chunk * per_unit + 1280 + 15) & ~15ull; }
    if (f_stage > 0) want = ((uint64_t)f_stage + 15) & ~15ull;
    want = std::min<uint64_t>(want, (K1_MAX_DYN_SMEM - lut_bytes) & ~15ull); // (overrides included; a chunk that does not fit reads global memory)
    const size_t smem = (size_t)(want + lut_bytes);
    if (nt == 512)
        l3_launch_k(k_huffman<512>, dim3((nunits + chunk - 1) / chunk), dim3(512), smem, st, pdl, arena, arena_bytes, units, u_lo,
                    nunits, T.huff_lut, T.huff_lut_len, (uint32_t)want, chunk, T.huff, T.quad_a, is_out, sf_out, nzv_out, zero_fill);
    else
        l3_launch_k(k_huffman<256>, dim3((nunits + chunk - 1) / chunk), dim3(256), smem, st, pdl, arena, arena_bytes, units, u_lo,
                    nunits, T.huff_lut, T.huff_lut_len, (uint32_t)want, chunk, T.huff, T.quad_a, is_out, sf_out, nzv_out, zero_fill);
}

size_t l3_huff_sort_ctl_bytes(void) { return sizeof(uint32_t) * HS_CTL_WORDS; }

// The sorted variant: counting sort of the wave's units, then one persistent CTA per SM (see k_huffman_sorted).
void l3_launch_huffman_sorted(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                              uint32_t nunits, uint32_t avg_unit_bytes, const L3DevTables &T, const L3HuffSort &scr,
                              int16_t *is_out, uint8_t *sf_out, uint8_t *nzv_out, int zero_fill, cudaStream_t st, bool pdl)
{
    if (!nunits) return;
    static std::atomic<unsigned long long> configured{0};
    if (l3_device_needs_setup(configured)) {
        cudaFuncSetAttribute(k_huffman_sorted<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_huffman_sorted<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        l3_device_setup_done(configured);
    }
    // tuning / test overrides (read per launch: the tests switch them inside one process)
    const char *e_head = getenv("MP3B_K1_HEADROOM"), *e_warps = getenv("MP3B_K1_WARPS"), *e_region = getenv("MP3B_K1_REGION");
    const double f_head = e_head ? atof(e_head) : 1.5;
    const long f_warps = e_warps ? atol(e_warps) : 0l, f_region = e_region ? atol(e_region) : 0l;
    const uint32_t lut_bytes = (uint32_t)(((uint64_t)T.huff_lut_len * sizeof(uint16_t) + 15) & ~15ull);
    const uint32_t avail = 220u * 1024u - lut_bytes; // (6.6 KB of static shared memory and the per-CTA reservation stay free)
    // a warp's part of shared memory: its 32 units' main data (+ 24 bytes of look-ahead and alignment each), with
    // headroom for the groups at the heavy end of the sorted order; as many warps as then fit, at most 32
    const uint32_t want = (uint32_t)(32.0 * (avg_unit_bytes + 24) * f_head);
    uint32_t warps = std::max(1u, std::min(32u, avail / std::max(want, 256u)));
    if (f_warps > 0) warps = (uint32_t)std::min(32l, f_warps);
    uint32_t region_words = (avail / warps / 16u * 16u) / 4u;
    if (f_region > 0) region_words = std::min<uint32_t>(region_words, (uint32_t)f_region); // (small: every group reads global memory)
    const uint32_t ngroups = (nunits + 31) / 32;
    const uint32_t sms = (uint32_t)std::max(1, scr.sm_count);
    // a small batch spreads over the SMs (fewer warps per CTA), a large one runs `warps` warps on every SM
    const uint32_t w_a = std::max(1u, std::min(warps, (ngroups + sms - 1) / sms));
    const uint32_t ctas_a = std::min(sms, (ngroups + w_a - 1) / w_a);
    // pass B stages a few words per unit and needs no code tables: always 32 warps, 1 KB each
    const uint32_t w_b = std::max(1u, std::min(32u, (ngroups + sms - 1) / sms));
    const uint32_t ctas_b = std::min(sms, (ngroups + w_b - 1) / w_b);
    const uint32_t region_b = f_region > 0 ? std::min<uint32_t>(1024u, (uint32_t)f_region) : 1024u; // words: 32 units x 128 bytes
    const uint32_t sort_ctas = (nunits + HS_BLOCK - 1) / HS_BLOCK;
    cudaMemsetAsync(scr.ctl, 0, sizeof(uint32_t) * HS_CTL_WORDS, st);
    l3_launch_k(k_hsort_local<0>, dim3(sort_ctas), dim3(HS_SORT_THREADS), 0, st, false, units, u_lo, nunits, scr.keys, scr.perm);
    l3_launch_k(k_huffman_sorted<0>, dim3(ctas_a), dim3(32 * w_a), (size_t)lut_bytes + (size_t)warps * region_words * 4, st,
                pdl, arena, arena_bytes, units, u_lo, nunits, scr.perm, scr.ctl, scr.state, scr.keys, T.huff_lut,
                T.huff_lut_len, lut_bytes, region_words, T.huff, T.quad_a, is_out, sf_out, nzv_out, zero_fill);
    l3_launch_k(k_hsort_local<1>, dim3(sort_ctas), dim3(HS_SORT_THREADS), 0, st, pdl, units, u_lo, nunits, scr.keys, scr.perm);
    l3_launch_k(k_huffman_sorted<1>, dim3(ctas_b), dim3(32 * w_b), (size_t)32 * region_b * 4, st, pdl, arena, arena_bytes,
                units, u_lo, nunits, scr.perm, scr.ctl + 1, scr.state, scr.keys, T.huff_lut, T.huff_lut_len, 0u,
                region_b, T.huff, T.quad_a, is_out, sf_out, nzv_out, zero_fill);
}

void l3_launch_huffman_warp(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                            uint32_t nunits, const L3DevTables &T, const L3HuffSort &scr, int16_t *is_out, uint8_t *sf_out,
                            uint8_t *nzv_out, int zero_fill, cudaStream_t st, bool pdl)
{
    if (!nunits) return;
    const uint32_t lut_bytes = (uint32_t)(((uint64_t)T.huff_lut_len * sizeof(uint16_t) + 15) & ~15ull);
    const uint32_t sms = (uint32_t)std::max(1, scr.sm_count);
    const uint32_t ctas = std::min(sms * 8u, (nunits + 7u) / 8u); // persistent: 8 CTAs of 8 warps per SM
    cudaMemsetAsync(scr.ctl, 0, sizeof(uint32_t) * HS_CTL_WORDS, st);
    l3_launch_k(k_huffman_warp, dim3(ctas), dim3(KW_THREADS), (size_t)lut_bytes, st, false, arena, arena_bytes, units, u_lo, nunits,
                scr.ctl, T.huff_lut, T.huff_lut_len, T.huff, T.quad_a, is_out, sf_out, nzv_out, zero_fill);
    (void)pdl;
}
