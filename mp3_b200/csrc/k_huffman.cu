// k_huffman.cu -- K1: scalefactor decode (a4) and Huffman big_values / count1 decode (a5).
//
// One thread per unit (granule-channel).  Decoding a unit is a serial bit walk (every codeword's
// position depends on the previous one), so the parallelism is across units: 1.57 M of them in
// the 1024-stream workload.  What the GPU mapping has to get right is
//   * table lookups: all 15 code books as one multi-level LUT (8-bit root, 6-bit sub levels,
//     8.8 KB) staged in shared memory, so a codeword costs one or two LDS, not a tree walk;
//   * convergence: one flat loop over pairs for all three regions (the table is selected by
//     data, not by control flow), then one loop over count1 quadruples, so the lanes of a warp
//     stay on the same instruction even though their tables and lengths differ;
//   * stores: spectral values are packed 8 per 16-byte store (a shift register of four words,
//     no dynamically indexed local array).
// Output is bit-exact by construction (integer work only) and is compared word for word with
// oracle/l3_oracle.c::huffman_decode / read_scalefactors in tests/test_gpu_stages.py.
// No reference code exists for this stage (/root/reference/README.md:1-84).
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

// Bit reader over the big-endian main-data arena.  Two consecutive 32-bit words are cached in
// registers; a peek is one funnel shift, a skip is one add plus a (rarely taken) refill when the
// position crosses into the next word.  Positions are 32-bit and relative to the word the unit
// starts in (the arena itself can exceed 2^32 bits).
struct BitReader {
    const uint32_t *words; // arena + first word of this unit
    uint32_t wlimit;       // words readable from `words`
    uint32_t w0, w1, w2;   // words[wi], words[wi + 1] in bit order; w2 = raw words[wi + 2], loaded one
                           // word ahead so that its latency is off the decode chain
    uint32_t wi;
    uint32_t pos, start;   // bit position relative to words[0]

    __device__ __forceinline__ uint32_t load_raw(uint32_t w) const { return w < wlimit ? __ldg(words + w) : 0u; }
    __device__ __forceinline__ uint32_t load(uint32_t w) const { return __byte_perm(load_raw(w), 0, 0x0123); }
    __device__ __forceinline__ void init(const uint8_t *arena, uint64_t arena_bytes, uint64_t bit_off)
    {
        const uint64_t w = bit_off >> 5, total = arena_bytes >> 2;
        words = reinterpret_cast<const uint32_t *>(arena) + w;
        const uint64_t left = total > w ? total - w : 0;
        wlimit = left > 0x7fffffffull ? 0x7fffffffu : (uint32_t)left;
        wi = 0;
        w0 = load(0);
        w1 = load(1);
        w2 = load_raw(2);
        pos = start = (uint32_t)(bit_off & 31);
    }
    __device__ __forceinline__ uint32_t consumed() const { return pos - start; }
    // the next 32 bits, left aligned
    __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_l(w1, w0, pos & 31); }
    // k in 1..32
    __device__ __forceinline__ uint32_t peek(int k) const { return peek32() >> (32 - k); }
    __device__ __forceinline__ void skip(int k) // k in 0..32
    {
        pos += (uint32_t)k;
        if ((pos >> 5) != wi) {
            wi++;
            w0 = w1;
            w1 = __byte_perm(w2, 0, 0x0123);
            w2 = load_raw(wi + 2);
        }
    }
    __device__ __forceinline__ uint32_t get(int k) // k in 0..16
    {
        if (k == 0) return 0;
        uint32_t v = peek(k);
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

constexpr int K1_THREADS = 256;

__global__ void __launch_bounds__(K1_THREADS)
k_huffman(const uint8_t *__restrict__ arena, uint64_t arena_bytes, const L3UnitDesc *__restrict__ units,
          uint32_t u_lo, uint32_t nunits, const uint16_t *__restrict__ g_lut, uint32_t lut_len,
          const L3HuffInfo *__restrict__ g_info, const uint8_t *__restrict__ g_quad,
          int16_t *__restrict__ is_out, uint8_t *__restrict__ sf_out, uint8_t *__restrict__ nzv_out, int zero_fill)
{
    extern __shared__ uint16_t s_lut[];
    __shared__ uint32_t s_info[32]; // base | root << 16 | linbits << 24
    __shared__ uint8_t s_quad[64];
    for (uint32_t i = threadIdx.x; i < lut_len; i += K1_THREADS) s_lut[i] = g_lut[i];
    if (threadIdx.x < 32)
        s_info[threadIdx.x] = g_info->base[threadIdx.x] | ((uint32_t)g_info->root[threadIdx.x] << 16) |
                              ((uint32_t)g_info->linbits[threadIdx.x] << 24);
    if (threadIdx.x < 64) s_quad[threadIdx.x] = g_quad[threadIdx.x];

    // Units of a CTA are regrouped by big_values so that the 32 lanes of a warp run loops of similar
    // length (lanes stay converged inside an iteration; what is left to lose is the loop count).
    __shared__ uint32_t s_hist[160];
    __shared__ uint16_t s_perm[K1_THREADS];
    const uint32_t cta_first = blockIdx.x * K1_THREADS;
    for (int i = threadIdx.x; i < 160; i += K1_THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t bucket = 159; // units past the end sort last
    if (cta_first + threadIdx.x < nunits) {
        const L3UnitDesc *dp = units + u_lo + cta_first + threadIdx.x;
        bucket = (dp->flags & L3F_VALID) ? (uint32_t)min((int)dp->big_values, 288) >> 1 : 0u;
    }
    atomicAdd(&s_hist[bucket], 1u);
    __syncthreads();
    if (threadIdx.x < 32) { // exclusive scan of 160 buckets: 5 per lane
        uint32_t v[5], sum = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) { v[k] = s_hist[threadIdx.x * 5 + k]; sum += v[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += t;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < 5; k++) { s_hist[threadIdx.x * 5 + k] = run; run += v[k]; }
    }
    __syncthreads();
    s_perm[atomicAdd(&s_hist[bucket], 1u)] = (uint16_t)threadIdx.x;
    __syncthreads();

    uint32_t u = cta_first + s_perm[threadIdx.x];
    if (u >= nunits) return;
    u += u_lo;
    const L3UnitDesc d = units[u];
    const bool valid = (d.flags & L3F_VALID) != 0;
    const int bt = d.flags & L3F_BT_MASK;
    const bool mixed = (d.flags & L3F_MIXED) != 0;
    const bool lsf = (d.hdr & L3H_LSF) != 0;
    const uint32_t p23 = d.p23len;

    BitReader br;
    br.init(arena, arena_bytes, d.bit_off);

    // ---------------------------------------------------------------- part 2: scalefactors (a4)
    uint32_t sfw[10];
#pragma unroll
    for (int i = 0; i < 10; i++) sfw[i] = 0;
    if (valid) {
        if (!lsf) {
            // slen1 / slen2 nibble tables indexed by scalefac_compress
            const unsigned long long SL1 = 0x4433322211130000ull, SL2 = 0x3232132132103210ull;
            const int s1 = (int)((SL1 >> (4 * d.sfc)) & 15), s2 = (int)((SL2 >> (4 * d.sfc)) & 15);
            if (bt == 2) {
                const int n1 = mixed ? 17 : 18, ntot = mixed ? 35 : 36;
#pragma unroll
                for (int b = 0; b < 36; b++) {
                    int w = b < n1 ? s1 : (b < ntot ? s2 : 0);
                    sfw[b >> 2] |= br.get(w) << (8 * (b & 3));
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
#pragma unroll
                for (int b = 0; b < 21; b++) {
                    const int grp = b < 6 ? 0 : (b < 11 ? 1 : (b < 16 ? 2 : 3));
                    uint32_t v;
                    if ((scfsi >> grp) & 1u) {
                        // b-th transmitted scalefactor of granule 0
                        int w0 = b < g0_n1 ? g0_s1 : (b < g0_ntot ? g0_s2 : 0);
                        uint64_t off = b < g0_n1 ? (uint64_t)(b * g0_s1)
                                                 : (uint64_t)(g0_n1 * g0_s1 + (b - g0_n1) * g0_s2);
                        v = bits_at(arena, arena_bytes, g0_bit + off, w0);
                    } else
                        v = br.get(b < 11 ? s1 : s2);
                    sfw[b >> 2] |= v << (8 * (b & 3));
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
#pragma unroll
            for (int b = 0; b < 40; b++) {
                int w = b < e0 ? sl0 : (b < e1 ? sl1 : (b < e2 ? sl2 : (b < e3 ? sl3 : 0)));
                uint32_t v = br.get(w);
                if (ist && w && v == (1u << w) - 1u) v |= 0x80u; // illegal intensity position
                sfw[b >> 2] |= v << (8 * (b & 3));
            }
        }
    }
    {
        uint2 *o = reinterpret_cast<uint2 *>(sf_out + (size_t)u * 40);
#pragma unroll
        for (int i = 0; i < 5; i++) o[i] = make_uint2(sfw[2 * i], sfw[2 * i + 1]);
    }

    // ---------------------------------------------------------------- part 3: Huffman (a5)
    uint4 *out = reinterpret_cast<uint4 *>(is_out + (size_t)u * 576);
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0; // shift register of packed (x, y) pairs
    int npairs = 0, nst = 0;
    auto push = [&](int x, int y) {
        q0 = q1; q1 = q2; q2 = q3;
        q3 = ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16);
        if (++npairs == 4) {
            out[nst++] = make_uint4(q0, q1, q2, q3);
            npairs = 0;
        }
    };

    const int bv2 = valid ? d.big_values * 2 : 0;
    const int r1 = d.r1, r2 = d.r2;
    const uint32_t i0 = s_info[d.tsel[0]], i1 = s_info[d.tsel[1]], i2 = s_info[d.tsel[2]];
    bool dead = !valid;
    for (int i = 0; i < bv2; i += 2) {
        // One pair per iteration, written without branches: lanes differ in table, code length and
        // escapes, but all of that is data (selects and predicated loads), not control flow.
        const uint32_t info = i < r1 ? i0 : (i < r2 ? i1 : i2);
        const int root = (info >> 16) & 0xff; // 0 = the empty book: (0, 0), no bits
        dead = dead || (root != 0 && br.consumed() >= p23);
        const bool act = root != 0 && !dead;
        const uint32_t base = info & 0xffffu;
        const int lin = (int)(info >> 24);
        // 64 bits of look-ahead: code (<= 19) + escapes and signs (<= 28)
        const uint32_t sh = br.pos & 31;
        const uint32_t hi = __funnelshift_l(br.w1, br.w0, sh);
        const uint32_t lo = __funnelshift_l(__byte_perm(br.w2, 0, 0x0123), br.w1, sh);
        uint32_t e = s_lut[base + ((hi >> 1) >> (31 - root))];
        int len = (e >> 8) & 15;
        if (e & 0x8000u) { // longer than the root: one second-level lookup covers the rest
            const int w2 = (e >> 11) & 15;
            e = s_lut[base + (e & 0x7ffu) + ((hi << root) >> (32 - w2))];
            len = root + ((e >> 8) & 15);
        }
        int x = (e >> 4) & 15, y = e & 15;
        const uint32_t rest = __funnelshift_l(lo, hi, len);
        const int lx = x == 15 ? lin : 0;
        x += (int)((rest >> 1) >> (31 - lx));
        int n = lx;
        const int sx = x != 0;
        if (((rest << n) >> 31) & sx) x = -x;
        n += sx;
        const int ly = y == 15 ? lin : 0;
        y += (int)(((rest << n) >> 1) >> (31 - ly));
        n += ly;
        const int sy = y != 0;
        if (((rest << n) >> 31) & sy) y = -y;
        n += sy;
        if (!act) { x = 0; y = 0; len = 0; n = 0; }
        br.skip(len);
        br.skip(n);
        push(x, y);
    }
    // count1 quadruples
    {
        const bool tab_b = (d.flags & L3F_C1TAB) != 0;
        int i = bv2;
        while (!dead && i <= 572 && br.consumed() < p23) {
            const uint32_t bits = br.peek32(); // code (<= 6 bits) + up to 4 sign bits
            int sym, len;
            if (tab_b) { sym = 15 - (int)(bits >> 28); len = 4; }
            else { uint32_t e = s_quad[bits >> 26]; sym = e & 15; len = e >> 4; }
            const uint32_t s4 = (bits << len) >> 28;
            int k = 0, v = 0, w = 0, x = 0, y = 0;
            if (sym & 8) { v = ((s4 >> (3 - k)) & 1) ? -1 : 1; k++; }
            if (sym & 4) { w = ((s4 >> (3 - k)) & 1) ? -1 : 1; k++; }
            if (sym & 2) { x = ((s4 >> (3 - k)) & 1) ? -1 : 1; k++; }
            if (sym & 1) { y = ((s4 >> (3 - k)) & 1) ? -1 : 1; k++; }
            br.skip(len + k);
            if (br.consumed() > p23) break; // overran part2_3_length: discard this quadruple
            push(v, w);
            push(x, y);
            i += 4;
        }
    }
    // flush the partial vector (zero padded).  The all-zero tail of the spectrum is not written:
    // nzv_out[u] tells the consumer how many 16-byte vectors (8 lines each) hold data, and it treats
    // the rest as zero.  zero_fill (staged pipeline, parity dumps) writes the tail anyway.
    if (npairs) {
        while (npairs) push(0, 0);
    }
    nzv_out[u] = (uint8_t)nst;
    if (zero_fill)
        for (int k = nst; k < 72; k++) out[k] = make_uint4(0, 0, 0, 0);
}

} // namespace

void l3_launch_huffman_range(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                             uint32_t nunits, const L3DevTables &T, int16_t *is_out, uint8_t *sf_out,
                             uint8_t *nzv_out, int zero_fill, cudaStream_t st)
{
    if (!nunits) return;
    size_t smem = (size_t)T.huff_lut_len * sizeof(uint16_t);
    k_huffman<<<(nunits + K1_THREADS - 1) / K1_THREADS, K1_THREADS, smem, st>>>(
        arena, arena_bytes, units, u_lo, nunits, T.huff_lut, T.huff_lut_len, T.huff, T.quad_a, is_out, sf_out,
        nzv_out, zero_fill);
}
