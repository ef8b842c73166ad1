// k_layer2.cu -- Layer II (MP2) frames -> subband samples (11172-3 2.4.1.6 / 2.4.3.3; SURVEY.md 8(f)
// rank 4).  Layer II has no Huffman coding, no bit reservoir and no hybrid filter bank: a frame is bit
// allocation, scfsi, scalefactors and 12 x 3 samples per subband, and its 36 time slots of 32 subband
// samples go straight into the same polyphase synthesis as Layer III (k_synth.cu) -- a frame is two
// "granules" of 18 slots, so the unit / PCM layout of the batch is unchanged.
//
// Mapping: one warp per frame, lane = subband.  Every field's position is a prefix sum of per-subband
// bit counts, so the three header sections and the sample section are located with warp scans and each
// lane reads its own bits; the frame's bytes are staged in shared memory first (coalesced).
// Checked against oracle/l3_oracle.c::l2_decode_frame (itself pinned by FFmpeg's mp2float).
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include <math.h>

#include "iso_tables_l2.h"
#include "kernels.h"

namespace {

constexpr int L2_WARPS = 4;
constexpr int L2_MAX_FRAME = 1732; // 384 kbit/s at 32 kHz: 1728 + padding

__constant__ uint8_t c_l2_rows[8][16];
__constant__ uint8_t c_l2_row_of_sb[5][30];
__constant__ int32_t c_l2_steps[17];
__constant__ int8_t c_l2_bits[17];

// Scalefactor 2^(1 - i / 3) for index i < 63, 0 for the forbidden index 63: a power of two times one of three
// cube-root factors, which is the correctly rounded table value (the index differs from lane to lane, and a
// lane-varying index into __constant__ memory is replayed once per distinct address).
__device__ __forceinline__ float l2_scf_of(uint32_t i)
{
    const uint32_t q = (i * 171u) >> 9, r = i - 3u * q; // i / 3, i % 3 for i < 64
    const float m = r == 0 ? 1.f : (r == 1 ? 0.79370052598409979f : 0.62996052494743658f);
    return i >= 63u ? 0.f : __int_as_float((128u - q) << 23) * m;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int *total)
{
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    *total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

// n (0..16) bits at bit position `pos` of the staged frame (bytes past the frame read as zero)
// The frame is staged in shared memory as big-endian 32-bit words, so a field is two word loads and one
// funnel shift.  Bits past the end of the frame read as zero (a damaged allocation can ask for far more
// bits than the frame holds): the staged frame is followed by >= 8 zero bytes and the position is clamped
// to its end.
__device__ __forceinline__ uint32_t l2_bits(const uint32_t *f, uint32_t pos, int n, uint32_t lim)
{
    pos = min(pos, lim);
    const uint32_t i = pos >> 5;
    const uint32_t w = __funnelshift_l(f[i + 1], f[i], pos & 31u);
    return n ? w >> (32 - n) : 0u;
}

// Stage `flen` bytes at src (any alignment) as big-endian words, zero-padded by at least three words.
__device__ __forceinline__ void l2_stage_frame(uint32_t *fw, const uint8_t *__restrict__ src, int flen, int lane)
{
    const int nw = (flen + 3) / 4 + 3;
    for (int i0 = 0; i0 < nw; i0 += 32) { // warp-uniform trip count
        const int i = i0 + lane;
        if (i < nw) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int b = 4 * i + k;
                w = (w << 8) | (b < flen ? (uint32_t)src[b] : 0u);
            }
            fw[i] = w;
        }
    }
}

__global__ void __launch_bounds__(L2_WARPS * 32)
k_layer2(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams, const L3FrameRec *__restrict__ frames,
         uint32_t nframes, float *__restrict__ sb_out)
{
    __shared__ uint32_t s_frame[L2_WARPS][L2_MAX_FRAME / 4 + 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * L2_WARPS + warp;
    if (f >= nframes) return;
    const L3FrameRec fr = frames[f];
    L3Hdr h;
    if (!l3_parse_hdr(fr.hdr, &h) || h.layer != 2) return;
    const L3StreamRec sr = streams[fr.stream];
    const uint8_t *src = raw + sr.raw_off + fr.rel_off;
    uint32_t *fb = s_frame[warp];
    const int flen = min(h.frame_len, L2_MAX_FRAME);
    l2_stage_frame(fb, src, flen, lane);
    __syncwarp();
    const uint32_t lim = (uint32_t)flen * 8u;

    const int nch = h.nch;
    const int tbl = l2_select_table(h.lsf, l3_sr_hz(h.sr_row), h.kbps, nch);
    const int sblimit = l2_sblimit_of(tbl);
    int bound = h.mode == 1 ? (h.mode_ext + 1) * 4 : sblimit;
    if (bound > sblimit || nch == 1) bound = sblimit;
    const bool on = lane < sblimit;
    const bool sep = lane < bound;                 // this subband carries separate codes per channel
    const int ncode = on ? (sep ? nch : 1) : 0;     // code sets of this subband
    const uint8_t *row = c_l2_rows[on ? c_l2_row_of_sb[tbl][lane] : 0];
    const int nbal = on ? row[0] : 0;
    uint32_t pos = (uint32_t)(4 + (h.crc ? 2 : 0)) * 8;
    int tot;
    // ---- bit allocation
    int off = warp_excl_scan(nbal * ncode, lane, &tot);
    int alloc[2] = {0, 0};
    if (on) {
        alloc[0] = (int)l2_bits(fb, pos + off, nbal, lim);
        alloc[1] = nch == 2 ? (sep ? (int)l2_bits(fb, pos + off + nbal, nbal, lim) : alloc[0]) : 0;
    }
    pos += (uint32_t)tot;
    // ---- scfsi: 2 bits per (subband, channel) that has samples
    const int na = (alloc[0] ? 1 : 0) + (alloc[1] ? 1 : 0);
    off = warp_excl_scan(2 * na, lane, &tot);
    int scfsi[2] = {0, 0};
    {
        uint32_t p = pos + off;
        for (int ch = 0; ch < nch; ch++)
            if (alloc[ch]) { scfsi[ch] = (int)l2_bits(fb, p, 2, lim); p += 2; }
    }
    pos += (uint32_t)tot;
    // ---- scalefactors: 3 / 2 / 1 / 2 indices of 6 bits by scfsi
    int nsf[2];
    for (int ch = 0; ch < 2; ch++) nsf[ch] = alloc[ch] ? (scfsi[ch] == 0 ? 3 : (scfsi[ch] == 2 ? 1 : 2)) : 0;
    off = warp_excl_scan(6 * (nsf[0] + nsf[1]), lane, &tot);
    float scf[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    {
        // one instruction stream for all lanes: an 18-bit window per channel, the indices picked by scfsi
        uint32_t p = pos + off;
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            const uint32_t w18 = l2_bits(fb, p, 18, lim);
            const uint32_t i0 = w18 >> 12, i1 = (w18 >> 6) & 63u, i2 = w18 & 63u;
            const int sel = scfsi[ch];
            const float a = l2_scf_of(i0), s1 = l2_scf_of(i1), s2 = l2_scf_of(i2);
            const float b = sel == 0 || sel == 3 ? s1 : a;                      // 0: a b c   1: a a b   2: a a a   3: a b b
            const float c = sel == 0 ? s2 : (sel == 2 ? a : s1);
            const bool on_ch = nsf[ch] != 0;
            scf[ch][0] = on_ch ? a : 0.f;
            scf[ch][1] = on_ch ? b : 0.f;
            scf[ch][2] = on_ch ? c : 0.f;
            p += 6u * (uint32_t)nsf[ch];
        }
    }
    pos += (uint32_t)tot;
    // ---- samples: 12 groups; within a group, subbands in order, channels inside a subband.  Everything that
    // depends only on the allocation is worked out once per frame: steps, field width, 1 / steps, and for
    // the grouped classes (3, 5, 9 steps: one codeword = three base-`steps` digits) a multiply-shift
    // reciprocal, exact for the 10-bit codewords.
    int steps[2], fbits[2], cbits[2];
    uint32_t rmul[2], rsh[2];
    float inv[2];
    for (int k = 0; k < 2; k++) {
        const int a = k < ncode ? alloc[k] : 0;
        const int q = a ? row[a] : 0;
        steps[k] = a ? c_l2_steps[q] : 1;
        fbits[k] = a ? c_l2_bits[q] : 0;                 // < 0: grouped
        cbits[k] = fbits[k] < 0 ? -fbits[k] : 3 * fbits[k];
        rmul[k] = steps[k] == 3 ? 0xAAABu : (steps[k] == 5 ? 0xCCCDu : 0xE38Fu);
        rsh[k] = steps[k] == 3 ? 17u : (steps[k] == 5 ? 18u : 19u);
        inv[k] = 1.f / (float)steps[k];
    }
    off = warp_excl_scan(cbits[0] + cbits[1], lane, &tot);
    const uint32_t fi = f - sr.frame_base;
    const size_t u0 = (size_t)(sr.unit_base - sr.sb_shift) + (size_t)fi * 2u * (size_t)nch; // dense index: Layer I / II streams only
#pragma unroll
    for (int part = 0; part < 3; part++) // scalefactor part: compile-time index into scf[][]
    for (int gr = part * 4; gr < part * 4 + 4; gr++) {
        uint32_t p = pos + (uint32_t)(gr * tot + off);
        float v[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        // one instruction stream for every lane, whatever its allocation class: a 64-bit window at the code
        // set, three fields of the class's width (a grouped class uses the first only and splits it), selects
        // instead of branches -- lanes of one warp hold different classes, a branch here ran 7 of 32 lanes
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t pc = min(p, lim), wi = pc >> 5, sh = pc & 31u;
            const uint32_t hi0 = __funnelshift_l(fb[wi + 1], fb[wi], sh), lo0 = __funnelshift_l(fb[wi + 2], fb[wi + 1], sh);
            const bool grouped = fbits[k] < 0;
            const int nb = grouped ? -fbits[k] : fbits[k]; // 0 (no samples) .. 16
            const uint32_t f0 = __funnelshift_l(hi0, 0u, nb);
            const uint32_t hi1 = __funnelshift_l(lo0, hi0, nb), lo1 = lo0 << nb;
            const uint32_t f1 = __funnelshift_l(hi1, 0u, nb);
            const uint32_t f2 = __funnelshift_l(__funnelshift_l(lo1, hi1, nb), 0u, nb);
            const uint32_t d1 = (f0 * rmul[k]) >> rsh[k], d2 = (d1 * rmul[k]) >> rsh[k];
            int code[3];
            code[0] = (int)(grouped ? f0 - d1 * (uint32_t)steps[k] : f0);
            code[1] = (int)(grouped ? d1 - d2 * (uint32_t)steps[k] : f1);
            code[2] = (int)(grouped ? d2 : f2);
            p += (uint32_t)cbits[k];
            const bool on_k = cbits[k] != 0;
            const float sc0 = scf[0][part] * inv[k], sc1 = scf[1][part] * inv[k], sck = scf[k][part] * inv[k];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float fr3 = on_k ? (float)(2 * code[i] + 1 - steps[k]) : 0.f;
                if (k == 0) { // the first code set feeds channel 0, and channel 1 too above the joint-stereo bound
                    v[0][i] = fr3 * (sep ? sck : sc0);
                    if (!sep) v[1][i] = fr3 * sc1;
                } else if (on_k) // a second code set exists only where the channels are coded separately
                    v[1][i] = fr3 * sck;
            }
        }
        const int g2 = gr >= 6 ? 1 : 0, t0 = (gr - 6 * g2) * 3; // slots 3 gr .. 3 gr + 2 of granule g2
        for (int ch = 0; ch < nch; ch++) {
            float *o = sb_out + (u0 + (size_t)g2 * nch + ch) * 576 + (size_t)t0 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 3; i++) o[i * 32] = v[ch][i];
        }
    }
}

// Layer I (11172-3 2.4.1.5 / 2.4.3.2): 4-bit allocations, one scalefactor, 12 samples of allocation + 1
// bits per subband.  A frame is 12 slots; a stream's slots are cut into 18-slot granules for the synthesis
// kernel, so slot s of frame f goes to granule (12 f + s) / 18.  The stream's last frame also zeroes the
// unused slots of the last granule.
__global__ void __launch_bounds__(L2_WARPS * 32)
k_layer1(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams, const L3FrameRec *__restrict__ frames,
         uint32_t nframes, float *__restrict__ sb_out)
{
    __shared__ uint32_t s_frame[L2_WARPS][L2_MAX_FRAME / 4 + 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * L2_WARPS + warp;
    if (f >= nframes) return;
    const L3FrameRec fr = frames[f];
    L3Hdr h;
    if (!l3_parse_hdr(fr.hdr, &h) || h.layer != 1) return;
    const L3StreamRec sr = streams[fr.stream];
    const uint8_t *src = raw + sr.raw_off + fr.rel_off;
    uint32_t *fb = s_frame[warp];
    const int flen = min(h.frame_len, L2_MAX_FRAME);
    l2_stage_frame(fb, src, flen, lane);
    __syncwarp();
    const uint32_t lim = (uint32_t)flen * 8u;

    const int nch = h.nch;
    const int bound = (h.mode == 1 && nch == 2) ? (h.mode_ext + 1) * 4 : 32;
    const bool sep = lane < bound;
    const int ncode = sep ? nch : 1;
    uint32_t pos = (uint32_t)(4 + (h.crc ? 2 : 0)) * 8;
    int tot;
    int off = warp_excl_scan(4 * ncode, lane, &tot);
    int alloc[2];
    alloc[0] = (int)l2_bits(fb, pos + off, 4, lim);
    alloc[1] = nch == 2 ? (sep ? (int)l2_bits(fb, pos + off + 4, 4, lim) : alloc[0]) : 0;
    pos += (uint32_t)tot;
    const int na = (alloc[0] ? 1 : 0) + (alloc[1] ? 1 : 0);
    off = warp_excl_scan(6 * na, lane, &tot);
    float scf[2] = {0.f, 0.f};
    {
        uint32_t p = pos + off;
        for (int ch = 0; ch < nch; ch++)
            if (alloc[ch]) { scf[ch] = l2_scf_of(l2_bits(fb, p, 6, lim)); p += 6; }
    }
    pos += (uint32_t)tot;
    // per code set: field width, steps and 1 / steps (0 for "no samples" and for the forbidden allocation 15),
    // worked out once; the sample loop is then one instruction stream for all lanes
    int cb[2], stp[2];
    float inv[2];
    for (int k = 0; k < 2; k++) {
        cb[k] = (k < ncode && alloc[k]) ? alloc[k] + 1 : 0;
        stp[k] = (1 << cb[k]) - 1;
        inv[k] = (cb[k] && alloc[k] != 15) ? 1.f / (float)stp[k] : 0.f;
    }
    off = warp_excl_scan(cb[0] + cb[1], lane, &tot);
    const uint32_t fi = f - sr.frame_base;
    for (int t = 0; t < 12; t++) {
        uint32_t p = pos + (uint32_t)(t * tot + off);
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int code = (int)l2_bits(fb, p, cb[k], lim);
            p += (uint32_t)cb[k];
            const float fr3 = (float)(2 * code + 1 - stp[k]) * inv[k];
            if (k == 0) {
                v[0] = fr3 * scf[0];
                if (!sep) v[1] = fr3 * scf[1];
            } else if (cb[1])
                v[1] = fr3 * scf[1];
        }
        const uint32_t slot = fi * 12u + (uint32_t)t;
        for (int ch = 0; ch < nch; ch++) {
            const size_t u = (size_t)(sr.unit_base - sr.sb_shift) + (size_t)(slot / 18u) * nch + ch;
            sb_out[u * 576 + (size_t)(slot % 18u) * 32 + lane] = v[ch];
        }
    }
    if (fi + 1 == sr.nframes) // the tail of the last granule
        for (uint32_t slot = sr.nframes * 12u; slot % 18u; slot++)
            for (int ch = 0; ch < nch; ch++) {
                const size_t u = (size_t)(sr.unit_base - sr.sb_shift) + (size_t)(slot / 18u) * nch + ch;
                sb_out[u * 576 + (size_t)(slot % 18u) * 32 + lane] = 0.f;
            }
}

} // namespace

void l3_layer2_init(void)
{
    cudaMemcpyToSymbol(c_l2_rows, l2_rows, sizeof l2_rows);
    cudaMemcpyToSymbol(c_l2_row_of_sb, l2_row_of_sb, sizeof l2_row_of_sb);
    cudaMemcpyToSymbol(c_l2_steps, l2_quant_steps, sizeof l2_quant_steps);
    cudaMemcpyToSymbol(c_l2_bits, l2_quant_bits, sizeof l2_quant_bits);
}

void l3_launch_layer2(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames, uint32_t nframes,
                      float *sb_out, cudaStream_t st)
{
    if (!nframes) return;
    k_layer2<<<(nframes + L2_WARPS - 1) / L2_WARPS, L2_WARPS * 32, 0, st>>>(raw, streams, frames, nframes, sb_out);
}

void l3_launch_layer1(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames, uint32_t nframes,
                      float *sb_out, cudaStream_t st)
{
    if (!nframes) return;
    k_layer1<<<(nframes + L2_WARPS - 1) / L2_WARPS, L2_WARPS * 32, 0, st>>>(raw, streams, frames, nframes, sb_out);
}
