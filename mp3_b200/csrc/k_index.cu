// k_index.cu -- K0: device-side frame indexer (stages a1-a3 of SURVEY.md section 8).
//
//   index_walk    thread per stream: walk the frame chain once; count frames and main-data bytes and
//                                    leave one L3FrameRec per frame in a scratch table
//   side_parse    thread per frame : gather the record into the dense frame table;
//                                    side info -> unit descriptors (+ bit-reservoir offsets)
//   payload_copy  warp per frame   : compact the frames' main-data bytes into one arena so
//                                    that every unit's bits are contiguous and addressable
//
// The chain walk is the only serial dependency of the whole decoder (a frame's position is known
// only after the previous header is read); it is latency-bound, one dependent DRAM access per
// frame, and parallel across streams only.  Everything else is embarrassingly parallel.
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include "kernels.h"
#include "walk_par.h"

namespace {

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Scratch slot of frame i of stream s: frames are at least 24 bytes long (MPEG-2 LSF, 8 kbit/s,
// 24 kHz), so a stream of raw_len bytes starting at raw_off never needs more than raw_len / 24 + 1
// slots, and raw_off / 24 + s is a collision-free base.
__device__ __forceinline__ uint64_t scratch_base(const L3StreamRec &r, uint32_t s) { return r.raw_off / 24 + s; }

// One walk per stream: counts frames and main-data bytes AND leaves one record per frame in the
// scratch table, so that no second walk is needed once the host has turned the counts into bases.
__global__ void k_index_walk(const uint8_t *__restrict__ raw, L3StreamRec *__restrict__ streams, int nstreams,
                             L3FrameRec *__restrict__ scratch)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    L3StreamRec r = streams[s];
    const uint8_t *buf = raw + r.raw_off;
    L3FrameRec *out = scratch + scratch_base(r, (uint32_t)s);
    uint32_t len = r.raw_len, p = l3_id3v2_len(buf, len), first = r.first_hdr, first_off = 0, n = 0, payload = 0;
    uint32_t end_off = p;
    uint32_t tag_kind = L3T_NONE, tag_frames = 0, tag_bytes = 0, tag_dp = 0;
    const bool streaming = (r.flags & L3S_STREAMING) != 0;
    // Fast path for the overwhelmingly common case: the header equals the previous one in every bit that
    // determines the frame geometry (sync, version, layer, protection, bitrate, sample rate, mode), so
    // the length is the previous base length plus this frame's padding bit -- no header arithmetic.
    const uint32_t GEOM = 0xFFFFFCC0u;
    uint32_t prev_w = 0, base_len = 0, overhead = 0;
    bool l2 = false; // Layer I / II frames carry no main data for the arena
    uint32_t pad_unit = 1; // bytes a set padding bit adds: one slot = 1 byte, or 4 bytes in Layer I
    while (p + 4 <= len) {
        L3Hdr h;
        uint32_t w = l3_load_be32(buf + p), flen;
        if (n && ((w ^ prev_w) & GEOM) == 0) {
            flen = base_len + ((w >> 9) & 1u) * pad_unit;
            if (p + flen > len) {
                if (streaming) break; // the rest of this frame has not arrived yet
                p++;
                continue;
            }
        } else {
            const int fa = l3_frame_at(buf, len, p, first, &h, &w);
            if (fa != 1 && !(fa == 3 && !streaming)) {
                if (fa >= 2 && streaming) break; // wait for the rest of the frame / the confirming next header
                p++;
                continue;
            }
            flen = (uint32_t)h.frame_len;
            prev_w = w;
            pad_unit = h.layer == 1 ? 4u : 1u;
            base_len = flen - ((w >> 9) & 1u) * pad_unit;
            overhead = 4u + (h.crc ? 2u : 0u) + (uint32_t)h.side_len;
            l2 = h.layer != 3;
            if (n == 0) {
                first = first ? first : w;
                first_off = p;
                if (r.skip_frames == 0) tag_kind = l3_parse_tag(buf + p, flen, &h, &tag_frames, &tag_bytes, &tag_dp);
                for (int j = 1; j < 8; j++) // the chain is latency-bound: pull the next headers towards L2
                    if (p + (uint32_t)j * flen < len) prefetch_l2(buf + p + j * flen);
            }
        }
        if (p + 8u * flen < len) prefetch_l2(buf + p + 8 * flen);
        L3FrameRec f;
        f.rel_off = p;
        f.payload_off = payload;
        f.hdr = w;
        f.stream = (uint32_t)s;
        out[n] = f;
        n++;
        payload += l2 ? 0u : flen - overhead;
        p += flen;
        end_off = p;
    }
    streams[s].end_off = end_off;
    streams[s].first_off = first_off;
    streams[s].first_hdr = first;
    streams[s].nframes = n;
    streams[s].payload_len = payload;
    streams[s].tag_kind = tag_kind;
    streams[s].tag_frames = tag_frames;
    streams[s].tag_bytes = tag_bytes;
    streams[s].tag_delay_pad = tag_dp;
}

// ---- the same walk, time-parallel inside a stream (walk_par.h has the algorithm; this is its mapping) ----
// Four small kernels, so that ONE long stream is spread over the whole chip instead of one CTA:
//   k_walk_first     a thread per stream: the first frame (phase 0)
//   k_walk_segments  a thread per segment slot of the batch: the speculative walks (phase 1) -- an hour of audio is
//                    some 14,000 segments of 4 KB, i.e. about ten frames per thread
//   k_walk_stitch    a CTA per stream: do the guesses chain up (else one thread repairs), exclusive scan of the
//                    segments' frame and main-data counts, the stream's totals (phases 2 and 3)
//   k_walk_compact   a warp per segment slot: its records to their dense positions
// Slot j of the batch belongs to the stream s with seg0(s) <= j < seg0(s + 1), seg0(s) = raw_off / seg + s (every
// stream has at least the slots its own segments need; k_walk_first leaves the table); slots beyond a stream's last
// segment stay empty.
constexpr int WP_THREADS = 128;
#ifndef WP_SPREAD
#define WP_SPREAD 4
#endif
static_assert(sizeof(L3WalkFirst) == 32 && sizeof(L3WalkSeg) == 16, "l3_walk_seg_bytes (kernels.h) sizes the buffer with these");

// the stream that owns slot j (the last one whose first slot is <= j)
__device__ __forceinline__ int wp_stream_of(const uint32_t *__restrict__ seg0, int nstreams, uint32_t j)
{
    int lo = 0, hi = nstreams - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(seg0 + mid) <= j) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ uint32_t wp_nseg(const L3WalkFirst &f, uint32_t len, uint32_t seg)
{
    return f.have ? (len - f.pf + seg - 1) / seg : 0u;
}

__global__ void k_walk_first(const uint8_t *__restrict__ raw, L3StreamRec *__restrict__ streams, int nstreams,
                             L3WalkFirst *__restrict__ firsts, uint32_t *__restrict__ seg0, uint32_t seg)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const L3StreamRec r = streams[s];
    seg0[s] = (uint32_t)(r.raw_off / seg) + (uint32_t)s;
    L3WalkFirst f;
    l3wp_first(raw + r.raw_off, r.raw_len, r.first_hdr, r.skip_frames, (r.flags & L3S_STREAMING) != 0, &f);
    firsts[s] = f;
    if (!f.have) { // no frame at all (or not yet): same outputs as the serial walk
        streams[s].end_off = f.end0;
        streams[s].first_off = 0;
        streams[s].first_hdr = f.first;
        streams[s].nframes = 0;
        streams[s].payload_len = 0;
        streams[s].tag_kind = L3T_NONE;
        streams[s].tag_frames = streams[s].tag_bytes = streams[s].tag_delay_pad = 0;
    }
}

__global__ void __launch_bounds__(WP_THREADS)
k_walk_segments(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams, int nstreams,
                const L3WalkFirst *__restrict__ firsts, const uint32_t *__restrict__ seg0,
                L3FrameRec *__restrict__ sparse, L3WalkSeg *__restrict__ segs, uint32_t nslots, uint32_t seg, uint32_t seg_cap)
{
    // Only every WP_SPREAD-th lane works.  The kernel is bound by latency, not by throughput: the lanes of a warp part
    // ways in the entry search (sync look-alikes in the main data), so a warp's time is the SUM of its lanes'
    // searches; spreading the segments over more warps shortens every warp's serial instruction stream.
    const uint32_t gt = blockIdx.x * WP_THREADS + threadIdx.x, j = gt / WP_SPREAD;
    if (gt % WP_SPREAD || j >= nslots) return;
    const int s = wp_stream_of(seg0, nstreams, j);
    const L3StreamRec r = streams[s];
    const L3WalkFirst f = firsts[s];
    const uint32_t t = j - seg0[s];
    L3WalkSeg e;
    e.start = e.exit = L3WP_NONE;
    e.n = e.payload = 0;
    if (t < wp_nseg(f, r.raw_len, seg))
        l3wp_segment(raw + r.raw_off, r.raw_len, f.pf, seg, wp_nseg(f, r.raw_len, seg), t, f.first,
                     (r.flags & L3S_STREAMING) != 0, (uint32_t)s, sparse + (size_t)j * seg_cap, &e);
    segs[j] = e;
}

__global__ void __launch_bounds__(WP_THREADS)
k_walk_stitch(const uint8_t *__restrict__ raw, L3StreamRec *__restrict__ streams, int nstreams,
              const L3WalkFirst *__restrict__ firsts, const uint32_t *__restrict__ seg0v,
              L3FrameRec *__restrict__ sparse, L3WalkSeg *__restrict__ segs, uint32_t seg, uint32_t seg_cap)
{
    const int s = blockIdx.x, tid = threadIdx.x;
    const L3WalkFirst f = firsts[s];
    if (!f.have) return;
    const L3StreamRec r = streams[s];
    const uint8_t *buf = raw + r.raw_off;
    const uint32_t len = r.raw_len, nseg = wp_nseg(f, len, seg);
    const size_t seg0 = seg0v[s];
    L3WalkSeg *sg = segs + seg0;
    L3FrameRec *sp = sparse + seg0 * seg_cap;
    __shared__ uint32_t sh_bad, sh_last, sh_from;
    __shared__ uint32_t sh_n[WP_THREADS], sh_pay[WP_THREADS];
    if (tid == 0) {
        sh_bad = L3WP_NONE; // the first segment at or behind sh_from that does not chain up
        sh_from = 0;
        sh_last = 0;        // 1 + the last segment that holds a frame
    }
    __syncthreads();
    // ---- phase 2: do the guesses chain up?  Where they do not, one thread follows the chain from the first such
    // segment until the guesses hold again (l3wp_repair_run); the search for the next one is parallel again.  A stream
    // that needs many such rounds (damage all over, segments shorter than its frames) gets one serial pass instead.
    for (int round = 0;; round++) {
        const uint32_t from = sh_from;
        uint32_t bad = L3WP_NONE;
        for (uint32_t t = from + tid; t < nseg; t += WP_THREADS)
            if (!l3wp_chained(sg, t, f.pf)) {
                bad = t;
                break;
            }
        if (bad != L3WP_NONE) atomicMin(&sh_bad, bad);
        __syncthreads();
        const uint32_t t0 = sh_bad;
        if (t0 == L3WP_NONE) break;
        __syncthreads();
        if (tid == 0) {
            sh_from = l3wp_repair_run(buf, len, f.pf, seg, nseg, f.first, (r.flags & L3S_STREAMING) != 0, (uint32_t)s, sp, seg_cap, sg,
                                      t0, round >= 32);
            sh_bad = L3WP_NONE;
        }
        __syncthreads();
    }
    // ---- phase 3: dense positions = exclusive scan of the segments' frame and main-data counts.  A thread owns a
    // run of consecutive segments; the segment records' start / exit fields (dead from here on) take the bases.
    const uint32_t per = (nseg + WP_THREADS - 1) / WP_THREADS, t0 = min(nseg, tid * per), t1 = min(nseg, t0 + per);
    uint32_t cn = 0, cp = 0, last = 0;
    for (uint32_t t = t0; t < t1; t++) {
        const uint32_t n = sg[t].n;
        cn += n;
        cp += sg[t].payload;
        if (n) last = t + 1;
    }
    sh_n[tid] = cn;
    sh_pay[tid] = cp;
    if (last) atomicMax(&sh_last, last);
    __syncthreads();
    uint32_t bn = 0, bp = 0, tot_n = 0, tot_p = 0;
    for (int k = 0; k < WP_THREADS; k++) {
        if (k == tid) { bn = tot_n; bp = tot_p; }
        tot_n += sh_n[k];
        tot_p += sh_pay[k];
    }
    for (uint32_t t = t0; t < t1; t++) {
        L3WalkSeg e = sg[t];
        e.start = bn;
        e.exit = bp;
        bn += e.n;
        bp += e.payload;
        sg[t] = e;
    }
    if (tid == 0) {
        uint32_t end_off = f.end0;
        if (sh_last) { // end of the last frame = header offset + length of the last record
            const uint32_t t = sh_last - 1;
            const L3FrameRec lf = sp[(size_t)t * seg_cap + (sg[t].n - 1)];
            L3Hdr h;
            l3_parse_hdr(lf.hdr, &h);
            end_off = lf.rel_off + (uint32_t)h.frame_len;
        }
        streams[s].end_off = end_off;
        streams[s].first_off = f.pf;
        streams[s].first_hdr = f.first;
        streams[s].nframes = tot_n;
        streams[s].payload_len = tot_p;
        streams[s].tag_kind = f.tag_kind;
        streams[s].tag_frames = f.tag_frames;
        streams[s].tag_bytes = f.tag_bytes;
        streams[s].tag_delay_pad = f.tag_delay_pad;
    }
}

__global__ void __launch_bounds__(WP_THREADS)
k_walk_compact(const L3StreamRec *__restrict__ streams, int nstreams, const uint32_t *__restrict__ seg0,
               const L3FrameRec *__restrict__ sparse, const L3WalkSeg *__restrict__ segs, L3FrameRec *__restrict__ dense,
               uint32_t nslots, uint32_t seg_cap)
{
    const uint32_t j = blockIdx.x * (WP_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= nslots) return;
    const L3WalkSeg e = segs[j];
    if (!e.n) return;
    const int s = wp_stream_of(seg0, nstreams, j);
    const L3FrameRec *src = sparse + (size_t)j * seg_cap;
    L3FrameRec *dst = dense + scratch_base(streams[s], (uint32_t)s) + e.start;
    for (uint32_t i = lane; i < e.n; i += 32) {
        L3FrameRec f = src[i];
        f.payload_off += e.exit;
        dst[i] = f;
    }
}

// Sequential big-endian bit reader over the side info (9..32 bytes), which the thread has staged in
// its own shared-memory slot: the global loads of a frame's side info are issued together (aligned
// words, all in flight at once) instead of byte after dependent byte.
constexpr int SP_THREADS = 128, SP_WORDS = 9, SP_STRIDE = 11; // 9 aligned words cover 32 bytes at any alignment
struct SideBits {
    const uint8_t *p, *end; // bytes at or beyond `end` read as zero
    unsigned long long acc;
    int n;
    __device__ __forceinline__ uint32_t byte() { return p < end ? *p++ : (p++, 0u); }
    // copy the side info [q, q + len) (clipped to the stream's end e) into `slot` and read from there
    __device__ __forceinline__ void stage(const uint8_t *q, const uint8_t *e, uint32_t len, uint32_t *slot)
    {
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 3u);
        const uint32_t *base = reinterpret_cast<const uint32_t *>(q - sh);
        const uint32_t avail = (uint32_t)max((ptrdiff_t)0, min((ptrdiff_t)len, e - q)); // side-info bytes that exist
        uint32_t w[SP_WORDS];
#pragma unroll
        for (int k = 0; k < SP_WORDS; k++) {
            // word k is needed if it holds a byte of the side info; a word that would reach past the
            // stream's end is fetched byte by byte (the buffer may end right there)
            const bool need = 4u * k < sh + avail;
            const bool whole = reinterpret_cast<const uint8_t *>(base + k + 1) <= e;
            uint32_t v = 0;
            if (need && whole) v = __ldg(base + k);
            else if (need)
                for (int j = 0; j < 4; j++) {
                    const uint8_t *bp = reinterpret_cast<const uint8_t *>(base + k) + j;
                    if (bp >= q && bp < e) v |= (uint32_t)*bp << (8 * j);
                }
            w[k] = v;
        }
#pragma unroll
        for (int k = 0; k < SP_WORDS; k++) slot[k] = w[k];
        init(reinterpret_cast<const uint8_t *>(slot) + sh, reinterpret_cast<const uint8_t *>(slot) + sh + avail);
    }
    __device__ __forceinline__ void init(const uint8_t *q, const uint8_t *e)
    {
        p = q;
        end = e;
        acc = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) acc = (acc << 8) | byte();
        n = 64;
    }
    __device__ __forceinline__ uint32_t get(int k) // k in 1..16
    {
        if (n < 32) {
            uint32_t w = byte() << 24;
            w |= byte() << 16;
            w |= byte() << 8;
            w |= byte();
            acc |= (unsigned long long)w << (32 - n);
            n += 32;
        }
        uint32_t v = (uint32_t)(acc >> (64 - k));
        acc <<= k;
        n -= k;
        return v;
    }
};

// Thread per frame: side info -> unit descriptors (a2, a3).  In gather mode (scratch != null) the
// frame record is fetched from the walk's scratch table (stream found by binary search over the
// frame bases) and also written to the dense frame table.
__global__ void k_side_parse(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams, int nstreams,
                             L3FrameRec *__restrict__ frames, const L3FrameRec *__restrict__ scratch,
                             uint32_t nframes, const uint16_t *__restrict__ sfb_long_all,
                             L3UnitDesc *__restrict__ units, uint32_t *__restrict__ gran_unit0,
                             uint32_t *__restrict__ concealed, int verify_crc)
{
    pdl_launch_dependents(); // the main-data compaction may be scheduled behind this grid's last wave (kernels.h)
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    L3FrameRec fr;
    if (scratch) {
        int lo = 0, hi = nstreams - 1; // last stream with frame_base <= f (empty streams share a base)
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (streams[mid].frame_base <= f) lo = mid;
            else hi = mid - 1;
        }
        // skip back over empty streams that share this base
        while (lo > 0 && streams[lo].nframes == 0) lo--;
        const L3StreamRec &r = streams[lo];
        fr = scratch[scratch_base(r, (uint32_t)lo) + (f - r.frame_base)];
        frames[f] = fr;
    } else
        fr = frames[f];
    const L3StreamRec sr = streams[fr.stream];
    L3Hdr h;
    l3_parse_hdr(fr.hdr, &h);
    const uint32_t fi = f - sr.frame_base; /* frame index inside its stream */
    if (h.layer == 1) {
        // Layer I: 12 slots per frame; the stream's granules are 18-slot pieces of its slot sequence, and the
        // frame that holds a granule's first slot writes its placeholder units
        const uint32_t g = (12u * fi + 17u) / 18u; // first granule starting at or after this frame's first slot
        if (18u * g < 12u * fi + 12u) {
            const uint32_t u0 = sr.unit_base + g * (uint32_t)h.nch;
            L3UnitDesc d;
            memset(&d, 0, sizeof d);
            d.hdr = (uint8_t)((h.sr_row << L3H_SR_SHIFT) | (h.nch == 2 ? L3H_STEREO : 0) | (h.lsf ? L3H_LSF : 0));
            d.stream = fr.stream;
            for (int ch = 0; ch < h.nch; ch++) {
                d.pos = (uint8_t)((ch ? L3P_CH : 0) | (g == 0 ? L3P_FIRST : 0));
                units[u0 + ch] = d;
            }
            gran_unit0[sr.gran_base + g] = u0 | (h.nch == 2 ? L3G_STEREO : 0u) | (g == 0 ? L3G_FIRST : 0u);
        }
        return;
    }
    if (h.layer == 2) {
        // Layer II: no side info; its units exist (two granules x channels per frame, the PCM layout is the
        // same) but stay invalid for the Huffman / back-end kernels: k_layer2 + the synthesis kernel fill them
        const uint32_t u0 = sr.unit_base + fi * (uint32_t)(2 * h.nch), g0 = sr.gran_base + fi * 2u;
        L3UnitDesc d;
        memset(&d, 0, sizeof d);
        d.hdr = (uint8_t)(L3H_LSF * 0 | (h.sr_row << L3H_SR_SHIFT) | (h.nch == 2 ? L3H_STEREO : 0) | (h.lsf ? L3H_LSF : 0));
        d.stream = fr.stream;
        for (int gr = 0; gr < 2; gr++) {
            for (int ch = 0; ch < h.nch; ch++) {
                d.pos = (uint8_t)((gr ? L3P_GR : 0) | (ch ? L3P_CH : 0) | ((fi == 0 && gr == 0) ? L3P_FIRST : 0));
                units[u0 + gr * h.nch + ch] = d;
            }
            gran_unit0[g0 + gr] = (u0 + (uint32_t)(gr * h.nch)) | (h.nch == 2 ? L3G_STEREO : 0u) |
                                  ((fi == 0 && gr == 0) ? L3G_FIRST : 0u);
        }
        return;
    }
    const uint16_t *sfb_long = sfb_long_all + h.sr_row * 23;
    __shared__ uint32_t s_side[SP_THREADS][SP_STRIDE];
    SideBits b;
    b.stage(raw + sr.raw_off + fr.rel_off + 4 + (h.crc ? 2 : 0), raw + sr.raw_off + sr.raw_len, (uint32_t)h.side_len,
            s_side[threadIdx.x]);
    const int nch = h.nch, ngr = h.ngr;
    uint32_t mdb, scfsi[2] = {0, 0};
    if (!h.lsf) {
        mdb = b.get(9);
        b.get(nch == 1 ? 5 : 3);
        for (int ch = 0; ch < nch; ch++) {
            const uint32_t v = b.get(4); // scfsi bits for groups 0..3, first transmitted = group 0
            scfsi[ch] = ((v >> 3) & 1) | (((v >> 2) & 1) << 1) | (((v >> 1) & 1) << 2) | ((v & 1) << 3);
        }
    } else {
        mdb = b.get(8);
        b.get(nch == 1 ? 1 : 2);
    }
    int valid = mdb <= fr.payload_off;
    if (verify_crc && h.crc) { // protected frame: a CRC mismatch conceals it like a lost reservoir does
        const uint8_t *fp = raw + sr.raw_off + fr.rel_off;
        const uint32_t avail = sr.raw_len - fr.rel_off; // >= frame_len >= 6 + side_len (checked by the walk)
        if (avail >= 6u + (uint32_t)h.side_len) {
            uint32_t crc = l3_crc16(0xffffu, fp + 2, 2);
            crc = l3_crc16(crc, fp + 6, (uint32_t)h.side_len);
            if (crc != (((uint32_t)fp[4] << 8) | fp[5])) valid = 0;
        }
    }
    if (!valid && fi >= sr.skip_frames) atomicAdd(concealed, 1u);
    uint64_t bit = valid ? (sr.payload_base + fr.payload_off - mdb) * 8ull : 0ull;
    uint8_t hdrbits = (uint8_t)((h.lsf ? L3H_LSF : 0) | (h.sr_row << L3H_SR_SHIFT) | (nch == 2 ? L3H_STEREO : 0));
    if (h.mode == 1) hdrbits |= (uint8_t)(((h.mode_ext & 2) ? L3H_MS : 0) | ((h.mode_ext & 1) ? L3H_IS : 0));
    const uint32_t u0 = sr.unit_base + fi * (uint32_t)(ngr * nch);
    const uint32_t g0 = sr.gran_base + fi * (uint32_t)ngr;
    for (int gr = 0; gr < ngr; gr++) {
        for (int ch = 0; ch < nch; ch++) {
            L3UnitDesc d;
            uint32_t p23 = b.get(12), bv = b.get(9), gg = b.get(8);
            uint32_t sfc = b.get(h.lsf ? 9 : 4), ws = b.get(1);
            uint32_t bt = 0, mixed = 0, t0, t1, t2 = 0, r0c = 0, r1c = 0, sbg0 = 0, sbg1 = 0, sbg2 = 0;
            if (ws) {
                const uint32_t v = b.get(13); // block_type 2, mixed 1, table_select 5 + 5
                bt = v >> 11;
                mixed = (v >> 10) & 1;
                t0 = (v >> 5) & 31;
                t1 = v & 31;
                const uint32_t g3 = b.get(9);
                sbg0 = g3 >> 6;
                sbg1 = (g3 >> 3) & 7;
                sbg2 = g3 & 7;
            } else {
                const uint32_t v = b.get(15);
                t0 = v >> 10;
                t1 = (v >> 5) & 31;
                t2 = v & 31;
                const uint32_t rc = b.get(7);
                r0c = rc >> 3;
                r1c = rc & 7;
            }
            uint32_t preflag = h.lsf ? 0 : b.get(1);
            const uint32_t tail = b.get(2);
            const uint32_t sfscale = tail >> 1, c1tab = tail & 1;
            if (h.lsf && !((hdrbits & L3H_IS) && ch == 1) && sfc >= 500) preflag = 1;
            if (bv > 288) bv = 288;
            uint32_t bv2 = bv * 2, r1, r2;
            if (ws) {
                r1 = (bt == 2 || !h.lsf) ? (h.sr_row == 8 ? 72 : 36) : (h.sr_row == 8 ? 108 : 54); // sfb_long[8] / 3 sfb_short[3]
                r2 = 576;
            } else {
                uint32_t a = r0c + 1, c = r0c + r1c + 2;
                if (a > 22) a = 22;
                if (c > 22) c = 22;
                r1 = sfb_long[a];
                r2 = sfb_long[c];
            }
            if (r1 > bv2) r1 = bv2;
            if (r2 > bv2) r2 = bv2;
            d.bit_off = bit;
            d.p23len = (uint16_t)(valid ? p23 : 0);
            d.big_values = (uint16_t)(valid ? bv : 0);
            d.r1 = (uint16_t)(valid ? r1 : 0);
            d.r2 = (uint16_t)(valid ? r2 : 0);
            d.sfc = (uint16_t)sfc;
            d.global_gain = (uint8_t)gg;
            d.tsel[0] = (uint8_t)t0; d.tsel[1] = (uint8_t)t1; d.tsel[2] = (uint8_t)t2;
            d.sbg[0] = (uint8_t)sbg0; d.sbg[1] = (uint8_t)sbg1; d.sbg[2] = (uint8_t)sbg2;
            d.flags = (uint8_t)(bt | (mixed ? L3F_MIXED : 0) | (preflag ? L3F_PREFLAG : 0) |
                                (sfscale ? L3F_SFSCALE : 0) | (c1tab ? L3F_C1TAB : 0) |
                                (valid ? L3F_VALID : 0) | (ws ? L3F_WS : 0));
            d.hdr = hdrbits;
            d.pos = (uint8_t)((gr ? L3P_GR : 0) | (ch ? L3P_CH : 0) | (scfsi[ch] << L3P_SCFSI_SHIFT) |
                              ((fi == 0 && gr == 0) ? L3P_FIRST : 0));
            d.stream = fr.stream;
            units[u0 + gr * nch + ch] = d;
            if (valid) bit += p23;
        }
        gran_unit0[g0 + gr] = (u0 + (uint32_t)(gr * nch)) | (nch == 2 ? L3G_STEREO : 0u) |
                              ((fi == 0 && gr == 0) ? L3G_FIRST : 0u);
    }
    // A frame's main data ends inside the frame (11172-3 2.4.2.7: the next frame's main_data_begin cannot
    // point forward).  Side info that claims more bits than that is damaged: conceal the frame, so that no
    // unit reads beyond its own stream's data.
    const uint64_t frame_end = (sr.payload_base + fr.payload_off + (uint32_t)(h.frame_len - 4 - (h.crc ? 2 : 0) - h.side_len)) * 8ull;
    if (valid && bit > frame_end) {
        if (fi >= sr.skip_frames) atomicAdd(concealed, 1u);
        for (int k = 0; k < ngr * nch; k++) {
            L3UnitDesc &d = units[u0 + k];
            d.bit_off = 0;
            d.p23len = d.big_values = d.r1 = d.r2 = 0;
            d.flags &= (uint8_t)~L3F_VALID;
        }
    }
}

// Main-data compaction.  A CTA takes 64 frames: 64 threads first resolve one frame each (record,
// stream record, header -> source, destination, length) into shared memory, so that the dependent
// loads of all 64 frames are in flight together instead of heading every warp's copy; then each warp
// copies eight frames.  Source and destination have arbitrary alignment: the destination is brought to
// a word boundary with a few byte stores, then each lane builds one aligned destination word from two
// aligned source words (byte permute), 128 bytes per warp instruction; a byte tail finishes.
constexpr int PC_FRAMES = 64, PC_THREADS = 256;
__global__ void __launch_bounds__(PC_THREADS)
k_payload_copy(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams,
               const L3FrameRec *__restrict__ frames, uint32_t nframes, uint8_t *__restrict__ arena)
{
    __shared__ const uint8_t *s_src[PC_FRAMES];
    __shared__ uint8_t *s_dst[PC_FRAMES];
    __shared__ uint32_t s_n[PC_FRAMES];
    pdl_launch_dependents();
    pdl_wait(); // frame records: written by k_side_parse
    const uint32_t f0 = blockIdx.x * PC_FRAMES, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < PC_FRAMES) {
        const uint32_t f = f0 + threadIdx.x;
        uint32_t n = 0;
        if (f < nframes) {
            const L3FrameRec fr = frames[f];
            const L3StreamRec &sr = streams[fr.stream];
            L3Hdr h;
            l3_parse_hdr(fr.hdr, &h);
            const uint32_t skip = 4u + (h.crc ? 2u : 0u) + (uint32_t)h.side_len;
            n = h.layer != 3 ? 0u : (uint32_t)h.frame_len - skip;
            s_src[threadIdx.x] = raw + sr.raw_off + fr.rel_off + skip;
            s_dst[threadIdx.x] = arena + sr.payload_base + fr.payload_off;
            if (f + 1 == sr.frame_base + sr.nframes) { // the stream's last frame: zero the padding behind its main data
                uint8_t *pad = arena + sr.payload_base + sr.payload_len;
                const uint32_t npad = ((sr.payload_len + L3_PAYLOAD_PAD + 15u) & ~15u) - sr.payload_len;
                for (uint32_t i = 0; i < npad; i++) pad[i] = 0;
            }
        }
        s_n[threadIdx.x] = n;
    }
    __syncthreads();
    // Each warp copies eight frames.  A frame's bytes: up to 15 single bytes until the destination is 16-byte aligned,
    // then 16-byte vectors -- a lane builds one from five aligned source words (byte permute; the source has any
    // alignment) --, then the rest as bytes.  Vectors that would need a source word beyond the frame's own bytes are
    // left to the byte loop: nothing is read past the frame.
    for (uint32_t k = warp; k < PC_FRAMES; k += PC_THREADS / 32) {
        const uint32_t n = s_n[k];
        if (!n) continue;
        const uint8_t *src = s_src[k];
        uint8_t *dst = s_dst[k];
        const uint32_t head = min(n, (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
        if (lane < head) dst[lane] = src[lane];
        const uint8_t *sb = src + head;
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sb) & 3u);
        const uint32_t rest = n - head;
        const uint32_t nv = rest >= 20u ? (rest - (sh ? 4u : 0u)) >> 4 : 0u;
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(sb - sh);
        uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
        const uint32_t sel = 0x3210u + 0x1111u * sh;
        for (uint32_t v0 = 0; v0 < nv; v0 += 32) { // (warp-uniform trip count)
            const uint32_t v = v0 + lane;
            if (v < nv) {
                const uint32_t *q = sw + 4 * v;
                const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = sh ? __ldg(q + 4) : 0u;
                dv[v] = make_uint4(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel),
                                   __byte_perm(w3, w4, sel));
            }
        }
        for (uint32_t i = head + nv * 16u + lane; i < n; i += 32) dst[i] = src[i];
    }
}

__global__ void k_publish_words(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst_host, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst_host[i] = src[i];
    __threadfence_system();
}

} // namespace

void l3_launch_publish(const void *src_dev, void *dst_pinned_host, size_t bytes, cudaStream_t st)
{
    const uint32_t n = (uint32_t)(bytes / 4);
    if (!n) return;
    const unsigned blocks = n > 65536 ? 64 : (n + 255) / 256;
    k_publish_words<<<blocks, 256, 0, st>>>(static_cast<const uint32_t *>(src_dev), static_cast<uint32_t *>(dst_pinned_host), n);
}
void l3_launch_index_walk(const uint8_t *raw, L3StreamRec *streams, int nstreams, L3FrameRec *scratch, cudaStream_t st)
{
    if (nstreams <= 0) return;
    k_index_walk<<<(nstreams + 31) / 32, 32, 0, st>>>(raw, streams, nstreams, scratch);
}
void l3_launch_index_walk_par(const uint8_t *raw, L3StreamRec *streams, int nstreams, L3FrameRec *dense, L3FrameRec *sparse,
                              void *segs, uint64_t nslots64, uint32_t seg_bytes, cudaStream_t st)
{
    if (nstreams <= 0) return;
    const uint32_t nslots = (uint32_t)nslots64; // (the caller sizes the segments so that a batch has a few million at most)
    L3WalkSeg *sg = static_cast<L3WalkSeg *>(segs);
    // behind the segment records: a first-frame record and the first slot of every stream (see l3_walk_seg_bytes)
    L3WalkFirst *firsts = reinterpret_cast<L3WalkFirst *>(sg + nslots);
    uint32_t *seg0 = reinterpret_cast<uint32_t *>(firsts + nstreams);
    const uint32_t cap = l3wp_seg_cap(seg_bytes);
    k_walk_first<<<(nstreams + 31) / 32, 32, 0, st>>>(raw, streams, nstreams, firsts, seg0, seg_bytes);
    k_walk_segments<<<(unsigned)(((uint64_t)nslots * WP_SPREAD + WP_THREADS - 1) / WP_THREADS), WP_THREADS, 0, st>>>(raw, streams, nstreams, firsts, seg0, sparse, sg, nslots,
                                                                             seg_bytes, cap);
    k_walk_stitch<<<nstreams, WP_THREADS, 0, st>>>(raw, streams, nstreams, firsts, seg0, sparse, sg, seg_bytes, cap);
    k_walk_compact<<<(nslots + WP_THREADS / 32 - 1) / (WP_THREADS / 32), WP_THREADS, 0, st>>>(streams, nstreams, seg0, sparse, sg, dense, nslots,
                                                                                      cap);
}
void l3_launch_side_parse(const uint8_t *raw, const L3StreamRec *streams, int nstreams, L3FrameRec *frames,
                          const L3FrameRec *scratch, uint32_t nframes, const L3DevTables &T, L3UnitDesc *units,
                          uint32_t *gran_unit0, uint32_t *concealed_counter, int verify_crc, cudaStream_t st)
{
    if (!nframes) return;
    k_side_parse<<<(nframes + SP_THREADS - 1) / SP_THREADS, SP_THREADS, 0, st>>>(raw, streams, nstreams, frames, scratch, nframes, T.sfb_long,
                                                        units, gran_unit0, concealed_counter, verify_crc);
}
void l3_launch_payload_copy(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames,
                            uint32_t nframes, uint8_t *arena, cudaStream_t st, bool pdl)
{
    if (!nframes) return;
    l3_launch_k(k_payload_copy, dim3((nframes + PC_FRAMES - 1) / PC_FRAMES), dim3(PC_THREADS), 0, st, pdl, raw, streams, frames,
                nframes, arena);
}
