/* walk_par.h -- the time-parallel frame walk, phase by phase, as host/device functions.
 *
 * The walk that finds a stream's frames (a1, a3) is a chain: a frame's position is known only after the previous
 * header is read.  But what the walk accepts at byte p depends on p and the stream's first header only, so the
 * chain can be cut: the bytes behind the first frame are split into segments of `seg` bytes; each segment is walked
 * from a GUESS of where the chain enters it (the first header of this stream that is followed by another one) into a
 * sparse table; then the guesses are checked in order -- segment t is right iff it starts where segment t - 1 left
 * -- and a segment whose guess was wrong (a header look-alike in front of the real frame, lost sync, a frame longer
 * than a segment) is walked again from the true position.  The dense table that results equals the serial walk's
 * (k_index_walk / host_index_stream) record for record.
 *
 * The CUDA kernels (k_index.cu: k_walk_first / _segments / _stitch / _compact) and the host emulation the CPU tests run
 * (tests/c/walk_emu.cpp) call the same functions, so the logic is tested without a GPU.
 * No reference code exists for this stage (/root/reference/README.md:1-84).
 */
#ifndef MP3B_WALK_PAR_H
#define MP3B_WALK_PAR_H

#include "l3_defs.h"

typedef struct L3WalkSeg {
    uint32_t start;    /* where the segment's walk began; L3WP_NONE: no frame start was guessed in it */
    uint32_t exit;     /* position after its last frame (>= the segment's end); L3WP_END: the chain ended in it */
    uint32_t n;        /* frames recorded */
    uint32_t payload;  /* main-data bytes of those frames */
} L3WalkSeg;

#define L3WP_NONE 0xffffffffu
#define L3WP_END 0xfffffffeu

/* frames that can START inside a segment of seg bytes (the shortest frame is 24 bytes) */
L3_HD uint32_t l3wp_seg_cap(uint32_t seg) { return seg / 24 + 2; }

typedef struct L3WalkFirst {
    uint32_t first;    /* the stream's first header (preset when streaming, else found) */
    uint32_t pf;       /* offset of the first frame */
    uint32_t have;     /* a first frame exists */
    uint32_t end0;     /* end_off to report when there is none (behind the ID3v2 tag) */
    uint32_t tag_kind, tag_frames, tag_bytes, tag_delay_pad;
} L3WalkFirst;

/* phase 0: the stream's first frame -- ID3v2 skip, sync search with confirmation, tag frame (as k_index_walk does) */
L3_HD void l3wp_first(const uint8_t *buf, uint32_t len, uint32_t first_preset, uint32_t skip_frames, int streaming,
                      L3WalkFirst *o)
{
    uint32_t p = l3_id3v2_len(buf, len), first = first_preset;
    o->have = 0;
    o->end0 = p;
    o->tag_kind = L3T_NONE;
    o->tag_frames = o->tag_bytes = o->tag_delay_pad = 0;
    while (p + 4 <= len) {
        L3Hdr h;
        uint32_t w;
        const int fa = l3_frame_at(buf, len, p, first, &h, &w);
        if (fa != 1 && !(fa == 3 && !streaming)) {
            if (fa >= 2 && streaming) break;
            p++;
            continue;
        }
        first = first ? first : w;
        if (skip_frames == 0)
            o->tag_kind = l3_parse_tag(buf + p, (uint32_t)h.frame_len, &h, &o->tag_frames, &o->tag_bytes, &o->tag_delay_pad);
        o->have = 1;
        break;
    }
    o->first = first;
    o->pf = p;
}

/* Walk from p while frames START below `limit`: the loop of k_index_walk with `first` known.  Returns the exit
 * position, or L3WP_END when the walk ended for good (end of the bytes, or an incomplete frame of a growing stream). */
L3_HD uint32_t l3wp_span(const uint8_t *buf, uint32_t len, uint32_t p, uint32_t limit, uint32_t first, int streaming,
                         uint32_t stream, L3FrameRec *out, uint32_t *n_out, uint32_t *payload_out)
{
    const uint32_t GEOM = 0xFFFFFCC0u;
    uint32_t prev_w = 0, base_len = 0, overhead = 0, pad_unit = 1, n = 0, payload = 0;
    int l2 = 0, ended = 1;
    while (p + 4 <= len) {
        if (p >= limit) { ended = 0; break; }
        L3Hdr h;
        uint32_t w = l3_load_be32(buf + p), flen;
        if (n && ((w ^ prev_w) & GEOM) == 0) {
            flen = base_len + ((w >> 9) & 1u) * pad_unit;
            if (p + flen > len) {
                if (streaming) break;
                p++;
                continue;
            }
        } else {
            const int fa = l3_frame_at(buf, len, p, first, &h, &w);
            if (fa != 1) { /* (`first` is set: the unconfirmed-first-frame case cannot occur) */
                if (fa >= 2 && streaming) break;
                p++;
                continue;
            }
            flen = (uint32_t)h.frame_len;
            prev_w = w;
            pad_unit = h.layer == 1 ? 4u : 1u;
            base_len = flen - ((w >> 9) & 1u) * pad_unit;
            overhead = 4u + (h.crc ? 2u : 0u) + (uint32_t)h.side_len;
            l2 = h.layer != 3;
        }
        L3FrameRec f;
        f.rel_off = p;
        f.payload_off = payload;
        f.hdr = w;
        f.stream = stream;
        out[n] = f;
        n++;
        payload += l2 ? 0u : flen - overhead;
        p += flen;
    }
    *n_out = n;
    *payload_out = payload;
    return ended ? L3WP_END : p;
}

/* phase 1: segment t of [pf, len) -- guess the entry, walk to the segment's end, records to `sp` (this segment's block) */
L3_HD void l3wp_segment(const uint8_t *buf, uint32_t len, uint32_t pf, uint32_t seg, uint32_t nseg, uint32_t t,
                        uint32_t first, int streaming, uint32_t stream, L3FrameRec *sp, L3WalkSeg *out)
{
    const uint32_t lo = pf + t * seg, hi = (t + 1 < nseg) ? lo + seg : L3WP_NONE;
    uint32_t g = lo;
    if (t) { /* the first header of this stream in the segment that is followed by another one */
        g = L3WP_NONE;
        for (uint32_t p = lo; p < hi && p + 4 <= len; p++) {
            if (buf[p] != 0xFF) continue; /* (cheap reject: 255 of 256 positions) */
            L3Hdr h, h2;
            uint32_t w;
            if (l3_frame_at(buf, len, p, first, &h, &w) != 1) continue;
            const uint32_t q = p + (uint32_t)h.frame_len;
            if (q + 4 <= len) {
                const uint32_t w2 = l3_load_be32(buf + q);
                if (!l3_parse_hdr(w2, &h2) || !l3_same_stream(w2, first)) continue;
            }
            g = p;
            break;
        }
    }
    out->start = g;
    out->exit = g;
    out->n = out->payload = 0;
    if (g != L3WP_NONE) out->exit = l3wp_span(buf, len, g, hi, first, streaming, stream, sp, &out->n, &out->payload);
}

/* phase 2 (one thread, only when some guess was wrong): follow the chain through the segments in order; a segment
 * the chain enters elsewhere than guessed is walked again from the true position, a segment it does not enter
 * (a long frame spans it, or the chain has ended) is voided.  `sparse` = the stream's blocks of `cap` records. */
L3_HD void l3wp_repair(const uint8_t *buf, uint32_t len, uint32_t pf, uint32_t seg, uint32_t nseg, uint32_t first,
                       int streaming, uint32_t stream, L3FrameRec *sparse, uint32_t cap, L3WalkSeg *sg)
{
    uint32_t cur = pf;
    int dead = 0;
    for (uint32_t t = 0; t < nseg; t++) {
        const uint32_t lo = pf + t * seg, hi = (t + 1 < nseg) ? lo + seg : L3WP_NONE;
        L3WalkSeg e = sg[t];
        if (dead || cur >= hi) { /* the chain does not enter this segment */
            if (e.n) {
                e.n = e.payload = 0;
                sg[t] = e;
            }
            continue;
        }
        if (e.start != cur) {
            e.start = cur;
            e.exit = l3wp_span(buf, len, cur, hi, first, streaming, stream, sparse + (size_t)t * cap, &e.n, &e.payload);
            sg[t] = e;
        }
        if (e.exit == L3WP_END) dead = 1;
        else cur = e.exit;
    }
}

/* the guess of segment t is consistent with its predecessor (phase 2's parallel pre-check) */
L3_HD int l3wp_chained(const L3WalkSeg *sg, uint32_t t, uint32_t pf)
{
    const uint32_t want = t ? sg[t - 1].exit : pf;
    if (want == L3WP_END) return sg[t].n == 0; /* the chain has ended: later segments must be empty */
    return sg[t].start == want;
}

#endif
