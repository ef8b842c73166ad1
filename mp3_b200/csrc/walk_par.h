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

/* the guess: is p the first header of this stream that is followed by another one? */
L3_HD int l3wp_entry_at(const uint8_t *buf, uint32_t len, uint32_t p, uint32_t first)
{
    L3Hdr h, h2;
    uint32_t w;
    if (l3_frame_at(buf, len, p, first, &h, &w) != 1) return 0;
    const uint32_t q = p + (uint32_t)h.frame_len;
    if (q + 4 <= len) {
        const uint32_t w2 = l3_load_be32(buf + q);
        if (!l3_parse_hdr(w2, &h2) || !l3_same_stream(w2, first)) return 0;
    }
    return 1;
}

L3_HD uint32_t l3wp_ctz(uint32_t x) /* x != 0 */
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}

/* 16 bytes from a 16-byte aligned address, as four little-endian words */
L3_HD void l3wp_load16(const uint8_t *p, uint32_t w[4])
{
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#else
    for (int k = 0; k < 4; k++)
        w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
#endif
}

/* The first entry point in [lo, hi) (positions p with p + 4 <= len), or L3WP_NONE.  A header starts with 0xFF: 255 of
 * 256 positions are rejected on that byte, sixteen positions per load where the address allows it. */
L3_HD uint32_t l3wp_guess(const uint8_t *buf, uint32_t len, uint32_t lo, uint32_t hi, uint32_t first)
{
    uint32_t p = lo;
    while (p < hi && p + 4 <= len) {
        if ((((uintptr_t)(buf + p)) & 15u) == 0 && hi - p >= 16 && len - p >= 19) { /* 16 positions, all inside both bounds */
            uint32_t w[4];
            l3wp_load16(buf + p, w);
            for (int k = 0; k < 4; k++) {
                /* bit 8 j + 7 set iff byte j of w[k] is 0xFF (exact: no carries between the bytes) */
                const uint32_t v = ~w[k];
                uint32_t ff = ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v | 0x7f7f7f7fu);
                while (ff) {
                    const uint32_t q = p + 4u * (uint32_t)k + (l3wp_ctz(ff) >> 3);
                    if (l3wp_entry_at(buf, len, q, first)) return q;
                    ff &= ff - 1;
                }
            }
            p += 16;
            continue;
        }
        if (buf[p] == 0xFF && l3wp_entry_at(buf, len, p, first)) return p;
        p++;
    }
    return L3WP_NONE;
}

/* phase 1: segment t of [pf, len) -- guess the entry, walk to the segment's end, records to `sp` (this segment's block) */
L3_HD void l3wp_segment(const uint8_t *buf, uint32_t len, uint32_t pf, uint32_t seg, uint32_t nseg, uint32_t t,
                        uint32_t first, int streaming, uint32_t stream, L3FrameRec *sp, L3WalkSeg *out)
{
    const uint32_t lo = pf + t * seg, hi = (t + 1 < nseg) ? lo + seg : L3WP_NONE;
    const uint32_t g = t ? l3wp_guess(buf, len, lo, hi, first) : lo;
    out->start = g;
    out->exit = g;
    out->n = out->payload = 0;
    if (g != L3WP_NONE) out->exit = l3wp_span(buf, len, g, hi, first, streaming, stream, sp, &out->n, &out->payload);
}

/* the guess of segment t is consistent with the chain arriving at `want` (the exit of its predecessor, pf for t = 0) */
L3_HD int l3wp_chained_to(const L3WalkSeg *e, uint32_t want)
{
    if (want == L3WP_END) return e->n == 0; /* the chain has ended: later segments are empty */
    return e->start == want;
}
/* phase 2's parallel pre-check */
L3_HD int l3wp_chained(const L3WalkSeg *sg, uint32_t t, uint32_t pf) { return l3wp_chained_to(&sg[t], t ? sg[t - 1].exit : pf); }

/* phase 2 (one thread, only when some guess was wrong): segment t0 is the first one that does not chain up (all before
 * it do).  Follow the chain from there: a segment the chain enters elsewhere than guessed is walked again from the
 * true position, a segment it does not enter (a long frame spans it, or the chain has ended) is voided; every
 * segment's `exit` becomes the chain's position behind it.  Stops at the first segment that chains up with what was
 * repaired -- from there on the pre-check's verdicts hold again -- and returns it (nseg: none); with `to_end` it goes
 * through all remaining segments instead.  `sparse` = the stream's blocks of `cap` records. */
L3_HD uint32_t l3wp_repair_run(const uint8_t *buf, uint32_t len, uint32_t pf, uint32_t seg, uint32_t nseg, uint32_t first,
                               int streaming, uint32_t stream, L3FrameRec *sparse, uint32_t cap, L3WalkSeg *sg, uint32_t t0,
                               int to_end)
{
    uint32_t cur = t0 ? sg[t0 - 1].exit : pf;
    for (uint32_t t = t0; t < nseg; t++) {
        const uint32_t lo = pf + t * seg, hi = (t + 1 < nseg) ? lo + seg : L3WP_NONE;
        L3WalkSeg e = sg[t];
        if (t > t0 && l3wp_chained_to(&e, cur) && (cur == L3WP_END || cur < hi)) {
            if (!to_end) return t;
            if (cur != L3WP_END) cur = e.exit;
            continue;
        }
        if (cur == L3WP_END || cur >= hi) { /* the chain does not enter this segment */
            e.start = L3WP_NONE;
            e.exit = cur;
            e.n = e.payload = 0;
            sg[t] = e;
            continue;
        }
        e.start = cur;
        e.exit = l3wp_span(buf, len, cur, hi, first, streaming, stream, sparse + (size_t)t * cap, &e.n, &e.payload);
        sg[t] = e;
        cur = e.exit;
    }
    return nseg;
}

#endif
