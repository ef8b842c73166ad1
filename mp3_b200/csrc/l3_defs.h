/* l3_defs.h -- record layouts and header arithmetic shared by the host indexer and the kernels.
 *
 * Stage ids (a1..a11) refer to SURVEY.md section 8(a).  The reference repository has no code for
 * any of them (/root/reference/README.md:1-84); the behaviour implemented is ISO/IEC 11172-3 /
 * 13818-3 as restated by oracle/l3_oracle.c.
 */
#ifndef MP3B_L3_DEFS_H
#define MP3B_L3_DEFS_H

#include <stdint.h>

#if defined(__CUDACC__)
#define L3_HD __host__ __device__ __forceinline__
#else
#define L3_HD static inline
#endif

/* One MPEG audio frame found by the indexer (a1, a3). */
typedef struct L3FrameRec {
    uint32_t rel_off;     /* byte offset of the header inside its stream */
    uint32_t payload_off; /* main-data bytes of this stream that precede this frame */
    uint32_t hdr;         /* the four header bytes, big endian */
    uint32_t stream;
} L3FrameRec;

/* One input stream of a batch. */
typedef struct L3StreamRec {
    uint64_t raw_off;      /* offset of the stream in the raw byte arena */
    uint64_t payload_base; /* offset of its compacted main data in the main-data arena */
    uint32_t raw_len;
    uint32_t first_off;    /* offset of the first frame (after ID3v2 / junk) */
    uint32_t first_hdr;    /* header of the first frame, 0 if none */
    uint32_t nframes;
    uint32_t payload_len;
    uint32_t frame_base;   /* global index of its first frame */
    uint32_t gran_base;    /* global index of its first granule */
    uint32_t unit_base;    /* global index of its first unit (granule-channel) */
    uint32_t end_off;      /* out: first byte not consumed by a complete frame (streaming) */
    uint32_t skip_frames;  /* in: leading frames that only re-derive state (already output earlier) */
    uint32_t flags;        /* in: L3S_* */
    /* out: the encoder's tag frame (Xing / Info / VBRI), parsed from the stream's first frame */
    uint32_t tag_kind;     /* L3T_*; bit 8 (L3T_LAME): a LAME extension with delay / padding follows */
    uint32_t tag_frames;   /* frame count field, 0 if absent */
    uint32_t tag_bytes;    /* byte count field, 0 if absent */
    uint32_t tag_delay_pad; /* encoder delay << 16 | padding (12 bits each; VBRI: delay only) */
    uint32_t sb_shift;     /* Layer I / II: unit_base minus the stream's base in the dense subband-sample buffer */
} L3StreamRec;

#define L3T_NONE 0u
#define L3T_XING 1u
#define L3T_INFO 2u
#define L3T_VBRI 3u
#define L3T_KIND_MASK 0xffu
#define L3T_LAME 0x100u

#define L3S_STREAMING 1u   /* more bytes may follow: stop at a valid header whose frame is not complete yet
                              instead of searching for a sync inside it; first_hdr may be preset */

/* A stream's main data in the arena is followed by zero bytes up to the next 16-byte boundary, at least this many. */
#define L3_PAYLOAD_PAD 8u

/* One granule-channel ("unit": 576 spectral lines), from the side info (a2, a3). 32 bytes. */
typedef struct __attribute__((aligned(16))) L3UnitDesc {
    uint64_t bit_off;     /* absolute bit offset of part2 in the main-data arena */
    uint16_t p23len;      /* part2_3_length */
    uint16_t big_values;  /* clamped to 288 */
    uint16_t r1, r2;      /* region 1 / 2 start lines, clamped to 2*big_values */
    uint16_t sfc;         /* scalefac_compress */
    uint8_t global_gain;
    uint8_t tsel[3];
    uint8_t sbg[3];       /* subblock_gain */
    uint8_t flags;        /* L3F_* */
    uint8_t hdr;          /* L3H_* */
    uint8_t pos;          /* L3P_* */
    uint32_t stream;
} L3UnitDesc;

#define L3F_BT_MASK 0x03      /* block_type */
#define L3F_MIXED 0x04
#define L3F_PREFLAG 0x08      /* effective preflag (MPEG-1 bit, or LSF scalefac_compress >= 500) */
#define L3F_SFSCALE 0x10
#define L3F_C1TAB 0x20
#define L3F_VALID 0x40        /* main data available (main_data_begin within the reservoir) */
#define L3F_WS 0x80           /* window_switching_flag */

#define L3H_LSF 0x01
#define L3H_SR_SHIFT 1        /* bits 1..4: sample-rate row 0..8 */
#define L3H_SR_MASK 15
#define L3H_STEREO 0x20       /* two channels */
#define L3H_MS 0x40           /* joint stereo with mode_ext MS bit */
#define L3H_IS 0x80           /* joint stereo with mode_ext intensity bit */

#define L3P_GR 0x01
#define L3P_CH 0x02
#define L3P_SCFSI_SHIFT 2     /* bits 2..5: scfsi[0..3], bit (2+k) = group k */
#define L3P_FIRST 0x40        /* first granule of its stream: no overlap / synthesis history */

/* gran_unit0[g]: unit index of channel 0 of global granule g, with flags in the top bits. */
#define L3G_UNIT_MASK 0x3FFFFFFFu
#define L3G_STEREO 0x40000000u
#define L3G_FIRST 0x80000000u

typedef struct L3Hdr {
    int lsf, sr_row, nch, mode, mode_ext, crc, frame_len, side_len, ngr;
    int layer; /* 3, 2 or 1: Layer I / II frames carry no side info / main data (side_len = 0) */
    int kbps;
    int spf;   /* samples per frame: 1152 / 576 (= ngr granules), or 384 for Layer I (ngr = 0: its granules are
                  18-slot pieces of the stream's slot sequence, see l3_stream_granules) */
} L3Hdr;

L3_HD int l3_kbps(int lsf, int idx)
{
    /* kbit/s divided by 8, one byte per bitrate_index 1..14 */
    const unsigned long long m1a = 0x0E0C0A0807060504ull; /* 4 5 6 7 8 10 12 14 */
    const unsigned long long m2a = 0x0807060504030201ull; /* 1 .. 8 */
    int i = idx - 1;
    unsigned v;
    if (!lsf) {
        const unsigned long long hi = 0x000028201C181410ull; /* 16 20 24 28 32 40 */
        v = i < 8 ? (unsigned)(m1a >> (8 * i)) & 0xff : (unsigned)(hi >> (8 * (i - 8))) & 0xff;
    } else {
        const unsigned long long hi = 0x00001412100E0C0Aull; /* 10 12 14 16 18 20 */
        v = i < 8 ? (unsigned)(m2a >> (8 * i)) & 0xff : (unsigned)(hi >> (8 * (i - 8))) & 0xff;
    }
    return (int)v * 8;
}

L3_HD int l3_sr_hz(int row)
{
    return row == 0 ? 44100 : row == 1 ? 48000 : row == 2 ? 32000 : row == 3 ? 22050 : row == 4 ? 24000
         : row == 5 ? 16000 : row == 6 ? 11025 : row == 7 ? 12000 : 8000;
}

/* a1: parse the 32 header bits (big-endian word).  Returns 0 if this is not a decodable
 * MPEG-1 / MPEG-2 LSF / MPEG-2.5 Layer III header (free format is rejected).  MPEG-2.5 (version
 * bits 00) is MPEG-2 LSF syntax at half the sample rates: rows 6..8. */
L3_HD int l3_parse_hdr(uint32_t w, L3Hdr *h)
{
    if ((w & 0xFFE00000u) != 0xFFE00000u) return 0;
    int ver = (w >> 19) & 3, layer = (w >> 17) & 3;
    if (layer == 0 || ver == 1) return 0; /* 01 = Layer III, 10 = Layer II, 11 = Layer I */
    int bri = (w >> 12) & 15, sri = (w >> 10) & 3;
    if (bri == 0 || bri == 15 || sri == 3) return 0;
    h->layer = 4 - layer;
    h->lsf = ver != 3;
    h->crc = !((w >> 16) & 1);
    h->mode = (w >> 6) & 3;
    h->mode_ext = (w >> 4) & 3;
    h->nch = h->mode == 3 ? 1 : 2;
    h->sr_row = sri + (ver == 3 ? 0 : ver == 2 ? 3 : 6);
    h->ngr = h->lsf ? 1 : 2;
    int pad = (w >> 9) & 1;
    if (h->layer == 1) { /* 384 samples per frame, slots of four bytes; bitrates 32 .. 448 (MPEG-1) in steps of 32 */
        /* LSF: kbit/s / 8 for bitrate_index 1..8 and 9..14: 32 48 56 64 80 96 112 128 | 144 160 176 192 224 256 */
        const unsigned long long lo = 0x100E0C0A08070604ull, hi = 0x0000201C18161412ull;
        h->kbps = h->lsf ? 8 * (int)((bri <= 8 ? (lo >> (8 * (bri - 1))) : (hi >> (8 * (bri - 9)))) & 0xff) : 32 * bri;
        h->ngr = 0;
        h->spf = 384;
        h->frame_len = (12000 * h->kbps / l3_sr_hz(h->sr_row) + pad) * 4;
        h->side_len = 0;
        return 1;
    }
    if (h->layer == 2) { /* 1152 samples per frame at every rate; MPEG-1 has its own Layer II bitrate table */
        /* kbit/s / 8 for bitrate_index 1..8 and 9..14: 32 48 56 64 80 96 112 128 | 160 192 224 256 320 384 */
        const unsigned long long lo = 0x100E0C0A08070604ull, hi = 0x00003028201C1814ull;
        h->kbps = h->lsf ? l3_kbps(1, bri) : 8 * (int)((bri <= 8 ? (lo >> (8 * (bri - 1))) : (hi >> (8 * (bri - 9)))) & 0xff);
        h->ngr = 2;
        h->spf = 1152;
        h->frame_len = 144000 * h->kbps / l3_sr_hz(h->sr_row) + pad;
        h->side_len = 0;
        return 1;
    }
    h->kbps = l3_kbps(h->lsf, bri);
    h->spf = h->ngr * 576;
    h->frame_len = (h->lsf ? 72000 : 144000) * h->kbps / l3_sr_hz(h->sr_row) + pad;
    h->side_len = h->lsf ? (h->nch == 1 ? 9 : 17) : (h->nch == 1 ? 17 : 32);
    return 1;
}

/* Granules (576 samples = 18 slots per channel) that n frames of this kind occupy. */
L3_HD uint64_t l3_stream_granules(const L3Hdr *h, uint64_t nframes)
{
    return h->layer == 1 ? (nframes * 12 + 17) / 18 : nframes * (uint64_t)h->ngr;
}

/* Two headers belong to the same stream when version, sample rate and channel count agree. */
L3_HD int l3_same_stream(uint32_t a, uint32_t b)
{
    /* version+layer bits 19..17, sample-rate bits 11..10; mono-ness = (mode == 3) */
    if (((a ^ b) & 0x001E0C00u) != 0) return 0;
    return (((a >> 6) & 3) == 3) == (((b >> 6) & 3) == 3);
}

/* Length of an ID3v2 tag at the start of a buffer, or 0. */
L3_HD uint32_t l3_id3v2_len(const uint8_t *b, uint32_t len)
{
    if (len >= 10 && b[0] == 'I' && b[1] == 'D' && b[2] == '3' && !((b[6] | b[7] | b[8] | b[9]) & 0x80)) {
        uint32_t n = ((uint32_t)b[6] << 21) | ((uint32_t)b[7] << 14) | ((uint32_t)b[8] << 7) | b[9];
        n += 10 + ((b[5] & 0x10) ? 10 : 0);
        if (n <= len) return n;
    }
    return 0;
}

L3_HD uint32_t l3_load_be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

/* CRC-16 of the audio frame (11172-3 2.4.3.1): generator x^16 + x^15 + x^2 + 1, initial value 0xFFFF,
 * most significant bit first, over the last two header bytes and the side info. */
L3_HD uint32_t l3_crc16(uint32_t crc, const uint8_t *p, uint32_t nbytes)
{
    for (uint32_t i = 0; i < nbytes; i++) {
        crc ^= (uint32_t)p[i] << 8;
        for (int k = 0; k < 8; k++) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x8005u) & 0xffffu : (crc << 1) & 0xffffu;
    }
    return crc;
}

/* The tag frame an encoder may put first (the container step just before the hot path, SURVEY.md
 * 8(f) rank 1 / 3).  "Xing" (VBR) or "Info" (CBR) sits right after the side info of an otherwise
 * empty Layer III frame: flags (bit 0 frames, 1 bytes, 2 TOC[100], 3 quality), then optionally the
 * LAME extension (9-byte version string, ..., 12-bit encoder delay and 12-bit padding at +21).
 * "VBRI" (Fraunhofer) sits 32 bytes after the header: version, delay, quality, bytes, frames.
 * `f` points at the frame's header, `flen` is the frame length.  Returns L3T_* | L3T_LAME. */
L3_HD uint32_t l3_parse_tag(const uint8_t *f, uint32_t flen, const L3Hdr *h, uint32_t *frames, uint32_t *bytes,
                            uint32_t *delay_pad)
{
    *frames = *bytes = *delay_pad = 0;
    for (int crc = 0; crc <= (h->crc ? 1 : 0); crc++) { /* writers differ on whether the CRC word shifts the tag */
        uint32_t o = 4u + (uint32_t)h->side_len + 2u * (uint32_t)crc;
        if (o + 8 > flen) break;
        const uint32_t id = l3_load_be32(f + o);
        if (id != 0x58696E67u /* Xing */ && id != 0x496E666Fu /* Info */) continue;
        const uint32_t fl = l3_load_be32(f + o + 4);
        uint32_t q = o + 8;
        if ((fl & 1) && q + 4 <= flen) { *frames = l3_load_be32(f + q); q += 4; }
        if ((fl & 2) && q + 4 <= flen) { *bytes = l3_load_be32(f + q); q += 4; }
        if (fl & 4) q += 100;
        if (fl & 8) q += 4;
        uint32_t kind = id == 0x58696E67u ? L3T_XING : L3T_INFO;
        /* LAME extension: a version string of printable characters ("LAME3.99r", "Lavc58.13", ...) */
        if (q + 24 <= flen && f[q] >= 0x20 && f[q] < 0x7f && f[q + 1] >= 0x20 && f[q + 1] < 0x7f) {
            const uint32_t d = ((uint32_t)f[q + 21] << 4) | (f[q + 22] >> 4);
            const uint32_t p = (((uint32_t)f[q + 22] & 15u) << 8) | f[q + 23];
            *delay_pad = (d << 16) | p;
            kind |= L3T_LAME;
        }
        return kind;
    }
    if (4u + 32u + 18u <= flen && l3_load_be32(f + 36) == 0x56425249u /* VBRI */) {
        *delay_pad = (((uint32_t)f[36 + 6] << 8) | f[36 + 7]) << 16;
        *bytes = l3_load_be32(f + 36 + 10);
        *frames = l3_load_be32(f + 36 + 14);
        return L3T_VBRI;
    }
    return L3T_NONE;
}

/* One step of the frame walk shared by the host and device indexers (the scan policy of
 * oracle/l3_oracle.c: l3o_decode): at byte p, is there a frame of this stream that fits?
 * `first` is the stream's first header (0 while none has been found). */
/* returns 1: a complete frame; 2: a valid header of this stream whose frame runs past the buffer; 0: none;
 * 3: a complete first-frame candidate whose confirming next header lies beyond the buffer -- a one-shot decode
 * accepts it (the buffer is the whole stream), an incremental one waits for the bytes (L3S_STREAMING) */
L3_HD int l3_frame_at(const uint8_t *buf, uint32_t len, uint32_t p, uint32_t first, L3Hdr *h, uint32_t *word)
{
    if (p + 4 > len) return 0;
    uint32_t w = l3_load_be32(buf + p);
    /* (the cheap test first: main data is full of 0xFF 0xFx pairs -- five per frame in the generator's streams --,
     * and the time-parallel walk tries every one of them as an entry point) */
    if (first && !l3_same_stream(w, first)) return 0;
    if (!l3_parse_hdr(w, h)) return 0;
    if (h->frame_len < 4 + (h->crc ? 2 : 0) + h->side_len) return 0;
    if (p + (uint32_t)h->frame_len > len) return 2;
    *word = w;
    if (!first) {
        /* the stream's first frame must be followed by a header of the same stream (when the bytes are
         * there to check): one valid-looking word inside junk or a tag does not start a stream */
        if (p + (uint32_t)h->frame_len + 4 > len) return 3;
        const uint32_t w2 = l3_load_be32(buf + p + (uint32_t)h->frame_len);
        L3Hdr h2;
        if (!l3_parse_hdr(w2, &h2) || !l3_same_stream(w, w2)) return 0;
    }
    return 1;
}

#endif
