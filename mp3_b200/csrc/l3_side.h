/* l3_side.h -- side-information parse (a2) and bit-reservoir resolution (a3): one frame in,
 * ngr*nch unit descriptors out.  Compiled for both host (MP3B_INDEX_HOST) and device
 * (MP3B_INDEX_DEVICE) so that the two indexers cannot drift apart; the independent check is
 * the oracle (tests compare descriptors-derived outputs bit-exactly).
 * Spec: ISO/IEC 11172-3 2.4.1.7 / 2.4.2.7, ISO/IEC 13818-3 2.4.1.7. */
#ifndef MP3B_L3_SIDE_H
#define MP3B_L3_SIDE_H

#include "l3_defs.h"

typedef struct L3SideBits {
    const uint8_t *p;
    uint32_t pos;
} L3SideBits;

L3_HD uint32_t l3_sb_get(L3SideBits *b, int n)
{
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        v = (v << 1) | ((b->p[b->pos >> 3] >> (7 - (b->pos & 7))) & 1u);
        b->pos++;
    }
    return v;
}

/* sfb_long: the 23 long-block band edges of this sample rate (for region boundaries).
 * frame_index_in_stream decides L3P_FIRST.  units_out receives ngr*nch descriptors in
 * [gr][ch] order.  Returns 1 if the frame's main data is available, 0 if it is concealed. */
L3_HD int l3_parse_side(const uint8_t *frame, const L3Hdr *h, const uint16_t *sfb_long,
                        uint64_t payload_base, uint32_t payload_off, uint32_t stream,
                        int first_frame, L3UnitDesc *units_out)
{
    L3SideBits b;
    b.p = frame + 4 + (h->crc ? 2 : 0);
    b.pos = 0;
    int nch = h->nch, ngr = h->ngr;
    uint32_t mdb, scfsi[2] = {0, 0};
    if (!h->lsf) {
        mdb = l3_sb_get(&b, 9);
        l3_sb_get(&b, nch == 1 ? 5 : 3);
        for (int ch = 0; ch < nch; ch++)
            for (int k = 0; k < 4; k++) scfsi[ch] |= l3_sb_get(&b, 1) << k;
    } else {
        mdb = l3_sb_get(&b, 8);
        l3_sb_get(&b, nch == 1 ? 1 : 2);
    }
    int valid = mdb <= payload_off;
    uint64_t bit = valid ? (payload_base + payload_off - mdb) * 8ull : 0ull;
    uint8_t hdrbits = (uint8_t)((h->lsf ? L3H_LSF : 0) | (h->sr_row << L3H_SR_SHIFT) | (nch == 2 ? L3H_STEREO : 0));
    if (h->mode == 1) hdrbits |= (uint8_t)(((h->mode_ext & 2) ? L3H_MS : 0) | ((h->mode_ext & 1) ? L3H_IS : 0));
    for (int gr = 0; gr < ngr; gr++)
        for (int ch = 0; ch < nch; ch++) {
            L3UnitDesc d;
            uint32_t p23 = l3_sb_get(&b, 12), bv = l3_sb_get(&b, 9), gg = l3_sb_get(&b, 8);
            uint32_t sfc = l3_sb_get(&b, h->lsf ? 9 : 4), ws = l3_sb_get(&b, 1);
            uint32_t bt = 0, mixed = 0, t0, t1, t2 = 0, r0c, r1c, sbg0 = 0, sbg1 = 0, sbg2 = 0;
            if (ws) {
                bt = l3_sb_get(&b, 2);
                mixed = l3_sb_get(&b, 1);
                t0 = l3_sb_get(&b, 5);
                t1 = l3_sb_get(&b, 5);
                sbg0 = l3_sb_get(&b, 3);
                sbg1 = l3_sb_get(&b, 3);
                sbg2 = l3_sb_get(&b, 3);
                r0c = r1c = 0;
            } else {
                t0 = l3_sb_get(&b, 5);
                t1 = l3_sb_get(&b, 5);
                t2 = l3_sb_get(&b, 5);
                r0c = l3_sb_get(&b, 4);
                r1c = l3_sb_get(&b, 3);
            }
            uint32_t preflag = h->lsf ? 0 : l3_sb_get(&b, 1);
            uint32_t sfscale = l3_sb_get(&b, 1), c1tab = l3_sb_get(&b, 1);
            if (h->lsf && !((hdrbits & L3H_IS) && ch == 1) && sfc >= 500) preflag = 1;
            if (bv > 288) bv = 288;
            uint32_t bv2 = bv * 2, r1, r2;
            if (ws) {
                r1 = (bt == 2 || !h->lsf) ? 36 : 54;
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
                              ((first_frame && gr == 0) ? L3P_FIRST : 0));
            d.stream = stream;
            units_out[gr * nch + ch] = d;
            if (valid) bit += p23;
        }
    return valid;
}

#endif
