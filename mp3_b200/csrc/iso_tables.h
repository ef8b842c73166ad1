/* ISO/IEC 11172-3 (MPEG-1) and 13818-3 (MPEG-2 LSF) Layer III constant tables.
 *
 * Plain C, header-only, shared by the host indexer, the CUDA table builder, the bitstream
 * generator and the CPU oracle.  Nothing here is derived from the reference repository
 * (lxm0851/mp3 ships no code: /root/reference/README.md:1-84); the values are the standard's,
 * as listed in SURVEY.md Appendix A, and are pinned by tests/test_tables_pin.py against the
 * independent libavcodec binary on the box.
 *
 * Sample-rate row order used everywhere: 0=44100 1=48000 2=32000 (MPEG-1),
 *                                         3=22050 4=24000 5=16000 (MPEG-2 LSF).
 */
#ifndef MP3B_ISO_TABLES_H
#define MP3B_ISO_TABLES_H

#include <stdint.h>
#include "iso_tables_gen.h"

#define L3_GRANULE 576
#define L3_SBLIMIT 32
#define L3_SSLIMIT 18

/* 11172-3 2.4.2.3 bitrate_index -> kbit/s, Layer III.  [0]=MPEG-1, [1]=MPEG-2 LSF. */
static const uint16_t l3_bitrate_kbps[2][16] = {
    {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0},
    {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0},
};

/* sampling_frequency index -> Hz, by row (see header comment). */
/* rows 0..2 MPEG-1, 3..5 MPEG-2 LSF, 6..8 MPEG-2.5 (the unofficial low-rate extension: version bits 00) */
static const uint32_t l3_sample_rate[9] = {44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000};

/* Annex B Table 3-B.8: scalefactor band edges, long blocks (23 edges = 22 bands). */
static const uint16_t l3_sfb_long[9][23] = {
    {0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576},
    {0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576},
    {0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576},
    {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
    {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 114, 136, 162, 194, 232, 278, 332, 394, 464, 540, 576},
    {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
    /* MPEG-2.5: 11.025 and 12 kHz use the 16 kHz partition, 8 kHz has its own */
    {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
    {0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
    {0, 12, 24, 36, 48, 60, 72, 88, 108, 132, 160, 192, 232, 280, 336, 400, 476, 566, 568, 570, 572, 574, 576},
};

/* Table 3-B.8: scalefactor band edges, short blocks (14 edges = 13 bands, per window). */
static const uint16_t l3_sfb_short[9][14] = {
    {0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192},
    {0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192},
    {0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192},
    {0, 4, 8, 12, 18, 24, 32, 42, 56, 74, 100, 132, 174, 192},
    {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 136, 180, 192},
    {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 134, 174, 192},
    {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 134, 174, 192},
    {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 134, 174, 192},
    {0, 8, 16, 24, 36, 52, 72, 96, 124, 160, 162, 164, 166, 192},
};

/* 2.4.2.7 scalefac_compress -> (slen1, slen2), MPEG-1. */
static const uint8_t l3_slen[2][16] = {
    {0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4},
    {0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3},
};

/* Table 3-B.6 preemphasis (pretab), long-block sfb 0..21. */
static const uint8_t l3_pretab[22] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0};

/* 13818-3 2.4.3.2: number of scalefactor bands per slen partition,
 * [table 0..5][0=long 1=short 2=mixed][partition 0..3]. */
static const uint8_t l3_lsf_nsfb[6][3][4] = {
    {{6, 5, 5, 5}, {9, 9, 9, 9}, {6, 9, 9, 9}},
    {{6, 5, 7, 3}, {9, 9, 12, 6}, {6, 9, 12, 6}},
    {{11, 10, 0, 0}, {18, 18, 0, 0}, {15, 18, 0, 0}},
    {{7, 7, 7, 0}, {12, 12, 12, 0}, {6, 15, 12, 0}},
    {{6, 6, 6, 3}, {12, 9, 9, 6}, {6, 12, 9, 6}},
    {{8, 8, 5, 0}, {15, 12, 9, 0}, {6, 18, 9, 0}},
};

/* Table 3-B.7: table_select 0..31 -> code book id (0 = all-zero book; 4 and 14 do not exist)
 * and number of linbits. */
static const uint8_t l3_book_of_table[32] = {
    0, 1, 2, 3, 0, 5, 6, 7, 8, 9, 10, 11, 12, 13, 0, 15,
    16, 16, 16, 16, 16, 16, 16, 16, 24, 24, 24, 24, 24, 24, 24, 24,
};
static const uint8_t l3_linbits_of_table[32] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13,
};

/* count1 quadruple books A (table 32) and B (table 33); index = v*8 + w*4 + x*2 + y. */
static const uint8_t l3_quad_hlen[2][16] = {
    {1, 4, 4, 5, 4, 6, 5, 6, 4, 5, 5, 6, 5, 6, 6, 6},
    {4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4},
};
static const uint8_t l3_quad_hcod[2][16] = {
    {1, 5, 4, 5, 6, 5, 4, 4, 7, 3, 6, 0, 7, 2, 3, 1},
    {15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0},
};

/* Table 3-B.9 alias-reduction coefficients c_i. */
static const double l3_alias_ci[8] = {-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037};

/* Accessor for a big_values code book: returns dimension (0 for the empty books). */
static inline int l3_book(int book, const uint8_t **hlen, const uint32_t **hcod)
{
    switch (book) {
#define L3_BOOK_CASE(id, dim) case id: *hlen = l3_hlen_##id; *hcod = l3_hcod_##id; return dim;
    L3_BOOK_CASE(1, 2) L3_BOOK_CASE(2, 3) L3_BOOK_CASE(3, 3) L3_BOOK_CASE(5, 4) L3_BOOK_CASE(6, 4)
    L3_BOOK_CASE(7, 6) L3_BOOK_CASE(8, 6) L3_BOOK_CASE(9, 6) L3_BOOK_CASE(10, 8) L3_BOOK_CASE(11, 8)
    L3_BOOK_CASE(12, 8) L3_BOOK_CASE(13, 16) L3_BOOK_CASE(15, 16) L3_BOOK_CASE(16, 16)
    L3_BOOK_CASE(24, 16)
#undef L3_BOOK_CASE
    default: *hlen = 0; *hcod = 0; return 0;
    }
}

/* Full 512-tap synthesis window D[i] (Table 3-B.3) from the q16 half table. */
static inline double l3_dwin(int i)
{
    if (i <= 256) return l3_dwin_q16[i] / 65536.0;
    int j = 512 - i;
    double v = l3_dwin_q16[j] / 65536.0;
    return (j & 63) ? -v : v;
}

#endif
