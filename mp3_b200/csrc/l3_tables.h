/* l3_tables.h -- decode-side lookup tables derived from the ISO tables, built once on the host
 * (tables_build.cpp) and uploaded to the device by the context. */
#ifndef MP3B_L3_TABLES_H
#define MP3B_L3_TABLES_H

#include <stdint.h>

#define L3_HUFF_ROOT_BITS 9
#define L3_HUFF_SUB_BITS 10 /* one second level covers every code longer than the root */
#define L3_HUFF_LUT_MAX 12288 /* entries; actual size is computed at build time */

/* Huffman LUT entry (uint16):
 *   leaf : bit15 = 0, bits 11..8 = code bits consumed at this level, bits 7..0 = (x << 4) | y
 *   link : bit15 = 1, bits 14..11 = width of the next level, bits 10..0 = offset of the next
 *          level from the book's base (the current level's full width is consumed). */
typedef struct L3HuffInfo {
    uint16_t base[32];   /* LUT offset of the root level, per table_select */
    uint8_t root[32];    /* root width in bits (0 = empty book) */
    uint8_t linbits[32];
} L3HuffInfo;

/* Band layouts: [sample-rate row 0..8][0 = long, 1 = short, 2 = mixed]. */
typedef struct L3BandTables {
    uint8_t line2band[9][3][576];
    uint16_t dst[9][3][576]; /* where bitstream line i goes after the short-block reorder (identity for long bands) */
    uint32_t lmap[9][3][576]; /* band | dst << 8: band and reordered position of a line, one word (fused back end) */
    uint16_t start[9][3][40];
    uint8_t width[9][3][40];
    int8_t win[9][3][40];    /* window 0..2 of a short band, -1 for a long band */
    uint8_t sfb[9][3][40];
    uint8_t nbands[9][3];
    uint8_t nlong[9][3];
} L3BandTables;

typedef struct L3HostTables {
    uint16_t huff_lut[L3_HUFF_LUT_MAX];
    uint32_t huff_lut_len;
    L3HuffInfo huff;
    uint8_t quad_a[64];      /* count1 book A: 6-bit peek -> (len << 4) | vwxy */
    L3BandTables bands;
    float pow43[8208];       /* |is|^(4/3), is = 0..8206 */
} L3HostTables;

#ifdef __cplusplus
extern "C" {
#endif
void l3_build_host_tables(L3HostTables *t);
#ifdef __cplusplus
}
#endif
#endif
