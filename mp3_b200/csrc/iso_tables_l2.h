/* iso_tables_l2.h -- ISO/IEC 11172-3 Annex B Layer II tables (3-B.2a..d bit allocation, 3-B.4 classes of
 * quantisation) and ISO/IEC 13818-3 Table B.1 (LSF allocation), in the compact row form
 *   row = { nbal, q(1), q(2), ..., q(2^nbal - 1) }   (allocation code 0 = no samples)
 * where q indexes the 17 quantisation classes below.  There is no copy of the standard on the build box:
 * the rows were read out of libavcodec's .rodata (the same library the Layer III tables were derived
 * from, tools/derive_tables.py) and tests/test_tables_pin.py re-checks them against it; the oracle and
 * FFmpeg's mp2float then have to agree sample for sample on generated streams.  Shared by the oracle,
 * the generator and the CUDA path (tables only, like iso_tables.h).
 */
#ifndef MP3B_ISO_TABLES_L2_H
#define MP3B_ISO_TABLES_L2_H

#include <stdint.h>

#if defined(__CUDACC__)
#define L2_HD __host__ __device__ __forceinline__
#else
#define L2_HD static inline
#endif

/* bitrate_index -> kbit/s, Layer II.  [0] = MPEG-1, [1] = MPEG-2 LSF / 2.5. */
static const uint16_t l2_bitrate_kbps[2][16] = {
    {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 0},
    {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0},
};

/* bitrate_index -> kbit/s, Layer I.  [0] = MPEG-1, [1] = MPEG-2 LSF / 2.5.  A Layer I frame is 384 samples:
 * (12 * bitrate / sample_rate + padding) slots of 4 bytes; 4-bit allocation per subband (0 = none, else
 * allocation + 1 bits per sample, 15 is forbidden), one 6-bit scalefactor, 12 samples. */
static const uint16_t l1_bitrate_kbps[2][16] = {
    {0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448, 0},
    {0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256, 0},
};

/* Table 3-B.4: number of steps of the 17 quantisation classes, and bits per codeword: a negative value
 * is a grouped class (3, 5, 9 steps): ONE codeword of that many bits carries three consecutive samples,
 * least significant digit (base `steps`) first. */
static const int32_t l2_quant_steps[17] = {3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};
static const int8_t l2_quant_bits[17] = {-5, -7, 3, -10, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

/* the distinct allocation rows */
#define L2_ROW_A {4, 0, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}
#define L2_ROW_B {4, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16}
#define L2_ROW_C {3, 0, 1, 2, 3, 4, 5, 16}
#define L2_ROW_D {2, 0, 1, 16}
#define L2_ROW_E {4, 0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}
#define L2_ROW_F {3, 0, 1, 3, 4, 5, 6, 7}
#define L2_ROW_G {4, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14}
#define L2_ROW_H {2, 0, 1, 3}
static const uint8_t l2_rows[8][16] = {L2_ROW_A, L2_ROW_B, L2_ROW_C, L2_ROW_D, L2_ROW_E, L2_ROW_F, L2_ROW_G, L2_ROW_H};

/* allocation table t: subband -> row id (index into l2_rows), for sb < l2_sblimit[t].
 * t = 0: 3-B.2a (27 subbands), 1: 3-B.2b (30), 2: 3-B.2c (8), 3: 3-B.2d (12), 4: LSF (30). */
static const uint8_t l2_sblimit[5] = {27, 30, 8, 12, 30};
static const uint8_t l2_row_of_sb[5][30] = {
    {0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3},
    {0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3},
    {4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7},
};

/* 11172-3 2.4.2.3 / 2.4.3.3: which allocation table a frame uses (sample rate in Hz, total bitrate in
 * kbit/s, channels); LSF and MPEG-2.5 always use table 4. */
L2_HD int l2_sblimit_of(int t) { return t == 0 ? 27 : (t == 2 ? 8 : (t == 3 ? 12 : 30)); }

L2_HD int l2_select_table(int lsf, int sample_rate, int kbps, int nch)
{
    if (lsf) return 4;
    const int per_ch = kbps / nch;
    if ((sample_rate == 48000 && per_ch >= 56) || (per_ch >= 56 && per_ch <= 80)) return 0;
    if (sample_rate != 48000 && per_ch >= 96) return 1;
    if (sample_rate != 32000 && per_ch <= 48) return 2;
    return 3;
}

#endif
