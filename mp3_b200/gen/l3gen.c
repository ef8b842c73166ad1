/* l3gen.c -- synthetic, spec-valid MPEG-1/2 Layer III bitstream generator.
 *
 * There is no MP3 encoder on the box and no corpus, so every test and
 * benchmark input is produced here: random but LEGAL frames -- exact part2_3_length, big_values
 * <= 288, region boundaries on scalefactor-band edges, code books 4/14 never selected,
 * main_data_begin never larger than the bytes actually left in the bit reservoir, legal window
 * switching sequences, MS / intensity joint stereo, MPEG-2 LSF scalefactor partitions, VBR.
 *
 * It shares only the ISO tables (../csrc/iso_tables.h) with the decoders: nothing of the decode
 * logic is reused, so a generator bug and a decoder bug cannot cancel silently -- and the FFmpeg
 * differential tests decode these streams with a third, independent implementation.
 *
 * The reference repository has no generator or sample audio to mirror
 * (/root/reference/README.md:1-84).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../csrc/iso_tables.h"
#include "../csrc/iso_tables_l2.h"

typedef struct {
    uint64_t seed;
    int32_t sample_rate;   /* 44100 48000 32000 | 22050 24000 16000 | 11025 12000 8000 (MPEG-2.5) */
    int32_t mode;          /* 0 stereo, 1 joint stereo, 2 dual channel, 3 mono */
    int32_t bitrate_kbps;  /* CBR rate; ignored when vbr_max_kbps > 0 */
    int32_t vbr_min_kbps, vbr_max_kbps;
    int32_t nframes;
    int32_t blocks;        /* 0 = long blocks only, 1 = random legal window switching */
    int32_t mixed_pct;     /* % of short granules with mixed_block_flag (MPEG-1 only) */
    int32_t reservoir;     /* 0 = main_data_begin always 0, 1 = use the bit reservoir */
    int32_t fill_lo_pct, fill_hi_pct; /* share of the available bits a frame uses */
    int32_t mode_ext_mask; /* joint stereo: bit v set => mode_ext value v may be drawn */
    int32_t crc;           /* 1 = protection on (CRC-16 word written) */
    int32_t lsf_avoid_illegal_ispos; /* LSF intensity: keep is_pos legal and <= 15 (what FFmpeg models) */
    int32_t level_lo_db, level_hi_db; /* target rms level range, dB below full scale */
    int32_t only_table;    /* >0: force this table_select everywhere */
    int32_t scfsi_pct;
    int32_t max_linbits_value; /* cap on escape magnitudes (<=8191+15); 0 = no cap */
    int32_t mixed_free;    /* 0 = mixed_block_flag is drawn once per run of short granules (what real
                              encoders do); 1 = drawn per granule, which also produces mixed -> pure
                              short transitions (legal to decode, but FFmpeg drops the long-window
                              tail of subbands 0-1 there, so differential tests keep this 0) */
    int32_t tag;           /* 0 = none; 1 = "Xing", 2 = "Info" tag frame first (frames, bytes, TOC, quality);
                              3 = "VBRI" (Fraunhofer) */
    int32_t tag_lame;      /* 1 = LAME extension after the Xing/Info fields (encoder delay / padding) */
    int32_t enc_delay, enc_padding; /* 0..4095 each */
    int32_t layer;         /* 0 / 3 = Layer III; 2 = Layer II (bitrate_kbps from the Layer II table, CBR;
                              mode, mode_ext_mask, crc, level_*_db, fill_*_pct apply) */
} l3gen_cfg;

/* ------------------------------------------------------------------ rng */
typedef struct { uint64_t s; } rng_t;
static uint64_t rnd64(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static int rndi(rng_t *r, int lo, int hi) /* inclusive */
{
    return lo + (int)(rnd64(r) % (uint64_t)(hi - lo + 1));
}
static double rndu(rng_t *r) { return (rnd64(r) >> 11) * (1.0 / 9007199254740992.0); }

/* ------------------------------------------------------------------ bit writer (into logical main-data stream) */
typedef struct {
    uint8_t *p;
    size_t cap_bits;
    size_t pos;
} bitw;
static void putbits(bitw *w, unsigned v, int n)
{
    for (int i = n - 1; i >= 0; i--) {
        if (w->pos < w->cap_bits && ((v >> i) & 1)) w->p[w->pos >> 3] |= (uint8_t)(0x80 >> (w->pos & 7));
        w->pos++;
    }
}

typedef struct {
    int part2_3_length, big_values, global_gain, scalefac_compress;
    int window_switching, block_type, mixed;
    int table_select[3], subblock_gain[3];
    int region0_count, region1_count;
    int preflag, scalefac_scale, count1table;
} gr_side;

typedef struct {
    int lsf, sr_row, nch, mode;
    const l3gen_cfg *cfg;
    rng_t rng;
    int bt_state[2];       /* previous block type per channel */
    int mixed_run[2];      /* mixed flag of the current run of short granules */
    uint8_t sf_gr0[2][40];
    int bt_gr0[2];
} gen_t;

static int pick_table(gen_t *g)
{
    if (g->cfg->only_table > 0) return g->cfg->only_table;
    for (;;) {
        int t = rndi(&g->rng, 0, 31);
        if (t == 4 || t == 14) continue;
        if (t == 0 && rndi(&g->rng, 0, 3)) continue; /* the empty book is legal but make it rarer */
        return t;
    }
}

static int next_block_type(gen_t *g, int ch)
{
    int prev = g->bt_state[ch], nx;
    if (!g->cfg->blocks) return 0;
    if (prev == 0 || prev == 3) nx = rndi(&g->rng, 0, 99) < 60 ? 0 : 1;
    else if (prev == 1) nx = 2;
    else nx = rndi(&g->rng, 0, 99) < 50 ? 2 : 3;
    g->bt_state[ch] = nx;
    return nx;
}

static int pair_bits(int tsel, int x, int y)
{
    int book = l3_book_of_table[tsel], lin = l3_linbits_of_table[tsel];
    const uint8_t *hl;
    const uint32_t *hc;
    int dim = l3_book(book, &hl, &hc);
    if (!dim) return 0;
    int ax = abs(x), ay = abs(y), cx = ax > 15 ? 15 : ax, cy = ay > 15 ? 15 : ay;
    if (!lin) { cx = ax; cy = ay; }
    int n = hl[cx * dim + cy];
    if (lin && cx == 15) n += lin;
    if (lin && cy == 15) n += lin;
    return n + (ax != 0) + (ay != 0);
}

static void put_pair(bitw *w, int tsel, int x, int y)
{
    int book = l3_book_of_table[tsel], lin = l3_linbits_of_table[tsel];
    const uint8_t *hl;
    const uint32_t *hc;
    int dim = l3_book(book, &hl, &hc);
    if (!dim) return;
    int ax = abs(x), ay = abs(y), cx = ax, cy = ay;
    if (lin) { if (cx > 15) cx = 15; if (cy > 15) cy = 15; }
    putbits(w, hc[cx * dim + cy], hl[cx * dim + cy]);
    if (lin && cx == 15) putbits(w, (unsigned)(ax - 15), lin);
    if (ax) putbits(w, x < 0, 1);
    if (lin && cy == 15) putbits(w, (unsigned)(ay - 15), lin);
    if (ay) putbits(w, y < 0, 1);
}

static int draw_value(gen_t *g, int tsel, double scale)
{
    int book = l3_book_of_table[tsel], lin = l3_linbits_of_table[tsel];
    const uint8_t *hl;
    const uint32_t *hc;
    int dim = l3_book(book, &hl, &hc);
    if (!dim) return 0;
    int maxv = dim - 1;
    if (lin) {
        maxv = 15 + (1 << lin) - 1;
        if (g->cfg->max_linbits_value > 0 && maxv > g->cfg->max_linbits_value) maxv = g->cfg->max_linbits_value;
    }
    int v;
    double u = rndu(&g->rng);
    if (u < 0.02) v = maxv;                       /* make sure the extremes are exercised */
    else if (u < 0.04 && lin) v = rndi(&g->rng, 15, maxv);
    else {
        v = (int)(-log(1.0 - rndu(&g->rng)) * scale);
        if (v > maxv) v = maxv;
    }
    return rndi(&g->rng, 0, 1) ? -v : v;
}

/* Generate one granule-channel into w (at most `budget` bits).  Returns bits written. */
static int gen_granule(gen_t *g, int gr, int ch, int block_type, int mixed, int mode_ext,
                       int budget, bitw *w, gr_side *s, int scfsi[4], int line_limit)
{
    const l3gen_cfg *c = g->cfg;
    rng_t *r = &g->rng;
    size_t start = w->pos;
    memset(s, 0, sizeof *s);
    for (int i = 0; i < 4; i++) scfsi[i] = 0;
    if (budget > 4095) budget = 4095;
    s->block_type = block_type;
    s->mixed = mixed;
    s->window_switching = block_type != 0;
    s->scalefac_scale = rndi(r, 0, 1);
    s->count1table = rndi(r, 0, 1);
    s->preflag = (!g->lsf && block_type != 2) ? (rndi(r, 0, 3) == 0) : 0;
    if (block_type == 2)
        for (int i = 0; i < 3; i++) s->subblock_gain[i] = rndi(r, 0, 3) ? rndi(r, 0, 7) : 0;
    for (int i = 0; i < 3; i++) s->table_select[i] = pick_table(g);
    if (s->window_switching) {
        s->table_select[2] = 0;
        s->region0_count = (block_type == 2 && !mixed) ? 8 : 7;
        s->region1_count = 36;
    } else {
        s->region0_count = rndi(r, 0, 15);
        int m = 20 - s->region0_count;
        s->region1_count = rndi(r, 0, m > 7 ? 7 : m);
    }

    /* ---- part 2: scalefactors (band order) */
    uint8_t sf[40];
    memset(sf, 0, sizeof sf);
    int ist = g->mode == 1 && (mode_ext & 1) && ch == 1;
    if (!g->lsf) {
        int sfc = rndi(r, 0, 15), s1, s2, nbits;
        for (;; sfc = 0) {
            s1 = l3_slen[0][sfc];
            s2 = l3_slen[1][sfc];
            if (block_type == 2) nbits = mixed ? 17 * s1 + 18 * s2 : 18 * s1 + 18 * s2;
            else nbits = 11 * s1 + 10 * s2;
            if (nbits <= budget || sfc == 0) break;
        }
        s->scalefac_compress = sfc;
        if (block_type == 2) {
            int n1 = mixed ? 17 : 18, ntot = mixed ? 35 : 36;
            for (int i = 0; i < ntot; i++) {
                int sl = i < n1 ? s1 : s2;
                sf[i] = (uint8_t)(sl ? rndi(r, 0, (1 << sl) - 1) : 0);
                putbits(w, sf[i], sl);
            }
        } else {
            static const int grp[5] = {0, 6, 11, 16, 21};
            int may_share = gr == 1 && g->bt_gr0[ch] != 2;
            for (int k = 0; k < 4; k++) {
                scfsi[k] = may_share && rndi(r, 0, 99) < c->scfsi_pct;
                for (int b = grp[k]; b < grp[k + 1]; b++) {
                    int sl = b < 11 ? s1 : s2;
                    if (scfsi[k]) sf[b] = g->sf_gr0[ch][b];
                    else {
                        sf[b] = (uint8_t)(sl ? rndi(r, 0, (1 << sl) - 1) : 0);
                        putbits(w, sf[b], sl);
                    }
                }
            }
        }
        if (gr == 0) { memcpy(g->sf_gr0[ch], sf, 40); g->bt_gr0[ch] = block_type; }
    } else {
        int bt = block_type == 2 ? (mixed ? 2 : 1) : 0;
        int sfc, slen[4], tbl, nbits;
        for (int tries = 0;; tries++) {
            sfc = tries < 8 ? rndi(r, 0, 511) : 0;
            int v = sfc;
            if (!ist) {
                if (v < 400) { slen[0] = (v >> 4) / 5; slen[1] = (v >> 4) % 5; slen[2] = (v & 15) >> 2; slen[3] = v & 3; tbl = 0; }
                else if (v < 500) { v -= 400; slen[0] = (v >> 2) / 5; slen[1] = (v >> 2) % 5; slen[2] = v & 3; slen[3] = 0; tbl = 1; }
                else { v -= 500; slen[0] = v / 3; slen[1] = v % 3; slen[2] = 0; slen[3] = 0; tbl = 2; }
            } else {
                v >>= 1;
                if (v < 180) { slen[0] = v / 36; slen[1] = (v % 36) / 6; slen[2] = v % 6; slen[3] = 0; tbl = 3; }
                else if (v < 244) { v -= 180; slen[0] = (v & 63) >> 4; slen[1] = (v & 15) >> 2; slen[2] = v & 3; slen[3] = 0; tbl = 4; }
                else { v -= 244; slen[0] = v / 3; slen[1] = v % 3; slen[2] = 0; slen[3] = 0; tbl = 5; }
            }
            nbits = 0;
            for (int k = 0; k < 4; k++) nbits += slen[k] * l3_lsf_nsfb[tbl][bt][k];
            if (nbits <= budget || tries >= 8) break;
        }
        s->scalefac_compress = sfc;
        int n = 0;
        for (int k = 0; k < 4; k++)
            for (int i = 0; i < l3_lsf_nsfb[tbl][bt][k]; i++) {
                int hi = slen[k] ? (1 << slen[k]) - 1 : 0;
                if (ist && c->lsf_avoid_illegal_ispos && hi > 0) {
                    hi--;                 /* never the 'illegal' position 2^slen - 1 ... */
                    if (hi > 15) hi = 15; /* ... nor positions FFmpeg's 16-entry table cannot express */
                }
                sf[n] = (uint8_t)rndi(r, 0, hi);
                putbits(w, sf[n], slen[k]);
                n++;
            }
    }

    /* ---- part 3: Huffman */
    const uint16_t *bl = l3_sfb_long[g->sr_row];
    int r1, r2;
    if (s->window_switching) { r1 = block_type == 2 ? 3 * l3_sfb_short[g->sr_row][3] : bl[8]; r2 = 576; }
    else { r1 = bl[s->region0_count + 1]; r2 = bl[s->region0_count + s->region1_count + 2]; }
    int bv_target = rndi(r, 0, 9) == 0 ? rndi(r, 0, 288) : rndi(r, 40, 200);
    if (bv_target * 2 > line_limit) bv_target = line_limit / 2;
    double scale[3];
    for (int i = 0; i < 3; i++) scale[i] = rndi(r, 0, 4) == 0 ? 8.0 + 30.0 * rndu(r) : 0.4 + 3.0 * rndu(r);
    double energy = 0;
    int i = 0, bv = 0;
    for (; bv < bv_target; bv++, i += 2) {
        int reg = i < r1 ? 0 : (i < r2 ? 1 : 2);
        int tsel = s->table_select[reg];
        int x = draw_value(g, tsel, scale[reg]), y = draw_value(g, tsel, scale[reg]);
        int nb = pair_bits(tsel, x, y);
        if ((int)(w->pos - start) + nb > budget) break;
        put_pair(w, tsel, x, y);
        energy += pow(fabs((double)x), 8.0 / 3.0) + pow(fabs((double)y), 8.0 / 3.0);
    }
    s->big_values = bv;
    int quads_target = rndi(r, 0, 3) == 0 ? 200 : rndi(r, 0, 60);
    const uint8_t *ql = l3_quad_hlen[s->count1table], *qc = l3_quad_hcod[s->count1table];
    for (int q = 0; q < quads_target && i + 4 <= 576 && i + 4 <= line_limit; q++, i += 4) {
        int sym = rndi(r, 0, 15), nb = ql[sym] + ((sym >> 3) & 1) + ((sym >> 2) & 1) + ((sym >> 1) & 1) + (sym & 1);
        if ((int)(w->pos - start) + nb > budget) break;
        putbits(w, qc[sym], ql[sym]);
        for (int k = 3; k >= 0; k--)
            if ((sym >> k) & 1) { putbits(w, (unsigned)rndi(r, 0, 1), 1); energy += 1.0; }
    }
    s->part2_3_length = (int)(w->pos - start);

    /* ---- level: choose global_gain so the decoded granule sits at a sane level */
    if (energy > 0) {
        double db = c->level_lo_db + (c->level_hi_db - c->level_lo_db) * rndu(r);
        double target = 2.0 * pow(10.0, -db / 10.0); /* sum(xr^2) = 2 * rms^2 */
        int gg = 210 + (int)floor(2.0 * log2(target / energy));
        if (gg < 0) gg = 0;
        if (gg > 255) gg = 255;
        s->global_gain = gg;
    } else
        s->global_gain = rndi(r, 100, 200);
    return s->part2_3_length;
}

static unsigned crc16_update(unsigned crc, const uint8_t *p, int nbits)
{
    for (int i = 0; i < nbits; i++) {
        unsigned bit = (p[i >> 3] >> (7 - (i & 7))) & 1u;
        unsigned top = (crc >> 15) & 1u;
        crc = (crc << 1) & 0xffff;
        if (top ^ bit) crc ^= 0x8005;
    }
    return crc;
}

static int bitrate_index(int lsf, int kbps)
{
    for (int i = 1; i < 15; i++)
        if (l3_bitrate_kbps[lsf][i] == kbps) return i;
    return -1;
}

/* Upper bound on the stream size for a config (for caller allocation). */
size_t l3gen_max_bytes(const l3gen_cfg *c)
{
    int lsf = c->sample_rate < 32000 && c->layer != 2 && c->layer != 1;
    int kb = c->vbr_max_kbps > 0 ? c->vbr_max_kbps : c->bitrate_kbps;
    size_t fl = (size_t)(lsf ? 72 : 144) * kb * 1000 / c->sample_rate + 1;
    return fl * (size_t)c->nframes + 16 + (c->tag ? 1500 : 0);
}

/* The tag frame encoders put first: a Layer III frame of the stream's own version / sample rate /
 * mode whose side info and main data are zero (it decodes to silence) and whose payload carries
 * "Xing"/"Info" (+ the LAME extension) or "VBRI".  Returns its length. */
static size_t put_tag_frame(const l3gen_cfg *c, int row, int lsf, int nch, int side_len, size_t audio_bytes,
                            uint8_t *out, size_t cap)
{
    int need = 4 + side_len + 120 + 36 + 4;
    if (c->tag == 3 && need < 4 + 32 + 26 + 4) need = 4 + 32 + 26 + 4;
    int bri = 0, flen = 0;
    for (int i = 1; i < 15; i++) {
        flen = (lsf ? 72 : 144) * l3_bitrate_kbps[lsf][i] * 1000 / (int)l3_sample_rate[row];
        if (flen >= need) { bri = i; break; }
    }
    if (!bri || (size_t)flen > cap) return 0;
    memset(out, 0, (size_t)flen);
    out[0] = 0xFF;
    out[1] = (uint8_t)(0xE0 | ((row >= 6 ? 0 : lsf ? 2 : 3) << 3) | (1 << 1) | 1);
    out[2] = (uint8_t)((bri << 4) | ((row % 3) << 2));
    out[3] = (uint8_t)(c->mode << 6);
    (void)nch;
    uint32_t frames = (uint32_t)c->nframes, bytes = (uint32_t)(audio_bytes + (size_t)flen);
    if (c->tag == 3) {
        uint8_t *t = out + 4 + 32;
        memcpy(t, "VBRI", 4);
        t[4] = 0; t[5] = 1;                                   /* version */
        t[6] = (uint8_t)(c->enc_delay >> 8); t[7] = (uint8_t)c->enc_delay;
        t[8] = 0; t[9] = 75;                                  /* quality */
        t[10] = (uint8_t)(bytes >> 24); t[11] = (uint8_t)(bytes >> 16); t[12] = (uint8_t)(bytes >> 8); t[13] = (uint8_t)bytes;
        t[14] = (uint8_t)(frames >> 24); t[15] = (uint8_t)(frames >> 16); t[16] = (uint8_t)(frames >> 8); t[17] = (uint8_t)frames;
        return (size_t)flen;
    }
    uint8_t *t = out + 4 + side_len;
    memcpy(t, c->tag == 1 ? "Xing" : "Info", 4);
    t[7] = 0x0F; /* frames | bytes | TOC | quality */
    t[8] = (uint8_t)(frames >> 24); t[9] = (uint8_t)(frames >> 16); t[10] = (uint8_t)(frames >> 8); t[11] = (uint8_t)frames;
    t[12] = (uint8_t)(bytes >> 24); t[13] = (uint8_t)(bytes >> 16); t[14] = (uint8_t)(bytes >> 8); t[15] = (uint8_t)bytes;
    for (int i = 0; i < 100; i++) t[16 + i] = (uint8_t)(i * 256 / 100);
    t[119] = 57; /* quality */
    if (c->tag_lame) {
        uint8_t *l = t + 120;
        memcpy(l, "LAME3.100", 9);
        l[9] = 0x04;  /* tag revision 0, VBR method 4 */
        l[10] = 190;  /* lowpass / 100 */
        l[21] = (uint8_t)(c->enc_delay >> 4);
        l[22] = (uint8_t)(((c->enc_delay & 15) << 4) | ((c->enc_padding >> 8) & 15));
        l[23] = (uint8_t)c->enc_padding;
    }
    return (size_t)flen;
}

/* Layer II: random bit allocation (thinned until the frame fits), random scfsi / scalefactors / sample
 * codes.  Returns bytes written, or 0 on a bad config / small buffer. */
static size_t l2gen_stream(const l3gen_cfg *c, uint8_t *out, size_t cap)
{
    rng_t rng;
    rng.s = c->seed * 0x2545F4914F6CDD1Dull + 0x7654321;
    int row = -1;
    for (int i = 0; i < 9; i++)
        if ((int)l3_sample_rate[i] == c->sample_rate) row = i;
    if (row < 0 || c->nframes <= 0) return 0;
    const int lsf = row >= 3, nch = c->mode == 3 ? 1 : 2;
    int bri = 0;
    for (int i = 1; i < 15; i++)
        if (l2_bitrate_kbps[lsf][i] == c->bitrate_kbps) bri = i;
    if (!bri) return 0;
    const int tbl = l2_select_table(lsf, c->sample_rate, c->bitrate_kbps, nch), sblimit = l2_sblimit[tbl];
    size_t o = 0;
    long pad_rest = 0;
    for (int f = 0; f < c->nframes; f++) {
        int pad = 0;
        long num = 144L * c->bitrate_kbps * 1000;
        pad_rest -= num % c->sample_rate;
        if (pad_rest < 0) { pad = 1; pad_rest += c->sample_rate; }
        const int flen = (int)(num / c->sample_rate) + pad;
        if (o + (size_t)flen > cap) return 0;
        uint8_t *fr = out + o;
        memset(fr, 0, (size_t)flen);
        int mode_ext = 0;
        if (c->mode == 1) mode_ext = rndi(&rng, 0, 3);
        fr[0] = 0xFF;
        fr[1] = (uint8_t)(0xE0 | ((row >= 6 ? 0 : lsf ? 2 : 3) << 3) | (2 << 1) | (c->crc ? 0 : 1));
        fr[2] = (uint8_t)((bri << 4) | ((row % 3) << 2) | (pad << 1));
        fr[3] = (uint8_t)((c->mode << 6) | (mode_ext << 4));
        int bound = c->mode == 1 ? (mode_ext + 1) * 4 : sblimit;
        if (bound > sblimit || nch == 1) bound = sblimit;
        const int hdr_bits = 32 + (c->crc ? 16 : 0);
        const long budget = ((long)flen * 8 - hdr_bits) * rndi(&rng, c->fill_lo_pct, c->fill_hi_pct) / 100;
        int alloc[2][32];
        memset(alloc, 0, sizeof alloc);
        for (int s = 0; s < sblimit; s++) {
            const uint8_t *rw = l2_rows[l2_row_of_sb[tbl][s]];
            int hi = (1 << rw[0]) - 1;
            for (int ch = 0; ch < (s < bound ? nch : 1); ch++) {
                int a = rndi(&rng, 0, 9) < 2 ? 0 : rndi(&rng, 1, hi);
                if (rndi(&rng, 0, 9) < 6 && a > 4) a = rndi(&rng, 1, 4); /* mostly coarse classes, sometimes the finest */
                alloc[ch][s] = a;
            }
            if (s >= bound) alloc[1][s] = alloc[0][s];
        }
        for (;;) { /* thin the allocation until the frame fits its budget */
            long bits = 0;
            for (int s = 0; s < sblimit; s++) {
                const uint8_t *rw = l2_rows[l2_row_of_sb[tbl][s]];
                bits += rw[0] * (s < bound ? nch : 1);
                for (int ch = 0; ch < nch; ch++)
                    if (alloc[ch][s]) bits += 2 + 18; /* scfsi + up to three scalefactors */
                for (int ch = 0; ch < (s < bound ? nch : 1); ch++)
                    if (alloc[ch][s]) {
                        int q = rw[alloc[ch][s]], b = l2_quant_bits[q];
                        bits += 12L * (b < 0 ? -b : 3 * b);
                    }
            }
            if (bits <= budget) break;
            int s = rndi(&rng, 0, sblimit - 1), ch = rndi(&rng, 0, nch - 1);
            if (s >= bound) { alloc[0][s] = alloc[1][s] = alloc[0][s] > 1 ? alloc[0][s] - 1 : 0; }
            else alloc[ch][s] = alloc[ch][s] > 1 ? alloc[ch][s] / 2 : 0;
        }
        bitw w = {fr, (size_t)flen * 8, (size_t)hdr_bits};
        for (int s = 0; s < sblimit; s++) {
            const uint8_t *rw = l2_rows[l2_row_of_sb[tbl][s]];
            for (int ch = 0; ch < (s < bound ? nch : 1); ch++) putbits(&w, (unsigned)alloc[ch][s], rw[0]);
        }
        int scfsi[2][32];
        for (int s = 0; s < sblimit; s++)
            for (int ch = 0; ch < nch; ch++)
                if (alloc[ch][s]) { scfsi[ch][s] = rndi(&rng, 0, 3); putbits(&w, (unsigned)scfsi[ch][s], 2); }
        const size_t crc_end = w.pos;
        /* scalefactor index: 2^(1 - i/3); level_lo_db .. level_hi_db below full scale, spread over the subbands */
        const int i_lo = 3 + c->level_lo_db / 2 + 8, i_hi = 3 + c->level_hi_db / 2 + 8;
        for (int s = 0; s < sblimit; s++)
            for (int ch = 0; ch < nch; ch++) {
                if (!alloc[ch][s]) continue;
                int n = scfsi[ch][s] == 0 ? 3 : (scfsi[ch][s] == 2 ? 1 : 2);
                for (int k = 0; k < n; k++) {
                    int v = rndi(&rng, i_lo, i_hi < 62 ? i_hi : 62);
                    if (rndi(&rng, 0, 99) == 0) v = rndi(&rng, 0, 62); /* the whole range now and then */
                    putbits(&w, (unsigned)v, 6);
                }
            }
        for (int gr = 0; gr < 12; gr++)
            for (int s = 0; s < sblimit; s++) {
                const uint8_t *rw = l2_rows[l2_row_of_sb[tbl][s]];
                for (int ch = 0; ch < (s < bound ? nch : 1); ch++) {
                    if (!alloc[ch][s]) continue;
                    int q = rw[alloc[ch][s]], steps = l2_quant_steps[q], b = l2_quant_bits[q];
                    unsigned code[3];
                    for (int i = 0; i < 3; i++) code[i] = (unsigned)rndi(&rng, 0, steps - 1);
                    if (b < 0) putbits(&w, code[0] + (unsigned)steps * (code[1] + (unsigned)steps * code[2]), -b);
                    else
                        for (int i = 0; i < 3; i++) putbits(&w, code[i], b);
                }
            }
        if (c->crc) { /* 11172-3 2.4.3.1: header bits 16..31, bit allocation and scfsi */
            unsigned crc = 0xffff;
            crc = crc16_update(crc, fr + 2, 16);
            crc = crc16_update(crc, fr + 6, (int)(crc_end - 48));
            fr[4] = (uint8_t)(crc >> 8);
            fr[5] = (uint8_t)crc;
        }
        o += (size_t)flen;
    }
    return o;
}

/* Layer I: random 4-bit allocations (thinned to the frame's budget), scalefactors, sample codes. */
static size_t l1gen_stream(const l3gen_cfg *c, uint8_t *out, size_t cap)
{
    rng_t rng;
    rng.s = c->seed * 0x2545F4914F6CDD1Dull + 0x1357911;
    int row = -1;
    for (int i = 0; i < 9; i++)
        if ((int)l3_sample_rate[i] == c->sample_rate) row = i;
    if (row < 0 || c->nframes <= 0) return 0;
    const int lsf = row >= 3, nch = c->mode == 3 ? 1 : 2;
    int bri = 0;
    for (int i = 1; i < 15; i++)
        if (l1_bitrate_kbps[lsf][i] == c->bitrate_kbps) bri = i;
    if (!bri) return 0;
    size_t o = 0;
    long pad_rest = 0;
    for (int f = 0; f < c->nframes; f++) {
        int pad = 0;
        long num = 12L * c->bitrate_kbps * 1000;
        pad_rest -= num % c->sample_rate;
        if (pad_rest < 0) { pad = 1; pad_rest += c->sample_rate; }
        const int flen = ((int)(num / c->sample_rate) + pad) * 4;
        if (o + (size_t)flen > cap) return 0;
        uint8_t *fr = out + o;
        memset(fr, 0, (size_t)flen);
        int mode_ext = c->mode == 1 ? rndi(&rng, 0, 3) : 0;
        fr[0] = 0xFF;
        fr[1] = (uint8_t)(0xE0 | ((row >= 6 ? 0 : lsf ? 2 : 3) << 3) | (3 << 1) | (c->crc ? 0 : 1));
        fr[2] = (uint8_t)((bri << 4) | ((row % 3) << 2) | (pad << 1));
        fr[3] = (uint8_t)((c->mode << 6) | (mode_ext << 4));
        const int bound = (c->mode == 1 && nch == 2) ? (mode_ext + 1) * 4 : 32;
        const int hdr_bits = 32 + (c->crc ? 16 : 0);
        const long budget = ((long)flen * 8 - hdr_bits) * rndi(&rng, c->fill_lo_pct, c->fill_hi_pct) / 100;
        int alloc[2][32];
        for (int s = 0; s < 32; s++) {
            for (int ch = 0; ch < (s < bound ? nch : 1); ch++) {
                int a = rndi(&rng, 0, 9) < 3 ? 0 : rndi(&rng, 1, 14);
                if (rndi(&rng, 0, 9) < 6 && a > 5) a = rndi(&rng, 1, 5);
                alloc[ch][s] = a;
            }
            if (s >= bound) alloc[1][s] = alloc[0][s];
            if (nch == 1) alloc[1][s] = 0;
        }
        for (;;) {
            long bits = 0;
            for (int s = 0; s < 32; s++) {
                bits += 4 * (s < bound ? nch : 1);
                for (int ch = 0; ch < nch; ch++)
                    if (alloc[ch][s]) bits += 6;
                for (int ch = 0; ch < (s < bound ? nch : 1); ch++)
                    if (alloc[ch][s]) bits += 12L * (alloc[ch][s] + 1);
            }
            if (bits <= budget) break;
            int s = rndi(&rng, 0, 31), ch = rndi(&rng, 0, nch - 1);
            if (s >= bound) alloc[0][s] = alloc[1][s] = alloc[0][s] > 1 ? alloc[0][s] - 1 : 0;
            else alloc[ch][s] = alloc[ch][s] > 1 ? alloc[ch][s] / 2 : 0;
        }
        bitw w = {fr, (size_t)flen * 8, (size_t)hdr_bits};
        for (int s = 0; s < 32; s++)
            for (int ch = 0; ch < (s < bound ? nch : 1); ch++) putbits(&w, (unsigned)alloc[ch][s], 4);
        const size_t crc_end = w.pos;
        const int i_lo = 3 + c->level_lo_db / 2 + 8, i_hi = 3 + c->level_hi_db / 2 + 8;
        for (int s = 0; s < 32; s++)
            for (int ch = 0; ch < nch; ch++)
                if (alloc[ch][s]) {
                    int v = rndi(&rng, i_lo, i_hi < 62 ? i_hi : 62);
                    if (rndi(&rng, 0, 99) == 0) v = rndi(&rng, 0, 62);
                    putbits(&w, (unsigned)v, 6);
                }
        for (int t = 0; t < 12; t++)
            for (int s = 0; s < 32; s++)
                for (int ch = 0; ch < (s < bound ? nch : 1); ch++)
                    if (alloc[ch][s]) {
                        int b = alloc[ch][s] + 1;
                        putbits(&w, (unsigned)rndi(&rng, 0, (1 << b) - 2), b); /* the all-ones code is forbidden */
                    }
        if (c->crc) { /* header bits 16..31 and the bit allocation */
            unsigned crc = 0xffff;
            crc = crc16_update(crc, fr + 2, 16);
            crc = crc16_update(crc, fr + 6, (int)(crc_end - 48));
            fr[4] = (uint8_t)(crc >> 8);
            fr[5] = (uint8_t)crc;
        }
        o += (size_t)flen;
    }
    return o;
}

/* Generate one stream.  Returns bytes written, or 0 on a bad config / small buffer. */
size_t l3gen_stream(const l3gen_cfg *c, uint8_t *out, size_t cap)
{
    if (c->layer == 2) return l2gen_stream(c, out, cap);
    if (c->layer == 1) return l1gen_stream(c, out, cap);
    gen_t g;
    memset(&g, 0, sizeof g);
    g.cfg = c;
    g.rng.s = c->seed * 0x2545F4914F6CDD1Dull + 0x1234567;
    int row = -1;
    for (int i = 0; i < 9; i++)
        if ((int)l3_sample_rate[i] == c->sample_rate) row = i;
    if (row < 0 || c->nframes <= 0) return 0;
    g.sr_row = row;
    g.lsf = row >= 3;
    g.mode = c->mode;
    g.nch = c->mode == 3 ? 1 : 2;
    int ngr = g.lsf ? 1 : 2, nch = g.nch;
    int side_len = g.lsf ? (nch == 1 ? 9 : 17) : (nch == 1 ? 17 : 32);
    int hdr_len = 4 + (c->crc ? 2 : 0);
    int max_mdb = g.lsf ? 255 : 511;

    /* logical main-data stream */
    size_t lcap = l3gen_max_bytes(c) + 1024;
    uint8_t *logical = (uint8_t *)calloc(lcap, 1);
    typedef struct { uint8_t hdr[4]; uint8_t side[32]; int payload; } fr_t;
    fr_t *fr = (fr_t *)calloc((size_t)c->nframes, sizeof(fr_t));
    size_t lstart = 0;   /* logical offset of this frame's payload */
    size_t wcur = 0;     /* first unused logical byte */
    long pad_rest = 0;

    for (int f = 0; f < c->nframes; f++) {
        int kbps = c->bitrate_kbps, pad = 0;
        if (c->vbr_max_kbps > 0) {
            int lo = bitrate_index(g.lsf, c->vbr_min_kbps), hi = bitrate_index(g.lsf, c->vbr_max_kbps);
            if (lo < 0 || hi < 0) { free(logical); free(fr); return 0; }
            kbps = l3_bitrate_kbps[g.lsf][rndi(&g.rng, lo, hi)];
            pad = rndi(&g.rng, 0, 1);
        } else {
            /* CBR padding: keep the long-run byte rate exact */
            long num = (long)(g.lsf ? 72 : 144) * kbps * 1000;
            pad_rest -= num % c->sample_rate;
            if (pad_rest < 0) { pad = 1; pad_rest += c->sample_rate; }
        }
        int bri = bitrate_index(g.lsf, kbps);
        if (bri < 0) { free(logical); free(fr); return 0; }
        int frame_len = (g.lsf ? 72 : 144) * kbps * 1000 / c->sample_rate + pad;
        int payload = frame_len - hdr_len - side_len;
        if (payload < 0) { free(logical); free(fr); return 0; }
        int mode_ext = 0;
        if (c->mode == 1) {
            int mask = c->mode_ext_mask & 15;
            if (!mask) mask = 15;
            do mode_ext = rndi(&g.rng, 0, 3); while (!((mask >> mode_ext) & 1));
        }
        /* reservoir */
        size_t tail = lstart - wcur;
        int mdb = 0;
        if (c->reservoir && tail > 0) {
            int m = tail > (size_t)max_mdb ? max_mdb : (int)tail;
            mdb = rndi(&g.rng, 0, 3) ? m : rndi(&g.rng, 0, m);
        }
        size_t dstart = lstart - (size_t)mdb;
        long avail = 8L * (mdb + payload);
        double fill = (c->fill_lo_pct + (c->fill_hi_pct - c->fill_lo_pct) * rndu(&g.rng)) / 100.0;
        long use = (long)(avail * fill);
        bitw w = {logical, lcap * 8, dstart * 8};
        size_t wstart = w.pos;

        gr_side gs[2][2];
        int scfsi[2][4];
        memset(scfsi, 0, sizeof scfsi);
        int nunits = ngr * nch, u = 0;
        for (int gr = 0; gr < ngr; gr++) {
            int bt_shared = 0, mixed_shared = 0;
            for (int ch = 0; ch < nch; ch++, u++) {
                int bt, mixed;
                if (ch == 1 && c->mode == 1) { bt = bt_shared; mixed = mixed_shared; g.bt_state[1] = g.bt_state[0]; }
                else {
                    int prev = g.bt_state[ch];
                    bt = next_block_type(&g, ch);
                    mixed = (bt == 2 && !g.lsf && rndi(&g.rng, 0, 99) < c->mixed_pct);
                    if (bt == 2 && !c->mixed_free) {
                        if (prev == 2) mixed = g.mixed_run[ch];
                        else g.mixed_run[ch] = mixed;
                    }
                    bt_shared = bt;
                    mixed_shared = mixed;
                }
                long used = (long)(w.pos - wstart);
                long left = use - used;
                int remaining_units = nunits - u;
                long share = left / remaining_units;
                if (remaining_units > 1) share = (long)(share * (0.6 + 0.8 * rndu(&g.rng)));
                if (share > left) share = left;
                if (share < 0) share = 0;
                int limit = 576;
                if (c->mode == 1 && (mode_ext & 1) && ch == 1) limit = rndi(&g.rng, 0, 3) ? rndi(&g.rng, 0, 400) : 576;
                int sc[4];
                gen_granule(&g, gr, ch, bt, mixed, mode_ext, (int)share, &w, &gs[gr][ch], sc, limit);
                if (gr == 1) memcpy(scfsi[ch], sc, sizeof sc);
            }
        }
        size_t used_bytes = (w.pos - wstart + 7) / 8;
        wcur = dstart + used_bytes;

        /* header */
        uint8_t *h = fr[f].hdr;
        int sri = row % 3;
        h[0] = 0xFF;
        h[1] = (uint8_t)(0xE0 | ((row >= 6 ? 0 : g.lsf ? 2 : 3) << 3) | (1 << 1) | (c->crc ? 0 : 1));
        h[2] = (uint8_t)((bri << 4) | (sri << 2) | (pad << 1));
        h[3] = (uint8_t)((c->mode << 6) | (mode_ext << 4));
        /* side info */
        uint8_t *sp = fr[f].side;
        memset(sp, 0, 32);
        bitw sw = {sp, (size_t)side_len * 8, 0};
        if (!g.lsf) {
            putbits(&sw, (unsigned)mdb, 9);
            putbits(&sw, 0, nch == 1 ? 5 : 3);
            for (int ch = 0; ch < nch; ch++)
                for (int k = 0; k < 4; k++) putbits(&sw, (unsigned)scfsi[ch][k], 1);
        } else {
            putbits(&sw, (unsigned)mdb, 8);
            putbits(&sw, 0, nch == 1 ? 1 : 2);
        }
        for (int gr = 0; gr < ngr; gr++)
            for (int ch = 0; ch < nch; ch++) {
                gr_side *s = &gs[gr][ch];
                putbits(&sw, (unsigned)s->part2_3_length, 12);
                putbits(&sw, (unsigned)s->big_values, 9);
                putbits(&sw, (unsigned)s->global_gain, 8);
                putbits(&sw, (unsigned)s->scalefac_compress, g.lsf ? 9 : 4);
                putbits(&sw, (unsigned)s->window_switching, 1);
                if (s->window_switching) {
                    putbits(&sw, (unsigned)s->block_type, 2);
                    putbits(&sw, (unsigned)s->mixed, 1);
                    putbits(&sw, (unsigned)s->table_select[0], 5);
                    putbits(&sw, (unsigned)s->table_select[1], 5);
                    for (int k = 0; k < 3; k++) putbits(&sw, (unsigned)s->subblock_gain[k], 3);
                } else {
                    for (int k = 0; k < 3; k++) putbits(&sw, (unsigned)s->table_select[k], 5);
                    putbits(&sw, (unsigned)s->region0_count, 4);
                    putbits(&sw, (unsigned)s->region1_count, 3);
                }
                if (!g.lsf) putbits(&sw, (unsigned)s->preflag, 1);
                putbits(&sw, (unsigned)s->scalefac_scale, 1);
                putbits(&sw, (unsigned)s->count1table, 1);
            }
        fr[f].payload = payload;
        lstart += (size_t)payload;
    }

    /* interleave headers / side info / payload slices into the physical stream */
    size_t o = 0, lp = 0;
    if (c->tag) {
        size_t audio = 0;
        for (int f = 0; f < c->nframes; f++) audio += (size_t)hdr_len + side_len + fr[f].payload;
        o = put_tag_frame(c, row, g.lsf, nch, side_len, audio, out, cap);
        if (!o) { free(logical); free(fr); return 0; }
    }
    for (int f = 0; f < c->nframes; f++) {
        size_t need = (size_t)hdr_len + side_len + fr[f].payload;
        if (o + need > cap) { free(logical); free(fr); return 0; }
        memcpy(out + o, fr[f].hdr, 4);
        if (c->crc) {
            unsigned crc = 0xffff;
            crc = crc16_update(crc, fr[f].hdr + 2, 16);
            crc = crc16_update(crc, fr[f].side, side_len * 8);
            out[o + 4] = (uint8_t)(crc >> 8);
            out[o + 5] = (uint8_t)crc;
        }
        memcpy(out + o + hdr_len, fr[f].side, (size_t)side_len);
        memcpy(out + o + hdr_len + side_len, logical + lp, (size_t)fr[f].payload);
        o += need;
        lp += (size_t)fr[f].payload;
    }
    free(logical);
    free(fr);
    return o;
}

/* =====================================================================================================================
 * l3enc_stream: a small ENCODER of real signals, so that the decoders (oracle, CUDA path, FFmpeg) also see streams
 * whose statistics come from audio and not from a random generator: s16 PCM in, MPEG-1 Layer III out (32 / 44.1 / 48
 * kHz, mono or plain stereo, CBR with the bit reservoir in use, long blocks, all scalefactors zero).  It is the
 * encoding process of ISO/IEC 11172-3 Annex C read backwards from the decoder: polyphase analysis with the window
 * C[i] = D[i] / 32, frequency inversion, 36-point MDCT with the sine window, alias butterflies, power-law quantiser
 * with one global gain per granule and channel found by bisection against the bit budget, Huffman table per region by
 * exhaustive bit count.  No psychoacoustic model: the quantisation noise is white inside a granule, which is enough for
 * what the tests ask of it -- a stream a third-party decoder reconstructs to the input signal within the noise the bit
 * rate allows, with real big_values / count1 / reservoir statistics.  Test infrastructure, like the generator above.
 * ===================================================================================================================== */
typedef struct {
    double fifo[512];      /* analysis input, newest sample first */
    double prev[32][18];   /* previous granule's subband samples (MDCT overlap) */
} enc_ch;

static void enc_analysis(enc_ch *e, const double *in32 /* 32 new samples, oldest first */, double *S /* 32 */)
{
    memmove(e->fifo + 32, e->fifo, sizeof(double) * 480);
    for (int i = 0; i < 32; i++) e->fifo[31 - i] = in32[i];
    double y[64];
    for (int i = 0; i < 64; i++) {
        double a = 0.0;
        for (int j = 0; j < 8; j++) {
            const int idx = i + 64 * j; /* C[i] = D[i] / 32 (11172-3 Table C.1 against Table B.3) */
            a += l3_dwin(idx) / 32.0 * e->fifo[idx];
        }
        y[i] = a;
    }
    for (int k = 0; k < 32; k++) {
        double a = 0.0;
        for (int i = 0; i < 64; i++) a += cos((2 * k + 1) * (i - 16) * M_PI / 64.0) * y[i];
        S[k] = a;
    }
}

static double enc_window(int bt, int i) /* the decoder's window shapes (11172-3 2.4.3.4.10.3) */
{
    switch (bt) {
    case 0: return sin(M_PI / 36.0 * (i + 0.5));
    case 1: return i < 18 ? sin(M_PI / 36.0 * (i + 0.5)) : (i < 24 ? 1.0 : (i < 30 ? sin(M_PI / 12.0 * (i - 18 + 0.5)) : 0.0));
    case 3: return i < 6 ? 0.0 : (i < 12 ? sin(M_PI / 12.0 * (i - 6 + 0.5)) : (i < 18 ? 1.0 : sin(M_PI / 36.0 * (i + 0.5))));
    default: return sin(M_PI / 12.0 * (i + 0.5));
    }
}

/* one granule of one channel: 576 PCM samples -> 576 spectral lines in BITSTREAM order (short blocks: per scalefactor
 * band the three windows one after the other).  bt = block type of the granule (2: three short windows, not mixed). */
static void enc_granule_spectrum(enc_ch *e, const double *pcm576, double *xr /* 576 */, int bt, int row)
{
    static double cur[32][18];
    for (int t = 0; t < 18; t++) {
        double S[32];
        enc_analysis(e, pcm576 + 32 * t, S);
        for (int sb = 0; sb < 32; sb++) cur[sb][t] = ((sb & 1) && (t & 1)) ? -S[sb] : S[sb]; /* frequency inversion */
    }
    for (int sb = 0; sb < 32; sb++) {
        double in[36];
        for (int n = 0; n < 18; n++) { in[n] = e->prev[sb][n]; in[18 + n] = cur[sb][n]; }
        if (bt != 2) {
            for (int k = 0; k < 18; k++) {
                double a = 0.0;
                for (int n = 0; n < 36; n++)
                    a += enc_window(bt, n) * in[n] * cos(M_PI / 72.0 * (2 * n + 1 + 18) * (2 * k + 1));
                xr[sb * 18 + k] = a / 9.0;
            }
        } else {
            for (int w = 0; w < 3; w++)
                for (int k = 0; k < 6; k++) {
                    double a = 0.0;
                    for (int n = 0; n < 12; n++)
                        a += enc_window(2, n) * in[6 + 6 * w + n] * cos(M_PI / 24.0 * (2 * n + 1 + 6) * (2 * k + 1));
                    xr[sb * 18 + 3 * k + w] = a / 3.0; /* the decoder's (reordered) layout: windows interleaved */
                }
        }
        memcpy(e->prev[sb], cur[sb], sizeof e->prev[sb]);
    }
    if (bt != 2) {
        for (int sb = 1; sb < 32; sb++) /* alias butterflies: the inverse rotation of the decoder's */
            for (int i = 0; i < 8; i++) {
                double ci = l3_alias_ci[i], cs = 1.0 / sqrt(1.0 + ci * ci), ca = ci / sqrt(1.0 + ci * ci);
                double lo = xr[sb * 18 - 1 - i], hi = xr[sb * 18 + i];
                xr[sb * 18 - 1 - i] = lo * cs + hi * ca;
                xr[sb * 18 + i] = hi * cs - lo * ca;
            }
    } else { /* back to transmission order: the inverse of the decoder's reorder */
        double t[576];
        memcpy(t, xr, sizeof t);
        const uint16_t *bs = l3_sfb_short[row];
        for (int sfb = 0; sfb < 13; sfb++) {
            const int s0 = 3 * bs[sfb], wd = bs[sfb + 1] - bs[sfb];
            for (int win = 0; win < 3; win++)
                for (int k = 0; k < wd; k++) xr[s0 + win * wd + k] = t[s0 + 3 * k + win];
        }
    }
}

static int enc_best_table(const int *ix, int lo, int hi, int *bits_out)
{
    int mx = 0;
    for (int i = lo; i < hi; i++) if (abs(ix[i]) > mx) mx = abs(ix[i]);
    if (lo >= hi || mx == 0) { *bits_out = 0; return 0; }
    int best = -1, bb = 1 << 30;
    for (int t = 1; t < 32; t++) {
        if (t == 4 || t == 14) continue;
        int book = l3_book_of_table[t], lin = l3_linbits_of_table[t];
        const uint8_t *hl;
        const uint32_t *hc;
        int dim = l3_book(book, &hl, &hc);
        if (!dim) continue;
        int cap = lin ? 15 + ((1 << lin) - 1) : dim - 1;
        if (mx > cap) continue;
        int b = 0;
        for (int i = lo; i < hi; i += 2) b += pair_bits(t, ix[i], ix[i + 1]);
        if (b < bb) { bb = b; best = t; }
    }
    *bits_out = bb;
    return best;
}

typedef struct { int bv2, c1end, r1, r2, a, c, tsel[3], c1tab, bits; } enc_plan;

/* bits of the quantised spectrum with the best tables; fills the plan.  bt != 0: window switching -- two regions, the
 * first one fixed by the standard (36 lines; 2.4.2.7 region0_count 7 / 8) */
static int enc_count(const int *ix, int row, enc_plan *p, int bt)
{
    int last = 576;
    while (last > 0 && ix[last - 1] == 0) last--;
    last = (last + 1) & ~1;
    int c1end = (last + 3) & ~3;
    if (c1end > 576) c1end = 576;
    /* count1 region: quadruples from the top whose magnitudes are all <= 1 */
    int bv2 = c1end;
    while (bv2 >= 4 && abs(ix[bv2 - 1]) <= 1 && abs(ix[bv2 - 2]) <= 1 && abs(ix[bv2 - 3]) <= 1 && abs(ix[bv2 - 4]) <= 1) bv2 -= 4;
    if (bv2 > 576) bv2 = 576;
    p->bv2 = bv2;
    p->c1end = c1end;
    /* region boundaries at scalefactor-band edges near the thirds of big_values */
    const uint16_t *sfb = l3_sfb_long[row];
    int a = 1, c = 2;
    for (int k = 1; k <= 16; k++) if (sfb[k] <= bv2 / 3 || k == 1) a = k;
    for (int k = a + 1; k <= a + 8 && k <= 22; k++) if (sfb[k] <= 2 * bv2 / 3 || k == a + 1) c = k;
    p->a = a;
    p->c = c;
    p->r1 = sfb[a] < bv2 ? sfb[a] : bv2;
    p->r2 = sfb[c] < bv2 ? sfb[c] : bv2;
    if (bt) {
        const int r1 = bt == 2 ? 3 * l3_sfb_short[row][3] : sfb[8];
        p->r1 = r1 < bv2 ? r1 : bv2;
        p->r2 = bv2;
    }
    int b0, b1, b2;
    p->tsel[0] = enc_best_table(ix, 0, p->r1, &b0);
    p->tsel[1] = enc_best_table(ix, p->r1, p->r2, &b1);
    p->tsel[2] = enc_best_table(ix, p->r2, bv2, &b2);
    if (p->tsel[0] < 0 || p->tsel[1] < 0 || p->tsel[2] < 0) return 1 << 30; /* a value no table can hold */
    int ca = 0, cb = 0;
    for (int i = bv2; i < c1end; i += 4) {
        int sym = (ix[i] != 0) << 3 | (ix[i + 1] != 0) << 2 | (ix[i + 2] != 0) << 1 | (ix[i + 3] != 0);
        int ns = (ix[i] != 0) + (ix[i + 1] != 0) + (ix[i + 2] != 0) + (ix[i + 3] != 0);
        ca += l3_quad_hlen[0][sym] + ns;
        cb += l3_quad_hlen[1][sym] + ns;
    }
    p->c1tab = cb < ca;
    p->bits = b0 + b1 + b2 + (cb < ca ? cb : ca);
    return p->bits;
}

static int enc_quantise(const double *xr, int gg, int *ix)
{
    const double step = pow(2.0, -(gg - 210) * 3.0 / 16.0);
    int mx = 0;
    for (int i = 0; i < 576; i++) {
        int q = (int)(pow(fabs(xr[i]), 0.75) * step + 0.4054);
        if (q > 8191 + 15) q = 8191 + 15 + 1; /* marks "too fine" */
        ix[i] = xr[i] < 0 ? -q : q;
        if (q > mx) mx = q;
    }
    return mx;
}

/* pcm: interleaved s16, nsamples per channel.  Returns bytes written (0: bad arguments / buffer too small). */
size_t l3enc_stream2(const int16_t *pcm, int nsamples, int nch, int sample_rate, int kbps, int use_short, uint8_t *out,
                     size_t cap);
size_t l3enc_stream(const int16_t *pcm, int nsamples, int nch, int sample_rate, int kbps, uint8_t *out, size_t cap)
{
    return l3enc_stream2(pcm, nsamples, nch, sample_rate, kbps, 0, out, cap);
}

/* use_short: 1 = window switching on attacks */
size_t l3enc_stream2(const int16_t *pcm, int nsamples, int nch, int sample_rate, int kbps, int use_short, uint8_t *out,
                     size_t cap)
{
    int row = -1;
    for (int i = 0; i < 3; i++) if ((int)l3_sample_rate[i] == sample_rate) row = i;
    const int bri = bitrate_index(0, kbps);
    if (row < 0 || bri < 0 || (nch != 1 && nch != 2) || nsamples <= 0) return 0;
    const int nframes = (nsamples + 1151) / 1152 + 1; /* one more frame flushes the filterbank's delay */
    const int side_len = nch == 1 ? 17 : 32;
    size_t lcap = (size_t)(144 * kbps * 1000 / sample_rate + 1) * nframes + 1024;
    uint8_t *logical = (uint8_t *)calloc(lcap, 1);
    typedef struct { uint8_t hdr[4]; uint8_t side[32]; int payload; } efr_t;
    efr_t *fr = (efr_t *)calloc((size_t)nframes, sizeof(efr_t));
    enc_ch *ech = (enc_ch *)calloc(2, sizeof(enc_ch));
    size_t lstart = 0, wcur = 0;
    long pad_rest = 0;
    double e_avg = 0.0;
    /* window switching: a granule whose energy jumps by more than 16 x between thirds (an attack) is coded with three
     * short windows; the granules around a run of short ones take the start / stop windows (0 -> 1 -> 2 .. 2 -> 3 -> 0).
     * The filterbank delays the signal by 480 samples and the MDCT looks one granule back: the attack detector reads the
     * PCM that far behind the granule it labels. */
    const int ngran = 2 * nframes;
    uint8_t *btype[2];
    for (int ch = 0; ch < 2; ch++) {
        btype[ch] = (uint8_t *)calloc((size_t)ngran + 2, 1);
        if (!use_short || ch >= nch) continue;
        double prev_e = 0.0;
        for (int g = 1; g < ngran; g++) {
            int attack = 0;
            for (int third = 0; third < 3; third++) {
                double en = 1e-9;
                for (int i = 0; i < 192; i++) {
                    const long n = (long)g * 576 + third * 192 + i - 480 - 288;
                    const double v = (n >= 0 && n < nsamples) ? pcm[n * nch + ch] / 32768.0 : 0.0;
                    en += v * v;
                }
                if (en > 16.0 * prev_e && en > 192 * 1e-4) attack = 1;
                prev_e = en;
            }
            if (attack) btype[ch][g] = 2; /* (for now: "wants short windows") */
        }
        /* legal sequence with one granule of look-ahead: 0 -> {0, 1}, 1 -> 2, 2 -> {2, 3}, 3 -> {0, 1} */
        int prev = 0;
        for (int g = 0; g < ngran; g++) {
            const int want = btype[ch][g] == 2, want_next = g + 1 < ngran && btype[ch][g + 1] == 2;
            int bt;
            if (prev == 1) bt = 2;
            else if (prev == 2) bt = (want || want_next) ? 2 : 3;
            else bt = want_next ? 1 : 0;
            btype[ch][g] = (uint8_t)bt;
            prev = bt;
        }
    }
    for (int f = 0; f < nframes; f++) {
        int pad = 0;
        long num = 144L * kbps * 1000;
        pad_rest -= num % sample_rate;
        if (pad_rest < 0) { pad = 1; pad_rest += sample_rate; }
        const int frame_len = 144 * kbps * 1000 / sample_rate + pad, payload = frame_len - 4 - side_len;
        size_t tail = lstart - wcur;
        const int mdb = tail > 511 ? 511 : (int)tail;
        const size_t dstart = lstart - (size_t)mdb;
        long avail = 8L * (mdb + payload);
        bitw w = {logical, lcap * 8, dstart * 8};
        const size_t wstart = w.pos;
        gr_side gs[2][2];
        int u = 0;
        for (int gr = 0; gr < 2; gr++)
            for (int ch = 0; ch < nch; ch++, u++) {
                double x[576], xr[576];
                double energy = 0.0;
                for (int i = 0; i < 576; i++) {
                    const long n = (long)f * 1152 + gr * 576 + i;
                    x[i] = n < nsamples ? pcm[n * nch + ch] / 32768.0 : 0.0;
                }
                const int bt = btype[ch][2 * f + gr];
                enc_granule_spectrum(&ech[ch], x, xr, bt, row);
                for (int i = 0; i < 576; i++) energy += xr[i] * xr[i];
                /* budget: the mean of what is left, more for loud granules, less for quiet ones (those feed the reservoir) */
                const long left = avail - (long)(w.pos - wstart);
                long budget = left / (2 * nch - u);
                e_avg = e_avg * 0.9 + energy * 0.1;
                double fac = e_avg > 0.0 ? sqrt(energy / e_avg) : 1.0;
                if (fac < 0.5) fac = 0.5;
                if (fac > 1.5) fac = 1.5;
                budget = (long)(budget * fac);
                if (budget > left) budget = left;
                if (budget > 4095) budget = 4095;
                /* finest global gain whose spectrum fits: bits fall as the gain rises */
                int ix[576], lo = 0, hi = 255, best_gg = 255;
                enc_plan plan, best_plan;
                memset(&best_plan, 0, sizeof best_plan);
                while (lo <= hi) {
                    const int gg = (lo + hi) / 2;
                    const int mx = enc_quantise(xr, gg, ix);
                    const int bits = mx > 8191 + 15 ? (1 << 30) : enc_count(ix, row, &plan, bt);
                    if (bits <= budget) { best_gg = gg; best_plan = plan; hi = gg - 1; }
                    else lo = gg + 1;
                }
                enc_quantise(xr, best_gg, ix);
                if (enc_count(ix, row, &best_plan, bt) > budget) { /* (not even the coarsest fits: silence) */
                    memset(ix, 0, sizeof ix);
                    enc_count(ix, row, &best_plan, bt);
                }
                gr_side *s = &gs[gr][ch];
                memset(s, 0, sizeof *s);
                s->global_gain = best_gg;
                s->big_values = best_plan.bv2 / 2;
                s->table_select[0] = best_plan.tsel[0];
                s->table_select[1] = best_plan.tsel[1];
                s->table_select[2] = best_plan.tsel[2];
                s->region0_count = best_plan.a - 1;
                s->region1_count = best_plan.c - best_plan.a - 1;
                s->count1table = best_plan.c1tab;
                s->window_switching = bt != 0;
                s->block_type = bt;
                const size_t p0 = w.pos;
                for (int i = 0; i < best_plan.bv2; i += 2)
                    put_pair(&w, best_plan.tsel[i < best_plan.r1 ? 0 : (i < best_plan.r2 ? 1 : 2)], ix[i], ix[i + 1]);
                for (int i = best_plan.bv2; i < best_plan.c1end; i += 4) {
                    const int sym = (ix[i] != 0) << 3 | (ix[i + 1] != 0) << 2 | (ix[i + 2] != 0) << 1 | (ix[i + 3] != 0);
                    putbits(&w, l3_quad_hcod[best_plan.c1tab][sym], l3_quad_hlen[best_plan.c1tab][sym]);
                    for (int k = 0; k < 4; k++) if (ix[i + k]) putbits(&w, ix[i + k] < 0, 1);
                }
                s->part2_3_length = (int)(w.pos - p0);
            }
        wcur = dstart + (w.pos - wstart + 7) / 8;
        uint8_t *h = fr[f].hdr;
        h[0] = 0xFF;
        h[1] = (uint8_t)(0xE0 | (3 << 3) | (1 << 1) | 1);
        h[2] = (uint8_t)((bri << 4) | (row << 2) | (pad << 1));
        h[3] = (uint8_t)((nch == 1 ? 3 : 0) << 6);
        bitw sw = {fr[f].side, (size_t)side_len * 8, 0};
        putbits(&sw, (unsigned)mdb, 9);
        putbits(&sw, 0, nch == 1 ? 5 : 3);
        for (int ch = 0; ch < nch; ch++) putbits(&sw, 0, 4);
        for (int gr = 0; gr < 2; gr++)
            for (int ch = 0; ch < nch; ch++) {
                const gr_side *s = &gs[gr][ch];
                putbits(&sw, (unsigned)s->part2_3_length, 12);
                putbits(&sw, (unsigned)s->big_values, 9);
                putbits(&sw, (unsigned)s->global_gain, 8);
                putbits(&sw, 0, 4);  /* scalefac_compress 0: no scalefactor bits */
                putbits(&sw, (unsigned)s->window_switching, 1);
                if (s->window_switching) {
                    putbits(&sw, (unsigned)s->block_type, 2);
                    putbits(&sw, 0, 1); /* mixed_block_flag */
                    putbits(&sw, (unsigned)s->table_select[0], 5);
                    putbits(&sw, (unsigned)s->table_select[1], 5);
                    putbits(&sw, 0, 9); /* subblock_gain x 3 */
                } else {
                    for (int k = 0; k < 3; k++) putbits(&sw, (unsigned)s->table_select[k], 5);
                    putbits(&sw, (unsigned)s->region0_count, 4);
                    putbits(&sw, (unsigned)s->region1_count, 3);
                }
                putbits(&sw, 0, 1);  /* preflag */
                putbits(&sw, 0, 1);  /* scalefac_scale */
                putbits(&sw, (unsigned)s->count1table, 1);
            }
        fr[f].payload = payload;
        lstart += (size_t)payload;
    }
    size_t o = 0, lp = 0;
    for (int f = 0; f < nframes; f++) {
        const size_t need = 4 + (size_t)side_len + (size_t)fr[f].payload;
        if (o + need > cap) { o = 0; break; }
        memcpy(out + o, fr[f].hdr, 4);
        memcpy(out + o + 4, fr[f].side, (size_t)side_len);
        memcpy(out + o + 4 + side_len, logical + lp, (size_t)fr[f].payload);
        o += need;
        lp += (size_t)fr[f].payload;
    }
    free(logical);
    free(fr);
    free(ech);
    free(btype[0]);
    free(btype[1]);
    return o;
}
