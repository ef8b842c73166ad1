/* l3_oracle.c -- scalar, double-precision, from-spec MPEG-1/2 Layer III decoder.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the CPU oracle for the B200 decode path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call it; nothing under mp3_b200/ links or imports it.
 *
 * What it restates.  The reference repository (lxm0851/mp3) contains no source code
 * (/root/reference/README.md:1-84 is prose; the only technical statement is "audio player",
 * README.md:2), so there is no reference file:line to follow.  PARITY UNPINNED on the reference
 * side.  The algorithm restated here is the normative decoding process of ISO/IEC 11172-3
 * (clauses 2.4.1-2.4.3, Annex B tables) and ISO/IEC 13818-3 (2.4.3.2, LSF), written in the most
 * literal form (direct-form IMDCT and polyphase matrixing, no fast transforms).  It is pinned
 * instead by: the FFmpeg `mp3float` decoder found on the box (tests/test_oracle_vs_ffmpeg.py),
 * the known-answer vectors of SURVEY.md section 8(c) (tests/test_oracle_kat.py), and the table pins
 * of tests/test_tables_pin.py.
 *
 * Stage numbering in comments (a1..a11) is SURVEY.md section 8(a).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "iso_tables.h"
#include "iso_tables_l2.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MAXCH 2

/* ---------------------------------------------------------------- precomputed constants
 * The transforms stay in direct (definition) form; only the trigonometric / power constants and
 * the code-book trees are tabulated once so the CPU baseline is not dominated by libm calls. */
static double T_cos36[36][18], T_cos12[12][6], T_win[4][36], T_syn[64][32], T_dwin[512];
static double T_pow43[8207], T_cs[8], T_ca[8];
typedef struct { int16_t child[2]; int16_t sym; } hnode; /* sym >= 0 at a leaf */
static hnode *T_tree[32];
static hnode T_quad[2][64];
static volatile int T_ready = 0;

static void tree_insert(hnode *t, int *n, unsigned code, int len, int sym)
{
    int cur = 0;
    for (int i = len - 1; i >= 0; i--) {
        int bit = (code >> i) & 1;
        if (t[cur].child[bit] < 0) {
            t[*n].child[0] = t[*n].child[1] = -1;
            t[*n].sym = -1;
            t[cur].child[bit] = (int16_t)(*n)++;
        }
        cur = t[cur].child[bit];
    }
    t[cur].sym = (int16_t)sym;
}

static double imdct_window(int bt, int i);

void l3o_init(void)
{
    if (T_ready) return;
    for (int i = 0; i < 36; i++)
        for (int k = 0; k < 18; k++) T_cos36[i][k] = cos(M_PI / 72.0 * (2 * i + 1 + 18) * (2 * k + 1));
    for (int i = 0; i < 12; i++)
        for (int k = 0; k < 6; k++) T_cos12[i][k] = cos(M_PI / 24.0 * (2 * i + 1 + 6) * (2 * k + 1));
    for (int bt = 0; bt < 4; bt++)
        for (int i = 0; i < 36; i++) T_win[bt][i] = (bt == 2 && i >= 12) ? 0.0 : imdct_window(bt, i);
    for (int i = 0; i < 64; i++)
        for (int k = 0; k < 32; k++) T_syn[i][k] = cos((16 + i) * (2 * k + 1) * M_PI / 64.0);
    for (int i = 0; i < 512; i++) T_dwin[i] = l3_dwin(i);
    for (int i = 0; i < 8207; i++) T_pow43[i] = pow((double)i, 4.0 / 3.0);
    for (int i = 0; i < 8; i++) {
        double ci = l3_alias_ci[i];
        T_cs[i] = 1.0 / sqrt(1.0 + ci * ci);
        T_ca[i] = ci / sqrt(1.0 + ci * ci);
    }
    for (int book = 0; book < 32; book++) {
        const uint8_t *hlen;
        const uint32_t *hcod;
        int dim = l3_book(book, &hlen, &hcod);
        T_tree[book] = 0;
        if (!dim) continue;
        hnode *t = (hnode *)malloc(sizeof(hnode) * 2 * dim * dim);
        int n = 1;
        t[0].child[0] = t[0].child[1] = -1;
        t[0].sym = -1;
        for (int i = 0; i < dim * dim; i++) tree_insert(t, &n, hcod[i], hlen[i], ((i / dim) << 4) | (i % dim));
        T_tree[book] = t;
    }
    for (int q = 0; q < 2; q++) {
        int n = 1;
        T_quad[q][0].child[0] = T_quad[q][0].child[1] = -1;
        T_quad[q][0].sym = -1;
        for (int i = 0; i < 16; i++) tree_insert(T_quad[q], &n, l3_quad_hcod[q][i], l3_quad_hlen[q][i], i);
    }
    T_ready = 1;
}

typedef struct {
    int lsf;        /* 0 = MPEG-1, 1 = MPEG-2 LSF or MPEG-2.5 */
    int sr_row;     /* row into l3_sfb_* / l3_sample_rate */
    int nch;
    int mode, mode_ext;
    int crc;
    int frame_len;
    int side_len;
    int bitrate_idx, padding;
    int layer;      /* 3 = Layer III, 2 = Layer II, 1 = Layer I */
} l3o_hdr;

typedef struct {
    int part2_3_length, big_values, global_gain, scalefac_compress;
    int window_switching, block_type, mixed;
    int table_select[3], subblock_gain[3];
    int region0_count, region1_count;
    int preflag, scalefac_scale, count1table;
} l3o_gr;

typedef struct {
    int main_data_begin;
    int scfsi[MAXCH][4];
    l3o_gr gr[2][MAXCH];
} l3o_side;

/* ---------------------------------------------------------------- bit reader */
typedef struct {
    const uint8_t *p;
    size_t nbits; /* total bits available */
    size_t pos;
} bitr;

static unsigned getbits(bitr *b, int n)
{
    unsigned v = 0;
    for (int i = 0; i < n; i++) {
        unsigned bit = 0;
        if (b->pos < b->nbits) bit = (b->p[b->pos >> 3] >> (7 - (b->pos & 7))) & 1u;
        b->pos++;
        v = (v << 1) | bit;
    }
    return v;
}

/* ---------------------------------------------------------------- a1: header */
int l3o_parse_header(const uint8_t *p, l3o_hdr *h)
{
    if (p[0] != 0xFF || (p[1] & 0xE0) != 0xE0) return 0;
    int ver = (p[1] >> 3) & 3;   /* 3 = MPEG-1, 2 = MPEG-2, 0 = MPEG-2.5, 1 = reserved */
    int layer = (p[1] >> 1) & 3; /* 1 = Layer III, 2 = Layer II, 3 = Layer I */
    if (layer == 0) return 0;
    if (ver == 1) return 0;
    h->layer = 4 - layer;
    h->lsf = (ver != 3); /* MPEG-2.5 = the LSF syntax at half the MPEG-2 sample rates (rows 6..8) */
    h->crc = !(p[1] & 1);
    h->bitrate_idx = p[2] >> 4;
    int sri = (p[2] >> 2) & 3;
    if (h->bitrate_idx == 0 || h->bitrate_idx == 15 || sri == 3) return 0;
    h->padding = (p[2] >> 1) & 1;
    h->mode = p[3] >> 6;
    h->mode_ext = (p[3] >> 4) & 3;
    h->nch = h->mode == 3 ? 1 : 2;
    h->sr_row = sri + (ver == 3 ? 0 : ver == 2 ? 3 : 6);
    int sr = (int)l3_sample_rate[h->sr_row];
    if (h->layer == 1) { /* 384 samples per frame, slots of four bytes */
        h->frame_len = (12 * l1_bitrate_kbps[h->lsf][h->bitrate_idx] * 1000 / sr + h->padding) * 4;
        h->side_len = 0;
        return 1;
    }
    if (h->layer == 2) { /* 11172-3 2.4.3.1: 1152 samples per frame at every sample rate, no side info */
        h->frame_len = 144 * l2_bitrate_kbps[h->lsf][h->bitrate_idx] * 1000 / sr + h->padding;
        h->side_len = 0;
        return 1;
    }
    int br = l3_bitrate_kbps[h->lsf][h->bitrate_idx] * 1000;
    h->frame_len = (h->lsf ? 72 : 144) * br / sr + h->padding;
    h->side_len = h->lsf ? (h->nch == 1 ? 9 : 17) : (h->nch == 1 ? 17 : 32);
    return 1;
}

/* ---------------------------------------------------------------- a2: side info */
static void parse_side(const uint8_t *p, const l3o_hdr *h, l3o_side *s)
{
    bitr b = {p, (size_t)h->side_len * 8, 0};
    int ngr = h->lsf ? 1 : 2;
    memset(s, 0, sizeof *s);
    if (!h->lsf) {
        s->main_data_begin = getbits(&b, 9);
        getbits(&b, h->nch == 1 ? 5 : 3);
        for (int ch = 0; ch < h->nch; ch++)
            for (int i = 0; i < 4; i++) s->scfsi[ch][i] = getbits(&b, 1);
    } else {
        s->main_data_begin = getbits(&b, 8);
        getbits(&b, h->nch == 1 ? 1 : 2);
    }
    for (int gr = 0; gr < ngr; gr++)
        for (int ch = 0; ch < h->nch; ch++) {
            l3o_gr *g = &s->gr[gr][ch];
            g->part2_3_length = getbits(&b, 12);
            g->big_values = getbits(&b, 9);
            g->global_gain = getbits(&b, 8);
            g->scalefac_compress = getbits(&b, h->lsf ? 9 : 4);
            g->window_switching = getbits(&b, 1);
            if (g->window_switching) {
                g->block_type = getbits(&b, 2);
                g->mixed = getbits(&b, 1);
                g->table_select[0] = getbits(&b, 5);
                g->table_select[1] = getbits(&b, 5);
                g->table_select[2] = 0;
                for (int w = 0; w < 3; w++) g->subblock_gain[w] = getbits(&b, 3);
                g->region0_count = (g->block_type == 2 && !g->mixed) ? 8 : 7;
                g->region1_count = 36;
            } else {
                for (int r = 0; r < 3; r++) g->table_select[r] = getbits(&b, 5);
                g->region0_count = getbits(&b, 4);
                g->region1_count = getbits(&b, 3);
                g->block_type = 0;
                g->mixed = 0;
            }
            g->preflag = h->lsf ? 0 : getbits(&b, 1);
            g->scalefac_scale = getbits(&b, 1);
            g->count1table = getbits(&b, 1);
        }
}

/* ---------------------------------------------------------------- band layout
 * A granule's spectrum, in bitstream (un-reordered) order, is a list of bands: long sfbs first,
 * then (short sfb, window) triples.  Scalefactors are transmitted in exactly this order, so
 * sf[b] is the scalefactor of band b; the trailing bands (long sfb 21, short sfb 12) have none. */
typedef struct {
    int nbands;
    int n_long;          /* number of leading long bands */
    int start[40];       /* first line of band b */
    int width[40];
    int win[40];         /* window of a short band, -1 for long */
    int sfb[40];         /* sfb number (long or short numbering) */
} bandmap;

static void make_bandmap(const l3o_hdr *h, const l3o_gr *g, bandmap *m)
{
    const uint16_t *bl = l3_sfb_long[h->sr_row], *bs = l3_sfb_short[h->sr_row];
    int n = 0, pos = 0;
    memset(m, 0, sizeof *m);
    if (g->block_type != 2) {
        for (int s = 0; s < 22; s++) {
            m->start[n] = bl[s]; m->width[n] = bl[s + 1] - bl[s]; m->win[n] = -1; m->sfb[n] = s; n++;
        }
        m->n_long = 22;
    } else {
        int s0 = 0;
        if (g->mixed) {
            int nl = h->lsf ? 6 : 8;
            for (int s = 0; s < nl; s++) {
                m->start[n] = bl[s]; m->width[n] = bl[s + 1] - bl[s]; m->win[n] = -1; m->sfb[n] = s; n++;
            }
            m->n_long = nl;
            pos = bl[nl]; /* 36 */
            s0 = 3;
        }
        for (int s = s0; s < 13; s++) {
            int w = bs[s + 1] - bs[s];
            for (int win = 0; win < 3; win++) {
                m->start[n] = pos; m->width[n] = w; m->win[n] = win; m->sfb[n] = s; n++;
                pos += w;
            }
        }
    }
    m->nbands = n;
}

/* ---------------------------------------------------------------- a4: scalefactors
 * sf[] is in band order (see bandmap).  Bit 7 of an entry marks "illegal intensity position"
 * for the right channel of an LSF intensity-stereo frame (13818-3 2.4.3.2). */
static void read_scalefactors(bitr *b, const l3o_hdr *h, const l3o_side *si, int gr, int ch,
                              const uint8_t *sf_gr0, uint8_t *sf, int *preflag_out)
{
    const l3o_gr *g = &si->gr[gr][ch];
    memset(sf, 0, 40);
    *preflag_out = g->preflag;
    if (!h->lsf) {
        int s1 = l3_slen[0][g->scalefac_compress], s2 = l3_slen[1][g->scalefac_compress];
        if (g->block_type == 2) {
            int n1 = g->mixed ? 17 : 18, ntot = g->mixed ? 35 : 36;
            for (int i = 0; i < ntot; i++) sf[i] = (uint8_t)getbits(b, i < n1 ? s1 : s2);
        } else {
            static const int grp[5] = {0, 6, 11, 16, 21};
            for (int k = 0; k < 4; k++) {
                int reuse = gr == 1 && si->scfsi[ch][k];
                for (int s = grp[k]; s < grp[k + 1]; s++)
                    sf[s] = reuse ? (uint8_t)(sf_gr0[s] & 0x7f) : (uint8_t)getbits(b, s < 11 ? s1 : s2);
            }
        }
        return;
    }
    /* LSF */
    int sfc = g->scalefac_compress, slen[4], tbl;
    int ist = (h->mode_ext & 1) && ch == 1 && h->mode == 1;
    if (!ist) {
        if (sfc < 400) {
            slen[0] = (sfc >> 4) / 5; slen[1] = (sfc >> 4) % 5; slen[2] = (sfc & 15) >> 2; slen[3] = sfc & 3; tbl = 0;
        } else if (sfc < 500) {
            sfc -= 400;
            slen[0] = (sfc >> 2) / 5; slen[1] = (sfc >> 2) % 5; slen[2] = sfc & 3; slen[3] = 0; tbl = 1;
        } else {
            sfc -= 500;
            slen[0] = sfc / 3; slen[1] = sfc % 3; slen[2] = 0; slen[3] = 0; tbl = 2;
            *preflag_out = 1;
        }
    } else {
        sfc >>= 1;
        if (sfc < 180) {
            slen[0] = sfc / 36; slen[1] = (sfc % 36) / 6; slen[2] = sfc % 6; slen[3] = 0; tbl = 3;
        } else if (sfc < 244) {
            sfc -= 180;
            slen[0] = (sfc & 63) >> 4; slen[1] = (sfc & 15) >> 2; slen[2] = sfc & 3; slen[3] = 0; tbl = 4;
        } else {
            sfc -= 244;
            slen[0] = sfc / 3; slen[1] = sfc % 3; slen[2] = 0; slen[3] = 0; tbl = 5;
        }
    }
    int bt = g->block_type == 2 ? (g->mixed ? 2 : 1) : 0;
    int n = 0;
    for (int k = 0; k < 4; k++) {
        int cnt = l3_lsf_nsfb[tbl][bt][k];
        for (int i = 0; i < cnt; i++) {
            unsigned v = slen[k] ? getbits(b, slen[k]) : 0;
            if (ist && slen[k] && v == (1u << slen[k]) - 1u) v |= 0x80;
            sf[n++] = (uint8_t)v;
        }
    }
}

/* ---------------------------------------------------------------- a5: Huffman */
static int decode_tree(bitr *b, const hnode *t)
{
    int cur = 0;
    while (t[cur].sym < 0) {
        int nx = t[cur].child[getbits(b, 1)];
        if (nx < 0) return 0; /* cannot happen: all books are complete prefix codes */
        cur = nx;
    }
    return t[cur].sym;
}

static void huffman_decode(bitr *b, size_t part3_end, const l3o_hdr *h, const l3o_gr *g, int16_t *is)
{
    memset(is, 0, 576 * sizeof(int16_t));
    int bv2 = g->big_values * 2;
    if (bv2 > 576) bv2 = 576;
    int r1, r2;
    if (g->window_switching) {
        /* region0_count is 8 (short: 9 short-window bands = 3 * sfb_short[3] lines) or 7 (8 long bands) */
        r1 = g->block_type == 2 ? 3 * l3_sfb_short[h->sr_row][3] : l3_sfb_long[h->sr_row][8];
        r2 = 576;
    } else {
        int a = g->region0_count + 1, c = g->region0_count + g->region1_count + 2;
        if (a > 22) a = 22;
        if (c > 22) c = 22;
        r1 = l3_sfb_long[h->sr_row][a];
        r2 = l3_sfb_long[h->sr_row][c];
    }
    if (r1 > bv2) r1 = bv2;
    if (r2 > bv2) r2 = bv2;
    int i = 0;
    for (; i < bv2; i += 2) {
        int tsel = g->table_select[i < r1 ? 0 : (i < r2 ? 1 : 2)];
        int book = l3_book_of_table[tsel], lin = l3_linbits_of_table[tsel];
        if (book == 0) continue; /* zero book: no bits */
        if (b->pos >= part3_end) break;
        int sym = decode_tree(b, T_tree[book]);
        int x = sym >> 4, y = sym & 15;
        if (lin && x == 15) x += getbits(b, lin);
        if (x && getbits(b, 1)) x = -x;
        if (lin && y == 15) y += getbits(b, lin);
        if (y && getbits(b, 1)) y = -y;
        is[i] = (int16_t)x;
        is[i + 1] = (int16_t)y;
    }
    if (i < bv2) return; /* ran out of bits inside big_values: rest stays zero */
    /* count1 */
    i = bv2;
    while (i <= 572 && b->pos < part3_end) {
        int sym = decode_tree(b, T_quad[g->count1table]);
        int v[4] = {(sym >> 3) & 1, (sym >> 2) & 1, (sym >> 1) & 1, sym & 1};
        for (int k = 0; k < 4; k++)
            if (v[k] && getbits(b, 1)) v[k] = -1;
        if (b->pos > part3_end) break; /* overran part2_3_length: discard this quadruple */
        for (int k = 0; k < 4; k++) is[i + k] = (int16_t)v[k];
        i += 4;
    }
}

/* ---------------------------------------------------------------- a6: requantise */
static void requantise(const l3o_hdr *h, const l3o_gr *g, const bandmap *m, const uint8_t *sf,
                       int preflag, const int16_t *is, double *xr)
{
    (void)h;
    double mult = g->scalefac_scale ? 1.0 : 0.5;
    for (int b = 0; b < m->nbands; b++) {
        double e;
        int s = sf[b] & 0x7f;
        if (m->win[b] < 0)
            e = (g->global_gain - 210) / 4.0 - mult * (s + (preflag ? l3_pretab[m->sfb[b]] : 0));
        else
            e = (g->global_gain - 210 - 8 * g->subblock_gain[m->win[b]]) / 4.0 - mult * s;
        double gain = pow(2.0, e);
        for (int i = m->start[b]; i < m->start[b] + m->width[b]; i++) {
            int v = is[i];
            double a = T_pow43[v < 0 ? -v : v] * gain;
            xr[i] = v < 0 ? -a : a;
        }
    }
}

/* ---------------------------------------------------------------- a7: stereo (un-reordered domain) */
static void stereo(const l3o_hdr *h, const l3o_gr *g1, const bandmap *m, const uint8_t *sf1,
                   double *l, double *r)
{
    int ms = (h->mode_ext & 2) != 0, is = (h->mode_ext & 1) != 0;
    const double isq2 = 1.0 / sqrt(2.0);
    if (h->mode != 1 || (!ms && !is)) return;
    if (!is) {
        for (int i = 0; i < 576; i++) {
            double a = l[i], b = r[i];
            l[i] = (a + b) * isq2;
            r[i] = (a - b) * isq2;
        }
        return;
    }
    int found[3] = {0, 0, 0}, found_long = 0, first_long_done = 0;
    for (int b = m->nbands - 1; b >= 0; b--) {
        int w = m->win[b];
        int *fnd;
        if (w >= 0) fnd = &found[w];
        else {
            if (!first_long_done) { found_long = found[0] | found[1] | found[2]; first_long_done = 1; }
            fnd = &found_long;
        }
        int s0 = m->start[b], n = m->width[b];
        int intensity = 0;
        double kl = 0, kr = 0;
        if (!*fnd) {
            for (int i = s0; i < s0 + n; i++)
                if (r[i] != 0.0) { *fnd = 1; break; }
            if (!*fnd) {
                /* last band of each kind borrows the previous band's position */
                int bsf = b;
                if (w >= 0 && m->sfb[b] == 12) bsf = b - 3;
                if (w < 0 && m->sfb[b] == 21) bsf = b - 1;
                int p = sf1[bsf];
                if (!h->lsf) {
                    if (p < 7) {
                        intensity = 1;
                        if (p == 6) { kl = 1.0; kr = 0.0; }
                        else { double t = tan(p * M_PI / 12.0); kl = t / (1.0 + t); kr = 1.0 / (1.0 + t); }
                    }
                } else if (!(p & 0x80)) {
                    intensity = 1;
                    double i0 = pow(2.0, -((g1->scalefac_compress & 1) + 1) / 4.0);
                    kl = kr = 1.0;
                    if (p & 1) kl = pow(i0, (p + 1) / 2);
                    else kr = pow(i0, p / 2);
                }
            }
        }
        if (intensity) {
            for (int i = s0; i < s0 + n; i++) { double a = l[i]; l[i] = a * kl; r[i] = a * kr; }
        } else if (ms) {
            for (int i = s0; i < s0 + n; i++) {
                double a = l[i], c = r[i];
                l[i] = (a + c) * isq2;
                r[i] = (a - c) * isq2;
            }
        }
    }
}

/* ---------------------------------------------------------------- a8: reorder + alias */
static void reorder(const bandmap *m, double *xr)
{
    double t[576];
    memcpy(t, xr, sizeof t);
    for (int b = m->n_long; b < m->nbands; b += 3) {
        int w = m->width[b], s0 = m->start[b];
        for (int win = 0; win < 3; win++)
            for (int k = 0; k < w; k++) xr[s0 + 3 * k + win] = t[s0 + win * w + k];
    }
}

static void alias_reduce(const l3o_gr *g, double *xr)
{
    int nb = 31;
    if (g->block_type == 2) nb = g->mixed ? 1 : 0;
    for (int sb = 1; sb <= nb; sb++)
        for (int i = 0; i < 8; i++) {
            double cs = T_cs[i], ca = T_ca[i];
            double lo = xr[sb * 18 - 1 - i], hi = xr[sb * 18 + i];
            xr[sb * 18 - 1 - i] = lo * cs - hi * ca;
            xr[sb * 18 + i] = hi * cs + lo * ca;
        }
}

/* ---------------------------------------------------------------- a9/a10: IMDCT, overlap, inversion */
static double imdct_window(int bt, int i)
{
    switch (bt) {
    case 0: return sin(M_PI / 36.0 * (i + 0.5));
    case 1:
        if (i < 18) return sin(M_PI / 36.0 * (i + 0.5));
        if (i < 24) return 1.0;
        if (i < 30) return sin(M_PI / 12.0 * (i - 18 + 0.5));
        return 0.0;
    case 3:
        if (i < 6) return 0.0;
        if (i < 12) return sin(M_PI / 12.0 * (i - 6 + 0.5));
        if (i < 18) return 1.0;
        return sin(M_PI / 36.0 * (i + 0.5));
    default: return sin(M_PI / 12.0 * (i + 0.5));
    }
}

static void hybrid(const l3o_gr *g, const double *xr, double *overlap /*[32][18]*/, double *sbs /*[18][32]*/)
{
    for (int sb = 0; sb < 32; sb++) {
        double out[36];
        int bt = g->block_type;
        if (bt == 2 && g->mixed && sb < 2) bt = 0;
        if (bt != 2) {
            for (int i = 0; i < 36; i++) {
                double s = 0;
                for (int k = 0; k < 18; k++)
                    s += xr[sb * 18 + k] * T_cos36[i][k];
                out[i] = s * T_win[bt][i];
            }
        } else {
            for (int i = 0; i < 36; i++) out[i] = 0;
            for (int w = 0; w < 3; w++)
                for (int i = 0; i < 12; i++) {
                    double s = 0;
                    for (int k = 0; k < 6; k++)
                        s += xr[sb * 18 + 3 * k + w] * T_cos12[i][k];
                    out[6 + 6 * w + i] += s * T_win[2][i];
                }
        }
        for (int i = 0; i < 18; i++) {
            double v = out[i] + overlap[sb * 18 + i];
            overlap[sb * 18 + i] = out[18 + i];
            if ((sb & 1) && (i & 1)) v = -v;
            sbs[i * 32 + sb] = v;
        }
    }
}

/* ---------------------------------------------------------------- a11: polyphase synthesis */
typedef struct {
    double v[1024];
} synth_state;

static void synth_slot(synth_state *st, const double *s /*[32]*/, double *pcm /*[32]*/)
{
    memmove(st->v + 64, st->v, 960 * sizeof(double));
    for (int i = 0; i < 64; i++) {
        double a = 0;
        for (int k = 0; k < 32; k++) a += T_syn[i][k] * s[k];
        st->v[i] = a;
    }
    for (int j = 0; j < 32; j++) {
        double a = 0;
        for (int i = 0; i < 8; i++) {
            a += st->v[128 * i + j] * T_dwin[64 * i + j];
            a += st->v[128 * i + 96 + j] * T_dwin[64 * i + 32 + j];
        }
        pcm[j] = a;
    }
}

/* ---------------------------------------------------------------- stream level */
typedef struct {
    size_t off;          /* byte offset of the header in the input */
    size_t payload_off;  /* offset of this frame's main-data bytes in the arena */
    l3o_hdr h;
} frame_ent;

typedef struct {
    int sample_rate, channels, lsf;
    long frames, units, samples; /* samples per channel */
    long concealed_frames;
} l3o_info;

/* Scan policy (shared, by specification, with the product's indexer -- see DESIGN.md):
 * an optional ID3v2 tag at offset 0 is skipped; then at position p a frame exists iff the 4 bytes
 * parse as a Layer III MPEG-1/2 header with the same version / sample rate / channel count as the
 * stream's first frame, and the whole frame fits in the buffer; otherwise p advances by one. */
static size_t id3v2_skip(const uint8_t *buf, size_t len)
{
    if (len >= 10 && buf[0] == 'I' && buf[1] == 'D' && buf[2] == '3' && !((buf[6] | buf[7] | buf[8] | buf[9]) & 0x80)) {
        size_t n = ((size_t)buf[6] << 21) | ((size_t)buf[7] << 14) | ((size_t)buf[8] << 7) | buf[9];
        n += 10 + ((buf[5] & 0x10) ? 10 : 0);
        if (n <= len) return n;
    }
    return 0;
}

/* ---------------------------------------------------------------- Layer II (11172-3 2.4.1.6, 2.4.3.3)
 * One frame = bit allocation, scfsi, scalefactors, then 12 groups of 3 samples for each of the
 * sblimit subbands (joint stereo: subbands >= bound carry one set of codes for both channels).
 * Requantisation (2.4.3.3.4): a code c of a class with `steps` levels is the fraction
 * (2 c + 1 - steps) / steps, times the scalefactor 2^(1 - index / 3) of its third of the frame.
 * Output: sb[ch][slot 0..35][subband 0..31]. */
static void l2_decode_frame(const uint8_t *frame, const l3o_hdr *h, double sb[MAXCH][36][32])
{
    int nch = h->nch;
    int kbps = l2_bitrate_kbps[h->lsf][h->bitrate_idx];
    int tbl = l2_select_table(h->lsf, (int)l3_sample_rate[h->sr_row], kbps, nch);
    int sblimit = l2_sblimit[tbl];
    int bound = h->mode == 1 ? (h->mode_ext + 1) * 4 : sblimit;
    if (bound > sblimit) bound = sblimit;
    if (nch == 1) bound = sblimit;
    bitr b = {frame, (size_t)h->frame_len * 8, (size_t)(4 + (h->crc ? 2 : 0)) * 8};
    int alloc[MAXCH][32], scfsi[MAXCH][32], scf[MAXCH][32][3];
    memset(alloc, 0, sizeof alloc);
    memset(scf, 0, sizeof scf);
    memset(sb, 0, sizeof(double) * MAXCH * 36 * 32);
    for (int s = 0; s < sblimit; s++) {
        const uint8_t *row = l2_rows[l2_row_of_sb[tbl][s]];
        if (s < bound)
            for (int ch = 0; ch < nch; ch++) alloc[ch][s] = (int)getbits(&b, row[0]);
        else
            alloc[0][s] = alloc[1][s] = (int)getbits(&b, row[0]);
    }
    for (int s = 0; s < sblimit; s++)
        for (int ch = 0; ch < nch; ch++)
            if (alloc[ch][s]) scfsi[ch][s] = (int)getbits(&b, 2);
    for (int s = 0; s < sblimit; s++)
        for (int ch = 0; ch < nch; ch++) {
            if (!alloc[ch][s]) continue;
            int *f = scf[ch][s];
            switch (scfsi[ch][s]) {
            case 0: f[0] = (int)getbits(&b, 6); f[1] = (int)getbits(&b, 6); f[2] = (int)getbits(&b, 6); break;
            case 1: f[0] = f[1] = (int)getbits(&b, 6); f[2] = (int)getbits(&b, 6); break;
            case 2: f[0] = f[1] = f[2] = (int)getbits(&b, 6); break;
            default: f[0] = (int)getbits(&b, 6); f[1] = f[2] = (int)getbits(&b, 6); break;
            }
        }
    for (int gr = 0; gr < 12; gr++)
        for (int s = 0; s < sblimit; s++) {
            int nc = s < bound ? nch : 1;
            for (int ch = 0; ch < nc; ch++) {
                int a = alloc[ch][s];
                if (!a) continue;
                const uint8_t *row = l2_rows[l2_row_of_sb[tbl][s]];
                int q = row[a], steps = l2_quant_steps[q], bits = l2_quant_bits[q];
                int code[3];
                if (bits < 0) {
                    unsigned c = getbits(&b, -bits);
                    code[0] = (int)(c % (unsigned)steps);
                    c /= (unsigned)steps;
                    code[1] = (int)(c % (unsigned)steps);
                    code[2] = (int)(c / (unsigned)steps);
                } else
                    for (int i = 0; i < 3; i++) code[i] = (int)getbits(&b, bits);
                for (int i = 0; i < 3; i++) {
                    double fr = (double)(2 * code[i] + 1 - steps) / (double)steps;
                    for (int c2 = ch; c2 < (s < bound ? ch + 1 : nch); c2++) {
                        int idx = scf[c2][s][gr >> 2];
                        sb[c2][gr * 3 + i][s] = idx < 63 ? fr * pow(2.0, 1.0 - idx / 3.0) : 0.0;
                    }
                }
            }
        }
}

/* ---------------------------------------------------------------- Layer I (11172-3 2.4.1.5, 2.4.3.2)
 * 4-bit allocation per subband and channel (joint stereo: shared above the bound), one 6-bit scalefactor
 * per allocated subband, then 12 samples of allocation + 1 bits; same requantisation formula as Layer II
 * with steps = 2^bits - 1.  Output: sb[ch][slot 0..11][subband]. */
static void l1_decode_frame(const uint8_t *frame, const l3o_hdr *h, double sb[MAXCH][36][32])
{
    int nch = h->nch;
    int bound = (h->mode == 1 && nch == 2) ? (h->mode_ext + 1) * 4 : 32;
    bitr b = {frame, (size_t)h->frame_len * 8, (size_t)(4 + (h->crc ? 2 : 0)) * 8};
    int alloc[MAXCH][32], scf[MAXCH][32];
    memset(sb, 0, sizeof(double) * MAXCH * 36 * 32);
    memset(alloc, 0, sizeof alloc);
    memset(scf, 0, sizeof scf);
    for (int s = 0; s < 32; s++) {
        if (s < bound)
            for (int ch = 0; ch < nch; ch++) alloc[ch][s] = (int)getbits(&b, 4);
        else
            alloc[0][s] = alloc[1][s] = (int)getbits(&b, 4);
    }
    for (int s = 0; s < 32; s++)
        for (int ch = 0; ch < nch; ch++)
            if (alloc[ch][s]) scf[ch][s] = (int)getbits(&b, 6);
    for (int t = 0; t < 12; t++)
        for (int s = 0; s < 32; s++) {
            int nc = s < bound ? nch : 1;
            for (int ch = 0; ch < nc; ch++) {
                int a = alloc[ch][s];
                if (!a) continue;
                int bits = a + 1, steps = (1 << bits) - 1;
                int code = (int)getbits(&b, bits);
                double fr = (double)(2 * code + 1 - steps) / (double)steps;
                if (a == 15) fr = 0.0; /* forbidden allocation: no defined value */
                for (int c2 = ch; c2 < (s < bound ? ch + 1 : nch); c2++)
                    sb[c2][t][s] = scf[c2][s] < 63 ? fr * pow(2.0, 1.0 - scf[c2][s] / 3.0) : 0.0;
            }
        }
}

/* 11172-3 2.4.3.1 error check: CRC-16, generator polynomial x^16 + x^15 + x^2 + 1, shift register preset
 * to all ones, fed bit by bit (MSB first) with header bits 16..31 and the side information.
 * Off by default (the word is skipped); l3o_set_verify_crc(1) makes a mismatch conceal the frame. */
static int g_verify_crc = 0;
void l3o_set_verify_crc(int on) { g_verify_crc = on; }
unsigned l3o_crc16(unsigned crc, const uint8_t *p, size_t nbits)
{
    for (size_t i = 0; i < nbits; i++) {
        unsigned in = (p[i >> 3] >> (7 - (i & 7))) & 1u;
        unsigned fb = ((crc >> 15) & 1u) ^ in;
        crc = (crc << 1) & 0xffffu;
        if (fb) crc ^= 0x8005u;
    }
    return crc;
}

/* Decode a whole buffer.  Any output pointer may be NULL.
 *   pcm      interleaved double, full scale +-1.0, cap_samples per channel
 *   dump_is  int16  [unit][576]   Huffman output (a5), bitstream order
 *   dump_sf  uint8  [unit][40]    scalefactors in band order (a4)
 *   dump_xr  double [unit][576]   after requant+stereo+reorder+alias (a6-a8)
 *   dump_sb  double [unit][18][32] subband samples after IMDCT/overlap/inversion (a9-a10)
 * units are ordered [frame][granule][channel].  Returns 0, or <0 on error. */
int l3o_decode(const uint8_t *buf, size_t len, l3o_info *info, double *pcm, size_t cap_samples,
               int16_t *dump_is, uint8_t *dump_sf, double *dump_xr, double *dump_sb)
{
    memset(info, 0, sizeof *info);
    size_t nfr = 0, capfr = 1024;
    frame_ent *fr = (frame_ent *)malloc(capfr * sizeof *fr);
    uint8_t *arena = (uint8_t *)malloc(len + 8);
    size_t arena_len = 0;
    l3o_hdr first;
    memset(&first, 0, sizeof first);
    int have_first = 0;
    size_t p = id3v2_skip(buf, len);
    while (p + 4 <= len) {
        l3o_hdr h;
        if (!l3o_parse_header(buf + p, &h) ||
            (have_first && (h.lsf != first.lsf || h.sr_row != first.sr_row || h.nch != first.nch || h.layer != first.layer)) ||
            h.frame_len < 4 + (h.crc ? 2 : 0) + h.side_len || p + (size_t)h.frame_len > len) {
            p++;
            continue;
        }
        if (!have_first) {
            /* sync confirmation: the first frame must be followed by a header of the same stream when
             * there are bytes to check (a lone valid-looking word in junk does not start a stream) */
            size_t q = p + (size_t)h.frame_len;
            l3o_hdr h2;
            if (q + 4 <= len && (!l3o_parse_header(buf + q, &h2) || h2.lsf != h.lsf || h2.sr_row != h.sr_row ||
                                 h2.nch != h.nch || h2.layer != h.layer)) {
                p++;
                continue;
            }
            first = h;
            have_first = 1;
        }
        if (nfr == capfr) { capfr *= 2; fr = (frame_ent *)realloc(fr, capfr * sizeof *fr); }
        fr[nfr].off = p;
        fr[nfr].h = h;
        fr[nfr].payload_off = arena_len;
        size_t skip = h.layer != 3 ? (size_t)h.frame_len : (size_t)(4 + (h.crc ? 2 : 0) + h.side_len);
        memcpy(arena + arena_len, buf + p + skip, h.frame_len - skip);
        arena_len += h.frame_len - skip;
        nfr++;
        p += h.frame_len;
    }
    memset(arena + arena_len, 0, 8);
    if (!have_first) { free(fr); free(arena); return -1; }
    int nch = first.nch, ngr = (first.lsf && first.layer == 3) ? 1 : 2;
    info->sample_rate = (int)l3_sample_rate[first.sr_row];
    info->channels = nch;
    info->lsf = first.lsf;
    info->frames = (long)nfr;
    info->units = (long)nfr * ngr * nch;
    info->samples = (long)nfr * ngr * 576;
    if (first.layer == 1) { /* 384 samples per frame: the units (576 samples) are granules of the slot sequence */
        info->samples = (long)nfr * 384;
        info->units = (long)((nfr * 12 + 17) / 18) * nch;
    }

    double(*overlap)[576] = (double(*)[576])calloc(MAXCH, sizeof(double[576]));
    synth_state *syn = (synth_state *)calloc(MAXCH, sizeof(synth_state));
    l3o_init();

    long unit = 0;
    for (size_t f = 0; f < nfr; f++) {
        const l3o_hdr *h = &fr[f].h;
        if (h->layer == 1) { /* Layer I: 12 slots per frame into the same synthesis */
            static __thread double l1sb[MAXCH][36][32];
            l1_decode_frame(buf + fr[f].off, h, l1sb);
            for (int t = 0; t < 12; t++) {
                size_t slot = f * 12 + (size_t)t;
                for (int ch = 0; ch < nch; ch++) {
                    long u = (long)(slot / 18) * nch + ch;
                    if (dump_sb) memcpy(dump_sb + u * 576 + (slot % 18) * 32, &l1sb[ch][t][0], 32 * sizeof(double));
                    double out[32];
                    synth_slot(&syn[ch], &l1sb[ch][t][0], out);
                    if (pcm)
                        for (int j = 0; j < 32; j++) {
                            size_t n = slot * 32 + (size_t)j;
                            if (n < cap_samples) pcm[n * nch + ch] = out[j];
                        }
                }
            }
            continue;
        }
        if (h->layer == 2) { /* Layer II: subband samples straight from the frame, then the same synthesis */
            static __thread double l2sb[MAXCH][36][32];
            l2_decode_frame(buf + fr[f].off, h, l2sb);
            for (int gr = 0; gr < 2; gr++) {
                for (int ch = 0; ch < nch; ch++) {
                    long u = unit + ch;
                    if (dump_is) memset(dump_is + u * 576, 0, 576 * sizeof(int16_t));
                    if (dump_sf) memset(dump_sf + u * 40, 0, 40);
                    if (dump_xr) memset(dump_xr + u * 576, 0, 576 * sizeof(double));
                    if (dump_sb) memcpy(dump_sb + u * 576, &l2sb[ch][gr * 18][0], 576 * sizeof(double));
                    size_t base = (f * 2 + gr) * 576;
                    for (int t = 0; t < 18; t++) {
                        double out[32];
                        synth_slot(&syn[ch], &l2sb[ch][gr * 18 + t][0], out);
                        if (pcm)
                            for (int j = 0; j < 32; j++) {
                                size_t n = base + t * 32 + j;
                                if (n < cap_samples) pcm[n * nch + ch] = out[j];
                            }
                    }
                }
                unit += nch;
            }
            continue;
        }
        l3o_side si;
        parse_side(buf + fr[f].off + 4 + (h->crc ? 2 : 0), h, &si);
        int ok = (size_t)si.main_data_begin <= fr[f].payload_off;
        if (g_verify_crc && h->crc) {
            const uint8_t *fp = buf + fr[f].off;
            unsigned crc = l3o_crc16(0xffffu, fp + 2, 16);
            crc = l3o_crc16(crc, fp + 6, (size_t)h->side_len * 8);
            if (crc != (((unsigned)fp[4] << 8) | fp[5])) ok = 0;
        }
        if (ok) { /* a frame's main data ends inside the frame (the next frame's main_data_begin cannot point
                   * forward, 11172-3 2.4.2.7): side info claiming more bits than that is damaged */
            size_t sum = 0;
            for (int gr = 0; gr < ngr; gr++)
                for (int ch = 0; ch < nch; ch++) sum += (size_t)si.gr[gr][ch].part2_3_length;
            size_t own = (size_t)h->frame_len - 4 - (h->crc ? 2 : 0) - (size_t)h->side_len;
            if ((fr[f].payload_off - si.main_data_begin) * 8 + sum > (fr[f].payload_off + own) * 8) ok = 0;
        }
        if (!ok) info->concealed_frames++;
        size_t bitpos = ok ? (fr[f].payload_off - si.main_data_begin) * 8 : 0;
        uint8_t sf[2][MAXCH][40];
        for (int gr = 0; gr < ngr; gr++) {
            int16_t is[MAXCH][576];
            double xr[MAXCH][576];
            int preflag[MAXCH];
            bandmap bm[MAXCH];
            for (int ch = 0; ch < nch; ch++) {
                const l3o_gr *g = &si.gr[gr][ch];
                make_bandmap(h, g, &bm[ch]);
                if (ok) {
                    bitr b = {arena, arena_len * 8, bitpos};
                    size_t end = bitpos + g->part2_3_length;
                    read_scalefactors(&b, h, &si, gr, ch, sf[0][ch], sf[gr][ch], &preflag[ch]);
                    huffman_decode(&b, end, h, g, is[ch]);
                    bitpos = end;
                } else {
                    memset(sf[gr][ch], 0, 40);
                    memset(is[ch], 0, sizeof is[ch]);
                    preflag[ch] = 0;
                }
                requantise(h, g, &bm[ch], sf[gr][ch], preflag[ch], is[ch], xr[ch]);
            }
            if (nch == 2 && ok) stereo(h, &si.gr[gr][1], &bm[1], sf[gr][1], xr[0], xr[1]);
            for (int ch = 0; ch < nch; ch++) {
                const l3o_gr *g = &si.gr[gr][ch];
                double sbs[576];
                if (g->block_type == 2) reorder(&bm[ch], xr[ch]);
                alias_reduce(g, xr[ch]);
                hybrid(g, xr[ch], overlap[ch], sbs);
                long u = unit + ch;
                if (dump_is) memcpy(dump_is + u * 576, is[ch], 576 * sizeof(int16_t));
                if (dump_sf) memcpy(dump_sf + u * 40, sf[gr][ch], 40);
                if (dump_xr) memcpy(dump_xr + u * 576, xr[ch], 576 * sizeof(double));
                if (dump_sb) memcpy(dump_sb + u * 576, sbs, 576 * sizeof(double));
                size_t base = (f * ngr + gr) * 576;
                for (int t = 0; t < 18; t++) {
                    double out[32];
                    synth_slot(&syn[ch], sbs + t * 32, out);
                    if (pcm)
                        for (int j = 0; j < 32; j++) {
                            size_t n = base + t * 32 + j;
                            if (n < cap_samples) pcm[n * nch + ch] = out[j];
                        }
                }
            }
            unit += nch;
        }
    }
    free(fr);
    free(arena);
    free(overlap);
    free(syn);
    return 0;
}

/* ---------------------------------------------------------------- tag frame (container step before a1)
 * Restated from the published layouts of the Xing/Info header (Xing Technology's VBR header SDK),
 * the LAME tag (LAME "Mp3 info tag rev 1" specification) and Fraunhofer's VBRI header; the
 * reference repository has no code for them (/root/reference/README.md:1-84).
 *   Xing/Info: 4-byte id right after the side info; 4-byte flags (1 frames, 2 bytes, 4 TOC, 8 quality);
 *              the fields present, in that order (4, 4, 100, 4 bytes).
 *   LAME ext : directly after; 9-byte encoder version, 1 rev/method, 1 lowpass, 8 replay gain,
 *              1 flags, 1 bitrate, then 3 bytes = 12-bit encoder delay, 12-bit end padding.
 *   VBRI     : always 32 bytes after the 4-byte header: id, version(2), delay(2), quality(2),
 *              bytes(4), frames(4), ...
 * Gapless window (what mpg123 and FFmpeg derive from the same fields): the tag frame is dropped; with
 * a LAME extension the next delay + 528 + 1 samples are dropped too and the padding comes off the end. */
typedef struct {
    int kind;     /* 0 none, 1 Xing, 2 Info, 3 VBRI */
    int has_lame;
    unsigned frames, bytes;
    int enc_delay, enc_padding;
    long first_sample, num_samples;
} l3o_tag;

static unsigned be32(const uint8_t *p) { return ((unsigned)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

int l3o_parse_tag(const uint8_t *buf, size_t len, l3o_tag *t)
{
    memset(t, 0, sizeof *t);
    size_t p = id3v2_skip(buf, len);
    l3o_hdr h;
    for (;; p++) { /* the first decodable frame */
        if (p + 4 > len) return -1;
        if (l3o_parse_header(buf + p, &h) && h.frame_len >= 4 + (h.crc ? 2 : 0) + h.side_len &&
            p + (size_t)h.frame_len <= len) {
            size_t q = p + (size_t)h.frame_len;
            l3o_hdr h2;
            if (q + 4 > len || (l3o_parse_header(buf + q, &h2) && h2.lsf == h.lsf && h2.sr_row == h.sr_row &&
                                h2.nch == h.nch && h2.layer == h.layer))
                break;
        }
    }
    const uint8_t *f = buf + p;
    size_t flen = (size_t)h.frame_len;
    /* count the frames of the stream the way l3o_decode does */
    long nfr = 0;
    {
        size_t q = p;
        while (q + 4 <= len) {
            l3o_hdr g;
            if (!l3o_parse_header(buf + q, &g) || g.lsf != h.lsf || g.sr_row != h.sr_row || g.nch != h.nch ||
                g.layer != h.layer ||
                g.frame_len < 4 + (g.crc ? 2 : 0) + g.side_len || q + (size_t)g.frame_len > len) {
                q++;
                continue;
            }
            nfr++;
            q += (size_t)g.frame_len;
        }
    }
    size_t at = 4 + (size_t)h.side_len;
    for (int pass = 0; pass < 2 && t->kind == 0; pass++, at += 2) {
        if (pass == 1 && !h.crc) break;
        if (at + 8 > flen) break;
        if (memcmp(f + at, "Xing", 4) == 0) t->kind = 1;
        else if (memcmp(f + at, "Info", 4) == 0) t->kind = 2;
        else continue;
        unsigned flags = be32(f + at + 4);
        size_t q = at + 8;
        if ((flags & 1) && q + 4 <= flen) { t->frames = be32(f + q); q += 4; }
        if ((flags & 2) && q + 4 <= flen) { t->bytes = be32(f + q); q += 4; }
        if (flags & 4) q += 100;
        if (flags & 8) q += 4;
        if (q + 24 <= flen && f[q] >= 0x20 && f[q] <= 0x7e && f[q + 1] >= 0x20 && f[q + 1] <= 0x7e) {
            t->has_lame = 1;
            t->enc_delay = (f[q + 21] << 4) | (f[q + 22] >> 4);
            t->enc_padding = ((f[q + 22] & 0x0f) << 8) | f[q + 23];
        }
    }
    if (t->kind == 0 && flen >= 4 + 32 + 18 && memcmp(f + 36, "VBRI", 4) == 0) {
        t->kind = 3;
        t->enc_delay = (f[42] << 8) | f[43];
        t->bytes = be32(f + 46);
        t->frames = be32(f + 50);
    }
    long spf = h.layer == 1 ? 384 : ((h.lsf && h.layer == 3) ? 576 : 1152), total = nfr * spf;
    long start = 0, count = total;
    if (t->kind) {
        start = spf;
        count = total - spf;
        if (t->has_lame) {
            start += t->enc_delay + 528 + 1;
            count -= t->enc_delay + t->enc_padding;
        }
    }
    if (start > total) start = total;
    if (count > total - start) count = total - start;
    if (count < 0) count = 0;
    t->first_sample = start;
    t->num_samples = count;
    return 0;
}

/* Table accessors so that tests can pin the ISO tables from Python. */
int l3o_book_entry(int book, int x, int y, int *hlen, unsigned *hcod)
{
    const uint8_t *l;
    const uint32_t *c;
    int dim = l3_book(book, &l, &c);
    if (!dim || x >= dim || y >= dim) return 0;
    *hlen = l[x * dim + y];
    *hcod = c[x * dim + y];
    return dim;
}
double l3o_dwin(int i) { return l3_dwin(i); }
int l3o_sfb_long(int row, int i) { return l3_sfb_long[row][i]; }
int l3o_sfb_short(int row, int i) { return l3_sfb_short[row][i]; }
