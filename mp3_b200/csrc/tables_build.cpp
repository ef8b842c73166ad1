// tables_build.cpp -- builds the decode-side lookup tables from the ISO tables (host, once).
#include <cmath>
#include <cstring>
#include <vector>

#include "iso_tables.h"
#include "l3_tables.h"

namespace {

struct Code { int len; uint32_t code; int sym; };

// Recursive multi-level LUT builder.  `lut` is the book-local table (offsets relative to 0).
void fill_level(std::vector<uint16_t> &lut, size_t at, int width, int consumed, const std::vector<Code> &items)
{
    const int size = 1 << width;
    std::vector<std::vector<Code>> deeper(size);
    for (const Code &c : items) {
        int rem = c.len - consumed;
        if (rem <= width) {
            // leaf: replicate over the don't-care bits
            uint32_t hi = (c.code & ((1u << rem) - 1u)) << (width - rem);
            for (uint32_t k = 0; k < (1u << (width - rem)); k++)
                lut[at + hi + k] = (uint16_t)((rem << 8) | c.sym);
        } else {
            uint32_t key = (c.code >> (rem - width)) & (size - 1);
            deeper[key].push_back(c);
        }
    }
    for (int key = 0; key < size; key++) {
        if (deeper[key].empty()) continue;
        int mx = 0;
        for (const Code &c : deeper[key]) mx = std::max(mx, c.len - consumed - width);
        int w = std::min(mx, L3_HUFF_SUB_BITS);
        size_t sub = lut.size();
        lut.resize(sub + ((size_t)1 << w), 0);
        lut[at + key] = (uint16_t)(0x8000u | (w << 11) | (uint16_t)sub);
        fill_level(lut, sub, w, consumed + width, deeper[key]);
    }
}

} // namespace

extern "C" void l3_build_host_tables(L3HostTables *t)
{
    std::memset(t, 0, sizeof *t);
    // ---- Huffman
    uint32_t total = 0;
    uint16_t book_base[32];
    uint8_t book_root[32];
    std::memset(book_base, 0, sizeof book_base);
    std::memset(book_root, 0, sizeof book_root);
    for (int book = 1; book < 32; book++) {
        const uint8_t *hlen;
        const uint32_t *hcod;
        int dim = l3_book(book, &hlen, &hcod);
        if (!dim) continue;
        std::vector<Code> items;
        int mx = 0;
        for (int i = 0; i < dim * dim; i++) {
            items.push_back({hlen[i], hcod[i], ((i / dim) << 4) | (i % dim)});
            mx = std::max(mx, (int)hlen[i]);
        }
        int root = std::min(mx, L3_HUFF_ROOT_BITS);
        std::vector<uint16_t> lut((size_t)1 << root, 0);
        fill_level(lut, 0, root, 0, items);
        if (total + lut.size() + 2 > L3_HUFF_LUT_MAX || lut.size() > 2048) return; // cannot happen
        std::memcpy(t->huff_lut + total, lut.data(), lut.size() * sizeof(uint16_t));
        book_base[book] = (uint16_t)total;
        book_root[book] = (uint8_t)root;
        total += (uint32_t)lut.size();
    }
    // the empty book (table_select 0, 4, 14) as a real one-bit table: two leaves (0, 0) of length 0,
    // so the decoder needs no special case for it
    t->huff_lut[total] = t->huff_lut[total + 1] = 0;
    const uint32_t empty_base = total;
    total += 2;
    t->huff_lut_len = total;
    for (int ts = 0; ts < 32; ts++) {
        int book = l3_book_of_table[ts];
        t->huff.base[ts] = book ? book_base[book] : (uint16_t)empty_base;
        t->huff.root[ts] = book ? book_root[book] : 1;
        t->huff.linbits[ts] = l3_linbits_of_table[ts];
    }
    for (int v = 0; v < 64; v++)
        for (int s = 0; s < 16; s++) {
            int len = l3_quad_hlen[0][s];
            if ((v >> (6 - len)) == l3_quad_hcod[0][s]) t->quad_a[v] = (uint8_t)((len << 4) | s);
        }
    // ---- band layouts
    for (int row = 0; row < 9; row++) {
        const uint16_t *bl = l3_sfb_long[row], *bs = l3_sfb_short[row];
        for (int lay = 0; lay < 3; lay++) {
            int n = 0, pos = 0, s0 = 0, nl = 0;
            L3BandTables &B = t->bands;
            if (lay == 0) nl = 22;
            if (lay == 2) { nl = row >= 3 ? 6 : 8; s0 = 3; }
            for (int s = 0; s < nl; s++, n++) {
                B.start[row][lay][n] = bl[s];
                B.width[row][lay][n] = (uint8_t)(bl[s + 1] - bl[s]);
                B.win[row][lay][n] = -1;
                B.sfb[row][lay][n] = (uint8_t)s;
            }
            pos = nl ? bl[nl] : 0;
            if (lay != 0)
                for (int s = s0; s < 13; s++)
                    for (int w = 0; w < 3; w++, n++) {
                        B.start[row][lay][n] = (uint16_t)pos;
                        B.width[row][lay][n] = (uint8_t)(bs[s + 1] - bs[s]);
                        B.win[row][lay][n] = (int8_t)w;
                        B.sfb[row][lay][n] = (uint8_t)s;
                        pos += bs[s + 1] - bs[s];
                    }
            B.nbands[row][lay] = (uint8_t)n;
            B.nlong[row][lay] = (uint8_t)nl;
            for (int b = 0; b < n; b++)
                for (int i = 0; i < B.width[row][lay][b]; i++) {
                    const int line = B.start[row][lay][b] + i, w = B.win[row][lay][b], wd = B.width[row][lay][b];
                    B.line2band[row][lay][line] = (uint8_t)b;
                    // short bands are transmitted window by window; the spectrum interleaves the windows
                    B.dst[row][lay][line] = (uint16_t)(w < 0 ? line : (B.start[row][lay][b] - w * wd) + 3 * i + w);
                    const uint32_t dl = B.dst[row][lay][line];
                    B.lmap[row][lay][line] = (uint32_t)b | (dl << 8);
                }
        }
    }
    for (int i = 0; i < 8208; i++) t->pow43[i] = (float)std::pow((double)i, 4.0 / 3.0);
}
