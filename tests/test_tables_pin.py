"""Pin the ISO tables compiled into the oracle against libavcodec's .rodata."""
import os
import struct
import sys

import pytest

import ffmpeg_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

SFB_LONG_WIDTH_441 = bytes([4, 4, 4, 4, 4, 4, 6, 6, 8, 8, 10, 12, 16, 20, 24, 28, 34, 42, 50, 54, 76, 158])
SFB_SHORT_WIDTH_441 = bytes([4, 4, 4, 4, 6, 8, 10, 12, 14, 18, 22, 30, 56])


def test_tables_are_complete_prefix_codes(oracle_mod):
    import ctypes
    L = oracle_mod.lib()
    for book, dim in ((1, 2), (2, 3), (3, 3), (5, 4), (6, 4), (7, 6), (8, 6), (9, 6), (10, 8), (11, 8), (12, 8),
                      (13, 16), (15, 16), (16, 16), (24, 16)):
        kraft = 0
        seen = set()
        for x in range(dim):
            for y in range(dim):
                hl, hc = ctypes.c_int(), ctypes.c_uint()
                assert L.l3o_book_entry(book, x, y, ctypes.byref(hl), ctypes.byref(hc)) == dim
                assert 1 <= hl.value <= 19 and hc.value < (1 << hl.value)
                kraft += 1 << (19 - hl.value)
                seen.add((hl.value, hc.value))
        assert kraft == 1 << 19 and len(seen) == dim * dim, book


def test_window_symmetry(oracle_mod):
    L = oracle_mod.lib()
    d = [L.l3o_dwin(i) for i in range(512)]
    assert d[0] == 0.0 and abs(d[256] - 75038 / 65536.0) == 0.0
    for i in range(1, 256):
        if i % 64:
            assert d[512 - i] == -d[i]
        else:
            assert d[512 - i] == d[i]
    assert all(abs(v * 65536 - round(v * 65536)) == 0 for v in d)


@pytest.mark.skipif(ffmpeg_ref.libavcodec_path() is None, reason="libavcodec not present")
def test_tables_match_libavcodec_rodata(oracle_mod):
    import ctypes
    import derive_tables
    books, win = derive_tables.extract(ffmpeg_ref.libavcodec_path())
    L = oracle_mod.lib()
    for bid, (dim, hlen, hcod) in books.items():
        for x in range(dim):
            for y in range(dim):
                hl, hc = ctypes.c_int(), ctypes.c_uint()
                L.l3o_book_entry(bid, x, y, ctypes.byref(hl), ctypes.byref(hc))
                assert (hl.value, hc.value) == (hlen[x * dim + y], hcod[x * dim + y])
    for i in range(257):
        assert L.l3o_dwin(i) == win[i] / 65536.0
    blob = open(ffmpeg_ref.libavcodec_path(), "rb").read()
    # band widths: 44.1 kHz long row
    i = blob.find(SFB_LONG_WIDTH_441)
    assert i >= 0
    # rows: 44.1 / 48 / 32 (MPEG-1), 22.05 / 24 / 16 (MPEG-2), 11.025 / 12 / 8 kHz (MPEG-2.5)
    rows = [blob[i + 22 * r: i + 22 * (r + 1)] for r in range(9)]
    for r in range(9):
        edges = [L.l3o_sfb_long(r, k) for k in range(23)]
        assert bytes(b - a for a, b in zip(edges, edges[1:])) == rows[r], r
    j = blob.find(SFB_SHORT_WIDTH_441)
    assert j >= 0
    for r in range(9):
        edges = [L.l3o_sfb_short(r, k) for k in range(14)]
        assert bytes(b - a for a, b in zip(edges, edges[1:])) == blob[j + 13 * r: j + 13 * (r + 1)], r


@pytest.mark.skipif(ffmpeg_ref.libavcodec_path() is None, reason="libavcodec not present")
def test_layer2_tables_match_libavcodec_rodata():
    """Layer II allocation rows / quantisation classes (mp3_b200/csrc/iso_tables_l2.h) against the arrays in
    libavcodec's .rodata they were read from."""
    import re
    src = open(os.path.join(ROOT, "mp3_b200", "csrc", "iso_tables_l2.h")).read()
    rows = {m.group(1): bytes(int(v) for v in m.group(2).split(","))
            for m in re.finditer(r"#define L2_ROW_(\w) \{([^}]*)\}", src)}
    blob = open(ffmpeg_ref.libavcodec_path(), "rb").read()
    t_ab = rows["A"] * 3 + rows["B"] * 8 + rows["C"] * 12 + rows["D"] * 7      # 3-B.2a/b: 27 / 30 subbands
    t_cd = rows["E"] * 2 + rows["F"] * 10                                      # 3-B.2c/d: 8 / 12 subbands
    t_lsf = rows["G"] * 4 + rows["F"] * 7 + rows["H"] * 19                     # 13818-3 B.1: 30 subbands
    for t in (t_ab, t_cd, t_lsf):
        assert blob.find(t) >= 0
    steps = [int(v) for v in re.search(r"l2_quant_steps\[17\] = \{([^}]*)\}", src).group(1).split(",")]
    bits = [int(v) for v in re.search(r"l2_quant_bits\[17\] = \{([^}]*)\}", src).group(1).split(",")]
    assert blob.find(struct.pack("<17i", *steps)) >= 0 and blob.find(struct.pack("<17i", *bits)) >= 0
    assert blob.find(struct.pack("<5i", 27, 30, 8, 12, 30)) >= 0
    # row ids per subband reproduce those concatenations
    ids = re.search(r"l2_row_of_sb\[5\]\[30\] = \{(.*?)\};", src, re.S).group(1)
    tabs = [[int(v) for v in r.split(",") if v.strip()] for r in re.findall(r"\{([^}]*)\}", ids)]
    names = "ABCDEFGH"
    for t, lim, want in ((0, 27, t_ab), (1, 30, t_ab), (2, 8, t_cd), (3, 12, t_cd), (4, 30, t_lsf)):
        got = b"".join(rows[names[i]] for i in tabs[t][:lim])
        assert want.startswith(got), t
