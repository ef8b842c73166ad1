// Host emulation of the time-parallel walk (k_walk_* in k_index.cu): the same phase functions (mp3_b200/csrc/walk_par.h), the CTA's threads run one
// after the other.  Built and driven by tests/test_walk_par_cpu.py; returns the dense frame table and how many segments
// had to be repaired.
#include <cstdint>
#include <cstring>
#include <vector>

#include "walk_par.h"

extern "C" int walk_emu(const uint8_t *buf, uint32_t len, uint32_t seg, int streaming, uint32_t first_preset,
                        uint32_t skip_frames, L3FrameRec *out, uint32_t cap_out, uint32_t *n_out, uint32_t *payload_out,
                        uint32_t *end_off, uint32_t *first_off, uint32_t *first_hdr, uint32_t *tag /* [4] */,
                        uint32_t *bad_segments)
{
    L3WalkFirst f;
    l3wp_first(buf, len, first_preset, skip_frames, streaming, &f);
    *n_out = *payload_out = 0;
    *end_off = f.end0;
    *first_off = 0;
    *first_hdr = f.first;
    tag[0] = tag[1] = tag[2] = tag[3] = 0;
    *bad_segments = 0;
    if (!f.have) return 0;
    const uint32_t nseg = (len - f.pf + seg - 1) / seg, cap = l3wp_seg_cap(seg);
    std::vector<L3WalkSeg> sg(nseg);
    std::vector<L3FrameRec> sparse((size_t)nseg * cap);
    for (uint32_t t = 0; t < nseg; t++)
        l3wp_segment(buf, len, f.pf, seg, nseg, t, f.first, streaming, 0, sparse.data() + (size_t)t * cap, &sg[t]);
    uint32_t bad = 0;
    for (uint32_t t = 0; t < nseg; t++) bad += !l3wp_chained(sg.data(), t, f.pf);
    *bad_segments = bad;
    // as k_walk_stitch does: repair run by run from the first segment that does not chain up, one pass to the end
    // after 32 rounds
    for (uint32_t from = 0, round = 0; from < nseg; round++) {
        uint32_t t0 = from;
        while (t0 < nseg && l3wp_chained(sg.data(), t0, f.pf)) t0++;
        if (t0 == nseg) break;
        from = l3wp_repair_run(buf, len, f.pf, seg, nseg, f.first, streaming, 0, sparse.data(), cap, sg.data(), t0, round >= 32);
    }
    uint32_t n = 0, pay = 0;
    for (uint32_t t = 0; t < nseg; t++) {
        if (sg[t].n > cap) return -2;
        for (uint32_t i = 0; i < sg[t].n; i++) {
            if (n >= cap_out) return -1;
            L3FrameRec r = sparse[(size_t)t * cap + i];
            r.payload_off += pay;
            out[n++] = r;
        }
        pay += sg[t].payload;
    }
    *n_out = n;
    *payload_out = pay;
    if (n) {
        L3Hdr h;
        l3_parse_hdr(out[n - 1].hdr, &h);
        *end_off = out[n - 1].rel_off + (uint32_t)h.frame_len;
    }
    *first_off = f.pf;
    tag[0] = f.tag_kind; tag[1] = f.tag_frames; tag[2] = f.tag_bytes; tag[3] = f.tag_delay_pad;
    return 0;
}
