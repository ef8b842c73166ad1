// k_index.cu -- K0: device-side frame indexer (stages a1-a3 of SURVEY.md section 8).
//
//   index_count   thread per stream: walk the frame chain, count frames and main-data bytes
//   index_fill    thread per stream: walk again, write one L3FrameRec per frame
//   side_parse    thread per frame : side info -> unit descriptors (+ bit-reservoir offsets)
//   payload_copy  warp per frame   : compact the frames' main-data bytes into one arena so
//                                    that every unit's bits are contiguous and addressable
//
// The chain walk is the only serial dependency of the whole decoder (a frame's position is known
// only after the previous header is read); it is latency-bound, one dependent DRAM access per
// frame, and parallel across streams only.  Everything else is embarrassingly parallel.
// No reference code exists for this stage (/root/reference/README.md:1-84).
#include "kernels.h"
#include "l3_side.h"

namespace {

__global__ void k_index_count(const uint8_t *__restrict__ raw, L3StreamRec *__restrict__ streams, int nstreams)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    L3StreamRec r = streams[s];
    const uint8_t *buf = raw + r.raw_off;
    uint32_t len = r.raw_len, p = l3_id3v2_len(buf, len), first = 0, first_off = 0, n = 0, payload = 0;
    while (p + 4 <= len) {
        L3Hdr h;
        uint32_t w;
        if (!l3_frame_at(buf, len, p, first, &h, &w)) { p++; continue; }
        if (!first) { first = w; first_off = p; }
        n++;
        payload += (uint32_t)(h.frame_len - 4 - (h.crc ? 2 : 0) - h.side_len);
        p += (uint32_t)h.frame_len;
    }
    streams[s].first_off = first_off;
    streams[s].first_hdr = first;
    streams[s].nframes = n;
    streams[s].payload_len = payload;
}

__global__ void k_index_fill(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams,
                             L3FrameRec *__restrict__ frames, int nstreams)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    L3StreamRec r = streams[s];
    if (!r.nframes) return;
    const uint8_t *buf = raw + r.raw_off;
    uint32_t len = r.raw_len, p = r.first_off, first = r.first_hdr, n = 0, payload = 0;
    while (p + 4 <= len && n < r.nframes) {
        L3Hdr h;
        uint32_t w;
        if (!l3_frame_at(buf, len, p, first, &h, &w)) { p++; continue; }
        L3FrameRec f;
        f.rel_off = p;
        f.payload_off = payload;
        f.hdr = w;
        f.stream = (uint32_t)s;
        frames[r.frame_base + n] = f;
        n++;
        payload += (uint32_t)(h.frame_len - 4 - (h.crc ? 2 : 0) - h.side_len);
        p += (uint32_t)h.frame_len;
    }
}

__global__ void k_side_parse(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams,
                             const L3FrameRec *__restrict__ frames, uint32_t nframes,
                             const uint16_t *__restrict__ sfb_long, L3UnitDesc *__restrict__ units,
                             uint32_t *__restrict__ gran_unit0, uint32_t *__restrict__ concealed)
{
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    L3FrameRec fr = frames[f];
    L3StreamRec sr = streams[fr.stream];
    L3Hdr h;
    l3_parse_hdr(fr.hdr, &h);
    uint32_t fi = f - sr.frame_base; /* frame index inside its stream */
    L3UnitDesc d[4];
    int ok = l3_parse_side(raw + sr.raw_off + fr.rel_off, &h, sfb_long + h.sr_row * 23, sr.payload_base,
                           fr.payload_off, fr.stream, fi == 0, d);
    if (!ok) atomicAdd(concealed, 1u);
    uint32_t u0 = sr.unit_base + fi * (uint32_t)(h.ngr * h.nch);
    uint32_t g0 = sr.gran_base + fi * (uint32_t)h.ngr;
    for (int i = 0; i < h.ngr * h.nch; i++) units[u0 + i] = d[i];
    for (int gr = 0; gr < h.ngr; gr++)
        gran_unit0[g0 + gr] = (u0 + (uint32_t)(gr * h.nch)) | (h.nch == 2 ? L3G_STEREO : 0u) |
                              ((fi == 0 && gr == 0) ? L3G_FIRST : 0u);
}

// One warp per frame; byte-granular because source and destination have arbitrary alignment.
__global__ void k_payload_copy(const uint8_t *__restrict__ raw, const L3StreamRec *__restrict__ streams,
                               const L3FrameRec *__restrict__ frames, uint32_t nframes, uint8_t *__restrict__ arena)
{
    uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= nframes) return;
    L3FrameRec fr = frames[f];
    L3StreamRec sr = streams[fr.stream];
    L3Hdr h;
    l3_parse_hdr(fr.hdr, &h);
    uint32_t skip = 4u + (h.crc ? 2u : 0u) + (uint32_t)h.side_len, n = (uint32_t)h.frame_len - skip;
    const uint8_t *src = raw + sr.raw_off + fr.rel_off + skip;
    uint8_t *dst = arena + sr.payload_base + fr.payload_off;
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

} // namespace

void l3_launch_index_count(const uint8_t *raw, L3StreamRec *streams, int nstreams, cudaStream_t st)
{
    if (nstreams <= 0) return;
    k_index_count<<<(nstreams + 63) / 64, 64, 0, st>>>(raw, streams, nstreams);
}
void l3_launch_index_fill(const uint8_t *raw, const L3StreamRec *streams, L3FrameRec *frames, int nstreams,
                          cudaStream_t st)
{
    if (nstreams <= 0) return;
    k_index_fill<<<(nstreams + 63) / 64, 64, 0, st>>>(raw, streams, frames, nstreams);
}
void l3_launch_side_parse(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames,
                          uint32_t nframes, const L3DevTables &T, L3UnitDesc *units, uint32_t *gran_unit0,
                          uint32_t *concealed_counter, cudaStream_t st)
{
    if (!nframes) return;
    k_side_parse<<<(nframes + 127) / 128, 128, 0, st>>>(raw, streams, frames, nframes, T.sfb_long, units,
                                                        gran_unit0, concealed_counter);
}
void l3_launch_payload_copy(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames,
                            uint32_t nframes, uint8_t *arena, cudaStream_t st)
{
    if (!nframes) return;
    uint32_t threads = 256, warps_per_block = threads / 32;
    k_payload_copy<<<(nframes + warps_per_block - 1) / warps_per_block, threads, 0, st>>>(raw, streams, frames,
                                                                                         nframes, arena);
}
