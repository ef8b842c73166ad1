// api.cu -- the C-ABI of libmp3b (include/mp3b.h): contexts, device memory, the decode
// orchestration and the host-side frame indexer.
//
// There is no reference interface to mirror (/root/reference/README.md:1-84 is prose; see the
// header of include/mp3b.h); the entry points are the ones SURVEY.md section 8(b) specifies.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "iso_tables.h"
#include "kernels.h"
#include "mp3b.h"

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { // retry exact
            e = cudaMalloc(&p, n);
            want = n;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

enum { EV_START = 0, EV_INDEX, EV_HUFF, EV_REQ, EV_IMDCT, EV_OVL, EV_SYNTH, EV_END, EV_COUNT };

} // namespace

struct mp3b_stream {
    mp3b_ctx *ctx;
    // bytes not yet retired: `ctx_frames` frames that were already output (kept so that the next call can
    // re-derive the bit reservoir, the overlap-add half and the synthesis history), then new bytes
    std::vector<uint8_t> pending;
    uint32_t ctx_frames = 0;
    uint32_t first_hdr = 0;      // the stream's first header once seen (identity for the sync search)
    int batch_index = -1;        // index in the last mp3b_decode() batch
    size_t cursor = 0;           // samples per channel of the last decode already fetched
    int64_t total_samples = 0;   // samples per channel emitted over the stream's life
    // A leading ID3v2 tag is dropped as its bytes arrive (it may be far larger than one enqueue: cover art),
    // so that the frame walk never searches a partly received tag body for syncs.
    bool id3_checked = false;    // the first 10 bytes of the stream have been looked at
    uint64_t id3_left = 0;       // tag bytes still to drop from coming enqueues
};

struct mp3b_ctx {
    int device = 0;
    mp3b_opts opts{};
    cudaStream_t stream = nullptr;     // the stream work is enqueued on
    cudaStream_t own_stream = nullptr; // created by the context
    cudaStream_t copy_stream = nullptr; // D2H of finished waves (PCM sink)
    cudaStream_t index_stream = nullptr; // run-ahead frame walk (opts.async_index)
    cudaEvent_t walk_done = nullptr, walk_t0 = nullptr, walk_t1 = nullptr;
    bool walk_timed = false;
    bool stage_timing = false;           // CUDA events between the kernels of a decode (mp3b_ctx_set_stage_timing)
    bool use_pdl = false;                // MP3B_PDL=1: programmatic dependent launch between them when no stage events are
                                         // recorded (measured: no gain -- cfg2 3.92 ms against 3.86 without, cfg3 5.93 / 5.95)
    int idx_cur = 0;                     // which set of (raw, stream records, walk scratch) the last call used
    std::vector<cudaEvent_t> wave_ev;
    cudaEvent_t copy_done = nullptr;
    cudaEvent_t staging_done = nullptr; // the last H2D out of the pinned staging tables has executed
    cudaEvent_t tiles2_done = nullptr;  // the same for the Layer I / II tile table
    void *sink = nullptr;
    uint64_t sink_cap = 0;
    cudaEvent_t ev[EV_COUNT]{};
    std::string err;

    // device tables
    DevBuf d_tables;
    L3DevTables T{};

    // per-batch device state
    // Two sets of the buffers the frame walk touches: the walk of call N+1 runs on its own stream while
    // call N's kernels are still queued, so it must not overwrite what call N's side-info / payload
    // kernels read.
    DevBuf d_raw[2], d_streams[2], d_scratch[2];
    DevBuf d_sparse[2], d_segs[2];         // time-parallel walk: per-segment records and bookkeeping
    int walk_mode = 0;                     // 0 = choose by batch shape, 1 = thread per stream, 2 = CTA per stream (MP3B_WALK)
    uint32_t walk_seg = 0;                 // bytes per segment of the time-parallel walk (MP3B_WALK_SEG); 0 = by stream length
    DevBuf d_frames, d_units, d_gran, d_arena, d_pcm, d_pcm2;
    // Two sets as well (same index): the tile table and the concealment counter of call N+1 are written on the index
    // stream, behind its walk, while call N's back end still reads call N's -- so that on the context's stream nothing
    // but kernels stands between two calls (the table uploads and the memset used to: 35 us of a 3.76-ms step).
    DevBuf d_tiles[2], d_counter[2];
    cudaEvent_t end_ev[2] = {nullptr, nullptr}; // the call that used set i has finished on the context's stream
    cudaEvent_t tables_done = nullptr;          // this call's tables are in place (index stream)
    int pcm_cur = 0;                       // PCM arena of the last decode (two alternate in sink mode)
    cudaEvent_t pcm_free[2] = {nullptr, nullptr}; // sink copies out of arena i have finished
    DevBuf &pcm() { return pcm_cur ? d_pcm2 : d_pcm; }
    const DevBuf &pcm() const { return pcm_cur ? d_pcm2 : d_pcm; }
    DevBuf d_is, d_sf, d_nzv, d_xr, d_imd, d_sb; // wave-sized intermediates
    DevBuf d_hkeys, d_hperm, d_hstate, d_hctl;           // sorted Huffman variant: keys, order, histogram / cursors (wave-sized)
    int k1_mode = -1;                            // Huffman kernel (MP3B_K1_MODE): -1 = by batch size (default: warp-per-unit for waves of
                                                 // <= 4096 units, chunked above), 0 = chunked, 1 = sorted persistent, 2 = warp-per-unit
    int sm_count = 148;
    DevBuf d_sb2, d_tiles2;                      // Layer II: subband samples of its streams, synthesis tiles
    PinBuf h_tiles2;
    PinBuf h_streams, h_frames, h_tiles, h_stage, h_counter;

    const uint8_t *raw_dev = nullptr; // d_raw or the caller's device buffer
    std::vector<mp3b_stream_info> infos;
    std::vector<mp3b_tag_info> tags;
    // resampled copy of the last batch
    DevBuf d_rs, d_rs_jobs, d_rs_taps;
    DevBuf d_rs_tcA, d_rs_tcpfx;         // tensor-core resampler: coefficient tiles of the cached rate pair, entry table
    L3RsTcPlan rs_tc_plan;
    long long rs_tc_key = -1;            // in_rate << 32 | out_rate of rs_tc_plan (and of d_rs_tcA's content); -2: pair not served
    int rs_tc_mode = -1;                 // MP3B_RS_TC: -1 by batch size, 0 never, 1 whenever the pair is served
    std::vector<L3ResampleJob> rs_jobs; // per stream (in_n = 0 for streams without audio)
    uint64_t rs_elems = 0;
    bool have_rs = false;
    // planar copy of the last batch
    DevBuf d_pl, d_pl_jobs;
    bool have_pl = false;
    DevBuf d_sg_jobs, d_sg_energy, d_sg_seg, d_sg_n;
    std::vector<L3SegJob> sg_jobs;   // per stream of the batch (nwin = 0: no audio)
    std::vector<long long> sg_seg;   // host copies, fetched on first use
    std::vector<int> sg_n, sg_row;  // sg_row[i]: stream i's row in the compacted job list, -1 if no audio
    bool have_sg = false, sg_on_host = false;
    // time-stretched copy of the last batch
    DevBuf d_ts, d_ts_jobs, d_ts_off;
    std::vector<L3StretchJob> ts_jobs;
    uint64_t ts_elems = 0;
    int ts_max_frames = 0;
    bool have_ts = false;
    uint64_t pcm_elems = 0;
    uint64_t arena_bytes = 0;
    uint32_t nstreams = 0, nframes = 0, ngran = 0, nunits = 0, ntiles = 0;
    uint64_t wave_units = 2u << 20;
    uint32_t tile_override = 0; // MP3B_FUSED_TILE: granules per fused tile (tests)
    bool poison = false;        // MP3B_DEBUG_POISON=1: every scratch / output buffer is filled with 0xFF before each
                                // decode, so that a read of something this call did not write cannot go unnoticed
                                // (buffers are grow-only and reused; the all-zero spectrum tails are never written)
    bool have_batch = false, timed = false, index_timed = false;
    mp3b_stats stats{};
    std::vector<mp3b_stream *> open_streams;
    bool stream_batch_pending = false; // mp3b_decode() issued, per-stream bookkeeping not done yet
    PinBuf h_frames_out;
    PinBuf h_gather;
    std::vector<uint64_t> gather_offsets;
};

namespace {

int fail(mp3b_ctx *c, cudaError_t e, const char *what)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    c->err = buf;
    cudaGetLastError();
    return MP3B_E_CUDA;
}

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail(ctx, e__, #call); \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int upload_tables(mp3b_ctx *ctx)
{
    L3HostTables *h = new L3HostTables;
    l3_build_host_tables(h);
    if (!h->huff_lut_len) { delete h; ctx->err = "table build failed"; return MP3B_E_INVAL; }
    // one allocation: [lut][huffinfo][quad][bands][pow43][sfb_long]
    size_t o_lut = 0, o_info = align_up(o_lut + sizeof h->huff_lut, 256);
    size_t o_quad = align_up(o_info + sizeof(L3HuffInfo), 256), o_bands = align_up(o_quad + 64, 256);
    size_t o_pow = align_up(o_bands + sizeof(L3BandTables), 256), o_sfb = align_up(o_pow + sizeof h->pow43, 256);
    size_t total = o_sfb + sizeof(uint16_t) * 9 * 23;
    std::vector<uint8_t> blob(total, 0);
    memcpy(blob.data() + o_lut, h->huff_lut, sizeof h->huff_lut);
    memcpy(blob.data() + o_info, &h->huff, sizeof(L3HuffInfo));
    memcpy(blob.data() + o_quad, h->quad_a, 64);
    memcpy(blob.data() + o_bands, &h->bands, sizeof(L3BandTables));
    memcpy(blob.data() + o_pow, h->pow43, sizeof h->pow43);
    memcpy(blob.data() + o_sfb, l3_sfb_long, sizeof(uint16_t) * 9 * 23);
    uint32_t lut_len = h->huff_lut_len;
    delete h;
    CK(ctx->d_tables.ensure(total));
    CK(cudaMemcpy(ctx->d_tables.p, blob.data(), total, cudaMemcpyHostToDevice));
    uint8_t *b = ctx->d_tables.as<uint8_t>();
    ctx->T.huff_lut = reinterpret_cast<const uint16_t *>(b + o_lut);
    ctx->T.huff_lut_len = lut_len;
    ctx->T.huff = reinterpret_cast<const L3HuffInfo *>(b + o_info);
    ctx->T.quad_a = b + o_quad;
    ctx->T.bands = reinterpret_cast<const L3BandTables *>(b + o_bands);
    ctx->T.pow43 = reinterpret_cast<const float *>(b + o_pow);
    ctx->T.sfb_long = reinterpret_cast<const uint16_t *>(b + o_sfb);
    l3_requant_init();
    l3_hybrid_init();
    l3_synth_init();
    l3_fused_init();
    l3_layer2_init();
    CK(cudaGetLastError());
    return MP3B_OK;
}

// Host frame indexer (MP3B_INDEX_HOST): the same walk as k_index_walk, on host threads.
void host_index_stream(const uint8_t *buf, L3StreamRec *r, std::vector<L3FrameRec> *out, uint32_t sidx)
{
    uint32_t len = r->raw_len, p = l3_id3v2_len(buf, len), first = r->first_hdr, first_off = 0, n = 0, payload = 0;
    uint32_t end_off = p;
    const bool streaming = (r->flags & L3S_STREAMING) != 0;
    while (p + 4 <= len) {
        L3Hdr h;
        uint32_t w;
        const int fa = l3_frame_at(buf, len, p, first, &h, &w);
        if (fa != 1 && !(fa == 3 && !streaming)) {
            if (fa >= 2 && streaming) break; // wait for the rest of the frame / the confirming next header
            p++;
            continue;
        }
        if (n == 0) {
            first = first ? first : w;
            first_off = p;
            if (r->skip_frames == 0)
                r->tag_kind = l3_parse_tag(buf + p, (uint32_t)h.frame_len, &h, &r->tag_frames, &r->tag_bytes, &r->tag_delay_pad);
        }
        L3FrameRec f;
        f.rel_off = p;
        f.payload_off = payload;
        f.hdr = w;
        f.stream = sidx;
        out->push_back(f);
        n++;
        payload += h.layer != 3 ? 0u : (uint32_t)(h.frame_len - 4 - (h.crc ? 2 : 0) - h.side_len);
        p += (uint32_t)h.frame_len;
        end_off = p;
    }
    r->end_off = end_off;
    r->first_off = first_off;
    r->first_hdr = first;
    r->nframes = n;
    r->payload_len = payload;
}

// The encoder's tag frame and the gapless window it implies: the tag frame itself is not audio, the
// encoder delay plus the decoder's own 528 + 1 samples are cut from the head, the padding from the tail
// (what mpg123 / FFmpeg do with the same fields).
void fill_tag_info(const L3StreamRec &r, const L3Hdr &h, mp3b_tag_info *tg)
{
    const int64_t spf = h.spf, total = (int64_t)r.nframes * spf;
    tg->kind = (int32_t)(r.tag_kind & L3T_KIND_MASK);
    tg->has_lame = (r.tag_kind & L3T_LAME) ? 1 : 0;
    tg->frames = r.tag_frames;
    tg->bytes = r.tag_bytes;
    tg->enc_delay = (int32_t)(r.tag_delay_pad >> 16);
    tg->enc_padding = (int32_t)(r.tag_delay_pad & 0xffffu);
    int64_t start = 0, count = total;
    if (tg->kind != 0) {
        start = spf;
        count = total - spf;
        if (tg->has_lame) {
            start += tg->enc_delay + 529;
            count -= (int64_t)tg->enc_delay + tg->enc_padding;
        }
    }
    start = std::min(start, total);
    count = std::max<int64_t>(0, std::min(count, total - start));
    tg->first_sample = start;
    tg->num_samples = count;
}

int nthreads_of(const mp3b_ctx *ctx)
{
    int n = ctx->opts.host_threads;
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : (n > 64 ? 64 : n);
}

template <class F> void parallel_for(int nthreads, size_t n, F f)
{
    if (nthreads <= 1 || n < 2) {
        for (size_t i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    auto body = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            f(i);
        }
    };
    for (int t = 0; t < nthreads; t++) th.emplace_back(body);
    for (auto &t : th) t.join();
}

// Per-stream hints of the incremental (open / enqueue / decode / fetch) interface.
struct StreamHint {
    uint32_t skip_frames; // leading frames already output by an earlier call: decoded for state only
    uint32_t first_hdr;   // the stream's first header, once known (stream identity for the sync search)
};

// The whole decode of one batch.  `base` holds the streams at offsets[i]..offsets[i+1].
int decode_impl(mp3b_ctx *ctx, const uint8_t *base, const uint64_t *offsets, int nstreams, int where,
                const StreamHint *hints = nullptr)
{
    if (!ctx || !offsets || nstreams < 0 || (nstreams > 0 && !base)) return MP3B_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->have_batch = false;
    ctx->have_rs = false;
    ctx->have_ts = false;
    ctx->have_pl = false;
    ctx->have_sg = false;
    ctx->infos.assign((size_t)nstreams, mp3b_stream_info{});
    ctx->tags.assign((size_t)nstreams, mp3b_tag_info{});
    ctx->stats = mp3b_stats{};
    ctx->nstreams = (uint32_t)nstreams;
    ctx->nframes = ctx->ngran = ctx->nunits = ctx->ntiles = 0;
    ctx->pcm_elems = 0;
    int64_t launches = 0;
    const uint64_t raw0 = nstreams ? offsets[0] : 0, raw_total = nstreams ? offsets[nstreams] - raw0 : 0;
    for (int i = 0; i < nstreams; i++)
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0xFFFFFF00ull) return MP3B_E_INVAL;
    const bool host_index = ctx->opts.indexer == MP3B_INDEX_HOST;
    if (host_index && where != MP3B_HOST) { ctx->err = "host indexer needs host-resident input"; return MP3B_E_INVAL; }
    // Run-ahead walk: the upload and the frame walk (the one serial chain, a few hundred microseconds of
    // latency on a handful of CTAs) go to a private stream, so they execute while the previous call's
    // Huffman / back-end kernels still occupy the context's stream, and the host round trip that follows
    // them costs the GPU nothing.
    const bool ahead = ctx->opts.async_index != 0 && !host_index;
    if (ahead) ctx->idx_cur ^= 1;
    DevBuf &d_raw = ctx->d_raw[ctx->idx_cur], &d_streams = ctx->d_streams[ctx->idx_cur],
           &d_scratch = ctx->d_scratch[ctx->idx_cur];
    cudaStream_t ist = ahead ? ctx->index_stream : st;
    // the call before the previous one used this set of buffers: nothing of this call may touch them before it is done
    if (ahead) CK(cudaStreamWaitEvent(ist, ctx->end_ev[ctx->idx_cur], 0));

    // the pinned staging tables (stream records, tiles, frames) are rewritten below: the previous call's
    // asynchronous uploads out of them must have executed (they are early in that call, so this is short)
    CK(cudaEventSynchronize(ctx->staging_done));
    CK(cudaEventRecord(ctx->ev[EV_START], st));
    // ---- raw bytes to the device
    if (where == MP3B_HOST) {
        CK(d_raw.ensure(raw_total + 64));
        if (raw_total) CK(cudaMemcpyAsync(d_raw.p, base + raw0, raw_total, cudaMemcpyHostToDevice, ist));
        ctx->raw_dev = d_raw.as<uint8_t>();
    } else
        ctx->raw_dev = base + raw0;

    // ---- stream records, frame count
    CK(ctx->h_streams.ensure(sizeof(L3StreamRec) * (size_t)std::max(nstreams, 1)));
    CK(d_streams.ensure(sizeof(L3StreamRec) * (size_t)std::max(nstreams, 1)));
    L3StreamRec *hs = ctx->h_streams.as<L3StreamRec>();
    for (int i = 0; i < nstreams; i++) {
        memset(&hs[i], 0, sizeof hs[i]);
        hs[i].raw_off = offsets[i] - raw0;
        hs[i].raw_len = (uint32_t)(offsets[i + 1] - offsets[i]);
        if (hints) {
            hs[i].skip_frames = hints[i].skip_frames;
            hs[i].first_hdr = hints[i].first_hdr;
            hs[i].flags = L3S_STREAMING;
        }
    }
    std::vector<std::vector<L3FrameRec>> host_frames;
    if (host_index) {
        host_frames.resize((size_t)nstreams);
        parallel_for(nthreads_of(ctx), (size_t)nstreams, [&](size_t i) {
            host_index_stream(base + offsets[i], &hs[i], &host_frames[i], (uint32_t)i);
        });
    } else if (nstreams) {
        CK(d_scratch.ensure(sizeof(L3FrameRec) * l3_index_scratch_records(raw_total, (uint64_t)nstreams)));
        CK(cudaMemcpyAsync(d_streams.p, hs, sizeof(L3StreamRec) * nstreams, cudaMemcpyHostToDevice, ist));
        if (ahead) CK(cudaEventRecord(ctx->walk_t0, ist));
        // Streams of some length are walked time-parallel (speculative segments, a thread each, stitched per stream:
        // four small kernels that use the whole chip -- an hour of audio in 0.2 ms instead of 107, the cfg2 batch in
        // 0.10 ms instead of 0.25); a batch of short streams has its parallelism across streams and takes the
        // thread-per-stream walk.  Both leave the same table.
        const bool par = ctx->walk_mode == 2 || (ctx->walk_mode == 0 && raw_total / (uint64_t)nstreams >= (64u << 10));
        if (par) {
            uint32_t seg = ctx->walk_seg;
            if (!seg) // a thread per segment: some ten frames each, and at most a million segments in the batch
                seg = (uint32_t)std::min<uint64_t>(65520, std::max<uint64_t>(4080, raw_total >> 20)) / 24 * 24;
            DevBuf &d_sparse = ctx->d_sparse[ctx->idx_cur], &d_segs = ctx->d_segs[ctx->idx_cur];
            CK(d_sparse.ensure(sizeof(L3FrameRec) * l3_walk_sparse_records(raw_total, (uint64_t)nstreams, seg)));
            CK(d_segs.ensure(l3_walk_seg_bytes(raw_total, (uint64_t)nstreams, seg)));
            l3_launch_index_walk_par(ctx->raw_dev, d_streams.as<L3StreamRec>(), nstreams, d_scratch.as<L3FrameRec>(),
                                     d_sparse.as<L3FrameRec>(), d_segs.p, l3_walk_segments(raw_total, (uint64_t)nstreams, seg), seg,
                                     ist);
            launches += 3; // (four kernels)
        } else
            l3_launch_index_walk(ctx->raw_dev, d_streams.as<L3StreamRec>(), nstreams, d_scratch.as<L3FrameRec>(), ist);
        launches++;
        l3_launch_publish(d_streams.p, hs, sizeof(L3StreamRec) * nstreams, ist);
        launches++;
        if (ahead) {
            CK(cudaEventRecord(ctx->walk_t1, ist));
            CK(cudaEventRecord(ctx->walk_done, ist));
            CK(cudaStreamWaitEvent(st, ctx->walk_done, 0));
        }
        CK(cudaStreamSynchronize(ist)); // the one host round trip: sizes of everything downstream
    }

    // ---- prefix sums, per-stream info, synthesis tiles
    const int G = l3_synth_tile_granules();
    uint64_t frames = 0, grans = 0, units = 0, units_l2 = 0, payload = 0, ntiles = 0;
    std::vector<uint32_t> sgran((size_t)nstreams, 0), sunit((size_t)nstreams, 0), sskip((size_t)nstreams, 0);
    std::vector<uint8_t> sl2((size_t)nstreams, 0); // Layer I / II streams: decoded by k_layer1/2 + the synthesis kernel
    bool any_l2 = false, any_layer[4] = {false, false, false, false}; // which layers the batch holds: kernels of absent ones are not launched
    for (int i = 0; i < nstreams; i++) {
        L3StreamRec &r = hs[i];
        mp3b_stream_info &inf = ctx->infos[(size_t)i];
        r.frame_base = (uint32_t)frames;
        r.gran_base = (uint32_t)grans;
        r.unit_base = (uint32_t)units;
        r.payload_base = payload;
        inf.pcm_offset = (int64_t)(units * 576);
        if (r.first_hdr) {
            L3Hdr h;
            l3_parse_hdr(r.first_hdr, &h);
            inf.sample_rate = l3_sr_hz(h.sr_row);
            inf.channels = h.nch;
            inf.lsf = h.lsf;
            // leading `skip` frames were output by an earlier call: they are decoded again only to
            // re-derive the reservoir / overlap / synthesis state, and produce no PCM
            const uint32_t skip = std::min(r.skip_frames, r.nframes);
            r.skip_frames = skip;
            inf.frames = r.nframes - skip;
            inf.samples = (int64_t)(r.nframes - skip) * h.spf;
            inf.pcm_offset += (int64_t)skip * h.spf * h.nch;
            // the encoder's tag frame and the gapless window it implies: the tag frame itself is not audio,
            // the encoder delay plus the decoder's own 528 + 1 samples are cut from the head, the padding
            // from the tail (what mpg123 / FFmpeg do with the same fields)
            {
                mp3b_tag_info &tg = ctx->tags[(size_t)i];
                fill_tag_info(r, h, &tg);
                if (ctx->opts.gapless && !hints) {
                    inf.pcm_offset += tg.first_sample * h.nch;
                    inf.samples = tg.num_samples;
                }
            }
            frames += r.nframes;
            uint64_t g = l3_stream_granules(&h, r.nframes);
            sgran[(size_t)i] = (uint32_t)g;
            sunit[(size_t)i] = (uint32_t)(g * h.nch);
            sskip[(size_t)i] = (uint32_t)((uint64_t)skip * h.spf / 576); // whole granules before the first new sample
            sl2[(size_t)i] = h.layer != 3;
            if (h.layer != 3) { // its subband samples go to a dense buffer of the Layer I / II streams only
                r.sb_shift = (uint32_t)(units - units_l2);
                units_l2 += g * h.nch;
            }
            any_l2 = any_l2 || h.layer != 3;
            any_layer[h.layer & 3] = true;
            if (h.layer == 3) ntiles += (g + G - 1) / G;
            grans += g;
            units += g * h.nch;
            // 8..23 bytes of zero padding follow each stream's main data (k_payload_copy writes it): the last
            // Huffman code of a damaged stream may read a few bits past the stream's end, and must see zeros
            payload = align_up(payload + r.payload_len + L3_PAYLOAD_PAD, 16);
        }
    }
    if (units > L3G_UNIT_MASK || frames > 0xFFFFFFF0ull) { ctx->err = "batch too large"; return MP3B_E_INVAL; }
    ctx->nframes = (uint32_t)frames;
    ctx->ngran = (uint32_t)grans;
    ctx->nunits = (uint32_t)units;
    ctx->ntiles = (uint32_t)ntiles;
    ctx->arena_bytes = align_up(payload + 64, 16);
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    ctx->pcm_elems = units * 576;

    const bool fused = ctx->opts.pipeline == MP3B_PIPE_FUSED;
    // fused back end: tile length chosen so that the grid has a few CTAs per SM-resident slot
    uint32_t GF = 8;
    if (fused) {
        // about three waves of tiles over the SMs' resident CTA slots (four per SM for s16, three for f32 output); the SM
        // count is the device's (cudaDevAttrMultiProcessorCount), not a constant
        uint64_t want = grans / ((uint64_t)ctx->sm_count * 12u);
        GF = (uint32_t)std::min<uint64_t>(128, std::max<uint64_t>(8, want / 4 * 4));
        if (ctx->tile_override) GF = ctx->tile_override;
        ntiles = 0;
        for (int i = 0; i < nstreams; i++) {
            uint64_t g = sgran[(size_t)i] - sskip[(size_t)i];
            if (!sl2[(size_t)i]) ntiles += (g + GF - 1) / GF;
        }
        ctx->ntiles = (uint32_t)ntiles;
    }
    CK(ctx->h_tiles.ensure(sizeof(uint4) * std::max<uint64_t>(ntiles, 1)));
    std::vector<uint64_t> tile_start((size_t)nstreams + 1, 0); // first tile of each stream
    {
        uint2 *t2 = ctx->h_tiles.as<uint2>();
        uint4 *t4 = ctx->h_tiles.as<uint4>();
        uint64_t k = 0;
        for (int i = 0; i < nstreams; i++) {
            const L3StreamRec &r = hs[i];
            tile_start[(size_t)i] = k;
            if (!r.nframes || sl2[(size_t)i]) continue;
            const uint32_t g = sgran[(size_t)i];
            if (fused)
                for (uint32_t a = sskip[(size_t)i]; a < g; a += GF)
                    t4[k++] = make_uint4(r.gran_base + a, std::min<uint32_t>(GF, g - a), std::min<uint32_t>(2u, a),
                                         sunit[(size_t)i] == g ? 1u : 0u /* mono */);
            else
                for (uint32_t a = 0; a < g; a += (uint32_t)G)
                    t2[k++] = make_uint2(r.gran_base + a, std::min<uint32_t>((uint32_t)G, g - a));
        }
        tile_start[(size_t)nstreams] = k;
    }

    // ---- device buffers
    CK(ctx->d_frames.ensure(sizeof(L3FrameRec) * std::max<uint64_t>(frames, 1)));
    CK(ctx->d_units.ensure(sizeof(L3UnitDesc) * std::max<uint64_t>(units, 1)));
    CK(ctx->d_gran.ensure(sizeof(uint32_t) * std::max<uint64_t>(grans, 1)));
    CK(ctx->d_arena.ensure(ctx->arena_bytes));
    DevBuf &d_tiles = ctx->d_tiles[ctx->idx_cur], &d_counter = ctx->d_counter[ctx->idx_cur];
    CK(d_tiles.ensure(sizeof(uint4) * std::max<uint64_t>(ntiles, 1)));
    CK(d_counter.ensure(64));
    CK(ctx->h_counter.ensure(64));

    const bool keep = ctx->opts.keep_stages != 0;
    const bool sink = ctx->sink != nullptr && !keep;
    if (sink && ctx->sink_cap < ctx->pcm_elems) { ctx->err = "PCM sink too small"; return MP3B_E_TRUNCATED; }
    // Sink mode alternates between two PCM arenas, so that this call's kernels need not wait for the
    // previous call's D2H copies (still draining on the copy stream): successive calls pipeline.
    ctx->pcm_cur = sink ? (ctx->pcm_cur ^ 1) : 0;
    CK(ctx->pcm().ensure(std::max<uint64_t>(ctx->pcm_elems * elem, 16)));
    if (sink) CK(cudaStreamWaitEvent(st, ctx->pcm_free[ctx->pcm_cur], 0)); // copies issued two calls ago
    void *const pcm_dev = ctx->pcm().p;
    uint64_t wave = keep ? std::max<uint64_t>(units, 1) : std::min<uint64_t>(std::max<uint64_t>(units, 1), ctx->wave_units);
    if (sink) // several waves so that the D2H of wave k overlaps the kernels of wave k+1
        wave = std::min<uint64_t>(wave, std::max<uint64_t>(units / 8, 32768));
    // a wave is a whole number of streams; size the intermediates for the largest wave
    std::vector<std::pair<int, int>> waves; // [first stream, last stream)
    uint64_t max_wave_units = 1;
    {
        int s0 = 0;
        while (s0 < nstreams) {
            int s1 = s0;
            uint64_t u = 0;
            while (s1 < nstreams) {
                uint64_t us = sunit[(size_t)s1];
                if (s1 > s0 && u + us > wave) break;
                u += us;
                s1++;
            }
            waves.emplace_back(s0, s1);
            max_wave_units = std::max(max_wave_units, u);
            s0 = s1;
        }
    }
    // Inside a wave the tile order is free: stereo tiles first, mono tiles after, so that the CTAs
    // resident on an SM at any time mostly run the same specialisation of the back end (instruction cache).
    if (fused)
        for (const auto &wv : waves) {
            uint4 *t4 = ctx->h_tiles.as<uint4>();
            std::stable_partition(t4 + tile_start[(size_t)wv.first], t4 + tile_start[(size_t)wv.second],
                                  [](const uint4 &t) { return t.w == 0u; });
        }
    CK(ctx->d_is.ensure(max_wave_units * 576 * sizeof(int16_t)));
    CK(ctx->d_sf.ensure(max_wave_units * 40));
    CK(ctx->d_nzv.ensure(max_wave_units + 16));
    CK(ctx->d_hctl.ensure(l3_huff_sort_ctl_bytes()));
    if (ctx->k1_mode == 1) {
        CK(ctx->d_hkeys.ensure(max_wave_units * sizeof(uint16_t) + 16));
        CK(ctx->d_hperm.ensure(max_wave_units * sizeof(uint32_t) + 16));
        CK(ctx->d_hstate.ensure(max_wave_units * sizeof(uint32_t) + 16));
        CK(ctx->d_hctl.ensure(l3_huff_sort_ctl_bytes()));
    }
    if (!fused) {
        CK(ctx->d_xr.ensure(max_wave_units * 576 * sizeof(float)));
        CK(ctx->d_imd.ensure(max_wave_units * 1152 * sizeof(float)));
        CK(ctx->d_sb.ensure(max_wave_units * 576 * sizeof(float)));
    }

    if (ctx->poison) {
        for (DevBuf *b : {&ctx->d_frames, &ctx->d_units, &ctx->d_gran, &ctx->d_arena, &d_tiles, &ctx->d_is, &ctx->d_sf,
                          &ctx->d_nzv, &ctx->d_hkeys, &ctx->d_hperm, &ctx->d_hstate, &ctx->d_xr, &ctx->d_imd, &ctx->d_sb, &ctx->d_sb2, &ctx->pcm()})
            if (b->p && b->cap) CK(cudaMemsetAsync(b->p, 0xFF, b->cap, st));
    }
    if (ahead && where == MP3B_HOST && !nstreams) CK(cudaStreamSynchronize(ist));
    // the tables go up on the index stream (behind the walk, under the previous call's kernels); poisoned buffers are
    // filled on the context's stream just above, so then the tables follow there
    cudaStream_t tst = (ahead && !ctx->poison) ? ist : st;
    if (nstreams) CK(cudaMemcpyAsync(d_streams.p, hs, sizeof(L3StreamRec) * nstreams, cudaMemcpyHostToDevice, tst));
    if (ntiles)
        CK(cudaMemcpyAsync(d_tiles.p, ctx->h_tiles.p, (fused ? sizeof(uint4) : sizeof(uint2)) * ntiles,
                           cudaMemcpyHostToDevice, tst));
    CK(cudaMemsetAsync(d_counter.p, 0, 64, tst));
    bool staging_recorded = false;
    if (!host_index) { CK(cudaEventRecord(ctx->staging_done, tst)); staging_recorded = true; }
    if (tst != st) {
        CK(cudaEventRecord(ctx->tables_done, tst));
        CK(cudaStreamWaitEvent(st, ctx->tables_done, 0));
    }
    CK(cudaMemsetAsync(ctx->d_arena.as<uint8_t>() + (ctx->arena_bytes - 48), 0, 48, st));

    // ---- frame table
    L3StreamRec *ds = d_streams.as<L3StreamRec>();
    L3FrameRec *df = ctx->d_frames.as<L3FrameRec>();
    if (host_index) {
        CK(ctx->h_frames.ensure(sizeof(L3FrameRec) * std::max<uint64_t>(frames, 1)));
        L3FrameRec *hf = ctx->h_frames.as<L3FrameRec>();
        parallel_for(nthreads_of(ctx), (size_t)nstreams, [&](size_t i) {
            if (!host_frames[i].empty())
                memcpy(hf + hs[i].frame_base, host_frames[i].data(), host_frames[i].size() * sizeof(L3FrameRec));
        });
        if (frames) CK(cudaMemcpyAsync(df, hf, sizeof(L3FrameRec) * frames, cudaMemcpyHostToDevice, st));
    }
    if (!staging_recorded) CK(cudaEventRecord(ctx->staging_done, st));
    L3UnitDesc *du = ctx->d_units.as<L3UnitDesc>();
    uint32_t *dg = ctx->d_gran.as<uint32_t>();
    if (frames) {
        l3_launch_side_parse(ctx->raw_dev, ds, nstreams, df, host_index ? nullptr : d_scratch.as<L3FrameRec>(),
                             (uint32_t)frames, ctx->T, du, dg, d_counter.as<uint32_t>(), ctx->opts.verify_crc, st);
        if (any_layer[3])
            l3_launch_payload_copy(ctx->raw_dev, ds, df, (uint32_t)frames, ctx->d_arena.as<uint8_t>(), st,
                                   ctx->use_pdl && !ctx->stage_timing);
        launches += any_layer[3] ? 2 : 1;
    }
    // Per-stage events only on request: an event between two kernels makes the second one wait for the first one's
    // last CTA; without them each kernel of the chain is launched as a programmatic dependent of the one before
    // (kernels.h), so its CTAs fill the SMs the predecessor's tail leaves idle.
    const bool stage_ev = ctx->stage_timing;
    const bool pdl = ctx->use_pdl && !stage_ev;
    if (stage_ev) CK(cudaEventRecord(ctx->ev[EV_INDEX], st));

    // ---- decode waves
    for (size_t w = 0; w < waves.size(); w++) {
        const int s0 = waves[w].first, s1 = waves[w].second;
        // unit / granule / tile ranges of the wave
        uint32_t u_lo = hs[s0].unit_base, g_lo = hs[s0].gran_base;
        uint32_t u_hi = s1 < nstreams ? hs[s1].unit_base : (uint32_t)units;
        uint32_t g_hi = s1 < nstreams ? hs[s1].gran_base : (uint32_t)grans;
        const uint64_t t_lo = tile_start[(size_t)s0], t_hi = tile_start[(size_t)s1];
        const uint32_t nu = u_hi - u_lo, ngw = g_hi - g_lo;
        if (!nu) continue;
        // wave-relative views of the intermediates (kernels index them by absolute unit id)
        int16_t *is = ctx->d_is.as<int16_t>() - (size_t)u_lo * 576;
        uint8_t *sf = ctx->d_sf.as<uint8_t>() - (size_t)u_lo * 40;
        float *xr = ctx->d_xr.as<float>() - (size_t)u_lo * 576;
        float *imd = ctx->d_imd.as<float>() - (size_t)u_lo * 1152;
        float *sb = ctx->d_sb.as<float>() - (size_t)u_lo * 576;
        uint8_t *nzv = ctx->d_nzv.as<uint8_t>() - (size_t)u_lo;
        if (any_layer[3] || !fused || keep) { // (the staged pipeline and stage dumps want every unit's arrays written)
            const uint32_t avg_unit = (uint32_t)((ctx->arena_bytes + units - 1) / std::max<uint64_t>(units, 1));
            // one short stream: a warp per unit (speculative decode at 32 bit positions) finishes sooner than a thread
            // per unit -- 0.037 ms against 0.060 for one 10-s stream, break-even at about four (profiles/r02_k1_modes.json)
            const int k1 = ctx->k1_mode >= 0 ? ctx->k1_mode : (nu <= 4096u ? 2 : 0);
            if (k1 == 1) {
                const L3HuffSort scr = {ctx->d_hkeys.as<uint16_t>(), ctx->d_hperm.as<uint32_t>(), ctx->d_hstate.as<uint32_t>(),
                                        ctx->d_hctl.as<uint32_t>(),
                                        ctx->sm_count};
                l3_launch_huffman_sorted(ctx->d_arena.as<uint8_t>(), ctx->arena_bytes, du, u_lo, nu, avg_unit, ctx->T, scr, is,
                                         sf, nzv, (!fused || keep) ? 1 : 0, st, pdl);
                launches += 3;
            } else if (k1 == 2) {
                const L3HuffSort scr = {nullptr, nullptr, nullptr, ctx->d_hctl.as<uint32_t>(), ctx->sm_count};
                l3_launch_huffman_warp(ctx->d_arena.as<uint8_t>(), ctx->arena_bytes, du, u_lo, nu, ctx->T, scr, is, sf, nzv,
                                       (!fused || keep) ? 1 : 0, st);
            } else
                l3_launch_huffman_range(ctx->d_arena.as<uint8_t>(), ctx->arena_bytes, du, u_lo, nu, avg_unit, ctx->T, is, sf,
                                        nzv, (!fused || keep) ? 1 : 0, st, pdl && frames);
        }
        if (waves.size() == 1 && stage_ev) CK(cudaEventRecord(ctx->ev[EV_HUFF], st));
        auto wave_done = [&]() -> int {
            if (!sink) return MP3B_OK;
            while (ctx->wave_ev.size() <= w) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->wave_ev.push_back(e);
            }
            CK(cudaEventRecord(ctx->wave_ev[w], st));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->wave_ev[w], 0));
            const size_t lo = (size_t)u_lo * 576 * elem, n = (size_t)nu * 576 * elem;
            CK(cudaMemcpyAsync(static_cast<char *>(ctx->sink) + lo, static_cast<char *>(pcm_dev) + lo, n,
                               cudaMemcpyDeviceToHost, ctx->copy_stream));
            return MP3B_OK;
        };
        if (fused) {
            l3_launch_backend(d_tiles.as<uint4>() + t_lo, (uint32_t)(t_hi - t_lo), dg, du, is, sf, nzv, ctx->T,
                              pcm_dev, ctx->opts.pcm_format, st, pdl && (any_layer[3] || keep));
            if (waves.size() == 1 && stage_ev) CK(cudaEventRecord(ctx->ev[EV_SYNTH], st));
            launches += (any_layer[3] || keep ? 1 : 0) + (t_hi > t_lo ? 1 : 0);
            if (int rc = wave_done()) return rc;
            continue;
        }
        l3_launch_requant_range(du, dg, g_lo, ngw, is, sf, ctx->T, xr, st);
        if (waves.size() == 1 && stage_ev) CK(cudaEventRecord(ctx->ev[EV_REQ], st));
        l3_launch_imdct_range(du, u_lo, nu, xr, imd, st);
        if (waves.size() == 1 && stage_ev) CK(cudaEventRecord(ctx->ev[EV_IMDCT], st));
        l3_launch_overlap_range(du, u_lo, nu, imd, sb, st);
        if (waves.size() == 1 && stage_ev) CK(cudaEventRecord(ctx->ev[EV_OVL], st));
        l3_launch_synth(d_tiles.as<uint2>() + t_lo, (uint32_t)(t_hi - t_lo), dg, sb, nullptr, pcm_dev,
                        ctx->opts.pcm_format, st);
        if (waves.size() == 1 && stage_ev) CK(cudaEventRecord(ctx->ev[EV_SYNTH], st));
        launches += 5;
        if (int rc = wave_done()) return rc;
    }
    // ---- Layer II streams: frames -> subband samples -> the same polyphase synthesis
    if (any_l2) {
        uint64_t u_min = ~0ull, u_max = 0, nt2 = 0;
        for (int i = 0; i < nstreams; i++)
            if (sl2[(size_t)i]) {
                u_min = std::min<uint64_t>(u_min, hs[i].unit_base);
                u_max = std::max<uint64_t>(u_max, (uint64_t)hs[i].unit_base + sunit[(size_t)i]);
                nt2 += (sgran[(size_t)i] - sskip[(size_t)i] + G - 1) / G;
            }
        const size_t sb2_cap0 = ctx->d_sb2.cap;
        CK(ctx->d_sb2.ensure(std::max<uint64_t>(units_l2 * 576 * sizeof(float), 16)));
        if (ctx->poison && ctx->d_sb2.cap != sb2_cap0) CK(cudaMemsetAsync(ctx->d_sb2.p, 0xFF, ctx->d_sb2.cap, st));
        // tiles {first granule, granules} and, behind them, one dense-buffer shift per tile
        CK(ctx->h_tiles2.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(nt2, 1)));
        CK(ctx->d_tiles2.ensure(sizeof(uint32_t) * 3 * std::max<uint64_t>(nt2, 1)));
        CK(cudaEventSynchronize(ctx->tiles2_done)); // h_tiles2 may still feed the previous call's upload
        uint2 *t2 = ctx->h_tiles2.as<uint2>();
        uint32_t *sh2 = reinterpret_cast<uint32_t *>(t2 + nt2);
        uint64_t k = 0;
        for (int i = 0; i < nstreams; i++)
            if (sl2[(size_t)i])
                for (uint32_t a = sskip[(size_t)i]; a < sgran[(size_t)i]; a += (uint32_t)G) {
                    sh2[k] = hs[i].sb_shift;
                    t2[k++] = make_uint2(hs[i].gran_base + a, std::min<uint32_t>((uint32_t)G, sgran[(size_t)i] - a));
                }
        float *sb2 = ctx->d_sb2.as<float>();
        if (any_layer[2]) l3_launch_layer2(ctx->raw_dev, ds, df, (uint32_t)frames, sb2, st);
        if (any_layer[1]) l3_launch_layer1(ctx->raw_dev, ds, df, (uint32_t)frames, sb2, st);
        if (nt2) {
            CK(cudaMemcpyAsync(ctx->d_tiles2.p, t2, sizeof(uint32_t) * 3 * nt2, cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(ctx->tiles2_done, st));
            l3_launch_synth(ctx->d_tiles2.as<uint2>(), (uint32_t)nt2, dg, sb2,
                            reinterpret_cast<const uint32_t *>(ctx->d_tiles2.as<uint2>() + nt2), pcm_dev,
                            ctx->opts.pcm_format, st);
        }
        launches += (any_layer[2] ? 1 : 0) + (any_layer[1] ? 1 : 0) + (nt2 ? 1 : 0);
        if (sink) { // the waves above did not cover these streams' PCM: one more copy, whole ranges
            while (ctx->wave_ev.size() <= waves.size()) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->wave_ev.push_back(e);
            }
            CK(cudaEventRecord(ctx->wave_ev[waves.size()], st));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->wave_ev[waves.size()], 0));
            const size_t lo = (size_t)u_min * 576 * elem, n = (size_t)(u_max - u_min) * 576 * elem;
            CK(cudaMemcpyAsync(static_cast<char *>(ctx->sink) + lo, static_cast<char *>(pcm_dev) + lo, n,
                               cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
    }
    if (sink) CK(cudaEventRecord(ctx->pcm_free[ctx->pcm_cur], ctx->copy_stream));
    l3_launch_publish(d_counter.p, ctx->h_counter.p, 4, st);
    launches++;
    CK(cudaEventRecord(ctx->ev[EV_END], st));
    if (ahead) CK(cudaEventRecord(ctx->end_ev[ctx->idx_cur], st));
    CK(cudaGetLastError());

    ctx->timed = waves.size() == 1 && stage_ev;
    ctx->index_timed = stage_ev;
    ctx->walk_timed = ahead && nstreams > 0;
    ctx->stats.streams = nstreams;
    ctx->stats.frames = (int64_t)frames;
    ctx->stats.granules = (int64_t)grans;
    ctx->stats.units = (int64_t)units;
    ctx->stats.bytes_in = (int64_t)raw_total;
    ctx->stats.pcm_bytes = (int64_t)(ctx->pcm_elems * elem);
    ctx->stats.kernel_launches = launches;
    ctx->have_batch = true;
    return MP3B_OK;
}

} // namespace

// ================================================================================ C-ABI
extern "C" {

int mp3b_abi_version(void) { return MP3B_ABI_VERSION; }

int mp3b_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void mp3b_opts_default(mp3b_opts *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->struct_size = sizeof *o;
    o->pcm_format = MP3B_PCM_S16;
    o->indexer = MP3B_INDEX_DEVICE;
    o->pipeline = MP3B_PIPE_FUSED;
    o->async_index = 1;
}

int mp3b_ctx_create(int device, const mp3b_opts *opts, mp3b_ctx **out)
{
    if (!out) return MP3B_E_INVAL;
    *out = nullptr;
    // options first (no device needed to reject a malformed request)
    mp3b_opts o;
    mp3b_opts_default(&o);
    if (opts) {
        size_t sz = std::min<size_t>(opts->struct_size, sizeof(mp3b_opts));
        if (sz < 8) return MP3B_E_INVAL;
        memcpy(&o, opts, sz);
        o.struct_size = sizeof(mp3b_opts);
        if ((o.pcm_format != MP3B_PCM_S16 && o.pcm_format != MP3B_PCM_F32) ||
            (o.indexer != MP3B_INDEX_DEVICE && o.indexer != MP3B_INDEX_HOST) ||
            (o.pipeline != MP3B_PIPE_FUSED && o.pipeline != MP3B_PIPE_STAGED) || o.host_threads < 0)
            return MP3B_E_INVAL;
    }
    int n = mp3b_device_count();
    if (n <= 0 || device < 0 || device >= n) return MP3B_E_CUDA; // no CPU fallback, by design
    mp3b_ctx *ctx = new (std::nothrow) mp3b_ctx;
    if (!ctx) return MP3B_E_NOMEM;
    ctx->device = device;
    ctx->opts = o;
    if (const char *w = getenv("MP3B_WAVE_UNITS")) {
        long long v = atoll(w);
        if (v > 0) ctx->wave_units = (uint64_t)v;
    }
    if (const char *w = getenv("MP3B_DEBUG_POISON")) ctx->poison = atoi(w) != 0;
    if (const char *w = getenv("MP3B_STAGE_TIMING")) ctx->stage_timing = atoi(w) != 0;
    if (const char *w = getenv("MP3B_PDL")) ctx->use_pdl = atoi(w) != 0;
    if (const char *w = getenv("MP3B_RS_TC")) ctx->rs_tc_mode = atoi(w) != 0 ? 1 : 0;
    if (const char *w = getenv("MP3B_K1_MODE")) ctx->k1_mode = !strcmp(w, "sorted") ? 1 : (!strcmp(w, "warp") ? 2 : (!strcmp(w, "chunk") ? 0 : -1));
    if (const char *w = getenv("MP3B_WALK")) ctx->walk_mode = !strcmp(w, "serial") ? 1 : (!strcmp(w, "par") ? 2 : 0);
    if (const char *w = getenv("MP3B_WALK_SEG")) {
        long v = atol(w);
        if (v >= 48 && v <= (1 << 24)) ctx->walk_seg = (uint32_t)v / 24 * 24;
    }
    if (const char *w = getenv("MP3B_FUSED_TILE")) {
        int v = atoi(w);
        if (v > 0 && v <= 4096) ctx->tile_override = (uint32_t)v;
    }
    auto bail = [&](int rc) { mp3b_ctx_destroy(ctx); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(MP3B_E_CUDA);
    if (cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->sm_count < 1)
        return bail(MP3B_E_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(MP3B_E_CUDA);
    ctx->stream = ctx->own_stream;
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(MP3B_E_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->index_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(MP3B_E_CUDA);
    if (cudaEventCreateWithFlags(&ctx->walk_done, cudaEventDisableTiming) != cudaSuccess) return bail(MP3B_E_CUDA);
    for (cudaEvent_t *e : {&ctx->end_ev[0], &ctx->end_ev[1], &ctx->tables_done})
        if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return bail(MP3B_E_CUDA);
    if (cudaEventCreate(&ctx->walk_t0) != cudaSuccess || cudaEventCreate(&ctx->walk_t1) != cudaSuccess)
        return bail(MP3B_E_CUDA);
    if (cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming) != cudaSuccess) return bail(MP3B_E_CUDA);
    if (cudaEventCreateWithFlags(&ctx->staging_done, cudaEventDisableTiming) != cudaSuccess) return bail(MP3B_E_CUDA);
    if (cudaEventCreateWithFlags(&ctx->tiles2_done, cudaEventDisableTiming) != cudaSuccess) return bail(MP3B_E_CUDA);
    for (auto &e : ctx->pcm_free)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(MP3B_E_CUDA);
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return bail(MP3B_E_CUDA);
    int rc = upload_tables(ctx);
    if (rc != MP3B_OK) return bail(rc);
    *out = ctx;
    return MP3B_OK;
}

void mp3b_ctx_destroy(mp3b_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto *s : ctx->open_streams) delete s;
    if (ctx->index_stream) cudaStreamSynchronize(ctx->index_stream);
    for (DevBuf *b : {&ctx->d_tables, &ctx->d_raw[0], &ctx->d_raw[1], &ctx->d_streams[0], &ctx->d_streams[1],
                      &ctx->d_scratch[0], &ctx->d_scratch[1], &ctx->d_sparse[0], &ctx->d_sparse[1], &ctx->d_segs[0],
                      &ctx->d_segs[1], &ctx->d_frames, &ctx->d_units, &ctx->d_gran,
                      &ctx->d_arena, &ctx->d_tiles[0], &ctx->d_tiles[1], &ctx->d_counter[0], &ctx->d_counter[1], &ctx->d_pcm, &ctx->d_pcm2, &ctx->d_rs, &ctx->d_rs_jobs, &ctx->d_rs_taps, &ctx->d_rs_tcA, &ctx->d_rs_tcpfx, &ctx->d_ts, &ctx->d_ts_jobs, &ctx->d_ts_off, &ctx->d_pl, &ctx->d_pl_jobs, &ctx->d_sg_jobs, &ctx->d_sg_energy, &ctx->d_sg_seg, &ctx->d_sg_n, &ctx->d_sb2, &ctx->d_tiles2, &ctx->d_is, &ctx->d_sf, &ctx->d_nzv, &ctx->d_hkeys, &ctx->d_hperm, &ctx->d_hstate, &ctx->d_hctl, &ctx->d_xr,
                      &ctx->d_imd, &ctx->d_sb})
        b->release();
    for (PinBuf *b : {&ctx->h_tiles2, &ctx->h_streams, &ctx->h_frames, &ctx->h_tiles, &ctx->h_stage, &ctx->h_counter, &ctx->h_gather,
                      &ctx->h_frames_out})
        b->release();
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (auto &e : ctx->wave_ev) cudaEventDestroy(e);
    if (ctx->index_stream) cudaStreamDestroy(ctx->index_stream);
    for (cudaEvent_t e : {ctx->walk_done, ctx->walk_t0, ctx->walk_t1, ctx->end_ev[0], ctx->end_ev[1], ctx->tables_done})
        if (e) cudaEventDestroy(e);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->staging_done) cudaEventDestroy(ctx->staging_done);
    if (ctx->tiles2_done) cudaEventDestroy(ctx->tiles2_done);
    for (auto &e : ctx->pcm_free)
        if (e) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int mp3b_ctx_set_stage_timing(mp3b_ctx *ctx, int on)
{
    if (!ctx) return MP3B_E_INVAL;
    ctx->stage_timing = on != 0;
    return MP3B_OK;
}

int mp3b_ctx_set_stream(mp3b_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return MP3B_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return MP3B_OK;
}

const char *mp3b_strerror(int status)
{
    switch (status) {
    case MP3B_OK: return "ok";
    case MP3B_E_INVAL: return "invalid argument";
    case MP3B_E_NOSYNC: return "no MPEG audio frame found";
    case MP3B_E_TRUNCATED: return "destination buffer too small";
    case MP3B_E_UNSUPPORTED: return "unsupported (free-format stream, or a sample-rate pair the resampler has no filter for)";
    case MP3B_E_CUDA: return "CUDA error (no usable device, or a runtime failure)";
    case MP3B_E_NOMEM: return "out of memory";
    case MP3B_E_STATE: return "call order violated";
    default: return "unknown status";
    }
}

const char *mp3b_last_error(const mp3b_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

void *mp3b_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void mp3b_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

static int finalize_stream_batch(mp3b_ctx *ctx);

int mp3b_index_stream_host(const uint8_t *bytes, size_t n, mp3b_frame_rec *frames, size_t cap_frames,
                           size_t *nframes, mp3b_stream_info *info, mp3b_tag_info *tag)
{
    if ((!bytes && n) || !nframes || n > 0xFFFFFF00ull) return MP3B_E_INVAL;
    static_assert(sizeof(mp3b_frame_rec) == sizeof(L3FrameRec), "frame record layouts must agree");
    L3StreamRec r;
    memset(&r, 0, sizeof r);
    r.raw_len = (uint32_t)n;
    std::vector<L3FrameRec> out;
    host_index_stream(bytes, &r, &out, 0);
    *nframes = out.size();
    if (info) memset(info, 0, sizeof *info);
    if (tag) memset(tag, 0, sizeof *tag);
    if (out.empty()) return MP3B_E_NOSYNC;
    L3Hdr h;
    l3_parse_hdr(r.first_hdr, &h);
    if (info) {
        info->sample_rate = l3_sr_hz(h.sr_row);
        info->channels = h.nch;
        info->lsf = h.lsf;
        info->frames = (int64_t)out.size();
        info->samples = (int64_t)out.size() * h.spf;
    }
    if (tag) fill_tag_info(r, h, tag);
    if (out.size() > cap_frames || (!frames && cap_frames)) return MP3B_E_TRUNCATED;
    for (size_t i = 0; i < out.size(); i++) {
        frames[i].offset = out[i].rel_off;
        frames[i].payload_offset = out[i].payload_off;
        frames[i].header = out[i].hdr;
        frames[i].reserved = 0;
    }
    return MP3B_OK;
}

int mp3b_seek_plan(const uint8_t *bytes, size_t n, const mp3b_frame_rec *frames, size_t nframes, int64_t target_sample,
                   mp3b_seek *out)
{
    if (!bytes || !frames || !nframes || !out || target_sample < 0) return MP3B_E_INVAL;
    L3Hdr h;
    if (!l3_parse_hdr(frames[0].header, &h)) return MP3B_E_INVAL;
    const uint64_t t = (uint64_t)target_sample / (uint64_t)h.spf;
    if (t >= nframes) return MP3B_E_INVAL;
    // Two granules of exact spectra in front of the target make its output exact: the overlap comes from the
    // granule before, the synthesis history (15 slots) from that granule's subband samples, which overlap
    // with the one before it.  That is one 1152-sample frame, or two shorter ones.
    const uint64_t warm = h.spf < 1152 ? 2 : 1;
    const uint64_t a = t > warm ? t - warm : 0;
    // ... and frame a needs the main data its main_data_begin reaches back for (Layer III only)
    uint64_t s = a;
    L3Hdr ha;
    if (l3_parse_hdr(frames[a].header, &ha) && ha.layer == 3) {
        const uint64_t side = (uint64_t)frames[a].offset + 4u + (ha.crc ? 2u : 0u);
        if (side + 2 > n) return MP3B_E_INVAL;
        const uint32_t mdb = ha.lsf ? bytes[side] : ((uint32_t)bytes[side] << 1) | (bytes[side + 1] >> 7);
        while (s > 0 && frames[a].payload_offset - frames[s].payload_offset < mdb) s--;
    }
    out->byte_offset = frames[s].offset;
    out->first_frame = (uint32_t)s;
    out->target_frame = (uint32_t)t;
    out->discard_samples = (int64_t)((t - s) * (uint64_t)h.spf + (uint64_t)target_sample % (uint64_t)h.spf);
    return MP3B_OK;
}

int mp3b_decode_packed(mp3b_ctx *ctx, const uint8_t *base, const uint64_t *offsets, int nstreams, int where)
{
    if (!ctx) return MP3B_E_INVAL;
    if (int rc = finalize_stream_batch(ctx)) return rc;
    for (auto *s : ctx->open_streams) s->batch_index = -1; // the bulk call replaces the streams' batch
    return decode_impl(ctx, base, offsets, nstreams, where);
}

int mp3b_decode_batch(mp3b_ctx *ctx, const uint8_t *const *bufs, const size_t *lens, int nstreams)
{
    if (!ctx || nstreams < 0 || (nstreams > 0 && (!bufs || !lens))) return MP3B_E_INVAL;
    if (int rc = finalize_stream_batch(ctx)) return rc;
    for (auto *s : ctx->open_streams) s->batch_index = -1;
    ctx->gather_offsets.assign((size_t)nstreams + 1, 0);
    uint64_t tot = 0;
    for (int i = 0; i < nstreams; i++) {
        if (lens[i] && !bufs[i]) return MP3B_E_INVAL;
        ctx->gather_offsets[(size_t)i] = tot;
        tot += lens[i];
    }
    ctx->gather_offsets[(size_t)nstreams] = tot;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream)); // the staging buffer may still feed a previous copy
    CK(ctx->h_gather.ensure(tot + 64));
    uint8_t *dst = ctx->h_gather.as<uint8_t>();
    parallel_for(nthreads_of(ctx), (size_t)nstreams, [&](size_t i) {
        if (lens[i]) memcpy(dst + ctx->gather_offsets[i], bufs[i], lens[i]);
    });
    return decode_impl(ctx, dst, ctx->gather_offsets.data(), nstreams, MP3B_HOST);
}

int mp3b_sync(mp3b_ctx *ctx)
{
    if (!ctx) return MP3B_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->have_batch) {
        ctx->stats.concealed_frames = *ctx->h_counter.as<uint32_t>();
        auto ms = [&](int a, int b) {
            float v = 0.f;
            if (cudaEventElapsedTime(&v, ctx->ev[a], ctx->ev[b]) != cudaSuccess) { cudaGetLastError(); v = 0.f; }
            return v;
        };
        ctx->stats.ms_total = ms(EV_START, EV_END);
        ctx->stats.ms_index = ctx->index_timed ? ms(EV_START, EV_INDEX) : 0.f;
        if (ctx->walk_timed) { // the run-ahead walk is not between those two events: add its own time
            float v = 0.f;
            if (cudaEventElapsedTime(&v, ctx->walk_t0, ctx->walk_t1) != cudaSuccess) { cudaGetLastError(); v = 0.f; }
            ctx->stats.ms_index += v;
        }
        if (ctx->timed && ctx->nunits && ctx->opts.pipeline == MP3B_PIPE_FUSED) {
            ctx->stats.ms_huffman = ms(EV_INDEX, EV_HUFF);
            ctx->stats.ms_fused = ms(EV_HUFF, EV_SYNTH);
        } else if (ctx->timed && ctx->nunits) {
            ctx->stats.ms_huffman = ms(EV_INDEX, EV_HUFF);
            ctx->stats.ms_requant = ms(EV_HUFF, EV_REQ);
            ctx->stats.ms_imdct = ms(EV_REQ, EV_IMDCT);
            ctx->stats.ms_overlap = ms(EV_IMDCT, EV_OVL);
            ctx->stats.ms_synth = ms(EV_OVL, EV_SYNTH);
        }
    }
    return MP3B_OK;
}

int mp3b_flush(mp3b_ctx *ctx)
{
    if (!ctx) return MP3B_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    return MP3B_OK;
}

int mp3b_batch_stream_info(const mp3b_ctx *ctx, int i, mp3b_stream_info *info)
{
    if (!ctx || !info) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->infos.size()) return MP3B_E_INVAL;
    *info = ctx->infos[(size_t)i];
    return info->frames ? MP3B_OK : MP3B_E_NOSYNC;
}

int mp3b_batch_tag_info(const mp3b_ctx *ctx, int i, mp3b_tag_info *info)
{
    if (!ctx || !info) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->tags.size()) return MP3B_E_INVAL;
    *info = ctx->tags[(size_t)i];
    return MP3B_OK;
}

int mp3b_batch_planar(mp3b_ctx *ctx)
{
    if (!ctx) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    ctx->have_pl = false;
    std::vector<L3PlanarJob> jl;
    std::vector<uint32_t> tjob, tfirst;
    for (size_t i = 0; i < ctx->infos.size(); i++) {
        const mp3b_stream_info &inf = ctx->infos[i];
        if (!inf.frames || inf.samples <= 0) continue;
        L3PlanarJob jb{};
        jb.off = inf.pcm_offset;
        jb.samples = inf.samples;
        jb.channels = inf.channels;
        tfirst.push_back((uint32_t)tjob.size());
        for (int64_t t = 0; t < (inf.samples + 2047) / 2048; t++) tjob.push_back((uint32_t)jl.size());
        jl.push_back(jb);
    }
    CK(ctx->d_pl.ensure(std::max<uint64_t>(ctx->pcm_elems * elem, 16)));
    const size_t o1 = align_up(jl.size() * sizeof(L3PlanarJob), 256), o2 = o1 + align_up(tjob.size() * 4, 256);
    CK(ctx->d_pl_jobs.ensure(std::max<size_t>(o2 + tfirst.size() * 4, 16)));
    if (!jl.empty()) {
        char *b = ctx->d_pl_jobs.as<char>();
        CK(cudaMemcpyAsync(b, jl.data(), jl.size() * sizeof(L3PlanarJob), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b + o1, tjob.data(), tjob.size() * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b + o2, tfirst.data(), tfirst.size() * 4, cudaMemcpyHostToDevice, st));
        l3_launch_planar(ctx->pcm().p, ctx->d_pl.p, ctx->opts.pcm_format, reinterpret_cast<const L3PlanarJob *>(b),
                         reinterpret_cast<const uint32_t *>(b + o1), reinterpret_cast<const uint32_t *>(b + o2),
                         (uint32_t)tjob.size(), st);
    }
    CK(cudaGetLastError());
    ctx->have_pl = true;
    return MP3B_OK;
}

int mp3b_batch_planar_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems)
{
    if (!ctx || !ptr) return MP3B_E_INVAL;
    if (!ctx->have_pl) return MP3B_E_STATE;
    *ptr = ctx->d_pl.p;
    if (nelems) *nelems = ctx->pcm_elems;
    return MP3B_OK;
}

int mp3b_batch_fetch_planar(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got)
{
    if (!ctx || (!dst && cap_elems)) return MP3B_E_INVAL;
    if (!ctx->have_pl) return MP3B_E_STATE;
    if (cap_elems < ctx->pcm_elems) return MP3B_E_TRUNCATED;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    CK(cudaSetDevice(ctx->device));
    if (ctx->pcm_elems)
        CK(cudaMemcpyAsync(dst, ctx->d_pl.p, ctx->pcm_elems * elem,
                           where == MP3B_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (got) *got = ctx->pcm_elems;
    return MP3B_OK;
}

int mp3b_batch_segments(mp3b_ctx *ctx, int threshold, int min_silence_ms, int min_sentence_ms)
{
    if (!ctx || threshold < 1 || threshold > 32767 || min_silence_ms < 10 || min_sentence_ms < 0) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int G = min_silence_ms / 10, S = std::max(1, min_sentence_ms / 10);
    const size_t ns = ctx->infos.size();
    ctx->have_sg = ctx->sg_on_host = false;
    ctx->sg_jobs.assign(ns, L3SegJob{});
    ctx->sg_row.assign(ns, -1);
    std::vector<L3SegJob> jl;
    uint64_t nwin = 0, nseg = 0;
    unsigned max_nwin = 0;
    for (size_t i = 0; i < ns; i++) {
        const mp3b_stream_info &inf = ctx->infos[i];
        if (!inf.frames || inf.sample_rate <= 0 || inf.samples <= 0) continue;
        L3SegJob &jb = ctx->sg_jobs[i];
        jb.off = inf.pcm_offset;
        jb.samples = inf.samples;
        jb.channels = inf.channels;
        jb.window = inf.sample_rate / 100;
        const uint64_t w = ((uint64_t)inf.samples + (uint64_t)jb.window - 1) / (uint64_t)jb.window;
        if (w > 0xFFFFFFF0ull || nwin + w > 0xFFFFFFF0ull) { ctx->err = "too many windows"; return MP3B_E_INVAL; }
        jb.win_base = (unsigned)nwin;
        jb.nwin = (unsigned)w;
        jb.seg_base = (unsigned)nseg;
        jb.seg_cap = (unsigned)(w / (uint64_t)(G + 1) + 1); // a sentence and the pause behind it take >= G + 1 windows
        nwin += w;
        nseg += jb.seg_cap;
        max_nwin = std::max(max_nwin, jb.nwin);
        ctx->sg_row[i] = (int)jl.size();
        jl.push_back(jb);
    }
    CK(ctx->d_sg_jobs.ensure(std::max<size_t>(jl.size() * sizeof(L3SegJob), 16)));
    CK(ctx->d_sg_energy.ensure(std::max<size_t>(nwin * 8, 16)));
    CK(ctx->d_sg_seg.ensure(std::max<size_t>(nseg * 16, 16)));
    CK(ctx->d_sg_n.ensure(std::max<size_t>(jl.size() * 4, 16)));
    if (!jl.empty()) {
        CK(cudaMemcpyAsync(ctx->d_sg_jobs.p, jl.data(), jl.size() * sizeof(L3SegJob), cudaMemcpyHostToDevice, st));
        l3_launch_segments(ctx->pcm().p, ctx->opts.pcm_format, ctx->d_sg_jobs.as<L3SegJob>(), (int)jl.size(), max_nwin,
                           ctx->d_sg_energy.as<unsigned long long>(), (unsigned long long)threshold * (unsigned long long)threshold,
                           G, S, ctx->d_sg_seg.as<long long>(), ctx->d_sg_n.as<int>(), st);
    }
    CK(cudaGetLastError());
    ctx->sg_seg.resize(nseg * 2);
    ctx->sg_n.resize(jl.size());
    ctx->have_sg = true;
    return MP3B_OK;
}

int mp3b_batch_fetch_segments(mp3b_ctx *ctx, int i, int64_t *dst, size_t cap, size_t *n)
{
    if (!ctx || !n) return MP3B_E_INVAL;
    if (!ctx->have_sg) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->sg_jobs.size()) return MP3B_E_INVAL;
    if (!ctx->sg_on_host) { // one copy of the whole (small) result per mp3b_batch_segments call
        CK(cudaSetDevice(ctx->device));
        if (!ctx->sg_n.empty()) {
            CK(cudaMemcpyAsync(ctx->sg_n.data(), ctx->d_sg_n.p, ctx->sg_n.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->sg_seg.data(), ctx->d_sg_seg.p, ctx->sg_seg.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->sg_on_host = true;
    }
    const L3SegJob &jb = ctx->sg_jobs[(size_t)i];
    *n = 0;
    if (!jb.nwin) return MP3B_OK;
    *n = (size_t)ctx->sg_n[(size_t)ctx->sg_row[(size_t)i]];
    if (*n > cap || (!dst && *n)) return MP3B_E_TRUNCATED;
    static_assert(sizeof(long long) == sizeof(int64_t), "segment pairs are int64");
    memcpy(dst, ctx->sg_seg.data() + 2 * (size_t)jb.seg_base, *n * 16);
    return MP3B_OK;
}

int mp3b_batch_window_energy(mp3b_ctx *ctx, int i, uint64_t *dst, size_t cap, size_t *n, int *window)
{
    if (!ctx || !n) return MP3B_E_INVAL;
    if (!ctx->have_sg) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->sg_jobs.size()) return MP3B_E_INVAL;
    const L3SegJob &jb = ctx->sg_jobs[(size_t)i];
    *n = jb.nwin;
    if (window) *window = jb.window;
    if (*n > cap || (!dst && *n)) return MP3B_E_TRUNCATED;
    CK(cudaSetDevice(ctx->device));
    if (*n) {
        CK(cudaMemcpyAsync(dst, ctx->d_sg_energy.as<unsigned long long>() + jb.win_base, *n * 8, cudaMemcpyDeviceToHost,
                           ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MP3B_OK;
}

int mp3b_batch_time_stretch(mp3b_ctx *ctx, int num, int den)
{
    if (!ctx || num <= 0 || den <= 0 || num > 64 * den || den > 64 * num) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    const size_t ns = ctx->infos.size();
    ctx->have_ts = false;
    ctx->ts_jobs.assign(ns, L3StretchJob{});
    std::vector<L3StretchJob> jl;
    uint64_t out_elems = 0;
    int64_t max_frames = 1;
    for (size_t i = 0; i < ns; i++) {
        const mp3b_stream_info &inf = ctx->infos[i];
        L3StretchJob &jb = ctx->ts_jobs[i];
        if (!inf.frames || inf.sample_rate <= 0 || inf.samples <= 0) continue;
        jb.in_off = inf.pcm_offset;
        jb.in_n = inf.samples;
        // every stream's output starts on a 16-byte boundary: k_stretch stores a stereo frame as one 32-bit
        // (s16) or 64-bit (f32) word, and an odd-length mono stream in front must not misalign it
        out_elems = align_up(out_elems, 8);
        jb.out_off = (long long)out_elems;
        jb.out_n = inf.samples * den / num;
        jb.channels = inf.channels;
        jb.hop = l3_stretch_hop(inf.sample_rate);
        out_elems += (uint64_t)jb.out_n * (uint64_t)inf.channels;
        max_frames = std::max<int64_t>(max_frames, (jb.out_n + jb.hop - 1) / jb.hop);
        jl.push_back(jb);
    }
    ctx->ts_max_frames = (int)max_frames;
    CK(ctx->d_ts.ensure(std::max<uint64_t>(out_elems * elem, 16)));
    CK(ctx->d_ts_jobs.ensure(std::max<size_t>(jl.size() * sizeof(L3StretchJob), 16)));
    CK(ctx->d_ts_off.ensure(std::max<size_t>(jl.size() * (size_t)max_frames * sizeof(int), 16)));
    if (!jl.empty()) {
        CK(cudaMemcpyAsync(ctx->d_ts_jobs.p, jl.data(), jl.size() * sizeof(L3StretchJob), cudaMemcpyHostToDevice, st));
        l3_launch_stretch(ctx->pcm().p, ctx->d_ts.p, ctx->opts.pcm_format, ctx->d_ts_jobs.as<L3StretchJob>(),
                          (int)jl.size(), num, den, ctx->d_ts_off.as<int>(), (int)max_frames, st);
    }
    CK(cudaGetLastError());
    ctx->ts_elems = out_elems;
    ctx->have_ts = true;
    return MP3B_OK;
}

int mp3b_batch_stretched_info(const mp3b_ctx *ctx, int i, int64_t *offset_elems, int64_t *samples)
{
    if (!ctx) return MP3B_E_INVAL;
    if (!ctx->have_ts) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->ts_jobs.size()) return MP3B_E_INVAL;
    if (offset_elems) *offset_elems = ctx->ts_jobs[(size_t)i].out_off;
    if (samples) *samples = ctx->ts_jobs[(size_t)i].out_n;
    return MP3B_OK;
}

int mp3b_batch_stretched_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems)
{
    if (!ctx || !ptr) return MP3B_E_INVAL;
    if (!ctx->have_ts) return MP3B_E_STATE;
    *ptr = ctx->d_ts.p;
    if (nelems) *nelems = ctx->ts_elems;
    return MP3B_OK;
}

int mp3b_batch_fetch_stretched(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got)
{
    if (!ctx || (!dst && cap_elems)) return MP3B_E_INVAL;
    if (!ctx->have_ts) return MP3B_E_STATE;
    if (cap_elems < ctx->ts_elems) return MP3B_E_TRUNCATED;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    CK(cudaSetDevice(ctx->device));
    if (ctx->ts_elems)
        CK(cudaMemcpyAsync(dst, ctx->d_ts.p, ctx->ts_elems * elem,
                           where == MP3B_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (got) *got = ctx->ts_elems;
    return MP3B_OK;
}

int mp3b_batch_stretch_offsets(mp3b_ctx *ctx, int i, int32_t *dst, size_t cap, size_t *n, int *hop)
{
    if (!ctx || !n) return MP3B_E_INVAL;
    if (!ctx->have_ts) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->ts_jobs.size()) return MP3B_E_INVAL;
    const L3StretchJob &jb = ctx->ts_jobs[(size_t)i];
    size_t row = 0; // jobs were compacted over streams with audio
    for (int k = 0; k < i; k++) row += ctx->ts_jobs[(size_t)k].in_n ? 1 : 0;
    *n = jb.in_n ? (size_t)((jb.out_n + jb.hop - 1) / jb.hop) : 0;
    if (hop) *hop = jb.hop;
    if (*n > cap || (!dst && *n)) return MP3B_E_TRUNCATED;
    CK(cudaSetDevice(ctx->device));
    if (*n) {
        CK(cudaMemcpyAsync(dst, ctx->d_ts_off.as<int>() + row * (size_t)ctx->ts_max_frames, *n * sizeof(int),
                           cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MP3B_OK;
}

int mp3b_resample_filter(int in_rate, int out_rate, float *taps, size_t cap, size_t *ncoef, int *L, int *M,
                         int *taps_per_phase)
{
    std::vector<float> hp;
    int l = 0, m = 0, t = 0, h = 0;
    if (in_rate == out_rate && in_rate > 0) { hp.assign(1, 1.f); l = m = t = 1; }
    else if (!l3_resample_design(in_rate, out_rate, &hp, &l, &m, &t, &h)) return MP3B_E_INVAL;
    if (ncoef) *ncoef = hp.size();
    if (L) *L = l;
    if (M) *M = m;
    if (taps_per_phase) *taps_per_phase = t;
    if (!taps || cap < hp.size()) return MP3B_E_TRUNCATED;
    memcpy(taps, hp.data(), hp.size() * sizeof(float));
    return MP3B_OK;
}

// The tensor-core path's plan for a rate pair, for verification without a GPU: the number of tile kinds, the padded
// window length K, and -- if `a` is given -- kind `kind`'s coefficient matrix [128][K] as c1 + c2 (the two fp16 pieces
// added in float), row-major.  MP3B_E_UNSUPPORTED: the pair is not served by that path.
int mp3b_resample_tc_plan(int in_rate, int out_rate, int kind, float *a, size_t cap, int *nkinds, int *kpad)
{
    std::vector<float> hp;
    int l = 0, m = 0, t = 0, h = 0;
    if (in_rate == out_rate || !l3_resample_design(in_rate, out_rate, &hp, &l, &m, &t, &h)) return MP3B_E_INVAL;
    L3RsTcPlan plan;
    if (t != 65 || !l3_resample_tc_plan(hp.data(), l, m, t, h, &plan)) return MP3B_E_UNSUPPORTED;
    if (nkinds) *nkinds = plan.NK;
    if (kpad) *kpad = plan.Kpad;
    if (!a) return MP3B_OK;
    if (kind < 0 || kind >= plan.NK) return MP3B_E_INVAL;
    if (cap < (size_t)128 * plan.Kpad) return MP3B_E_TRUNCATED;
    auto half_to_float = [](uint16_t v) -> float { // IEEE binary16 -> binary32
        const uint32_t s = (uint32_t)(v >> 15) << 31, e = (v >> 10) & 31u, f = v & 1023u;
        uint32_t bits;
        if (e == 0) {
            if (!f) bits = s;
            else { // subnormal: normalise
                int sh = 0;
                uint32_t ff = f;
                while (!(ff & 1024u)) { ff <<= 1; sh++; }
                bits = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | ((ff & 1023u) << 13);
            }
        } else if (e == 31) bits = s | 0x7f800000u | (f << 13);
        else bits = s | ((e - 15 + 127) << 23) | (f << 13);
        float r;
        memcpy(&r, &bits, 4);
        return r;
    };
    const uint16_t *hi = plan.A.data() + (size_t)kind * 2 * 128 * plan.Kpad, *lo = hi + (size_t)128 * plan.Kpad;
    for (int r = 0; r < 128; r++)
        for (int c = 0; c < plan.Kpad; c++) {
            // the MMA's K-major core-matrix layout (k_resample_tc.cu canon_off16): 8 rows x 8 elements per 128 bytes
            const size_t off = ((size_t)(c >> 3) * (128 >> 3) + (size_t)(r >> 3)) * 64 + (size_t)(r & 7) * 8 + (size_t)(c & 7);
            a[(size_t)r * plan.Kpad + c] = half_to_float(hi[off]) + half_to_float(lo[off]);
        }
    return MP3B_OK;
}

int mp3b_batch_resample(mp3b_ctx *ctx, int out_rate)
{
    if (!ctx || out_rate <= 0) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    const size_t ns = ctx->infos.size();
    ctx->have_rs = false;
    ctx->rs_jobs.assign(ns, L3ResampleJob{});
    // output layout: streams back to back; one filter (and one launch) per distinct input rate
    std::vector<int> rates;
    uint64_t out_elems = 0;
    for (size_t i = 0; i < ns; i++) {
        const mp3b_stream_info &inf = ctx->infos[i];
        L3ResampleJob &jb = ctx->rs_jobs[i];
        if (!inf.frames || inf.sample_rate <= 0 || inf.samples <= 0) continue;
        const int g = std::__gcd(inf.sample_rate, out_rate);
        const int64_t L = out_rate / g, M = inf.sample_rate / g;
        jb.in_off = inf.pcm_offset;
        jb.in_n = inf.samples;
        out_elems = align_up(out_elems, 8); // 16-byte aligned stream starts, as in the stretched arena
        jb.out_off = (long long)out_elems;
        jb.out_n = (inf.samples * L + M - 1) / M;
        jb.channels = inf.channels;
        out_elems += (uint64_t)jb.out_n * (uint64_t)inf.channels;
        if (std::find(rates.begin(), rates.end(), inf.sample_rate) == rates.end()) rates.push_back(inf.sample_rate);
    }
    CK(ctx->d_rs.ensure(std::max<uint64_t>(out_elems * elem, 16)));
    // the pageable staging vectors below are consumed by synchronous copies (cudaMemcpyAsync from pageable
    // memory returns after the source has been read), so they may die at the end of each iteration
    size_t job_bytes = 0, tap_bytes = 0;
    std::vector<std::vector<L3ResampleJob>> jl(rates.size());
    std::vector<std::vector<float>> taps(rates.size());
    std::vector<int> pL(rates.size()), pM(rates.size()), pT(rates.size()), pH(rates.size());
    for (size_t r = 0; r < rates.size(); r++) {
        for (size_t i = 0; i < ns; i++)
            if (ctx->rs_jobs[i].in_n && ctx->infos[i].sample_rate == rates[r]) jl[r].push_back(ctx->rs_jobs[i]);
        if (rates[r] == out_rate) { taps[r].assign(1, 1.f); pL[r] = pM[r] = pT[r] = 1; pH[r] = 0; }
        else if (!l3_resample_design(rates[r], out_rate, &taps[r], &pL[r], &pM[r], &pT[r], &pH[r])) {
            ctx->err = "unsupported sample-rate pair";
            return MP3B_E_UNSUPPORTED;
        }
        job_bytes += align_up(jl[r].size() * sizeof(L3ResampleJob), 256);
        tap_bytes += align_up(taps[r].size() * sizeof(float), 256);
    }
    CK(ctx->d_rs_jobs.ensure(std::max<size_t>(job_bytes, 16)));
    CK(ctx->d_rs_taps.ensure(std::max<size_t>(tap_bytes, 16)));
    size_t jo = 0, to = 0;
    for (size_t r = 0; r < rates.size(); r++) {
        char *dj = ctx->d_rs_jobs.as<char>() + jo, *dt = ctx->d_rs_taps.as<char>() + to;
        CK(cudaMemcpyAsync(dj, jl[r].data(), jl[r].size() * sizeof(L3ResampleJob), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dt, taps[r].data(), taps[r].size() * sizeof(float), cudaMemcpyHostToDevice, st));
        long long mx = 0;
        for (const auto &j : jl[r]) mx = std::max(mx, j.out_n);
        // Stereo s16 streams of a large batch go through the tensor cores (k_resample_tc.cu); the rest -- mono, float
        // PCM, rate pairs whose window does not fit, small batches -- through the FP32 kernels.
        int mask = 3;
        if (ctx->opts.pcm_format == MP3B_PCM_S16 && ctx->rs_tc_mode != 0 && rates[r] != out_rate && pT[r] == 65) {
            const long long key = ((long long)rates[r] << 32) | (long long)out_rate;
            if (ctx->rs_tc_key != key) {
                ctx->rs_tc_key = -1;
                if (l3_resample_tc_plan(taps[r].data(), pL[r], pM[r], pT[r], pH[r], &ctx->rs_tc_plan)) {
                    CK(ctx->d_rs_tcA.ensure(ctx->rs_tc_plan.A.size() * sizeof(uint16_t)));
                    CK(cudaMemcpyAsync(ctx->d_rs_tcA.p, ctx->rs_tc_plan.A.data(), ctx->rs_tc_plan.A.size() * sizeof(uint16_t),
                                       cudaMemcpyHostToDevice, st));
                    ctx->rs_tc_key = key;
                } else
                    ctx->rs_tc_plan = L3RsTcPlan{};
            }
            if (ctx->rs_tc_key == key) {
                std::vector<uint32_t> pfx;
                const unsigned long long total = l3_resample_tc_prefix(ctx->rs_tc_plan, jl[r].data(), (int)jl[r].size(), &pfx);
                // worth it from a few groups per CTA on (one short stream is faster on the FP32 kernel)
                if (total && (ctx->rs_tc_mode == 1 || total >= 4096ull)) {
                    CK(ctx->d_rs_tcpfx.ensure(pfx.size() * sizeof(uint32_t)));
                    CK(cudaMemcpyAsync(ctx->d_rs_tcpfx.p, pfx.data(), pfx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                    uint32_t mxe = 0;
                    for (int k = 0; k < ctx->rs_tc_plan.NK; k++)
                        mxe = std::max(mxe, pfx[(size_t)k * (jl[r].size() + 1) + jl[r].size()]);
                    l3_launch_resample_tc(ctx->pcm().p, ctx->d_rs.p, reinterpret_cast<const L3ResampleJob *>(dj), (int)jl[r].size(),
                                          ctx->d_rs_tcpfx.as<uint32_t>(), mxe, ctx->d_rs_tcA.as<uint16_t>(), ctx->rs_tc_plan,
                                          ctx->sm_count, st);
                    mask = 1;
                }
            }
        }
        l3_launch_resample(ctx->pcm().p, ctx->d_rs.p, ctx->opts.pcm_format, reinterpret_cast<const L3ResampleJob *>(dj),
                           (int)jl[r].size(), mx, reinterpret_cast<const float *>(dt), pL[r], pM[r], pT[r], pH[r], st, mask);
        jo += align_up(jl[r].size() * sizeof(L3ResampleJob), 256);
        to += align_up(taps[r].size() * sizeof(float), 256);
    }
    CK(cudaGetLastError());
    ctx->rs_elems = out_elems;
    ctx->have_rs = true;
    return MP3B_OK;
}

int mp3b_batch_resampled_info(const mp3b_ctx *ctx, int i, int64_t *offset_elems, int64_t *samples)
{
    if (!ctx) return MP3B_E_INVAL;
    if (!ctx->have_rs) return MP3B_E_STATE;
    if (i < 0 || (size_t)i >= ctx->rs_jobs.size()) return MP3B_E_INVAL;
    if (offset_elems) *offset_elems = ctx->rs_jobs[(size_t)i].out_off;
    if (samples) *samples = ctx->rs_jobs[(size_t)i].out_n;
    return MP3B_OK;
}

int mp3b_batch_resampled_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems)
{
    if (!ctx || !ptr) return MP3B_E_INVAL;
    if (!ctx->have_rs) return MP3B_E_STATE;
    *ptr = ctx->d_rs.p;
    if (nelems) *nelems = ctx->rs_elems;
    return MP3B_OK;
}

int mp3b_batch_fetch_resampled(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got)
{
    if (!ctx || (!dst && cap_elems)) return MP3B_E_INVAL;
    if (!ctx->have_rs) return MP3B_E_STATE;
    if (cap_elems < ctx->rs_elems) return MP3B_E_TRUNCATED;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    CK(cudaSetDevice(ctx->device));
    if (ctx->rs_elems)
        CK(cudaMemcpyAsync(dst, ctx->d_rs.p, ctx->rs_elems * elem,
                           where == MP3B_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (got) *got = ctx->rs_elems;
    return MP3B_OK;
}

int mp3b_batch_pcm_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems)
{
    if (!ctx || !ptr || !nelems) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    *ptr = ctx->pcm().p;
    *nelems = ctx->pcm_elems;
    return MP3B_OK;
}

int mp3b_batch_fetch_pcm(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got)
{
    if (!ctx || (!dst && cap_elems)) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    if (cap_elems < ctx->pcm_elems) return MP3B_E_TRUNCATED;
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    CK(cudaSetDevice(ctx->device));
    if (ctx->pcm_elems)
        CK(cudaMemcpyAsync(dst, ctx->pcm().p, ctx->pcm_elems * elem,
                           where == MP3B_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (got) *got = ctx->pcm_elems;
    return MP3B_OK;
}

int mp3b_set_pcm_sink(mp3b_ctx *ctx, void *host_dst, uint64_t cap_elems)
{
    if (!ctx) return MP3B_E_INVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->sink = host_dst;
    ctx->sink_cap = host_dst ? cap_elems : 0;
    return MP3B_OK;
}

int mp3b_get_stats(const mp3b_ctx *ctx, mp3b_stats *st)
{
    if (!ctx || !st) return MP3B_E_INVAL;
    *st = ctx->stats;
    return MP3B_OK;
}

// ---------------------------------------------------------------------------- stream interface
// Incremental: every mp3b_decode() emits PCM for the frames completed since the previous call.  A
// stream keeps, besides the not yet decodable tail, the last few frames it already output: enough
// main data behind them to serve any main_data_begin (511 / 255 bytes) plus the two granules the
// fused back end replays to re-derive its overlap-add and synthesis state.  Nothing else is carried
// between calls, so a resumed stream produces exactly the samples a one-shot decode would.
static int finalize_stream_batch(mp3b_ctx *ctx)
{
    if (!ctx->stream_batch_pending) return MP3B_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream_batch_pending = false;
    const L3StreamRec *hs = ctx->h_streams.as<L3StreamRec>();
    const L3FrameRec *fr = ctx->h_frames_out.as<L3FrameRec>();
    for (auto *s : ctx->open_streams) {
        if (s->batch_index < 0) continue;
        const L3StreamRec &r = hs[s->batch_index];
        const mp3b_stream_info &inf = ctx->infos[(size_t)s->batch_index];
        s->total_samples += inf.samples;
        size_t keep_from = 0;
        if (r.nframes) {
            if (!s->first_hdr) s->first_hdr = r.first_hdr;
            L3Hdr h;
            l3_parse_hdr(r.first_hdr, &h);
            // W frames = the two granules the back end replays; `need` = the reach of main_data_begin
            // (Layer II frames are self-contained)
            const uint32_t nfr = r.nframes, W = h.spf < 1152 ? 2u : 1u, need = h.layer != 3 ? 0u : (h.lsf ? 255u : 511u);
            const L3FrameRec *f = fr + r.frame_base;
            const uint32_t wf = nfr > W ? nfr - W : 0; // first frame the back end will replay
            uint32_t idx = wf;
            while (idx > 0 && nfr - idx < 192 && f[wf].payload_off - f[idx].payload_off < need) idx--;
            keep_from = f[idx].rel_off;
            s->ctx_frames = nfr - idx;
        } else if (s->pending.size() > (1u << 20)) {
            keep_from = s->pending.size() - 4096; // megabytes without a single frame: drop the junk
        }
        if (keep_from) s->pending.erase(s->pending.begin(), s->pending.begin() + (ptrdiff_t)keep_from);
    }
    return MP3B_OK;
}

int mp3b_stream_open(mp3b_ctx *ctx, mp3b_stream **out)
{
    if (!ctx || !out) return MP3B_E_INVAL;
    if (int rc = finalize_stream_batch(ctx)) return rc;
    mp3b_stream *s = new (std::nothrow) mp3b_stream;
    if (!s) return MP3B_E_NOMEM;
    s->ctx = ctx;
    ctx->open_streams.push_back(s);
    *out = s;
    return MP3B_OK;
}

void mp3b_stream_close(mp3b_stream *s)
{
    if (!s) return;
    finalize_stream_batch(s->ctx);
    auto &v = s->ctx->open_streams;
    // the other streams keep their batch_index: it indexes the last decode's tables (ctx->infos), not
    // open_streams, so their unfetched PCM stays fetchable; mp3b_decode() reassigns every index anyway
    v.erase(std::remove(v.begin(), v.end(), s), v.end());
    delete s;
}

int mp3b_stream_enqueue(mp3b_stream *s, const uint8_t *bytes, size_t n)
{
    if (!s || (n && !bytes)) return MP3B_E_INVAL;
    if (s->id3_left) { // still inside the stream's leading ID3v2 tag
        const size_t d = (size_t)std::min<uint64_t>(s->id3_left, n);
        s->id3_left -= d;
        bytes += d;
        n -= d;
    }
    try {
        s->pending.insert(s->pending.end(), bytes, bytes + n);
    } catch (...) {
        return MP3B_E_NOMEM;
    }
    if (!s->id3_checked && !s->pending.empty()) {
        const std::vector<uint8_t> &pb = s->pending;
        static const uint8_t magic[3] = {'I', 'D', '3'};
        if (memcmp(pb.data(), magic, std::min<size_t>(3, pb.size())) != 0) s->id3_checked = true; // no tag
        else if (pb.size() >= 10) { // (fewer bytes of a possible tag header: decide when they are all here)
            s->id3_checked = true;
            if (!((pb[6] | pb[7] | pb[8] | pb[9]) & 0x80)) {
                uint64_t len = ((uint64_t)pb[6] << 21) | ((uint64_t)pb[7] << 14) | ((uint64_t)pb[8] << 7) | pb[9];
                len += 10 + ((pb[5] & 0x10) ? 10 : 0);
                const size_t d = (size_t)std::min<uint64_t>(len, pb.size());
                s->id3_left = len - d;
                s->pending.erase(s->pending.begin(), s->pending.begin() + (ptrdiff_t)d);
            }
        }
    }
    return MP3B_OK;
}

int mp3b_decode(mp3b_ctx *ctx)
{
    if (!ctx) return MP3B_E_INVAL;
    if (int rc = finalize_stream_batch(ctx)) return rc;
    const int n = (int)ctx->open_streams.size();
    std::vector<StreamHint> hints((size_t)n);
    ctx->gather_offsets.assign((size_t)n + 1, 0);
    uint64_t tot = 0;
    for (int i = 0; i < n; i++) {
        mp3b_stream *s = ctx->open_streams[(size_t)i];
        s->batch_index = i;
        s->cursor = 0;
        hints[(size_t)i].skip_frames = s->ctx_frames;
        hints[(size_t)i].first_hdr = s->first_hdr;
        ctx->gather_offsets[(size_t)i] = tot;
        tot += s->pending.size();
    }
    ctx->gather_offsets[(size_t)n] = tot;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->h_gather.ensure(tot + 64));
    uint8_t *dst = ctx->h_gather.as<uint8_t>();
    for (int i = 0; i < n; i++) {
        const auto &pb = ctx->open_streams[(size_t)i]->pending;
        if (!pb.empty()) memcpy(dst + ctx->gather_offsets[(size_t)i], pb.data(), pb.size());
    }
    int rc = decode_impl(ctx, dst, ctx->gather_offsets.data(), n, MP3B_HOST, hints.data());
    if (rc != MP3B_OK) {
        for (auto *s : ctx->open_streams) s->batch_index = -1;
        return rc;
    }
    // the frame table comes back for the bookkeeping done in finalize_stream_batch()
    CK(ctx->h_frames_out.ensure(sizeof(L3FrameRec) * std::max<uint32_t>(ctx->nframes, 1)));
    if (ctx->nframes)
        CK(cudaMemcpyAsync(ctx->h_frames_out.p, ctx->d_frames.p, sizeof(L3FrameRec) * ctx->nframes,
                           cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stream_batch_pending = true;
    return MP3B_OK;
}

int mp3b_stream_get_info(const mp3b_stream *s, mp3b_stream_info *info)
{
    if (!s || !info) return MP3B_E_INVAL;
    if (s->batch_index < 0 || !s->ctx->have_batch) return MP3B_E_STATE;
    if (int rc = finalize_stream_batch(s->ctx)) return rc;
    *info = s->ctx->infos[(size_t)s->batch_index];
    info->total_samples = s->total_samples;
    return s->first_hdr ? MP3B_OK : MP3B_E_NOSYNC;
}

int mp3b_stream_pcm_device_ptr(const mp3b_stream *s, const void **ptr, size_t *nsamples)
{
    if (!s || !ptr || !nsamples) return MP3B_E_INVAL;
    if (s->batch_index < 0 || !s->ctx->have_batch) return MP3B_E_STATE;
    const mp3b_stream_info &inf = s->ctx->infos[(size_t)s->batch_index];
    const int elem = s->ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    *ptr = s->ctx->pcm().as<uint8_t>() + (size_t)inf.pcm_offset * elem;
    *nsamples = (size_t)inf.samples;
    return MP3B_OK;
}

int mp3b_stream_fetch_pcm(mp3b_stream *s, void *dst, size_t cap_samples, int where, size_t *got)
{
    if (!s || (!dst && cap_samples)) return MP3B_E_INVAL;
    mp3b_ctx *ctx = s->ctx;
    if (s->batch_index < 0 || !ctx->have_batch) return MP3B_E_STATE;
    if (int rc = finalize_stream_batch(ctx)) return rc;
    const mp3b_stream_info &inf = ctx->infos[(size_t)s->batch_index];
    const int elem = ctx->opts.pcm_format == MP3B_PCM_S16 ? 2 : 4;
    size_t left = (size_t)inf.samples - std::min<size_t>(s->cursor, (size_t)inf.samples);
    size_t n = std::min(left, cap_samples);
    CK(cudaSetDevice(ctx->device));
    if (n) {
        const uint8_t *src = ctx->pcm().as<uint8_t>() + ((size_t)inf.pcm_offset + s->cursor * inf.channels) * elem;
        CK(cudaMemcpyAsync(dst, src, n * inf.channels * elem,
                           where == MP3B_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    s->cursor += n;
    if (got) *got = n;
    return MP3B_OK;
}

// ---------------------------------------------------------------------------- debug stages
int mp3b_debug_stage(mp3b_ctx *ctx, int stage, void *dst, uint64_t cap_bytes, uint32_t *elem_size, uint64_t *count)
{
    if (!ctx) return MP3B_E_INVAL;
    if (!ctx->have_batch) return MP3B_E_STATE;
    const void *src = nullptr;
    uint32_t es = 1;
    uint64_t n = 0;
    switch (stage) {
    case MP3B_STAGE_FRAMES: src = ctx->d_frames.p; es = 16; n = ctx->nframes; break;
    case MP3B_STAGE_UNITDESC: src = ctx->d_units.p; es = 32; n = ctx->nunits; break;
    case MP3B_STAGE_MAINDATA: src = ctx->d_arena.p; es = 1; n = ctx->arena_bytes; break;
    case MP3B_STAGE_IS: src = ctx->d_is.p; es = 2; n = (uint64_t)ctx->nunits * 576; break;
    case MP3B_STAGE_SF: src = ctx->d_sf.p; es = 1; n = (uint64_t)ctx->nunits * 40; break;
    case MP3B_STAGE_XR: src = ctx->d_xr.p; es = 4; n = (uint64_t)ctx->nunits * 576; break;
    case MP3B_STAGE_SB: src = ctx->d_sb.p; es = 4; n = (uint64_t)ctx->nunits * 576; break;
    default: return MP3B_E_INVAL;
    }
    if (stage >= MP3B_STAGE_IS && stage <= MP3B_STAGE_SB && !ctx->opts.keep_stages) {
        ctx->err = "intermediates are only kept with opts.keep_stages = 1";
        return MP3B_E_STATE;
    }
    if (elem_size) *elem_size = es;
    if (count) *count = n;
    if (!dst) return MP3B_OK;
    if (cap_bytes < n * es) return MP3B_E_TRUNCATED;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n) CK(cudaMemcpy(dst, src, n * es, cudaMemcpyDeviceToHost));
    return MP3B_OK;
}

} // extern "C"
