/* mp3b.h -- C-ABI of libmp3b, the B200-native batched MPEG-1 / 2 / 2.5 Layer III (and Layer II / I) decoder.
 *
 * Reference interface replaced: NONE EXISTS.  The reference repository (lxm0851/mp3) exposes no
 * plugin / operator / FFI boundary and names no decoder library: /root/reference/README.md:1-84
 * is prose (the only technical statement is "audio player", README.md:2; the only input format is
 * "user-uploaded audio", README.md:71).  The boundary below is therefore the one SURVEY.md
 * section 8(b) specifies for the inferred hot path (open stream / enqueue bytes / decode / fetch
 * PCM, plus a bulk batch entry), in the shape a player written in a scripting language would
 * bind through ctypes/cffi (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C types only; no C++ types, exceptions, errno or signals cross this boundary;
 *   - every function returning int returns MP3B_OK (0) or a negative mp3b_status; mp3b_ctx_create rejects
 *     option values outside the enums below with MP3B_E_INVAL before it touches a device;
 *   - one mp3b_ctx per GPU, owned by one thread; different contexts share nothing, so N GPUs are
 *     N contexts in N threads or processes with no communication (no NCCL);
 *   - corrupt or undecodable frames are never fatal: they are skipped (sync search) or
 *     concealed as silence, and counted in mp3b_stats;
 *   - there is NO CPU fallback: if no CUDA device is usable, mp3b_ctx_create fails with
 *     MP3B_E_CUDA.
 */
#ifndef MP3B_H
#define MP3B_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP3B_ABI_VERSION 1

typedef enum mp3b_status {
    MP3B_OK = 0,
    MP3B_E_INVAL = -1,       /* bad argument */
    MP3B_E_NOSYNC = -2,      /* no MPEG audio frame found in a stream */
    MP3B_E_TRUNCATED = -3,   /* destination buffer too small */
    MP3B_E_UNSUPPORTED = -4, /* free format, unusable sample-rate pair */
    MP3B_E_CUDA = -5,        /* CUDA runtime error; mp3b_last_error() has the text */
    MP3B_E_NOMEM = -6,
    MP3B_E_STATE = -7        /* call order violated (e.g. fetch before decode) */
} mp3b_status;

typedef enum mp3b_pcm_format {
    MP3B_PCM_S16 = 0, /* interleaved int16, round-to-nearest, saturated */
    MP3B_PCM_F32 = 1  /* interleaved float32, full scale +-1.0, not clipped */
} mp3b_pcm_format;

typedef enum mp3b_where { MP3B_HOST = 0, MP3B_DEVICE = 1 } mp3b_where;

typedef enum mp3b_indexer {
    MP3B_INDEX_DEVICE = 0, /* frame scan + side-info parse on the GPU (default) */
    MP3B_INDEX_HOST = 1    /* frame scan on host threads, table uploaded */
} mp3b_indexer;

typedef enum mp3b_pipeline {
    MP3B_PIPE_FUSED = 0,  /* default: fewest HBM round trips */
    MP3B_PIPE_STAGED = 1  /* one kernel per stage, every intermediate in HBM (debug / parity) */
} mp3b_pipeline;

typedef struct mp3b_opts {
    uint32_t struct_size;  /* = sizeof(mp3b_opts); lets the struct grow compatibly */
    int32_t pcm_format;    /* mp3b_pcm_format */
    int32_t indexer;       /* mp3b_indexer */
    int32_t pipeline;      /* mp3b_pipeline */
    int32_t host_threads;  /* host indexer / gather threads; 0 = hardware concurrency */
    int32_t keep_stages;   /* 1 = keep intermediates addressable through mp3b_debug_stage */
    int32_t async_index;   /* 1 (default) = the frame walk of a decode call runs ahead on a private CUDA
                              stream, overlapping the previous call's kernels.  Device-resident input must
                              then be completely written when the call is made (it is not ordered behind
                              work the caller queued on the context's stream).  0 = strictly stream-ordered. */
    int32_t gapless;       /* 1 = mp3b_stream_info.samples / pcm_offset of the batch interface describe the
                              gapless window of mp3b_tag_info instead of everything decoded.  Default 0. */
    int32_t verify_crc;    /* 1 = check the CRC-16 of protected frames (header protection bit 0); a frame
                              that fails is concealed as silence and counted.  Default 0: the word is
                              skipped, which is what common decoders do. */
} mp3b_opts;

typedef struct mp3b_stream_info {
    int32_t sample_rate;
    int32_t channels;
    int32_t lsf;            /* 0 = MPEG-1, 1 = MPEG-2 LSF or MPEG-2.5 (sample_rate tells which) */
    int32_t reserved;       /* (a Layer II stream reports the same fields; 1152 samples per frame at every rate) */
    int64_t frames;
    int64_t samples;        /* per channel */
    int64_t concealed_frames;
    int64_t pcm_offset;     /* element offset of this stream's PCM in the batch PCM arena */
    int64_t total_samples;  /* stream interface: samples per channel emitted since mp3b_stream_open */
} mp3b_stream_info;

/* The encoder's tag frame (the first frame of many files: "Xing" / "Info", optionally with the LAME
 * extension, or "VBRI") and the gapless window it implies inside the stream's decoded PCM: the tag
 * frame is not audio; with LAME delay / padding fields the first enc_delay + 528 + 1 samples and the
 * last enc_padding samples are encoder / decoder latency, not signal. */
typedef struct mp3b_tag_info {
    int32_t kind;          /* 0 = no tag, 1 = Xing, 2 = Info, 3 = VBRI */
    int32_t has_lame;      /* delay / padding fields present */
    uint32_t frames;       /* frame count announced by the tag (0 if absent) */
    uint32_t bytes;        /* byte count announced by the tag (0 if absent) */
    int32_t enc_delay, enc_padding;
    int64_t first_sample;  /* window start, samples per channel from the start of the decoded stream */
    int64_t num_samples;   /* window length, samples per channel */
} mp3b_tag_info;

typedef struct mp3b_stats {
    int64_t streams, frames, granules, units; /* unit = one granule-channel (576 lines) */
    int64_t bytes_in, pcm_bytes;
    int64_t concealed_frames;
    int64_t kernel_launches;  /* kernels launched by the last decode call */
    float ms_index, ms_huffman, ms_requant, ms_imdct, ms_overlap, ms_synth, ms_fused, ms_total;
} mp3b_stats;

typedef struct mp3b_ctx mp3b_ctx;
typedef struct mp3b_stream mp3b_stream;

/* ---- context ------------------------------------------------------------------------------ */
int mp3b_abi_version(void);
int mp3b_device_count(void);
void mp3b_opts_default(mp3b_opts *o);
int mp3b_ctx_create(int device, const mp3b_opts *opts, mp3b_ctx **out);
void mp3b_ctx_destroy(mp3b_ctx *ctx);
/* Run all work of the context on the caller's CUDA stream (a cudaStream_t passed as void*), so that
 * the caller's events / graphs / other kernels order with the decode.  NULL restores the
 * context's own stream.  The caller keeps ownership of the stream. */
int mp3b_ctx_set_stream(mp3b_ctx *ctx, void *cuda_stream);
/* Per-stage timing (mp3b_stats.ms_index / ms_huffman / ms_fused ...): on = CUDA events are recorded between the
 * kernels of a decode.  Off (default; MP3B_STAGE_TIMING=1 in the environment turns it on) = only ms_total is
 * measured (0.03 ms less per 1,024-stream step), and with MP3B_PDL=1 every kernel of the chain is launched as a
 * programmatic dependent of the one before it (measured: no further gain, hence off by default). */
int mp3b_ctx_set_stage_timing(mp3b_ctx *ctx, int on);
const char *mp3b_strerror(int status);
const char *mp3b_last_error(const mp3b_ctx *ctx);

/* Pinned host memory (page-locked) so that H2D / D2H copies run at PCIe speed and async. */
void *mp3b_host_alloc(size_t bytes);
void mp3b_host_free(void *p);

/* ---- bulk batch decode ----------------------------------------------------------------------
 * Streams are given either as an array of host pointers (gathered by the library) or packed in one
 * buffer with offsets (offsets[nstreams] = end of the last stream), host or device resident.
 * The call enqueues everything on the context's CUDA stream; mp3b_sync() waits for it.
 * Results stay in the context (device memory) until the next decode call or mp3b_ctx_destroy. */
int mp3b_decode_batch(mp3b_ctx *ctx, const uint8_t *const *bufs, const size_t *lens, int nstreams);
int mp3b_decode_packed(mp3b_ctx *ctx, const uint8_t *base, const uint64_t *offsets, int nstreams,
                       int where /* mp3b_where: where `base` lives */);
int mp3b_sync(mp3b_ctx *ctx);
/* Asynchronous join: makes the context's CUDA stream wait for the sink copies issued so far (see
 * mp3b_set_pcm_sink), without blocking the host.  An event the caller records on that stream after
 * mp3b_flush() therefore covers the D2H transfers too. */
int mp3b_flush(mp3b_ctx *ctx);

int mp3b_batch_stream_info(const mp3b_ctx *ctx, int stream_index, mp3b_stream_info *info);
int mp3b_batch_tag_info(const mp3b_ctx *ctx, int stream_index, mp3b_tag_info *info);
/* Whole-batch PCM arena: streams back to back in input order, interleaved channels; stream i starts at
 * mp3b_stream_info.pcm_offset (a Layer I stream, whose frames are 384 samples, is padded to a multiple of
 * 576 samples per channel). */
int mp3b_batch_pcm_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems);
/* Copy the whole arena (nelems elements of the context's pcm_format) to `dst`; async when dst is
 * pinned; call mp3b_sync() before reading it. */
int mp3b_batch_fetch_pcm(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got);
int mp3b_get_stats(const mp3b_ctx *ctx, mp3b_stats *st);
/* Streaming PCM sink: register a host buffer (pinned memory from mp3b_host_alloc for full PCIe speed)
 * of cap_elems PCM elements.  Every following decode call splits the batch into waves and copies each
 * wave's PCM to host_dst + its arena offset on a second CUDA stream while the next wave decodes, so
 * the D2H transfer overlaps the kernels -- and, because two PCM arenas alternate, the next decode call's
 * upload and kernels as well.  mp3b_sync() waits for the copies; mp3b_flush() orders them into the
 * context's stream.  The device PCM of a call stays valid until the call after the next one.
 * host_dst = NULL removes the sink.  A too-small sink makes the decode call fail with TRUNCATED. */
int mp3b_set_pcm_sink(mp3b_ctx *ctx, void *host_dst, uint64_t cap_elems);

/* ---- stream interface (open / enqueue / decode / fetch) -------------------------------------
 * Incremental: bytes may be enqueued in pieces of any size; each mp3b_decode() decodes the frames
 * completed since the previous call (for every open stream, as one batch) and mp3b_stream_fetch_pcm
 * returns exactly their samples -- concatenated over the calls, the PCM equals a one-shot decode of
 * the whole stream.  PCM of a call that is not fetched before the next mp3b_decode() is dropped; closing one
 * stream does not touch the others' unfetched PCM.
 * frames / samples in mp3b_stream_info describe the last call, total_samples the stream's life.
 * A leading ID3v2 tag is dropped as its bytes arrive, however large.  A stream's first frame is emitted once the
 * header of the frame behind it has arrived (sync confirmation; a lone look-alike in junk must not fix the stream's
 * identity), so a stream that consists of a single frame decodes in the batch interface only. */
int mp3b_stream_open(mp3b_ctx *ctx, mp3b_stream **out);
void mp3b_stream_close(mp3b_stream *s);
/* Appends raw MP3 bytes; the library copies, the caller may free its buffer on return. */
int mp3b_stream_enqueue(mp3b_stream *s, const uint8_t *bytes, size_t n);
/* Decodes everything enqueued so far on every open stream of the context (async), as one batch. */
int mp3b_decode(mp3b_ctx *ctx);
int mp3b_stream_get_info(const mp3b_stream *s, mp3b_stream_info *info);
/* Copies up to cap_samples (per channel) of not-yet-fetched PCM, advancing the stream's read
 * cursor; *got = samples per channel copied. */
int mp3b_stream_fetch_pcm(mp3b_stream *s, void *dst, size_t cap_samples, int where, size_t *got);
/* Zero-copy view of the stream's decoded PCM (valid until the next decode call / close). */
int mp3b_stream_pcm_device_ptr(const mp3b_stream *s, const void **ptr, size_t *nsamples);

/* ---- sample-rate conversion of the decoded batch ----------------------------------------------
 * The output step after the decode: every stream of the last batch (its gapless window when
 * opts.gapless is set) is converted to out_rate by a polyphase Kaiser-windowed-sinc FIR (FP32
 * accumulation, 32 zero crossings each side, pass band to 0.82 of the lower Nyquist frequency, about
 * -90 dB from that Nyquist frequency up), into a second arena of the
 * context's pcm_format, streams in input order, each starting on a 16-byte boundary.  Stream i then holds
 * ceil(samples * out_rate / rate) frames at mp3b_batch_resampled_info's offset.  Asynchronous on the context's stream; the arena
 * stays valid until the next decode or resample call. */
int mp3b_batch_resample(mp3b_ctx *ctx, int out_rate);
int mp3b_batch_resampled_info(const mp3b_ctx *ctx, int stream_index, int64_t *offset_elems, int64_t *samples);
int mp3b_batch_resampled_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems);
int mp3b_batch_fetch_resampled(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got);
/* The FIR used for in_rate -> out_rate, for verification: out_rate / in_rate = L / M in lowest terms;
 * taps[p * taps_per_phase + j] = h[p + j L], h centred at half * L with taps_per_phase = 2 half + 1.
 * Returns MP3B_OK, MP3B_E_TRUNCATED (cap too small; *ncoef = needed) or MP3B_E_INVAL. */
/* The tensor-core resampler's plan for a rate pair (verification, no GPU needed): tiles of 128 outputs come in
 * *nkinds kinds; a tile of kind k starting at output n0 = 128 t (t = k mod *nkinds) is y[n0 + r] = sum_c a[r][c] x[base + c],
 * base = (n0 M + half L) div L - (taps_per_phase - 1), c < *kpad.  a (optional): kind `kind`'s matrix, [128][*kpad] floats.
 * MP3B_E_UNSUPPORTED: the pair is not served by that path (mp3b_batch_resample then uses the FP32 kernels). */
int mp3b_resample_tc_plan(int in_rate, int out_rate, int kind, float *a, size_t cap, int *nkinds, int *kpad);
int mp3b_resample_filter(int in_rate, int out_rate, float *taps, size_t cap, size_t *ncoef, int *L, int *M,
                         int *taps_per_phase);

/* ---- planar copy of the decoded batch -------------------------------------------------------------
 * The decoder's arena is interleaved; this makes a second arena in which stream i (same pcm_offset, same
 * format) is laid out [channel][sample]: channel c of stream i starts at pcm_offset + c * samples.
 * Asynchronous on the context's stream; valid until the next decode or planar call. */
int mp3b_batch_planar(mp3b_ctx *ctx);
int mp3b_batch_planar_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems);
int mp3b_batch_fetch_planar(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got);

/* ---- slow / fast playback of the decoded batch without pitch change -----------------------------
 * Waveform-similarity overlap-add (WSOLA): every stream of the last batch (its gapless window when
 * opts.gapless is set) is stretched to floor(samples * den / num) frames, speed = num / den (1/2 = half
 * speed, the "slow listening" of the reference's README).  Segments of 2 * hop samples (hop = 512 / 256 /
 * 128 by sample rate) are re-spaced and each is shifted by up to hop / 2 samples to where it best
 * continues the previous one (exact integer cross-correlation of an 8-bit alignment signal), then
 * cross-faded with a Hann window.  Result in a third arena of the context's pcm_format, streams in input
 * order, each starting on a 16-byte boundary (mp3b_batch_stretched_info gives the offset); asynchronous on the
 * context's stream; valid until the next decode or stretch call. */
int mp3b_batch_time_stretch(mp3b_ctx *ctx, int speed_num, int speed_den);
int mp3b_batch_stretched_info(const mp3b_ctx *ctx, int stream_index, int64_t *offset_elems, int64_t *samples);
int mp3b_batch_stretched_device_ptr(const mp3b_ctx *ctx, const void **ptr, uint64_t *nelems);
int mp3b_batch_fetch_stretched(mp3b_ctx *ctx, void *dst, uint64_t cap_elems, int where, uint64_t *got);
/* The alignment offsets chosen for stream i (one per output segment of `hop` frames), for verification. */
int mp3b_batch_stretch_offsets(mp3b_ctx *ctx, int stream_index, int32_t *dst, size_t cap, size_t *n, int *hop);

/* ---- sentence boundaries of the last batch ------------------------------------------------------
 * The other half of the reference's "repeat each sentence" feature (/root/reference/README.md:46): where the
 * pauses are.  Integer arithmetic only, so the result is exactly reproducible: the PCM as s16 (float PCM is
 * rounded), mono mix (L + R) >> 1, windows of sample_rate / 100 samples (10 ms), window energy = sum of
 * squares; a window is silent iff energy <= threshold^2 x its length (threshold = RMS level in s16 units,
 * 1 .. 32767; 328 is about -40 dBFS); a pause is a run of >= min_silence_ms / 10 silent windows; a sentence
 * is a maximal run of windows that starts and ends voiced and holds no pause; sentences shorter than
 * min_sentence_ms are dropped.  mp3b_batch_fetch_segments copies stream i's sentences to dst as
 * {first sample, end sample} pairs (per-channel sample indices relative to the stream's PCM, i.e. to
 * mp3b_stream_info.pcm_offset); *n = the number of pairs (MP3B_E_TRUNCATED if cap pairs are too few).
 * mp3b_batch_window_energy returns the window energies the decision was made on. */
int mp3b_batch_segments(mp3b_ctx *ctx, int threshold, int min_silence_ms, int min_sentence_ms);
int mp3b_batch_fetch_segments(mp3b_ctx *ctx, int stream_index, int64_t *dst, size_t cap_pairs, size_t *n);
int mp3b_batch_window_energy(mp3b_ctx *ctx, int stream_index, uint64_t *dst, size_t cap, size_t *n, int *window);

/* ---- host-side frame index of one stream (no GPU involved) ---------------------------------
 * The frame walk of the host indexer (MP3B_INDEX_HOST) as a utility: sync search past ID3v2 / junk,
 * header validation, stream consistency, tag frame.  frames[i] = {byte offset of the header, main-data
 * bytes of the stream before this frame, the 4 header bytes big-endian, 0}.  Returns MP3B_OK,
 * MP3B_E_NOSYNC if no Layer I / II / III frame was found, MP3B_E_TRUNCATED if cap_frames was too small
 * (*nframes then holds the number needed). */
typedef struct mp3b_frame_rec { uint32_t offset, payload_offset, header, reserved; } mp3b_frame_rec;
int mp3b_index_stream_host(const uint8_t *bytes, size_t n, mp3b_frame_rec *frames, size_t cap_frames,
                           size_t *nframes, mp3b_stream_info *info, mp3b_tag_info *tag);

/* ---- seek (no GPU involved) ------------------------------------------------------------------
 * Where to start feeding a stream so that the PCM from `target_sample` on (per-channel sample index
 * of the un-trimmed decode, i.e. before any gapless window) is bit-identical to a decode from the
 * stream's start: far enough back that the target frame's two preceding granules decode from complete
 * main data (bit reservoir: up to 511 bytes behind their headers), which re-derives the overlap and the
 * synthesis history exactly.  Decode bytes[byte_offset ..] as a stream of its own (batch or
 * mp3b_stream_open / enqueue) and drop the first discard_samples samples per channel; the frames in
 * front of the target whose own reservoir is missing are concealed (and counted) -- they only warm up
 * state.  `frames` is the table mp3b_index_stream_host returned for these bytes.  This is the
 * "repeat the sentence" seek of the reference's player (/root/reference/README.md:46). */
typedef struct mp3b_seek {
    uint64_t byte_offset;     /* offset of the header of first_frame */
    uint32_t first_frame;     /* frame to start feeding from */
    uint32_t target_frame;    /* frame that holds target_sample */
    int64_t discard_samples;  /* samples per channel to drop from the decode of bytes[byte_offset ..] */
} mp3b_seek;
int mp3b_seek_plan(const uint8_t *bytes, size_t n, const mp3b_frame_rec *frames, size_t nframes, int64_t target_sample,
                   mp3b_seek *out);

/* ---- debug / parity access to intermediates (keep_stages = 1) -------------------------------
 * stage: see mp3b_stage.  Copies the whole stage array for the last batch to host memory.
 * *elem_size receives the element size in bytes, *count the number of elements. */
typedef enum mp3b_stage {
    MP3B_STAGE_FRAMES = 0,   /* uint32[4] per frame: rel_off, payload_off, header, stream */
    MP3B_STAGE_UNITDESC = 1, /* 32-byte unit descriptors */
    MP3B_STAGE_IS = 2,       /* int16[576] per unit: Huffman output */
    MP3B_STAGE_SF = 3,       /* uint8[40] per unit: scalefactors in band order */
    MP3B_STAGE_XR = 4,       /* float[576] per unit: after requant/stereo/reorder/alias */
    MP3B_STAGE_SB = 5,       /* float[18][32] per unit: subband samples */
    MP3B_STAGE_MAINDATA = 6  /* uint8: compacted main-data arena */
} mp3b_stage;
int mp3b_debug_stage(mp3b_ctx *ctx, int stage, void *dst, uint64_t cap_bytes, uint32_t *elem_size,
                     uint64_t *count);

#ifdef __cplusplus
}
#endif
#endif /* MP3B_H */
