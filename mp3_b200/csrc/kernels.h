/* kernels.h -- launch wrappers of the CUDA kernels (one translation unit per stage). */
#ifndef MP3B_KERNELS_H
#define MP3B_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "l3_defs.h"
#include "l3_tables.h"

/* cudaFuncSetAttribute is per device: launchers that raise a kernel's dynamic shared-memory limit do it
 * once per device of the process (contexts on several GPUs may live in one process, on several threads). */
#ifdef __cplusplus
#include <atomic>
/* usage: if (l3_device_needs_setup(mask)) { cudaFuncSetAttribute(...); l3_device_setup_done(mask); }
 * (two threads may both run the set-up: harmless; neither launches before one of them has finished) */
static inline unsigned long long l3_device_bit(void)
{
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
}
static inline bool l3_device_needs_setup(const std::atomic<unsigned long long> &mask) { return (mask.load() & l3_device_bit()) == 0; }
static inline void l3_device_setup_done(std::atomic<unsigned long long> &mask) { mask.fetch_or(l3_device_bit()); }

/* Programmatic dependent launch (sm_90+): with `pdl` the kernel may be scheduled while the kernel before it in the
 * stream is still draining -- its CTAs run their prologue (tables into shared memory, barriers) on the SMs the
 * predecessor's tail leaves idle and block in pdl_wait() until the predecessor has completed and its writes are
 * visible.  Every kernel of the decode chain calls pdl_launch_dependents() first thing and pdl_wait() before its first
 * access to anything an earlier kernel wrote; without the launch attribute both are no-ops. */
#include <utility>
template <typename... KArgs, typename... Args>
static inline cudaError_t l3_launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                      Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
#endif

/* Device-resident tables (pointers into one allocation owned by the context). */
struct L3DevTables {
    const uint16_t *huff_lut;
    uint32_t huff_lut_len;
    const L3HuffInfo *huff;
    const uint8_t *quad_a;
    const L3BandTables *bands;
    const float *pow43;
    const uint16_t *sfb_long; /* [9][23] */
};

/* K0: device frame indexer (a1-a3) */
/* scratch: at least l3_index_scratch_records(raw_total, nstreams) records */
static inline uint64_t l3_index_scratch_records(uint64_t raw_total, uint64_t nstreams)
{
    return raw_total / 24 + 2 * nstreams + 2;
}
/* device -> pinned host through kernel stores (no DMA engine involved); bytes is a multiple of 4 */
void l3_launch_publish(const void *src_dev, void *dst_pinned_host, size_t bytes, cudaStream_t st);
void l3_launch_index_walk(const uint8_t *raw, L3StreamRec *streams, int nstreams, L3FrameRec *scratch,
                          cudaStream_t st);
/* The time-parallel walk (segments of seg_bytes walked speculatively by a thread each, then stitched per stream: four
 * kernels): same dense table as l3_launch_index_walk.  nslots = l3_walk_segments(); sparse: l3_walk_sparse_records()
 * records; segs: l3_walk_seg_bytes() bytes (the segment records and, behind them, one first-frame record and one first-slot index per stream). */
static inline uint64_t l3_walk_segments(uint64_t raw_total, uint64_t nstreams, uint32_t seg_bytes)
{
    return raw_total / seg_bytes + nstreams + 2;
}
static inline uint64_t l3_walk_sparse_records(uint64_t raw_total, uint64_t nstreams, uint32_t seg_bytes)
{
    return l3_walk_segments(raw_total, nstreams, seg_bytes) * (uint64_t)(seg_bytes / 24 + 2); /* l3wp_seg_cap() records per segment */
}
static inline uint64_t l3_walk_seg_bytes(uint64_t raw_total, uint64_t nstreams, uint32_t seg_bytes)
{
    return 16 * l3_walk_segments(raw_total, nstreams, seg_bytes) + (32 + 4) * nstreams + 32;
}
void l3_launch_index_walk_par(const uint8_t *raw, L3StreamRec *streams, int nstreams, L3FrameRec *dense, L3FrameRec *sparse,
                              void *segs, uint64_t nslots, uint32_t seg_bytes, cudaStream_t st);
/* scratch == NULL: `frames` is already dense (host indexer) */
void l3_launch_side_parse(const uint8_t *raw, const L3StreamRec *streams, int nstreams, L3FrameRec *frames,
                          const L3FrameRec *scratch, uint32_t nframes, const L3DevTables &T, L3UnitDesc *units,
                          uint32_t *gran_unit0, uint32_t *concealed_counter, int verify_crc, cudaStream_t st);
void l3_launch_payload_copy(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames,
                            uint32_t nframes, uint8_t *arena, cudaStream_t st, bool pdl = false);

/* K1: scalefactor + Huffman / count1 decode (a4, a5) */
/* arena_bytes: readable size of the main-data arena (a multiple of 4; reads beyond it return 0) */
/* units [u_lo, u_lo + nunits); output arrays are indexed by absolute unit id */
/* avg_unit_bytes: average main-data bytes per unit of the batch (sizes the shared-memory bit stage) */
void l3_launch_huffman_range(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                             uint32_t nunits, uint32_t avg_unit_bytes, const L3DevTables &T, int16_t *is_out,
                             uint8_t *sf_out,
                             uint8_t *nzv_out /* [unit]: 16-byte vectors of is_out that hold data */,
                             int zero_fill /* also write the all-zero tail */, cudaStream_t st, bool pdl = false);

/* The sorted variant (default; k_huffman.cu): the wave's units are counting-sorted by (big_values, block type) and
 * decoded by one persistent CTA per SM whose warps pull groups of 32 equal-length units.  Scratch per wave:
 * see L3HuffSort (ctl is zeroed by the launcher). */
struct L3HuffSort {
    uint16_t *keys;   /* [nunits] sort keys (pass A, then pass B) */
    uint32_t *perm;   /* [nunits] sorted order (pass A, then pass B) */
    uint32_t *state;  /* [nunits] per unit behind the pairs: bits consumed | line index << 13 */
    uint32_t *ctl;    /* l3_huff_sort_ctl_bytes(): histograms, cursors, group counters */
    int sm_count;
};
size_t l3_huff_sort_ctl_bytes(void);
void l3_launch_huffman_sorted(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                              uint32_t nunits, uint32_t avg_unit_bytes, const L3DevTables &T, const L3HuffSort &scr,
                              int16_t *is_out, uint8_t *sf_out, uint8_t *nzv_out, int zero_fill, cudaStream_t st,
                              bool pdl = false);

/* The warp-per-unit variant (speculative decode at 32 bit positions, chain followed by shuffles); uses scr.ctl only */
void l3_launch_huffman_warp(const uint8_t *arena, uint64_t arena_bytes, const L3UnitDesc *units, uint32_t u_lo,
                            uint32_t nunits, const L3DevTables &T, const L3HuffSort &scr, int16_t *is_out, uint8_t *sf_out,
                            uint8_t *nzv_out, int zero_fill, cudaStream_t st, bool pdl = false);

/* K2: requantise + stereo + reorder + alias reduction (a6-a8) */
/* granules [g_lo, g_lo + ngranules) */
void l3_launch_requant_range(const L3UnitDesc *units, const uint32_t *gran_unit0, uint32_t g_lo, uint32_t ngranules,
                             const int16_t *is_in, const uint8_t *sf_in, const L3DevTables &T, float *xr_out,
                             cudaStream_t st);

/* K3a: IMDCT + window (a9); K3b: overlap-add + frequency inversion (a10) */
void l3_requant_init(void);
void l3_hybrid_init(void);
void l3_launch_imdct_range(const L3UnitDesc *units, uint32_t u_lo, uint32_t nunits, const float *xr, float *imd,
                           cudaStream_t st);
void l3_launch_overlap_range(const L3UnitDesc *units, uint32_t u_lo, uint32_t nunits, const float *imd, float *sb,
                             cudaStream_t st);

/* Layer II frames -> subband samples [unit][18][32] (k_layer2.cu); frames of other layers are skipped */
void l3_layer2_init(void);
void l3_launch_layer2(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames, uint32_t nframes,
                      float *sb_out, cudaStream_t st);
void l3_launch_layer1(const uint8_t *raw, const L3StreamRec *streams, const L3FrameRec *frames, uint32_t nframes,
                      float *sb_out, cudaStream_t st);

/* K4: polyphase synthesis (a11) */
void l3_synth_init(void);
/* tiles[i] = {first global granule, number of granules (<= l3_synth_tile_granules())}; a tile
 * never spans two streams */
int l3_synth_tile_granules(void);
void l3_launch_synth(const uint2 *tiles, uint32_t ntiles, const uint32_t *gran_unit0, const float *sb,
                     const uint32_t *tile_shift /* null: sb is indexed by unit */, void *pcm,
                     int pcm_format, cudaStream_t st);

/* KF: fused back end (a6-a11).  tiles[i] = {first granule to output, granules to output,
 * warm-up granules before it (0..2) that re-derive the overlap / synthesis state, 0}. */
void l3_fused_init(void);
void l3_launch_backend(const uint4 *tiles, uint32_t ntiles, const uint32_t *gran_unit0, const L3UnitDesc *units,
                       const int16_t *is_in, const uint8_t *sf_in, const uint8_t *nzv_in, const L3DevTables &T,
                       void *pcm, int pcm_format, cudaStream_t st, bool pdl = false);


/* Sample-rate conversion of decoded PCM (k_resample.cu).  One job per stream; offsets in elements
 * of the arena's sample type, lengths in frames (samples per channel). */
#ifdef __cplusplus
#include <vector>
struct L3ResampleJob {
    long long in_off, in_n, out_off, out_n;
    int channels, pad;
};
/* Planar copy of the PCM arena (k_planar.cu): one job per stream with audio, tiles of 2048 frames. */
struct L3PlanarJob {
    long long off, samples; /* element offset of the stream in both arenas; frames */
    int channels, pad;
};
void l3_launch_planar(const void *in, void *out, int pcm_format, const L3PlanarJob *jobs, const uint32_t *tile_job,
                      const uint32_t *tile_first, uint32_t ntiles, cudaStream_t st);

/* Sentence boundaries (k_segments.cu): one job per stream with audio. */
struct L3SegJob {
    long long off, samples;      /* element offset of the stream in the PCM arena; frames */
    int channels, window;        /* samples per 10-ms window */
    unsigned win_base, nwin;     /* the stream's windows in the energy array */
    unsigned seg_base, seg_cap;  /* the stream's slots in the segment array */
};
void l3_launch_segments(const void *pcm, int pcm_format, const L3SegJob *jobs, int njobs, unsigned max_nwin,
                        unsigned long long *energy, unsigned long long thr2, int G, int S, long long *seg, int *nseg,
                        cudaStream_t st);

/* Time-scale modification (k_stretch.cu).  One job (and one CTA) per stream. */
struct L3StretchJob {
    long long in_off, in_n, out_off, out_n;
    int channels, hop;
};
int l3_stretch_hop(int sample_rate);
void l3_launch_stretch(const void *in, void *out, int pcm_format, const L3StretchJob *jobs, int njobs, int num, int den,
                       int *offsets_out, int max_frames, cudaStream_t st);
size_t l3_resample_design(int in_rate, int out_rate, std::vector<float> *hp, int *L, int *M, int *taps, int *half);
/* channel_mask: bit 0 = convert the mono jobs, bit 1 = the stereo jobs (those may have gone to the tensor-core path) */
void l3_launch_resample(const void *in, void *out, int pcm_format, const L3ResampleJob *jobs, int njobs,
                        long long max_out_n, const float *hp, int L, int M, int taps, int half, cudaStream_t st,
                        int channel_mask = 3);
/* Tensor-core path for stereo s16 jobs (k_resample_tc.cu): tiles of 128 outputs as [128 x K] x [K x 64] products. */
struct L3RsTcPlan {
    int L = 0, M = 0, taps = 0, half = 0, NK = 0, Kpad = 0;
    std::vector<uint16_t> A; /* [NK][hi, lo][128 x Kpad] fp16 bit patterns in the MMA's shared-memory layout */
};
bool l3_resample_tc_plan(const float *hp, int L, int M, int taps, int half, L3RsTcPlan *plan);
unsigned long long l3_resample_tc_prefix(const L3RsTcPlan &plan, const L3ResampleJob *jobs, int njobs,
                                         std::vector<uint32_t> *prefix);
void l3_launch_resample_tc(const void *in, void *out, const L3ResampleJob *jobs_dev, int njobs, const uint32_t *prefix_dev,
                           uint32_t max_entries_per_kind, const uint16_t *A_dev, const L3RsTcPlan &plan, int sm_count,
                           cudaStream_t st);
#endif

#endif
