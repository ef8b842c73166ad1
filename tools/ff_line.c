/* ff_line.c -- a C loop around FFmpeg's mp3float decoder, for the CPU lines of bench.py ("also
 * time FFmpeg mp3float ... as a second CPU line").  The Python driver tests/ffmpeg_ref.py spends several
 * microseconds of interpreter time per frame and holds the GIL between frames, which would cap a multi-thread
 * line; here a whole stream is decoded inside one foreign call.
 *
 * libavcodec ships without headers on this image (it is bundled in the opencv wheel), so the few functions are
 * resolved with dlsym from the libraries tests/ffmpeg_ref.py has already loaded RTLD_GLOBAL, and the struct
 * fields are touched at the hand-verified offsets of SURVEY.md 8(c): AVPacket.data @24, .size @32;
 * AVFrame.nb_samples @112.  Not part of the product and not part of the oracle: a baseline timing aid.
 * Build: gcc -O2 -shared -fPIC tools/ff_line.c -o tools/libffline.so -ldl */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef void *(*fn_find)(const char *);
typedef void *(*fn_alloc_ctx)(void *);
typedef int (*fn_open)(void *, void *, void *);
typedef void *(*fn_pkt_alloc)(void);
typedef void (*fn_pkt_free)(void **);
typedef int (*fn_send)(void *, void *);
typedef int (*fn_recv)(void *, void *);
typedef void (*fn_free_ctx)(void **);
typedef void *(*fn_frame_alloc)(void);
typedef void (*fn_frame_unref)(void *);
typedef void (*fn_frame_free)(void **);

static struct {
    fn_find find; fn_alloc_ctx alloc_ctx; fn_open open; fn_pkt_alloc pkt_alloc; fn_pkt_free pkt_free; fn_send send;
    fn_recv recv; fn_free_ctx free_ctx; fn_frame_alloc frame_alloc; fn_frame_unref frame_unref; fn_frame_free frame_free;
    int ok;
} F;

int ffl_init(void)
{
    if (F.ok) return 0;
    F.find = (fn_find)dlsym(RTLD_DEFAULT, "avcodec_find_decoder_by_name");
    F.alloc_ctx = (fn_alloc_ctx)dlsym(RTLD_DEFAULT, "avcodec_alloc_context3");
    F.open = (fn_open)dlsym(RTLD_DEFAULT, "avcodec_open2");
    F.pkt_alloc = (fn_pkt_alloc)dlsym(RTLD_DEFAULT, "av_packet_alloc");
    F.pkt_free = (fn_pkt_free)dlsym(RTLD_DEFAULT, "av_packet_free");
    F.send = (fn_send)dlsym(RTLD_DEFAULT, "avcodec_send_packet");
    F.recv = (fn_recv)dlsym(RTLD_DEFAULT, "avcodec_receive_frame");
    F.free_ctx = (fn_free_ctx)dlsym(RTLD_DEFAULT, "avcodec_free_context");
    F.frame_alloc = (fn_frame_alloc)dlsym(RTLD_DEFAULT, "av_frame_alloc");
    F.frame_unref = (fn_frame_unref)dlsym(RTLD_DEFAULT, "av_frame_unref");
    F.frame_free = (fn_frame_free)dlsym(RTLD_DEFAULT, "av_frame_free");
    if (!F.find || !F.alloc_ctx || !F.open || !F.pkt_alloc || !F.pkt_free || !F.send || !F.recv || !F.free_ctx ||
        !F.frame_alloc || !F.frame_unref || !F.frame_free)
        return -1;
    F.ok = 1;
    return 0;
}

/* Decode the frames data[off[i] .. off[i+1]) of one stream with one decoder context.  Returns the samples per
 * channel produced, or < 0.  (The PCM stays in FFmpeg's frame buffers: decode throughput is what is timed.) */
long ffl_decode_stream(const uint8_t *data, const uint32_t *off, int nframes, const char *codec_name)
{
    if (ffl_init() != 0) return -1;
    void *codec = F.find(codec_name);
    if (!codec) return -2;
    void *ctx = F.alloc_ctx(codec);
    if (!ctx || F.open(ctx, codec, NULL) != 0) return -3;
    void *pkt = F.pkt_alloc(), *frm = F.frame_alloc();
    uint8_t *buf = (uint8_t *)malloc(4096 + 64);
    long total = 0;
    for (int i = 0; i < nframes; i++) {
        const uint32_t n = off[i + 1] - off[i];
        if (n > 4096) continue;
        memcpy(buf, data + off[i], n);
        memset(buf + n, 0, 64);
        *(uint8_t **)((char *)pkt + 24) = buf;
        *(int *)((char *)pkt + 32) = (int)n;
        if (F.send(ctx, pkt) != 0) continue;
        if (F.recv(ctx, frm) != 0) continue;
        total += *(int *)((char *)frm + 112);
        F.frame_unref(frm);
    }
    *(uint8_t **)((char *)pkt + 24) = NULL;
    *(int *)((char *)pkt + 32) = 0;
    free(buf);
    F.pkt_free(&pkt);
    F.frame_free(&frm);
    F.free_ctx(&ctx);
    return total;
}
