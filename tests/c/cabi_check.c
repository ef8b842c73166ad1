/* A plain C99 caller of libmp3b.so: proves that include/mp3b.h is a C header and that the boundary is a
 * C-ABI (no C++ types, no exceptions).  Without a GPU it exercises what does not need one: version,
 * option defaults, error strings, the "no CPU fallback" failure of mp3b_ctx_create, the host frame index
 * and the resampling filter.  With a GPU (argv[1] = "gpu") it also decodes silent frames (KAT-0) and the
 * survey's single-spectral-line known answer (KAT-1: peak -0.8535741 at sample 922, energy 288.006, spot
 * values to 1e-6, support [1, 1631], right channel silent), in float and in s16, and through the
 * open / enqueue / decode / fetch interface.
 * Build: gcc -std=c99 -pedantic -Wall -Werror -Iinclude tests/c/cabi_check.c -Lmp3_b200 -lmp3b */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mp3b.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

/* [KAT-1, KAT-0, KAT-0]; is[0] = +1 in granule 0 / channel 0, global_gain 210, long block. */
static int kat1_check(void)
{
    static const unsigned char side[32] = {0x00, 0x00, 0x00, 0x04, 0x01, 0x69, 0x00, 0x21, 0x08, 0x00, 0x00, 0x00,
                                           0x0D, 0x20, 0x04, 0x21, 0x00, 0x00, 0x00, 0x01, 0xA4, 0x00, 0x84, 0x20,
                                           0x00, 0x00, 0x00, 0x34, 0x80, 0x10, 0x84, 0x00};
    static const struct { int at; double v; } spot[] = {{240, 0.04973330349}, {528, -0.07968214154},
        {800, -0.74641698599}, {922, -0.8535740971565247}, {1152, -0.50985199213}, {1300, -0.12107632309}};
    unsigned char stream[3 * 417];
    static float pcm[3 * 1152 * 2], pcm2[3 * 1152 * 2];
    static short s16[3 * 1152 * 2];
    const unsigned char *bufs[1];
    size_t lens[1], i, got_s = 0, at = 0;
    uint64_t got = 0;
    double energy = 0.0, peak = 0.0;
    mp3b_opts o;
    mp3b_ctx *ctx = NULL;
    mp3b_stream *st = NULL;
    mp3b_stream_info info;
    int first = -1, last = -1, pipe;

    memset(stream, 0, sizeof stream);
    for (i = 0; i < 3; i++) {
        stream[i * 417 + 0] = 0xFF; stream[i * 417 + 1] = 0xFB; stream[i * 417 + 2] = 0x90;
    }
    memcpy(stream + 4, side, 32);
    stream[36] = 0x50;
    bufs[0] = stream; lens[0] = sizeof stream;
    for (pipe = 0; pipe < 2; pipe++) { /* fused (default) and staged pipelines */
        mp3b_opts_default(&o);
        o.pcm_format = MP3B_PCM_F32;
        o.pipeline = pipe ? MP3B_PIPE_STAGED : MP3B_PIPE_FUSED;
        CHECK(mp3b_ctx_create(0, &o, &ctx) == MP3B_OK);
        CHECK(mp3b_decode_batch(ctx, bufs, lens, 1) == MP3B_OK);
        CHECK(mp3b_batch_fetch_pcm(ctx, pcm, 3 * 1152 * 2, MP3B_HOST, &got) == MP3B_OK && mp3b_sync(ctx) == MP3B_OK);
        CHECK(got == 3 * 1152 * 2);
        energy = 0.0; peak = 0.0; first = last = -1; at = 0;
        for (i = 0; i < 3 * 1152; i++) {
            const double l = pcm[2 * i];
            CHECK(pcm[2 * i + 1] == 0.0f); /* right channel: exactly silent */
            energy += l * l;
            if (fabs(l) > fabs(peak)) { peak = l; at = i; }
            if (l != 0.0) { if (first < 0) first = (int)i; last = (int)i; }
        }
        CHECK(at == 922 && fabs(peak - -0.8535740971565247) < 2e-6);
        CHECK(fabs(energy - 288.0059985) < 1e-3);
        CHECK(first >= 1 && first <= 4 && last <= 1631 && last >= 1600); /* FFmpeg and the oracle: exactly [1, 1631] */
        for (i = 0; i < sizeof spot / sizeof spot[0]; i++) CHECK(fabs((double)pcm[2 * spot[i].at] - spot[i].v) < 3e-6); /* FFmpeg's float32 answers */
        if (pipe == 0) {
            /* the same bytes through open / enqueue (two pieces) / decode / fetch: bit-identical PCM */
            CHECK(mp3b_stream_open(ctx, &st) == MP3B_OK);
            CHECK(mp3b_stream_enqueue(st, stream, 500) == MP3B_OK);
            CHECK(mp3b_stream_enqueue(st, stream + 500, sizeof stream - 500) == MP3B_OK);
            CHECK(mp3b_decode(ctx) == MP3B_OK && mp3b_sync(ctx) == MP3B_OK);
            CHECK(mp3b_stream_get_info(st, &info) == MP3B_OK && info.samples == 3 * 1152 && info.channels == 2);
            CHECK(mp3b_stream_fetch_pcm(st, pcm2, 3 * 1152, MP3B_HOST, &got_s) == MP3B_OK && got_s == 3 * 1152);
            CHECK(memcmp(pcm, pcm2, sizeof pcm) == 0);
            mp3b_stream_close(st);
        }
        mp3b_ctx_destroy(ctx);
        ctx = NULL;
    }
    /* s16: round to nearest of the float answer */
    CHECK(mp3b_ctx_create(0, NULL, &ctx) == MP3B_OK);
    CHECK(mp3b_decode_batch(ctx, bufs, lens, 1) == MP3B_OK);
    CHECK(mp3b_batch_fetch_pcm(ctx, s16, 3 * 1152 * 2, MP3B_HOST, &got) == MP3B_OK && mp3b_sync(ctx) == MP3B_OK);
    CHECK(s16[2 * 922] == -27970 && s16[2 * 922 + 1] == 0); /* round(-0.8535741 * 32768) */
    for (i = 0; i < 3 * 1152 * 2; i++) CHECK(abs((int)s16[i] - (int)lrint((double)pcm[i] * 32768.0)) <= 1);
    mp3b_ctx_destroy(ctx);
    return 0;
}

int main(int argc, char **argv)
{
    unsigned char stream[3 * 417];
    size_t i, nframes = 0, ncoef = 0;
    mp3b_frame_rec fr[8];
    mp3b_stream_info info;
    mp3b_tag_info tag;
    mp3b_opts o;
    mp3b_ctx *ctx = NULL;
    int L, M, T, rc;

    CHECK(mp3b_abi_version() == MP3B_ABI_VERSION);
    mp3b_opts_default(&o);
    CHECK(o.struct_size == sizeof o && o.pcm_format == MP3B_PCM_S16 && o.async_index == 1);
    CHECK(strlen(mp3b_strerror(MP3B_E_CUDA)) > 0 && strlen(mp3b_strerror(-12345)) > 0);

    /* three silent MPEG-1 Layer III frames: FF FB 90 00 + zeros */
    memset(stream, 0, sizeof stream);
    for (i = 0; i < 3; i++) {
        stream[i * 417 + 0] = 0xFF; stream[i * 417 + 1] = 0xFB; stream[i * 417 + 2] = 0x90;
    }
    rc = mp3b_index_stream_host(stream, sizeof stream, fr, 8, &nframes, &info, &tag);
    CHECK(rc == MP3B_OK && nframes == 3 && fr[1].offset == 417 && fr[2].header == 0xFFFB9000u);
    CHECK(info.sample_rate == 44100 && info.channels == 2 && info.samples == 3 * 1152 && tag.kind == 0);
    CHECK(mp3b_index_stream_host(stream, 3, fr, 8, &nframes, NULL, NULL) == MP3B_E_NOSYNC);
    CHECK(mp3b_resample_filter(44100, 48000, NULL, 0, &ncoef, &L, &M, &T) == MP3B_E_TRUNCATED);
    CHECK(L == 160 && M == 147 && ncoef == (size_t)L * (size_t)T);

    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        const unsigned char *bufs[1];
        size_t lens[1];
        short pcm[3 * 1152 * 2];
        uint64_t got = 0;
        mp3b_stats st;
        CHECK(mp3b_ctx_create(0, NULL, &ctx) == MP3B_OK);
        bufs[0] = stream; lens[0] = sizeof stream;
        CHECK(mp3b_decode_batch(ctx, bufs, lens, 1) == MP3B_OK);
        CHECK(mp3b_batch_fetch_pcm(ctx, pcm, 3 * 1152 * 2, MP3B_HOST, &got) == MP3B_OK && mp3b_sync(ctx) == MP3B_OK);
        CHECK(got == 3 * 1152 * 2);
        for (i = 0; i < got; i++) CHECK(pcm[i] == 0);
        CHECK(mp3b_get_stats(ctx, &st) == MP3B_OK && st.frames == 3 && st.kernel_launches > 0);
        mp3b_ctx_destroy(ctx);
        ctx = NULL;
        if (kat1_check() != 0) return 1;
        printf("cabi_check: gpu ok\n");
    } else {
        /* no usable GPU here: the product must refuse, not fall back */
        if (mp3b_device_count() == 0) CHECK(mp3b_ctx_create(0, NULL, &ctx) == MP3B_E_CUDA && ctx == NULL);
        printf("cabi_check: ok\n");
    }
    return 0;
}
