/* A plain C99 caller of libmp3b.so: proves that include/mp3b.h is a C header and that the boundary is a
 * C-ABI (no C++ types, no exceptions).  Without a GPU it exercises what does not need one: version,
 * option defaults, error strings, the "no CPU fallback" failure of mp3b_ctx_create, the host frame index
 * and the resampling filter.  With a GPU (argv[1] = "gpu") it also decodes a silent frame.
 * Build: gcc -std=c99 -pedantic -Wall -Werror -Iinclude tests/c/cabi_check.c -Lmp3_b200 -lmp3b */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mp3b.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

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
        printf("cabi_check: gpu ok\n");
    } else {
        /* no usable GPU here: the product must refuse, not fall back */
        if (mp3b_device_count() == 0) CHECK(mp3b_ctx_create(0, NULL, &ctx) == MP3B_E_CUDA && ctx == NULL);
        printf("cabi_check: ok\n");
    }
    return 0;
}
