/* A host with no Python and no CUDA headers: loads a checkpoint dump and a batch of clips, runs encode_audio + decode_audio
 * through the step-level C ABI (include/l3ac_b200.h: l3ac_create / l3ac_encode_host / l3ac_decode_host) and writes the
 * token indices and the decoded waveform.  Built with plain gcc by tests/test_codec_gpu.py, which compares the outputs
 * with the Python host's.
 *
 *   codec_driver <weights.bin> <audio.bin> <out.bin> [repeats]
 *
 * weights.bin: "L3ACW1\0\0", l3ac_codec_config, int32 n, then n x { int32 name_len, name, int64 numel, float data[numel] }
 * audio.bin:   int32 B, int32 T, float audio[B*T]
 * out.bin:     int32 B, int32 T_tok, int32 hop, int32 indices[B*T_tok], float audio[B*T_tok*hop]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "l3ac_b200.h"

static void die(const char* what) {
    fprintf(stderr, "codec_driver: %s (%s)\n", what, l3ac_last_error());
    exit(1);
}

static void rd(void* dst, size_t n, FILE* f) {
    if (fread(dst, 1, n, f) != n) die("short read");
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s weights.bin audio.bin out.bin [repeats]\n", argv[0]);
        return 2;
    }
    const int repeats = argc > 4 ? atoi(argv[4]) : 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) die("cannot open the weights file");
    char magic[8];
    rd(magic, 8, f);
    if (memcmp(magic, "L3ACW1\0\0", 8) != 0) die("bad magic");
    l3ac_codec_config cfg;
    rd(&cfg, sizeof cfg, f);
    int32_t n;
    rd(&n, 4, f);
    l3ac_tensor* tensors = calloc((size_t)n, sizeof *tensors);
    for (int i = 0; i < n; ++i) {
        int32_t len;
        int64_t numel;
        rd(&len, 4, f);
        char* name = calloc((size_t)len + 1, 1);
        rd(name, (size_t)len, f);
        rd(&numel, 8, f);
        float* data = malloc((size_t)numel * 4);
        rd(data, (size_t)numel * 4, f);
        tensors[i].name = name;
        tensors[i].data = data;
        tensors[i].numel = numel;
    }
    fclose(f);

    f = fopen(argv[2], "rb");
    if (!f) die("cannot open the audio file");
    int32_t B, T;
    rd(&B, 4, f);
    rd(&T, 4, f);
    float* audio = malloc((size_t)B * T * 4);
    rd(audio, (size_t)B * T * 4, f);
    fclose(f);

    l3ac_codec* codec = NULL;
    if (l3ac_create(&cfg, tensors, n, &codec) != L3AC_OK) die("l3ac_create");
    const int hop = l3ac_hop_length(codec);
    const int T_tok = (T + hop - 1) / hop;
    int32_t* indices = malloc((size_t)B * T_tok * 4);
    float* wav = malloc((size_t)B * T_tok * hop * 4);
    double best = 1e30;
    for (int r = 0; r < repeats; ++r) {
        const double t0 = now_ms();
        if (l3ac_encode_host(codec, audio, B, T, indices, NULL) != L3AC_OK) die("l3ac_encode_host");
        if (l3ac_decode_host(codec, indices, B, T_tok, wav) != L3AC_OK) die("l3ac_decode_host");
        const double dt = now_ms() - t0;
        if (dt < best) best = dt;
    }
    printf("{\"B\": %d, \"T\": %d, \"hop\": %d, \"ms_encode_decode\": %.3f, \"audio_s_per_s\": %.1f, \"launches\": %lld}\n", B, T, hop, best,
           (double)B * T / 16000.0 / (best * 1e-3), l3ac_launch_count(codec));

    f = fopen(argv[3], "wb");
    if (!f) die("cannot open the output file");
    int32_t hdr[3] = {B, T_tok, hop};
    fwrite(hdr, 4, 3, f);
    fwrite(indices, 4, (size_t)B * T_tok, f);
    fwrite(wav, 4, (size_t)B * T_tok * hop, f);
    fclose(f);
    if (l3ac_destroy(codec) != L3AC_OK) die("l3ac_destroy");
    return 0;
}
