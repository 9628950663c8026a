/* A plain C99 consumer of include/imgcorr.h: proves that the boundary is a C ABI (no C++ / torch types), that the header
 * compiles as C with -Wall -Wextra -Werror -pedantic, and that the library links and runs without Python.
 *   abi_smoke            : no-GPU checks (version, error reporting, argument validation)
 *   abi_smoke gpu H W N  : the full chain through imgcorr_correct_host on N synthetic uint16 frames; prints a checksum
 *                          of the output that tests/test_gpu_parity.py compares with the Python path. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "imgcorr.h"

static int fail(const char* what) {
    fprintf(stderr, "FAIL %s: %s\n", what, imgcorr_last_error());
    return 1;
}

int main(int argc, char** argv) {
    imgcorr_ctx* ctx = NULL;
    if (imgcorr_version() != IMGCORR_VERSION) return fail("version");
    if (imgcorr_ctx_create(0, 0, 16, &ctx) != IMGCORR_ERR_INVALID || ctx != NULL) return fail("bad shape accepted");
    if (strlen(imgcorr_last_error()) == 0) return fail("no message for a failure");
    if (imgcorr_set_option(NULL, IMGCORR_OPT_CHAIN_GROUP, 4) != IMGCORR_ERR_INVALID) return fail("null ctx accepted");
    if (imgcorr_ctx_destroy(NULL) != IMGCORR_OK) return fail("destroy(NULL)");
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
        int n = imgcorr_device_count();
        printf("abi ok, devices=%d\n", n);
        if (n <= 0 && imgcorr_ctx_create(0, 16, 16, &ctx) == IMGCORR_OK) return fail("context without a device");
        return 0;
    }
    {
        const int H = argc > 2 ? atoi(argv[2]) : 64, W = argc > 3 ? atoi(argv[3]) : 128, N = argc > 4 ? atoi(argv[4]) : 3;
        const size_t npx = (size_t)H * W;
        uint16_t* raw = NULL;
        float *out = NULL, *dark, *flat;
        double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1}, P[9], dist[5] = {-0.2, 0.05, 1e-3, -1e-3, 0.0}, sum = 0.0;
        size_t i;
        unsigned s = 12345u;
        if (imgcorr_ctx_create(0, H, W, &ctx) != IMGCORR_OK) return fail("ctx_create");
        if (imgcorr_host_alloc(npx * N * sizeof(uint16_t), (void**)&raw) != IMGCORR_OK) return fail("host_alloc raw");
        if (imgcorr_host_alloc(npx * N * sizeof(float), (void**)&out) != IMGCORR_OK) return fail("host_alloc out");
        dark = (float*)malloc(npx * sizeof(float));
        flat = (float*)malloc(npx * sizeof(float));
        for (i = 0; i < npx * N; ++i) { s = s * 1664525u + 1013904223u; raw[i] = (uint16_t)(20000u + (s >> 20)); }
        for (i = 0; i < npx; ++i) { dark[i] = 100.0f + (float)(i % 7); flat[i] = 0.5f + (float)(i % 11) * 0.04f; }
        K[0] = K[4] = (double)W; K[2] = W / 2.0; K[5] = H / 2.0;
        memcpy(P, K, sizeof P);
        P[0] = P[4] = 0.9 * W;                                  /* any invertible new camera matrix */
        if (imgcorr_set_dark(ctx, dark, NULL, 0.0, 16, 0) != IMGCORR_OK) return fail("set_dark");
        if (imgcorr_set_flat(ctx, flat, 0) != IMGCORR_OK) return fail("set_flat");
        if (imgcorr_set_lens(ctx, K, dist, P) != IMGCORR_OK) return fail("set_lens");
        if (imgcorr_correct_host(ctx, raw, IMGCORR_U16, out, IMGCORR_F32, N, 0.1, 3,
                                 IMGCORR_DO_DARK | IMGCORR_DO_FLAT | IMGCORR_DO_NAN_TO_NUM, 1, 0.0, 0, 0, W, H) != IMGCORR_OK)
            return fail("correct_host");
        for (i = 0; i < npx * N; ++i) sum += out[i];
        printf("checksum %.6f launches %lld\n", sum, imgcorr_launch_count(ctx));
        free(dark);
        free(flat);
        imgcorr_host_free(raw);
        imgcorr_host_free(out);
        if (imgcorr_ctx_destroy(ctx) != IMGCORR_OK) return fail("ctx_destroy");
    }
    return 0;
}
