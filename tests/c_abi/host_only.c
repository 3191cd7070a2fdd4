/* A plain C99 client of include/orion_b200.h: what a C (or cgo / FFI) host sees.  Only [host-only] entry points are
 * exercised unless a CUDA device is present; without one, constructors must fail with ORION_B200_ERR_NO_DEVICE
 * (there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "orion_b200.h"

int main(void) {
    if (orion_b200_abi_version() != ORION_B200_ABI_VERSION) { printf("abi version mismatch\n"); return 1; }
    size_t n = orion_b200_fir_lowpass_design(2.4e6f, 100e3f, 38400.0f, NULL, 0);
    if (n != 63) { printf("expected 63 taps, got %zu\n", n); return 2; }
    float *taps = (float *)malloc(n * sizeof(float));
    orion_b200_fir_lowpass_design(2.4e6f, 100e3f, 38400.0f, taps, n);
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += taps[i];
    if (fabsf(sum - 1.0f) > 1e-5f || taps[0] != 0.0f || taps[n - 1] != 0.0f) { printf("bad design: sum %g\n", sum); return 3; }
    float c[5];
    orion_b200_lp_biquad_design(300e3f, 13500.0f, c);
    if (!(c[0] > 0.0f && c[3] < 0.0f)) { printf("bad biquad\n"); return 4; }

    orion_b200_block *b = NULL;
    int st = orion_b200_fm_demod_create(48e3f, 2.5e3f, 5e3f, &b);
    if (orion_b200_device_count() <= 0) {
        if (st != ORION_B200_ERR_NO_DEVICE || b != NULL) { printf("expected NO_DEVICE, got %d\n", st); return 5; }
        printf("host-only ok (no device: %s)\n", orion_b200_status_string(st));
        free(taps);
        return 0;
    }
    if (st != ORION_B200_OK) { printf("create failed: %s\n", orion_b200_status_string(st)); return 6; }
    enum { N = 4096 };
    orion_b200_c32 *iq = (orion_b200_c32 *)calloc(N, sizeof(orion_b200_c32));
    float *audio = (float *)calloc(N, sizeof(float));
    for (int i = 0; i < N; ++i) { iq[i].re = cosf(0.05f * i); iq[i].im = sinf(0.05f * i); }
    size_t r = 0, w = 0;
    st = orion_b200_block_process(b, iq, N, audio, N, &r, &w);
    if (st != ORION_B200_OK || r != N || w != N) { printf("process failed: %d %zu %zu\n", st, r, w); return 7; }
    orion_b200_work_report wr = orion_b200_block_plan(b, 1000, 10);
    if (wr.in_read != 10 || wr.out_written != 10) { printf("bad plan\n"); return 8; }
    orion_b200_block_destroy(b);
    printf("device ok: %zu items through FmQuadratureDemod, audio[%d] = %g\n", w, N - 1, audio[N - 1]);
    free(iq); free(audio); free(taps);
    return 0;
}
