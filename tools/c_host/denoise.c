/* A host WITHOUT Python or PyTorch running the hot path through the C ABI only (include/sdb200.h, plan-level entry).
 *
 *   gcc -O2 -I include tools/c_host/denoise.c -o denoise -L stable-diffusion-pytorch_b200 -lsdb200 -Wl,-rpath,$PWD/stable-diffusion-pytorch_b200
 *
 *   denoise <engine> loop    <latent.f32> <context.f32> <steps>    <out.f32>   the sampling loop of models/diffusion.py:223-236
 *   denoise <engine> forward <x.f32>      <context.f32> <timestep> <out.f32>   one UNet.forward (models/unet/unet.py:431-443)
 *
 * <engine> is written by DenoiseLoop.export_engine() / StepProgram.export_engine() (sdk_plan_save): launch lists, tensor-core
 * descriptors with their tuned tilings, packed weights, timestep / coefficient tables.  Inputs and outputs are raw little-endian
 * fp32 files in the reference's layouts (latent NCHW, context [B][77][D]). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sdb200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != SDK_OK) {                                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sdk_last_error());              \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

static void* read_file(const char* path, int64_t* bytes) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return NULL; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* buf = malloc(n > 0 ? (size_t)n : 1);
    if (buf && n > 0 && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
    fclose(f);
    *bytes = n;
    return buf;
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int main(int argc, char** argv) {
    if (argc != 7 || (strcmp(argv[2], "loop") != 0 && strcmp(argv[2], "forward") != 0)) {
        fprintf(stderr, "usage: %s <engine> loop|forward <latent.f32> <context.f32> <steps|timestep> <out.f32>\n", argv[0]);
        return 2;
    }
    const int loop = strcmp(argv[2], "loop") == 0;
    const long arg = atol(argv[5]);
    void* plan = NULL;
    void* stream = NULL;
    CHECK(sdk_plan_load(argv[1], &plan));
    CHECK(sdk_stream_create(&stream));

    int64_t nx = 0, nc = 0, cap = 0;
    void* x = read_file(argv[3], &nx);
    void* ctx = read_file(argv[4], &nc);
    if (!x || !ctx) return 1;
    CHECK(sdk_plan_upload(plan, "x", x, nx, stream));
    CHECK(sdk_plan_upload(plan, "context", ctx, nc, stream));
    CHECK(sdk_plan_launch(plan, 1, stream));                   /* context program: cross-attention K/V, once per prompt */

    const char* result = loop ? "x" : "out";
    void* dptr = NULL;
    CHECK(sdk_plan_region(plan, result, &dptr, &cap));
    void* out = malloc((size_t)cap);
    if (!out) return 1;
    double t0 = 0.0, t1 = 0.0;
    if (loop) {
        const int32_t zero = 0;
        CHECK(sdk_plan_upload(plan, "counter", &zero, 4, stream));
        CHECK(sdk_plan_launch(plan, 4, stream));               /* first step eagerly (sets kernel attributes), then a CUDA graph */
        CHECK(sdk_stream_sync(stream));
        if (arg > 1) CHECK(sdk_plan_capture(plan, 4, stream));
        t0 = now_ms();
        for (long i = 1; i < arg; ++i) CHECK(sdk_plan_launch(plan, 4, stream));
        CHECK(sdk_stream_sync(stream));
        t1 = now_ms();
        printf("loop: %ld steps, %d launches per step, %.3f ms per step after the first (host clock, graph replay)\n", arg,
               sdk_plan_num_launches(plan, 4), arg > 1 ? (t1 - t0) / (double)(arg - 1) : 0.0);
    } else {
        const int64_t t = (int64_t)arg;
        CHECK(sdk_plan_upload(plan, "timestep", &t, 8, stream));
        CHECK(sdk_plan_launch(plan, 0, stream));
        printf("forward: timestep %ld, %d launches\n", arg, sdk_plan_num_launches(plan, 0));
    }
    CHECK(sdk_plan_download(plan, result, out, cap, stream));
    CHECK(sdk_stream_sync(stream));
    FILE* f = fopen(argv[6], "wb");
    if (!f || fwrite(out, 1, (size_t)cap, f) != (size_t)cap) { fprintf(stderr, "cannot write %s\n", argv[6]); return 1; }
    fclose(f);
    CHECK(sdk_stream_destroy(stream));
    CHECK(sdk_plan_destroy(plan));
    free(x); free(ctx); free(out);
    return 0;
}
