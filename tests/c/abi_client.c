/* tests/c/abi_client.c -- a plain C99 host, the reference's own language (heterogeneous_blur.c is C with OpenCL calls),
 * built against include/b200blur.h and linked with libb200blur.so by tests/test_cabi_cpu.py.  It makes the host-side
 * calls a ported heterogeneous_blur.c / split_image_blur.c makes before any device work: version, device discovery with
 * the reference's "no device" exit (:181-184), the per-batch split (:446-458), the split row (split_image_blur.c:144-154)
 * and the launch geometry of both Approach 2 parts (:401, :414).  Prints one line per check; exit code 0 = all as expected.
 * With a GPU present it also blurs one 320x240 image through write -> blur -> read and prints a checksum. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200blur.h"

#define CHECK(cond, what)                                              \
    do {                                                               \
        if (!(cond)) {                                                 \
            printf("FAIL %s (%s)\n", what, b200blur_last_error());     \
            return 1;                                                  \
        }                                                              \
        printf("ok %s\n", what);                                       \
    } while (0)

int main(void)
{
    int n_dev = -1, rc, n_first = 0, n_second = 0, split_row = 0;
    int64_t begin = 0, count = 0;
    b200blur_launch top, bottom;
    static unsigned char fake_in[16], fake_out[16]; /* only their (aligned) addresses are used */
    const int W = 320, H = 240, C = 3;

    CHECK(strncmp(b200blur_version(), "b200blur ", 9) == 0, "version");
    rc = b200blur_device_count(&n_dev);
    CHECK((rc == B200BLUR_OK && n_dev > 0) || rc == B200BLUR_ERR_NO_DEVICE, "device discovery");
    printf("devices %d\n", rc == B200BLUR_OK ? n_dev : 0);

    /* heterogeneous_blur.c:446-458 with the logs' known answer: batch of 35 at ratio 0.728 -> 25 + 10 */
    CHECK(b200blur_ratio_split_images(35, 0.728f, 0, &n_first, &n_second) == B200BLUR_OK && n_first + n_second == 35,
          "ratio split of a batch");
    printf("split %d %d\n", n_first, n_second);
    CHECK(b200blur_partition(5000, 8, 7, &begin, &count) == B200BLUR_OK && begin + count == 5000 && count == 625,
          "even partition over 8 GPUs");
    /* split_image_blur.c:144: (int)(240 * (1 - 0.837)) = 39 */
    CHECK(b200blur_ratio_split_row(H, 0.837f, &split_row) == B200BLUR_OK && split_row == 39, "split row 39");

    /* the two Approach 2 launches: rows incl. halo as kernel height, keep the part's own rows */
    CHECK(b200blur_launch_rows(&top, fake_in, fake_out, W, split_row + 1, C, 0, split_row, 35,
                               (size_t)(split_row + 1) * W * C, (size_t)split_row * W * C) == B200BLUR_OK &&
              top.rows == split_row && top.halo_top == NULL && top.halo_bottom != NULL,
          "A2 top part geometry");
    CHECK(b200blur_launch_rows(&bottom, fake_in, fake_out, W, H - split_row + 1, C, 1, H - split_row, 35,
                               (size_t)(H - split_row + 1) * W * C, (size_t)(H - split_row) * W * C) == B200BLUR_OK &&
              bottom.rows == H - split_row && bottom.halo_top != NULL && bottom.halo_bottom == NULL,
          "A2 bottom part geometry");
    CHECK(b200blur_launch_rows(&top, fake_in, fake_out, W, H, C, 1, H, 1, 0, 0) == B200BLUR_ERR_INVALID, "a row range outside the input is rejected");

    if (rc == B200BLUR_OK && n_dev > 0) {
        b200blur_ctx *ctx = NULL;
        void *d_in = NULL, *d_out = NULL, *h_in = NULL, *h_out = NULL;
        const size_t bytes = (size_t)W * H * C;
        size_t i;
        unsigned long long sum = 1469598103934665603ull;
        b200blur_launch l;
        b200blur_event ev;
        double ms = 0;
        CHECK(b200blur_ctx_create(0, 1, &ctx) == B200BLUR_OK, "context");
        CHECK(b200blur_dev_alloc(ctx, bytes, &d_in) == B200BLUR_OK && b200blur_dev_alloc(ctx, bytes, &d_out) == B200BLUR_OK, "device buffers");
        CHECK(b200blur_host_alloc(bytes, &h_in) == B200BLUR_OK && b200blur_host_alloc(bytes, &h_out) == B200BLUR_OK, "pinned buffers");
        for (i = 0; i < bytes; i++) ((unsigned char *)h_in)[i] = (unsigned char)((i * 2654435761u) >> 13);
        CHECK(b200blur_launch_rows(&l, d_in, d_out, W, H, C, 0, H, 1, bytes, bytes) == B200BLUR_OK, "launch geometry");
        CHECK(b200blur_enqueue_write(ctx, 0, d_in, h_in, bytes, NULL) == B200BLUR_OK, "write");
        CHECK(b200blur_enqueue_blur(ctx, 0, &l, &ev) == B200BLUR_OK, "blur");
        CHECK(b200blur_enqueue_read(ctx, 0, h_out, d_out, bytes, NULL) == B200BLUR_OK, "read");
        CHECK(b200blur_finish(ctx, 0) == B200BLUR_OK, "finish");
        CHECK(b200blur_event_ms(ctx, ev, &ms) == B200BLUR_OK && ms > 0, "kernel time");
        for (i = 0; i < bytes; i++) sum = (sum ^ ((unsigned char *)h_out)[i]) * 1099511628211ull;
        printf("input_rule (i*2654435761)>>13\nchecksum %016llx\n", sum);
        b200blur_event_release(ctx, ev);
        {
            /* the whole batch loop in one call (heterogeneous_blur.c:418-600): 8 replicas of the image, batch_size 3,
             * through the multi-GPU entry point with a one-context list; every replica must equal the output above */
            enum { N = 8 };
            void *s_in = NULL, *s_out = NULL;
            b200blur_ctx *list[1];
            b200blur_ctx *ctx3 = NULL;
            b200blur_stats st;
            int k, same = 1;
            CHECK(b200blur_ctx_create(0, 3, &ctx3) == B200BLUR_OK, "context with three queues");
            list[0] = ctx3;
            CHECK(b200blur_host_alloc(N * bytes, &s_in) == B200BLUR_OK && b200blur_host_alloc(N * bytes, &s_out) == B200BLUR_OK,
                  "pinned stream buffers");
            for (k = 0; k < N; k++) memcpy((unsigned char *)s_in + k * bytes, h_in, bytes);
            memset(s_out, 0, N * bytes);
            CHECK(b200blur_run_host_multi(list, 1, s_in, s_out, W, H, C, N, 3, &st) == B200BLUR_OK && st.images == N,
                  "stream engine");
            for (k = 0; k < N; k++) same = same && memcmp((unsigned char *)s_out + k * bytes, h_out, bytes) == 0;
            CHECK(same, "every replica equals the single-image result");
            b200blur_host_free(s_in);
            b200blur_host_free(s_out);
            CHECK(b200blur_ctx_destroy(ctx3) == B200BLUR_OK, "second context destroyed");
        }
        b200blur_dev_free(ctx, d_in);
        b200blur_dev_free(ctx, d_out);
        b200blur_host_free(h_in);
        b200blur_host_free(h_out);
        CHECK(b200blur_ctx_destroy(ctx) == B200BLUR_OK, "context destroyed");
    }
    printf("done\n");
    return 0;
}
