/*
 * oracle/gaussian_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's one hot path (3x3 Gaussian blur of interleaved uint8 images and the
 * two work-distribution schemes around it).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path (the C-ABI library in
 * include/b200blur.h) never links, loads or calls it, and fails loudly without its CUDA code.
 *
 * Parity status: PINNED.  There are no golden vectors in the reference (it has no tests), so the pin is
 * the reference kernel itself: oracle/_ref/libgaussian_ref.so is gaussian_kernel.cl compiled unmodified
 * (oracle/ref_driver.c + oracle/cl_shim.h) and tests/test_oracle.py checks this restatement against it
 * bit-for-bit on random and edge-case inputs, plus against the committed fixtures in tests/golden/ that
 * were generated from it (tests/golden/make_golden.py) and the distribution known-answers taken from the
 * reference's run logs (data/approach1/35_run_1.txt:50,:57; data/approach2/35_run_1.txt:16-18).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* gaussian_kernel.cl:36-41 -- the weight table, as written there (fp32 dyadic fractions). */
static const float kWeightsF[3][3] = {
    {0.0625f, 0.125f, 0.0625f},
    {0.125f,  0.25f,  0.125f},
    {0.0625f, 0.125f, 0.0625f},
};
/* Same table times 16. */
static const int kWeightsI[3][3] = {{1, 2, 1}, {2, 4, 2}, {1, 2, 1}};

/*
 * One work-item of gaussian_blur, gaussian_kernel.cl:29-71.
 * Loop order (c outer, ky, kx inner), clamp via max(0,min(n,dim-1)) (:56-57), index formula (:60),
 * fp32 accumulate of uchar*weight (:63) and the truncating (unsigned char) cast (:70) are kept as they are.
 */
static inline void work_item_float(const unsigned char *input, unsigned char *output,
                                   int width, int height, int channels, int x, int y)
{
    if (x >= width || y >= height) return;                      /* :33 */
    for (int c = 0; c < channels; c++) {                        /* :44 */
        float sum = 0.0f;                                       /* :45 */
        for (int ky = -1; ky <= 1; ky++) {                      /* :48 */
            for (int kx = -1; kx <= 1; kx++) {                  /* :49 */
                int nx = x + kx;                                /* :52 */
                int ny = y + ky;                                /* :53 */
                nx = clampi(nx, 0, width - 1);                  /* :56 */
                ny = clampi(ny, 0, height - 1);                 /* :57 */
                int index = (ny * width + nx) * channels + c;   /* :60 */
                sum += input[index] * kWeightsF[ky + 1][kx + 1];/* :63 */
            }
        }
        int out_index = (y * width + x) * channels + c;         /* :69 */
        output[out_index] = (unsigned char)sum;                 /* :70 */
    }
}

/* Whole NDRange of gaussian_blur on one image (heterogeneous_blur.c:397-400, :507: global = roundup16, the
 * out-of-range work-items return at gaussian_kernel.cl:33, so iterating 0..W-1 x 0..H-1 is equivalent). */
ORACLE_API void oracle_blur_image_f32(const unsigned char *in, unsigned char *out, int width, int height, int channels)
{
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
            work_item_float(in, out, width, height, channels, x, y);
}

/* Integer form: out = (sum_k w_int * p) >> 4.  Equal to the fp32 form because every product and partial sum
 * is a multiple of 2^-4 below 2^8; tests assert the equality. */
ORACLE_API void oracle_blur_image_int(const unsigned char *in, unsigned char *out, int width, int height, int channels)
{
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            for (int c = 0; c < channels; c++) {
                int sum = 0;
                for (int ky = -1; ky <= 1; ky++) {
                    int ny = clampi(y + ky, 0, height - 1);
                    for (int kx = -1; kx <= 1; kx++) {
                        int nx = clampi(x + kx, 0, width - 1);
                        sum += in[((size_t)ny * width + nx) * channels + c] * kWeightsI[ky + 1][kx + 1];
                    }
                }
                out[((size_t)y * width + x) * channels + c] = (unsigned char)(sum >> 4);
            }
        }
    }
}

/*
 * A batch of images, each an independent launch of the kernel (heterogeneous_blur.c:482-535: one
 * write/kernel/read triplet per image).  Image i lives at in + i*in_stride / out + i*out_stride.
 * OpenMP over images x rows: this is the CPU baseline bench.py times ("port" kind).
 * use_int != 0 selects the integer form.
 */
ORACLE_API void oracle_blur_batch(const unsigned char *in, unsigned char *out, int width, int height, int channels,
                                  long n_images, size_t in_stride, size_t out_stride, int use_int)
{
    long total_rows = n_images * (long)height;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < total_rows; r++) {
        long i = r / height;
        int y = (int)(r % height);
        const unsigned char *src = in + (size_t)i * in_stride;
        unsigned char *dst = out + (size_t)i * out_stride;
        if (!use_int) {
            for (int x = 0; x < width; x++) work_item_float(src, dst, width, height, channels, x, y);
        } else {
            for (int x = 0; x < width; x++) {
                for (int c = 0; c < channels; c++) {
                    int sum = 0;
                    for (int ky = -1; ky <= 1; ky++) {
                        int ny = clampi(y + ky, 0, height - 1);
                        for (int kx = -1; kx <= 1; kx++) {
                            int nx = clampi(x + kx, 0, width - 1);
                            sum += src[((size_t)ny * width + nx) * channels + c] * kWeightsI[ky + 1][kx + 1];
                        }
                    }
                    dst[((size_t)y * width + x) * channels + c] = (unsigned char)(sum >> 4);
                }
            }
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline must still use all host cores. */
ORACLE_API void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORACLE_API int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------------
 * Approach 1 distribution, heterogeneous_blur.c:446-458.
 * mode 0 = both, 1 = cpu only, 2 = gpu only (:51).  float multiply then (int) truncation (:450).
 * ---------------------------------------------------------------------------------------------------- */
ORACLE_API void oracle_a1_batch_split(int batch_count, float gpu_ratio, int mode, int *n_cpu, int *n_gpu)
{
    int num_images_cpu = 0, num_images_gpu = 0;
    if (mode == 0) {
        num_images_gpu = (int)(batch_count * gpu_ratio);        /* :450 */
        num_images_cpu = batch_count - num_images_gpu;          /* :451 */
    } else if (mode == 1) {
        num_images_cpu = batch_count;                           /* :453 */
    } else if (mode == 2) {
        num_images_gpu = batch_count;                           /* :457 */
    }
    *n_cpu = num_images_cpu;
    *n_gpu = num_images_gpu;
}

/* Whole-run totals of Approach 1: NUM_BATCHES = ceil(N/B) (:86), last batch short (:423-427),
 * totals accumulated per batch (:460-461). */
ORACLE_API void oracle_a1_totals(int num_images, int batch_size, float gpu_ratio, int mode,
                                 int *num_batches, int *total_cpu, int *total_gpu)
{
    int nb = (num_images + batch_size - 1) / batch_size;        /* :86 */
    int tc = 0, tg = 0;
    for (int batch = 0; batch < nb; batch++) {
        int batch_start = batch * batch_size;                   /* :423 */
        int batch_count = batch_size;
        if (batch_start + batch_count > num_images) batch_count = num_images - batch_start; /* :425-427 */
        int c, g;
        oracle_a1_batch_split(batch_count, gpu_ratio, mode, &c, &g);
        tc += c; tg += g;
    }
    *num_batches = nb; *total_cpu = tc; *total_gpu = tg;
}

/* ------------------------------------------------------------------------------------------------------
 * Approach 2 geometry, split_image_blur.c:144-166 (HALO = 1, :70).
 * ---------------------------------------------------------------------------------------------------- */
ORACLE_API void oracle_a2_geometry(int height, float gpu_ratio, int *split_row_out,
                                   int *cpu_input_rows, int *cpu_output_rows,
                                   int *gpu_input_rows, int *gpu_output_rows)
{
    const int HALO = 1;
    int split_row = (int)(height * (1.0f - gpu_ratio));         /* :144 */
    if (split_row < HALO) split_row = HALO;                     /* :147-150 */
    if (split_row > height - HALO) split_row = height - HALO;   /* :151-154 */
    *split_row_out = split_row;
    *cpu_input_rows = split_row + HALO;                         /* :157 */
    *cpu_output_rows = split_row;                               /* :158 */
    *gpu_input_rows = (height - split_row) + HALO;              /* :163 */
    *gpu_output_rows = height - split_row;                      /* :164 */
}

/*
 * Approach 2 composition for one image, split_image_blur.c:503-541: two launches of the unmodified kernel,
 * each on its rows plus one halo row with `height = rows incl. halo` (:401, :414); the top part keeps its
 * first cpu_output_size bytes (:526), the bottom part is read back from byte offset HALO*W*C (:537-539).
 * Device buffers are modelled by two scratch allocations sized like :371-384.
 */
ORACLE_API int oracle_split_image(const unsigned char *in, unsigned char *out, int width, int height, int channels,
                                  int split_row)
{
    const int HALO = 1;
    size_t row_bytes = (size_t)width * channels;
    int cpu_input_rows = split_row + HALO, cpu_output_rows = split_row;
    int gpu_input_rows = (height - split_row) + HALO, gpu_output_rows = height - split_row;
    unsigned char *dev_cpu_out = (unsigned char *)malloc(row_bytes * cpu_input_rows);
    unsigned char *dev_gpu_out = (unsigned char *)malloc(row_bytes * gpu_input_rows);
    if (!dev_cpu_out || !dev_gpu_out) { free(dev_cpu_out); free(dev_gpu_out); return -1; }
    const unsigned char *cpu_in = in;                                         /* :511 */
    const unsigned char *gpu_in = in + (size_t)(split_row - HALO) * row_bytes;/* :516 */
    oracle_blur_image_f32(cpu_in, dev_cpu_out, width, cpu_input_rows, channels);
    oracle_blur_image_f32(gpu_in, dev_gpu_out, width, gpu_input_rows, channels);
    memcpy(out, dev_cpu_out, row_bytes * cpu_output_rows);                    /* :526 */
    memcpy(out + (size_t)split_row * row_bytes, dev_gpu_out + HALO * row_bytes,
           row_bytes * gpu_output_rows);                                      /* :517, :537-539 */
    free(dev_cpu_out); free(dev_gpu_out);
    return 0;
}

/*
 * The same composition generalised to G row bands: band k owns rows
 * [k*H/G, (k+1)*H/G); it runs the unmodified kernel on its rows plus one halo row on each interior side and
 * drops the halo rows' outputs.  G == 2 with H/2 == split_row reproduces oracle_split_image.
 * Returns -1 if a band would be empty.
 */
ORACLE_API int oracle_band_split(const unsigned char *in, unsigned char *out, int width, int height, int channels,
                                 int n_bands)
{
    size_t row_bytes = (size_t)width * channels;
    for (int k = 0; k < n_bands; k++) {
        int r0 = (int)((long)k * height / n_bands), r1 = (int)((long)(k + 1) * height / n_bands);
        if (r1 <= r0) return -1;
        int top = (k > 0) ? 1 : 0, bot = (k < n_bands - 1) ? 1 : 0;
        int in_rows = (r1 - r0) + top + bot;
        unsigned char *dev_out = (unsigned char *)malloc(row_bytes * in_rows);
        if (!dev_out) return -1;
        oracle_blur_image_f32(in + (size_t)(r0 - top) * row_bytes, dev_out, width, in_rows, channels);
        memcpy(out + (size_t)r0 * row_bytes, dev_out + (size_t)top * row_bytes, row_bytes * (r1 - r0));
        free(dev_out);
    }
    return 0;
}
