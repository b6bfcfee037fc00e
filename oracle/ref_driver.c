/*
 * oracle/ref_driver.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Executes the reference's own kernel source: REF_KERNEL_PATH (= /root/reference/gaussian_kernel.cl, passed by
 * oracle/Makefile) is #included below, unmodified, behind oracle/cl_shim.h.  The functions here play the role
 * of clEnqueueNDRangeKernel for it: global size roundup16(W) x roundup16(H) as in heterogeneous_blur.c:397-400,
 * one call of gaussian_blur() per work-item, ids delivered through get_global_id().
 * Built only where /root/reference exists; output oracle/_ref/libgaussian_ref.so (git-ignored, travels to the
 * GPU box with the snapshot).  Used to pin oracle/gaussian_oracle.c, to generate tests/golden/, and as
 * bench.py's cpu_baseline of kind "reference".
 */
#include <stddef.h>
#include "cl_shim.h"

#ifndef REF_KERNEL_PATH
#error "REF_KERNEL_PATH must point at the reference's gaussian_kernel.cl"
#endif
#include REF_KERNEL_PATH

#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_API __attribute__((visibility("default")))

static inline int roundup16(int v) { return ((v + 15) / 16) * 16; }

/* One NDRange launch on one image: what clEnqueueNDRangeKernel(…, 2, global, local 16x16) does. */
REF_API void ref_gaussian_blur_ndrange(const unsigned char *input, unsigned char *output,
                                       int width, int height, int channels)
{
    int gx = roundup16(width), gy = roundup16(height);
    for (int y = 0; y < gy; y++) {
        for (int x = 0; x < gx; x++) {
            cl_shim_gid[0] = x;
            cl_shim_gid[1] = y;
            gaussian_blur(input, output, width, height, channels);
        }
    }
}

/* A stream of images, one NDRange launch each (heterogeneous_blur.c:482-535), spread over the host cores
 * (images x 16-row work-group rows in parallel, like a CPU OpenCL runtime spreads work-groups). */
REF_API void ref_gaussian_blur_batch(const unsigned char *input, unsigned char *output,
                                     int width, int height, int channels, long n_images,
                                     size_t in_stride, size_t out_stride)
{
    int gx = roundup16(width), gy = roundup16(height);
    long groups_y = gy / 16;
    long total = n_images * groups_y;
#pragma omp parallel for schedule(static)
    for (long t = 0; t < total; t++) {
        long i = t / groups_y;
        int y0 = (int)(t % groups_y) * 16;
        const unsigned char *src = input + (size_t)i * in_stride;
        unsigned char *dst = output + (size_t)i * out_stride;
        for (int y = y0; y < y0 + 16; y++) {
            for (int x = 0; x < gx; x++) {
                cl_shim_gid[0] = x;
                cl_shim_gid[1] = y;
                gaussian_blur(src, dst, width, height, channels);
            }
        }
    }
}

REF_API void ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

REF_API int ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
