/*
 * oracle/cl_shim.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Just enough OpenCL C vocabulary for gcc to compile the reference's gaussian_kernel.cl UNMODIFIED, from
 * where it lies under /root/reference, as plain C (see oracle/ref_driver.c).  No OpenCL runtime exists in
 * this image (no CL/cl.h, no ICD, no PoCL), so this is how the reference's own kernel source is executed
 * here.  Nothing of the reference is copied: the .cl file is #included by path at build time and the
 * result goes to oracle/_ref/ (git-ignored).
 */
#ifndef ORACLE_CL_SHIM_H
#define ORACLE_CL_SHIM_H

#define __kernel
#define __global
#define __constant const
#define __local
#define __private

/* Work-item ids of the work-item being executed by the calling thread (set by the NDRange loop). */
static __thread int cl_shim_gid[3];
static inline int get_global_id(unsigned dim) { return cl_shim_gid[dim]; }

/* OpenCL C integer built-ins used by the kernel (gaussian_kernel.cl:56-57), on int. */
static inline int cl_shim_min(int a, int b) { return a < b ? a : b; }
static inline int cl_shim_max(int a, int b) { return a > b ? a : b; }
#define min(a, b) cl_shim_min((a), (b))
#define max(a, b) cl_shim_max((a), (b))

#endif
