/*
 * b200blur.h -- C ABI of the B200-native 3x3 Gaussian-blur stream engine.
 *
 * This header is the drop-in boundary for the one hot path of CC834/Heterogeneous-OpenCL-Image-Processing-Engine:
 * it replaces, call for call, the OpenCL host API that heterogeneous_blur.c and split_image_blur.c use inline
 * from main().  Plain C: opaque handles, raw pointers and sizes, int status returns.
 * No PyTorch, no OpenCL, no CPU fallback -- every compute entry point runs hand-written sm_100a CUDA and fails
 * with B200BLUR_ERR_NO_DEVICE when no CUDA device is present.
 *
 * Reference citations are `file:line` relative to the reference repository root
 * (A1 = heterogeneous_blur.c, A2 = split_image_blur.c, K = gaussian_kernel.cl).
 *
 * Threading: one host thread per GPU may call into distinct contexts concurrently (the reference is single
 * threaded with two asynchronous queues, A1:201-212).  A context must not be used by two threads at once, with ONE
 * exception: b200blur_enqueue_wait_peer(ctx, queue, src, ev) may name an event of a context `src` that another
 * thread is using at that moment.  That is safe because a context's event pool is fixed storage (slots are appended,
 * never moved) and `ev` must be an event the owner has already enqueued and not yet released -- the caller orders
 * that with its own host-side rendezvous (split_image_blur.cpp does, with two barriers per batch).
 * b200blur_last_error() is thread-local.
 */
#ifndef B200BLUR_H
#define B200BLUR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B200BLUR_API
#else
#define B200BLUR_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------------ status */
enum {
    B200BLUR_OK = 0,
    B200BLUR_ERR_INVALID = -1,   /* bad argument (NULL, negative size, misuse)                                */
    B200BLUR_ERR_NO_DEVICE = -2, /* no CUDA device / device index out of range (A1:181-184 "Could not find")  */
    B200BLUR_ERR_CUDA = -3,      /* a CUDA runtime call failed; b200blur_last_error() carries its message      */
    B200BLUR_ERR_NOMEM = -4,     /* allocation failed (A1:434-437)                                            */
    B200BLUR_ERR_PEER = -5       /* peer access between the two devices is not available                      */
};

/* Message of the last failure on the calling thread ("" if none).  Replaces cl_error()'s code+string (A1:25-30). */
B200BLUR_API const char *b200blur_last_error(void);
/* "b200blur <semver> sm_100a" */
B200BLUR_API const char *b200blur_version(void);

/* ---------------------------------------------------------------------------------------- device discovery
 * Replaces clGetPlatformIDs / clGetDeviceIDs / clGetDeviceInfo (A1:144-191, A2:179-226). */
B200BLUR_API int b200blur_device_count(int *count);
B200BLUR_API int b200blur_device_name(int device, char *buf, size_t buf_len);
/* sm count, compute capability major*10+minor, total global memory bytes; any pointer may be NULL. */
B200BLUR_API int b200blur_device_props(int device, int *sm_count, int *cc, size_t *global_mem_bytes);

/* ------------------------------------------------------------------------------------------------ context
 * One per GPU.  Replaces clCreateContext + clCreateCommandQueueWithProperties(PROFILING_ENABLE) +
 * clCreateProgramWithSource/clBuildProgram/clCreateKernel (A1:194-322, A2:229-353): the kernel is compiled ahead
 * of time into this library, so there is no .cl file to find at run time (A1:222-226).
 * `n_queues` in-order queues (CUDA streams) are created; queue 0..n_queues-1 are addressed by index.
 * Pass n_queues <= 0 for the default (4). */
typedef struct b200blur_ctx b200blur_ctx;
B200BLUR_API int b200blur_ctx_create(int device, int n_queues, b200blur_ctx **ctx);
B200BLUR_API int b200blur_ctx_destroy(b200blur_ctx *ctx); /* clRelease* (A1:727-744) */
B200BLUR_API int b200blur_ctx_device(const b200blur_ctx *ctx);
B200BLUR_API int b200blur_ctx_num_queues(const b200blur_ctx *ctx);
/* The CUstream/cudaStream_t behind a queue, as an opaque pointer (for callers that bring their own events). */
B200BLUR_API void *b200blur_ctx_queue_handle(const b200blur_ctx *ctx, int queue);

/* ------------------------------------------------------------------------------------------------- memory
 * dev_alloc/dev_free replace clCreateBuffer/clReleaseMemObject (A1:341-353, :727-730).
 * host_alloc/host_free replace the per-batch malloc/free of the staging buffers (A1:431-432, :596-597) with
 * page-locked memory so H2D/D2H run asynchronously at link speed; host_register pins a buffer the caller
 * already owns. */
B200BLUR_API int b200blur_dev_alloc(b200blur_ctx *ctx, size_t bytes, void **dptr);
B200BLUR_API int b200blur_dev_free(b200blur_ctx *ctx, void *dptr);
B200BLUR_API int b200blur_host_alloc(size_t bytes, void **hptr);
B200BLUR_API int b200blur_host_free(void *hptr);
B200BLUR_API int b200blur_host_register(void *hptr, size_t bytes);
B200BLUR_API int b200blur_host_unregister(void *hptr);

/* ------------------------------------------------------------------------------------------------- events
 * Every enqueue can return an event handle that records the command's start and end on the device, like the
 * cl_event of a PROFILING_ENABLE queue (A1:502-533).  event_ms = (END - START) in ms, the quantity summed at
 * A1:544-579.  Handles are small integers owned by the context; release returns them to its pool (at most 32768
 * live events per context, B200BLUR_ERR_NOMEM beyond; an enqueue that fails gives the event it took back). */
typedef int32_t b200blur_event;
#define B200BLUR_NO_EVENT ((b200blur_event *)0)
B200BLUR_API int b200blur_event_ms(b200blur_ctx *ctx, b200blur_event ev, double *ms); /* clGetEventProfilingInfo */
B200BLUR_API int b200blur_event_release(b200blur_ctx *ctx, b200blur_event ev);        /* clReleaseEvent          */
/* clEnqueueMarker: an event that completes when everything enqueued on `queue` so far has completed. */
B200BLUR_API int b200blur_enqueue_marker(b200blur_ctx *ctx, int queue, b200blur_event *ev);
/* Device time in ms from the END of `from` to the END of `to` (both must have completed or be in flight). */
B200BLUR_API int b200blur_events_elapsed_ms(b200blur_ctx *ctx, b200blur_event from, b200blur_event to, double *ms);
/* Make `queue` wait for the END of `ev` (cross-queue dependency; OpenCL's event wait list). */
B200BLUR_API int b200blur_enqueue_wait(b200blur_ctx *ctx, int queue, b200blur_event ev);
/* Same, for an event that belongs to ANOTHER context (another GPU): orders a band's kernel after its neighbours'
 * uploads when halo rows are read from peer memory.  The event must already have been enqueued by its owner and must
 * stay unreleased until this call returns; the owner may keep enqueueing other commands meanwhile (see Threading). */
B200BLUR_API int b200blur_enqueue_wait_peer(b200blur_ctx *ctx, int queue, b200blur_ctx *src, b200blur_event ev);

/* ----------------------------------------------------------------------------------------------- transfers
 * Asynchronous, in order on `queue`.  Host memory should be pinned (host_alloc/host_register); the caller keeps
 * it valid until b200blur_finish, exactly as for CL_FALSE writes (A1:502, :538).
 * enqueue_read takes a raw device pointer, so the "non-zero device offset" read of A2:537 is pointer arithmetic. */
B200BLUR_API int b200blur_enqueue_write(b200blur_ctx *ctx, int queue, void *dst_dev, const void *src_host,
                                        size_t bytes, b200blur_event *ev); /* clEnqueueWriteBuffer */
B200BLUR_API int b200blur_enqueue_read(b200blur_ctx *ctx, int queue, void *dst_host, const void *src_dev,
                                       size_t bytes, b200blur_event *ev);  /* clEnqueueReadBuffer  */
/* Strided forms: `rows` runs of `row_bytes`, source/destination advancing by their own pitch (one call moves
 * one row-range of every image of a batch, e.g. a row band or a halo row). */
B200BLUR_API int b200blur_enqueue_write_2d(b200blur_ctx *ctx, int queue, void *dst_dev, size_t dst_pitch,
                                           const void *src_host, size_t src_pitch, size_t row_bytes, size_t rows,
                                           b200blur_event *ev);
B200BLUR_API int b200blur_enqueue_read_2d(b200blur_ctx *ctx, int queue, void *dst_host, size_t dst_pitch,
                                          const void *src_dev, size_t src_pitch, size_t row_bytes, size_t rows,
                                          b200blur_event *ev);
B200BLUR_API int b200blur_finish(b200blur_ctx *ctx, int queue); /* clFinish (A1:538-539) */
B200BLUR_API int b200blur_finish_all(b200blur_ctx *ctx);

/* -------------------------------------------------------------------------------------------- kernel launch
 * One launch blurs a row band of `rows` rows of EVERY image of a batch (the reference launches one image at a
 * time, A1:507).  The arguments of K:19-25 map as: input -> in (+ halo_top/halo_bottom), output -> out,
 * width -> width, channels -> channels, height -> rows + (halo_top != NULL) + (halo_bottom != NULL).
 *
 *   out[i][r][x][c] = ( sum_{ky,kx} w[ky][kx] * src_i(r+ky, clamp(x+kx, 0, width-1), c) ) >> 4,   r in [0, rows)
 *   src_i(-1, .)   = halo_top    ? row i of halo_top    : src_i(0, .)          (clamp, K:57)
 *   src_i(rows, .) = halo_bottom ? row i of halo_bottom : src_i(rows-1, .)
 *   w = {1,2,1; 2,4,2; 1,2,1}  (K:36-41 times 16; identical to the fp32 form)
 *
 * Rows are tight: pitch = width*channels bytes.  Image i of `in` starts at in + i*in_image_stride, and likewise
 * for out / halo_top / halo_bottom with their own strides.  Halo pointers may address another GPU's memory
 * (peer-enabled or IPC-opened): the kernel then loads those rows over NVLink itself -- that is Approach 2's
 * halo exchange fused into the stencil.
 * in == out (in place) is not allowed.  The vectorised path needs channels <= 4 and, on the OUTPUT side, 16-byte aligned
 * pointers/strides and a row pitch that is a multiple of 16 (tight rows with width*channels % 16 == 0, or out_row_pitch
 * set).  The INPUT side (in, halo rows) may then be tight rows of any length and alignment up to 4096 bytes per row --
 * the kernel copies aligned supersets and re-aligns them in shared memory -- or 16-byte pitched/aligned rows of any
 * length.  Anything else runs the generic path (same results).
 */
typedef struct b200blur_launch {
    const void *in;
    void *out;
    int32_t width;
    int32_t channels;
    int32_t rows;
    int32_t reserved; /* must be 0 */
    int64_t n_images;
    size_t in_image_stride;
    size_t out_image_stride;
    const void *halo_top;
    size_t halo_top_stride;
    const void *halo_bottom;
    size_t halo_bottom_stride;
    /* Row pitch in bytes of `in` / `out` (0 = tight, width*channels).  A pitch that is a multiple of 16 lets ANY width
     * run on the vectorised path: bytes between width*channels and the pitch are padding (read, never meaningful;
     * the output's padding is overwritten with don't-care bytes).  The stream engines use this to re-pitch odd-width
     * images with strided copies.  Halo rows are single rows and have no pitch. */
    size_t in_row_pitch;
    size_t out_row_pitch;
} b200blur_launch;

/* Fill a launch that is exactly "the reference kernel with height = in_height on buffer `in`, keeping output rows
 * [first_row, first_row + n_rows)" written tightly at `out` (row first_row lands at out + 0).
 *   whole image (A1:366-389):            in_height = H,               first_row = 0, n_rows = H
 *   A2 top part (A2:401, :520-527):      in_height = split_row + 1,   first_row = 0, n_rows = split_row
 *   A2 bottom part (A2:414, :530-541):   in_height = H - split_row + 1, first_row = 1, n_rows = H - split_row
 * Rows of `in` next to the kept range become the halo rows; at the ends of the buffer the clamp applies. */
B200BLUR_API int b200blur_launch_rows(b200blur_launch *l, const void *in, void *out, int width, int in_height,
                                      int channels, int first_row, int n_rows, int64_t n_images,
                                      size_t in_image_stride, size_t out_image_stride);

/* Same with explicit row pitches (0 = tight): rows of `in` are in_row_pitch bytes apart, rows of `out` out_row_pitch. */
B200BLUR_API int b200blur_launch_rows_pitched(b200blur_launch *l, const void *in, void *out, int width, int in_height,
                                              int channels, int first_row, int n_rows, int64_t n_images,
                                              size_t in_image_stride, size_t out_image_stride, size_t in_row_pitch,
                                              size_t out_row_pitch);

/* clSetKernelArg x5 + clEnqueueNDRangeKernel (A1:366-389, :507): asynchronous, in order on `queue`. */
B200BLUR_API int b200blur_enqueue_blur(b200blur_ctx *ctx, int queue, const b200blur_launch *launch,
                                       b200blur_event *ev);
/* The same for a LIST of launches that are independent of each other (batches): when they share one geometry (width,
 * rows, channels, image and halo strides; tight rows of a 16-byte-multiple length, aligned pointers) they run as ONE
 * kernel launch with per-batch descriptors -- each launch keeps its own in / out / halo pointers and image count, work
 * units never span two of them, and the tail of one overlaps the start of the next instead of costing a launch ramp and
 * drain each (Approach 2 bands of a small image on many GPUs: 47 us per pass as separate launches).
 * Other lists are enqueued launch by launch.  Asynchronous, in order on `queue`; `ev` times the whole list.  The
 * descriptor table belongs to the context: a call on another queue starts after the previous call's kernel. */
B200BLUR_API int b200blur_enqueue_blur_batches(b200blur_ctx *ctx, int queue, const b200blur_launch *launches, int n_launches,
                                               b200blur_event *ev);
/* Which device code a launch would run: 1 = vectorised sm_100a stencil, 0 = generic path. */
B200BLUR_API int b200blur_launch_is_vectorised(const b200blur_launch *launch);
/* Select the kernel variant of the vectorised path (for profiling/tests): 0 = auto, 1 = register/shuffle
 * stencil, 2 = TMA-bulk staged persistent stencil.  Returns the previous value. */
B200BLUR_API int b200blur_set_kernel_variant(b200blur_ctx *ctx, int variant);
/* Host-side plan of the vectorised kernel for rows of `row_bytes` = width*channels bytes (introspection for tests;
 * needs no GPU): out[0] = live 16-byte chunks per row, out[1] = 1 when the row ends inside a chunk (pitched rows),
 * out[2] = 1 when the chunk before the last needs its right-neighbour word patched too, out[3..8] = PRMT selectors
 * for the six window words {wl, w0..w3, wr} of the last chunk, out[9] = selector for the wr word of the chunk before. */
B200BLUR_API int b200blur_plan_row_edge(int row_bytes, int channels, uint32_t out[10]);
/* Host-side work plan of the streamed kernel for `n_images` images of `rows` x `width` x `channels` (row pitch 0 = tight)
 * on a GPU that keeps `resident_ctas` CTAs of it resident (introspection for tests; needs no GPU).  `feed` != 0 plans the
 * FEED form (n_images = images per batch at most).  out[] = {chunks per row, chunks per column block, column blocks,
 * images per group, seg, nseg, seg_fine, nseg_fine, image blocks, coarse image blocks, first fine group, groups, margin,
 * threads per CTA, dynamic shared memory bytes, rows end inside a chunk}.  Groups [0, first fine group) cut image blocks
 * [0, coarse) into nseg segments of seg rows (x column blocks); the remaining groups cut the remaining image blocks
 * into nseg_fine segments of seg_fine rows -- group order: image block, then segment, then column block. */
B200BLUR_API int b200blur_plan_groups(int width, int rows, int channels, int64_t n_images, size_t row_pitch,
                                      int resident_ctas, int feed, int64_t out[16]);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
B200BLUR_API int64_t b200blur_ctx_launch_count(const b200blur_ctx *ctx);

/* ----------------------------------------------------------------------------------- work distribution (L4)
 * Even contiguous partition of `n_items` over `n_parts`; the first n_items % n_parts parts get one more.
 * Replaces num_images_gpu = (int)(batch_count * gpu_ratio) (A1:449-451) for whole images and
 * split_row = (int)(H * (1 - gpu_ratio)) (A2:144) for row bands. */
B200BLUR_API int b200blur_partition(int64_t n_items, int n_parts, int part, int64_t *begin, int64_t *count);
/* The reference's own ratio arithmetic, kept for CLI compatibility and reporting (float multiply, truncation). */
B200BLUR_API int b200blur_ratio_split_images(int batch_count, float gpu_ratio, int mode, int *n_first, int *n_second);
B200BLUR_API int b200blur_ratio_split_row(int height, float gpu_ratio, int *split_row);

/* ------------------------------------------------------------------------------------------- stream engines
 * The batch loop of A1:418-600 as one call.  Times are device times from events (ms), summed per stage like
 * A1:544-579; wall_ms is host wall-clock around the whole call. */
typedef struct b200blur_stats {
    double wall_ms;
    double h2d_ms;    /* "Transfer IN"  */
    double kernel_ms; /* "Kernel execution" */
    double d2h_ms;    /* "Transfer OUT" */
    int64_t images;
    int64_t launches;
    int64_t h2d_bytes;
    int64_t d2h_bytes;
} b200blur_stats;

/* Device-resident: `d_in`/`d_out` hold n_images tight images in HBM.  Images are processed `batch_size` at a time:
 *   coalesce == 1  consecutive batches are fused into as few launches as possible (they are independent);
 *   coalesce == 0  one work descriptor per batch through the feed kernel (see b200blur_feed_*): one kernel launch per
 *                  call, work units never span batches, every batch completes on its own like the reference's
 *                  per-batch sync (A1:538); a repeated identical request re-uses the descriptor table on the device;
 *   coalesce == 2  one kernel LAUNCH per batch, spread over all queues of the context (forked from and joined back
 *                  into queue 0) and replayed as a CUDA graph when repeated -- the launch-bound form, kept to measure.
 * The call orders like one operation on queue 0.  stats == NULL makes it asynchronous (no host synchronisation).
 * Widths with width*channels % 16 != 0 are re-pitched through a scratch pair owned by the context so that they still
 * run on the vectorised kernel. */
B200BLUR_API int b200blur_run_resident(b200blur_ctx *ctx, const void *d_in, void *d_out, int width, int height,
                                       int channels, int64_t n_images, int batch_size, int coalesce,
                                       b200blur_stats *stats);
/* End to end: `h_in`/`h_out` are HOST buffers (pinned for full speed) of n_images tight images.  Chunks of
 * `batch_size` images flow H2D -> blur -> D2H through a ring of device buffers on separate queues so the three
 * stages of different chunks overlap (the reference serialises them per image).  Batches are
 * independent, so small batches are fused (and very large ones cut) into ~64 MB transfer chunks; stats->launches
 * reports the kernels actually launched.  Odd widths are re-pitched by the strided copies on the way in and out. */
B200BLUR_API int b200blur_run_host(b200blur_ctx *ctx, const void *h_in, void *h_out, int width, int height,
                                   int channels, int64_t n_images, int batch_size, b200blur_stats *stats);
/* The same stream over SEVERAL GPUs in one call -- Approach 1's "split every batch between the devices" (A1:446-458)
 * without a ratio: one host thread per context runs the pipeline above, and the pipelines TAKE transfer chunks from one
 * shared counter whenever a ring slot is free, so a GPU behind a slower path to host memory simply moves fewer chunks
 * (dynamic scheduling instead of the hand-tuned gpu_ratio, A1:713-722).  `ctxs` are n_ctx distinct
 * contexts (usually one per GPU), each with >= 3 queues and used by no other thread during the call; h_in / h_out
 * must be visible to every device (b200blur_host_alloc memory is).  stats, if given, has n_ctx entries: stats[k].images
 * is what context k moved, wall_ms its own wall clock.  n_ctx == 1 is b200blur_run_host. */
B200BLUR_API int b200blur_run_host_multi(b200blur_ctx *const *ctxs, int n_ctx, const void *h_in, void *h_out, int width,
                                         int height, int channels, int64_t n_images, int batch_size,
                                         b200blur_stats *stats);

/* ---------------------------------------------------------------------------------------------------- feed
 * The batch loop of A1:418-600 as a RESIDENT kernel: instead of one kernel launch per batch (per image in the
 * reference, A1:507), one persistent kernel per GPU pulls per-batch descriptors from a ring the host appends to and
 * reports every batch's completion individually.  `batch_size` keeps the reference's meaning -- the unit the host
 * stages, submits and waits for (A1:431-442, :482-539) -- without a launch per batch: at 35 images a launch carries
 * ~2.4 us of HBM work, at 1 image 60 ns, far below launch latency.
 *
 *   create(ctx, W, H, C, max_batch, capacity)  geometry is fixed per feed; batches hold 1..max_batch tight images
 *   start                                      launches the resident kernel on the feed's own stream
 *   submit(d_in, d_out, n) -> ticket           stages one batch descriptor (device pointers, 16-byte aligned); blocks only
 *                                              when `capacity` batches are in flight
 *   flush                                      publishes everything staged so far to the kernel (two small async copies)
 *   wait(ticket) / completed(ticket)           clFinish for ONE batch: its output is complete and visible to the host,
 *                                              to copies and to kernels enqueued afterwards
 *   stop                                       publishes, lets the kernel drain and exit; start may be called again
 * Ordering with other work is by the host: submit a batch after its input is in place (e.g. after b200blur_event_ms /
 * b200blur_finish on the upload), read its output after wait().  A kernel that sees no new batch for
 * B200BLUR_FEED_TIMEOUT_MS (default 5000) stops by itself and the feed reports B200BLUR_ERR_CUDA -- a dead host never
 * hangs the GPU.  While a feed runs its CTAs occupy the SMs: other kernels on the device wait for it to stop, and so do
 * calls that wait for an idle device (dev_alloc / dev_free: allocate before start).  Copies on other queues proceed.
 * Needs channels <= 4 and width*channels >= 256 and a multiple of 16 (else B200BLUR_ERR_INVALID: use enqueue_blur). */
typedef struct b200blur_feed b200blur_feed;
B200BLUR_API int b200blur_feed_create(b200blur_ctx *ctx, int width, int height, int channels, int max_batch_images,
                                      int capacity, b200blur_feed **feed);
B200BLUR_API int b200blur_feed_destroy(b200blur_feed *feed);
B200BLUR_API int b200blur_feed_start(b200blur_feed *feed);
B200BLUR_API int b200blur_feed_submit(b200blur_feed *feed, const void *d_in, void *d_out, int n_images, int64_t *ticket);
B200BLUR_API int b200blur_feed_flush(b200blur_feed *feed);
B200BLUR_API int b200blur_feed_wait(b200blur_feed *feed, int64_t ticket);
B200BLUR_API int b200blur_feed_completed(b200blur_feed *feed, int64_t ticket, int *done);
B200BLUR_API int b200blur_feed_stop(b200blur_feed *feed);
B200BLUR_API int64_t b200blur_feed_submitted(const b200blur_feed *feed);

/* -------------------------------------------------------------------------------- multi-GPU (Approach 2 bands)
 * No reference counterpart: the reference's devices only meet in host memory (A2:511-517). */
/* Enable direct loads/stores from ctx `a`'s device into ctx `b`'s memory and vice versa (same process). */
B200BLUR_API int b200blur_peer_enable(b200blur_ctx *a, b200blur_ctx *b);
/* Cross-process form (one process per GPU): export a device allocation as a 64-byte handle, open it elsewhere. */
#define B200BLUR_IPC_HANDLE_BYTES 64
B200BLUR_API int b200blur_ipc_export(b200blur_ctx *ctx, void *dptr, unsigned char handle[B200BLUR_IPC_HANDLE_BYTES]);
B200BLUR_API int b200blur_ipc_open(b200blur_ctx *ctx, const unsigned char handle[B200BLUR_IPC_HANDLE_BYTES], void **dptr);
B200BLUR_API int b200blur_ipc_close(b200blur_ctx *ctx, void *dptr);

#ifdef __cplusplus
}
#endif
#endif /* B200BLUR_H */
