// b200blur.cu -- implementation of the C ABI declared in include/b200blur.h.
//
// The thin layer that replaces the OpenCL plumbing of heterogeneous_blur.c / split_image_blur.c
// (clCreateContext / clCreateCommandQueue / clCreateBuffer / clEnqueue{Write,NDRange,Read} / clFinish /
// clGetEventProfilingInfo) with the CUDA runtime, plus the two stream engines that replace
// the batch loop of heterogeneous_blur.c:418-600.  CUDA runtime only: no PyTorch, no OpenCL, no CPU fallback.
#include "b200blur.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "blur_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? B200BLUR_ERR_NOMEM                         \
                        : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ||             \
                           e_ == cudaErrorInvalidDevice)                                               \
                            ? B200BLUR_ERR_NO_DEVICE                                                   \
                            : B200BLUR_ERR_CUDA;                                                       \
            return fail(code_, "%d - %s failed: %s", (int)e_, #expr, cudaGetErrorString(e_));          \
        }                                                                                              \
    } while (0)

struct EventSlot {
    cudaEvent_t start = nullptr, end = nullptr;
    bool in_use = false;
};

}  // namespace

struct b200blur_ctx {
    int device = -1;
    int sm_count = 0;
    std::vector<cudaStream_t> queues;
    std::vector<EventSlot> events;
    std::vector<int> free_events;
    int kernel_variant = 0;   // 0 auto, 1 register/shuffle strips, 2 TMA-bulk streamed
    int64_t launches = 0;
    // tuning knobs of the streamed kernel (0 = automatic); set from B200BLUR_V2_* at context creation
    int v2_threads = 0, v2_seg = 0, v2_cfg = 0, v2_ctas_per_sm = 0;
    // ring of device buffers owned by b200blur_run_host
    struct Slot {
        uint8_t *d_in = nullptr, *d_out = nullptr;
        uint8_t *t_in = nullptr, *t_out = nullptr;  // tight staging for odd widths (re-pitched on the device)
        cudaEvent_t ev[6] = {};  // h2d start/end, kernel start/end, d2h start/end
    };
    std::vector<Slot> ring;
    size_t ring_slot_bytes = 0, ring_tight_bytes = 0;
    // work counters of the streamed kernel: two 64-bit words per queue, zero between launches
    unsigned long long *d_work = nullptr;
    // per-kernel launch facts (max dynamic smem attribute set, resident CTAs/SM), cached: both calls are slow
    struct KernelInfo { const void *fn; int block; size_t smem; int per_sm; };
    std::vector<KernelInfo> kernel_info;
    // scratch pair for re-pitching odd-width resident streams (b200blur_run_resident)
    uint8_t *scratch_in = nullptr, *scratch_out = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t fork_event = nullptr;          // fork/join of the per-batch launches of b200blur_run_resident
    std::vector<cudaEvent_t> join_events;
    // CUDA graph of the last per-batch launch sequence (launch-bound loop: hundreds of small kernels)
    struct GraphKey {
        const void *in = nullptr; void *out = nullptr;
        int w = 0, h = 0, c = 0, batch = 0; int64_t n = 0;
        bool operator==(const GraphKey &o) const
        { return in == o.in && out == o.out && w == o.w && h == o.h && c == o.c && batch == o.batch && n == o.n; }
    } graph_key;
    int graph_seen = 0;                        // times graph_key was requested without a graph
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_launches = 0;
};

namespace {

int ctx_check(const b200blur_ctx *ctx)
{
    if (!ctx) return fail(B200BLUR_ERR_INVALID, "context is NULL");
    return B200BLUR_OK;
}

int queue_check(const b200blur_ctx *ctx, int queue)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (queue < 0 || queue >= (int)ctx->queues.size())
        return fail(B200BLUR_ERR_INVALID, "queue %d out of range [0,%d)", queue, (int)ctx->queues.size());
    return B200BLUR_OK;
}

int event_begin(b200blur_ctx *ctx, int queue, b200blur_event *ev, int *slot_out)
{
    *slot_out = -1;
    if (!ev) return B200BLUR_OK;
    int idx;
    if (!ctx->free_events.empty()) {
        idx = ctx->free_events.back();
        ctx->free_events.pop_back();
    } else {
        EventSlot s;
        CU_TRY(cudaEventCreate(&s.start));
        CU_TRY(cudaEventCreate(&s.end));
        ctx->events.push_back(s);
        idx = (int)ctx->events.size() - 1;
    }
    ctx->events[idx].in_use = true;
    CU_TRY(cudaEventRecord(ctx->events[idx].start, ctx->queues[queue]));
    *slot_out = idx;
    *ev = idx;
    return B200BLUR_OK;
}

int event_end(b200blur_ctx *ctx, int queue, int slot)
{
    if (slot < 0) return B200BLUR_OK;
    CU_TRY(cudaEventRecord(ctx->events[slot].end, ctx->queues[queue]));
    return B200BLUR_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

size_t in_pitch_of(const b200blur_launch *l) { return l->in_row_pitch ? l->in_row_pitch : (size_t)l->width * l->channels; }
size_t out_pitch_of(const b200blur_launch *l) { return l->out_row_pitch ? l->out_row_pitch : (size_t)l->width * l->channels; }

bool launch_vectorised(const b200blur_launch *l)
{
    if (l->channels < 1 || l->channels > 4) return false;
    if (in_pitch_of(l) % 16 != 0 || out_pitch_of(l) % 16 != 0) return false;
    if (!aligned16(l->in) || !aligned16(l->out)) return false;
    if (l->n_images > 1 && (l->in_image_stride % 16 || l->out_image_stride % 16)) return false;
    if (l->halo_top && (!aligned16(l->halo_top) || (l->n_images > 1 && l->halo_top_stride % 16))) return false;
    if (l->halo_bottom && (!aligned16(l->halo_bottom) || (l->n_images > 1 && l->halo_bottom_stride % 16)))
        return false;
    return true;
}

int launch_validate(const b200blur_launch *l)
{
    if (!l) return fail(B200BLUR_ERR_INVALID, "launch is NULL");
    if (l->width < 0 || l->rows < 0 || l->n_images < 0 || l->channels < 1)
        return fail(B200BLUR_ERR_INVALID, "negative size or channels < 1 (width %d rows %d channels %d n_images %lld)",
                    l->width, l->rows, l->channels, (long long)l->n_images);
    if (l->reserved != 0) return fail(B200BLUR_ERR_INVALID, "launch.reserved must be 0");
    const size_t row_bytes = (size_t)l->width * l->channels;
    if (row_bytes > 0x7fffffffULL || in_pitch_of(l) > 0x7fffffffULL || out_pitch_of(l) > 0x7fffffffULL)
        return fail(B200BLUR_ERR_INVALID, "row pitch exceeds 2^31-1 bytes");
    if (in_pitch_of(l) < row_bytes || out_pitch_of(l) < row_bytes)
        return fail(B200BLUR_ERR_INVALID, "row pitch smaller than width*channels");
    const bool empty = l->width == 0 || l->rows == 0 || l->n_images == 0;
    if (!empty && (!l->in || !l->out)) return fail(B200BLUR_ERR_INVALID, "in/out pointer is NULL");
    if (!empty && l->in == l->out) return fail(B200BLUR_ERR_INVALID, "in-place blur (in == out) is not supported");
    return B200BLUR_OK;
}

b200blur::BandParams to_params(const b200blur_launch *l)
{
    b200blur::BandParams p;
    p.in = static_cast<const uint8_t *>(l->in);
    p.out = static_cast<uint8_t *>(l->out);
    p.halo_top = static_cast<const uint8_t *>(l->halo_top);
    p.halo_bot = static_cast<const uint8_t *>(l->halo_bottom);
    p.in_stride = l->in_image_stride;
    p.out_stride = l->out_image_stride;
    p.top_stride = l->halo_top_stride;
    p.bot_stride = l->halo_bottom_stride;
    p.row_bytes = l->width * l->channels;
    p.pitch = (int)in_pitch_of(l);
    p.out_pitch = (int)out_pitch_of(l);
    p.rows = l->rows;
    p.width = l->width;
    p.channels = l->channels;
    p.n_images = l->n_images;
    return p;
}

template <int RS>
void launch_strip(const b200blur::BandParams &p, cudaStream_t s, long long img0, long long n)
{
    const int cpr = p.row_bytes / 16;
    const int n_strips = (p.rows + RS - 1) / RS;
    const long long units = (long long)cpr * n_strips;
    const int block = 256;
    b200blur::BandParams q = p;
    q.in += (size_t)img0 * p.in_stride;
    q.out += (size_t)img0 * p.out_stride;
    if (q.halo_top) q.halo_top += (size_t)img0 * p.top_stride;
    if (q.halo_bot) q.halo_bot += (size_t)img0 * p.bot_stride;
    q.n_images = n;
    dim3 grid((unsigned)n, (unsigned)((units + block - 1) / block));
    switch (p.channels) {
        case 1: b200blur::blur_strip_kernel<1, RS><<<grid, block, 0, s>>>(q, cpr, n_strips); break;
        case 2: b200blur::blur_strip_kernel<2, RS><<<grid, block, 0, s>>>(q, cpr, n_strips); break;
        case 3: b200blur::blur_strip_kernel<3, RS><<<grid, block, 0, s>>>(q, cpr, n_strips); break;
        default: b200blur::blur_strip_kernel<4, RS><<<grid, block, 0, s>>>(q, cpr, n_strips); break;
    }
}

using StreamKernel = void (*)(const b200blur::StreamParams);

struct StreamCfg {
    int rb, ns;
    StreamKernel fn[4];       // by channels-1: rows end on a chunk boundary
    StreamKernel fn_edge[4];  // by channels-1: rows end inside a chunk (pitched rows)
};

template <int RB, int NS>
constexpr StreamCfg make_cfg()
{
    return StreamCfg{RB, NS,
                     {b200blur::blur_stream_kernel<1, RB, NS, false>, b200blur::blur_stream_kernel<2, RB, NS, false>,
                      b200blur::blur_stream_kernel<3, RB, NS, false>, b200blur::blur_stream_kernel<4, RB, NS, false>},
                     {b200blur::blur_stream_kernel<1, RB, NS, true>, b200blur::blur_stream_kernel<2, RB, NS, true>,
                      b200blur::blur_stream_kernel<3, RB, NS, true>, b200blur::blur_stream_kernel<4, RB, NS, true>}};
}

// {rows per slot, slots}; index 0 is the default (B200BLUR_V2_CFG selects another for tuning runs)
// (round-1 sweep over {8,4} {8,3} {4,3} {4,4} {4,6} {16,2} {8,2}: all within 3 % once work is handed out dynamically)
const StreamCfg kStreamCfgs[] = {make_cfg<8, 4>(), make_cfg<8, 3>()};
constexpr int kNumStreamCfgs = sizeof(kStreamCfgs) / sizeof(kStreamCfgs[0]);

// Whether the streamed (variant 2) kernel can run this launch: rows wide enough for bulk copies to pay.
bool stream_eligible(const b200blur::BandParams &p) { return p.row_bytes >= 256; }

// PRMT selectors that apply the right-edge clamp to the {wl, w, wr} window of the chunk holding the end of a row whose
// length is not a multiple of 16 (see StreamParams): window byte idx takes the byte C positions earlier when it is one
// of the C bytes just past the end of the row.  Selector nibbles index the 8 bytes of (previous word, this word).
void edge_selectors(b200blur::StreamParams &sp, int row_bytes, int channels)
{
    const int v = row_bytes - (sp.cpr - 1) * 16;  // bytes of the row inside its last chunk, 1..16
    sp.edge_general = (v != 16);
    sp.edge_prev = 0;
    sp.sel_prev = 0x7654;
    for (int m = 0; m < 6; m++) sp.sel_last[m] = 0x7654;
    if (!sp.edge_general) return;
    const int t0 = v + 4;  // window index of the first byte past the end of the row (window starts 4 bytes before the chunk)
    for (int m = 1; m < 6; m++) {
        uint32_t sel = 0;
        for (int b = 0; b < 4; b++) {
            const int idx = 4 * m + b;
            const int src = (idx >= t0 && idx < t0 + channels) ? idx - channels : idx;
            sel |= (uint32_t)(src - 4 * (m - 1)) << (4 * b);
        }
        sp.sel_last[m] = sel;
    }
    if (v < 4 && sp.cpr >= 2) {  // the end of the row is within the first word of the last chunk = the wr of the chunk before
        sp.edge_prev = 1;
        uint32_t sel = 0;
        for (int b = 0; b < 4; b++) {
            const int idx = 20 + b;
            const int src = (b >= v && b < v + channels) ? idx - channels : idx;
            sel |= (uint32_t)(src - 16) << (4 * b);
        }
        sp.sel_prev = sel;
    }
}

int launch_stream(b200blur_ctx *ctx, const b200blur::BandParams &p, cudaStream_t s, int queue)
{
    b200blur::StreamParams sp;
    sp.b = p;
    sp.cpr = (p.row_bytes + 15) / 16;   // live chunks per row; bytes past row_bytes up to the pitch are padding
    edge_selectors(sp, p.row_bytes, p.channels);
    const StreamCfg &cfg = kStreamCfgs[(ctx->v2_cfg >= 0 && ctx->v2_cfg < kNumStreamCfgs) ? ctx->v2_cfg : 0];
    int threads;
    if (p.pitch <= 4096) {
        // full-width rows: a CTA step covers `ipc` images side by side; pick the block size that wastes fewest lanes
        sp.cb = sp.cpr;
        sp.ncb = 1;
        sp.margin = 0;
        const int prefer = ctx->v2_threads > 0 ? ctx->v2_threads : 128;
        int best_t = 0;
        double best_score = -1.0;
        for (int t = 64; t <= 256; t += 32) {
            if (t < sp.cb) continue;
            const int ipc = t / sp.cb;
            const double eff = (double)(ipc * sp.cb) / t;
            const double score = eff - 0.0005 * (t > prefer ? t - prefer : prefer - t);
            if (score > best_score) { best_score = score; best_t = t; }
        }
        threads = best_t;
        sp.ipc = threads / sp.cb;
    } else {
        threads = ctx->v2_threads > 0 ? ctx->v2_threads : 128;
        if (threads > 256) threads = 256;
        sp.cb = threads;
        sp.ncb = (sp.cpr + sp.cb - 1) / sp.cb;
        sp.margin = 16;
        sp.ipc = 1;
    }
    sp.sstride = sp.margin ? sp.cb * 16 + 2 * sp.margin : p.pitch;   // full-width: rows land exactly as they lie in memory
    sp.slot_bytes = sp.ipc * cfg.rb * sp.sstride;
    const size_t smem = 16 + (size_t)cfg.ns * sp.slot_bytes + 16 + 24 * cfg.ns;
    const int block = threads + 32;  // + the producer warp
    if (smem > 220 * 1024) return fail(B200BLUR_ERR_INVALID, "streamed kernel needs %zu B of shared memory", smem);
    StreamKernel fn = sp.edge_general ? cfg.fn_edge[p.channels - 1] : cfg.fn[p.channels - 1];
    int per_sm = 0;
    for (auto &ki : ctx->kernel_info)
        if (ki.fn == (const void *)fn && ki.block == block && ki.smem == smem) per_sm = ki.per_sm;
    if (per_sm == 0) {
        // opt in to the largest dynamic shared-memory size once per kernel (any later, smaller request is covered)
        CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, smem));
        if (per_sm < 1) return fail(B200BLUR_ERR_CUDA, "streamed kernel does not fit on an SM");
        ctx->kernel_info.push_back({(const void *)fn, block, smem, per_sm});
    }
    if (ctx->v2_ctas_per_sm > 0 && per_sm > ctx->v2_ctas_per_sm) per_sm = ctx->v2_ctas_per_sm;
    const long long slots = (long long)ctx->sm_count * per_sm;
    sp.img_blocks = (p.n_images + sp.ipc - 1) / sp.ipc;
    // Work unit = `seg` output rows of `ipc` images (or of one column block): ~48 KB in + 48 KB out for full-width
    // rows, ~128 KB for column blocks; the band is cut into equal segments of about that size.
    int seg;
    if (ctx->v2_seg > 0) {
        seg = ctx->v2_seg;
    } else {
        const double unit_bytes = sp.ncb == 1 ? 48.0 * 1024 : 128.0 * 1024;
        const double row_bytes = (double)sp.ipc * sp.cb * 16;
        long long want = (long long)(unit_bytes / row_bytes + 0.5);
        if (want < 6) want = 6;
        // very small launches only: shorter units until there is one per SM.  (Shrinking further to "fill" every CTA
        // slot makes small launches slower: 143 launches of 35 images take 0.85 ms with 6-row units, 0.44 ms with 24.)
        while (want > 6 && sp.img_blocks * sp.ncb * ((p.rows + want - 1) / want) < ctx->sm_count) want = (want + 1) / 2;
        const long long nseg = (p.rows + want - 1) / want;
        seg = (int)((p.rows + nseg - 1) / nseg);
    }
    if (seg > p.rows) seg = p.rows;
    sp.seg = seg;
    sp.nseg = (p.rows + seg - 1) / seg;
    sp.n_groups = sp.img_blocks * sp.nseg * sp.ncb;
    if (sp.n_groups >= 0x7fffffffLL || (long long)sp.nseg * sp.ncb >= 0x7fffffffLL)
        return fail(B200BLUR_ERR_INVALID, "too many work units for one launch (%lld)", sp.n_groups);
    sp.work = ctx->d_work + 2 * queue;
    const long long grid = sp.n_groups < slots ? sp.n_groups : slots;
    fn<<<(unsigned)grid, block, smem, s>>>(sp);
    return B200BLUR_OK;
}

// Launches the device code for one b200blur_launch on stream s.  Returns the number of kernels launched.
int do_launch(b200blur_ctx *ctx, int queue, const b200blur_launch *l, int *n_kernels)
{
    cudaStream_t s = ctx->queues[queue];
    *n_kernels = 0;
    if (l->width == 0 || l->rows == 0 || l->n_images == 0) return B200BLUR_OK;  // nothing to do
    b200blur::BandParams p = to_params(l);
    const bool vec = launch_vectorised(l);
    if (vec && stream_eligible(p) && ctx->kernel_variant != 1) {
        if (int rc = launch_stream(ctx, p, s, queue)) return rc;
        ++*n_kernels;
    } else if (vec && p.row_bytes % 16 == 0) {
        // strip height: tall strips amortise the two halo rows; short strips expose more threads for small batches
        const long long cpr = p.row_bytes / 16;
        const long long threads_rs16 = cpr * ((p.rows + 15) / 16) * p.n_images;
        const bool small = threads_rs16 < (long long)ctx->sm_count * 1024;
        const long long max_grid_y = 65535;
        if ((cpr * ((p.rows + 3) / 4) + 255) / 256 > max_grid_y)
            return fail(B200BLUR_ERR_INVALID, "image too large for one launch (%lld chunks x %d rows)", cpr, p.rows);
        const long long chunk = 0x7fffffffLL;
        for (long long i0 = 0; i0 < p.n_images; i0 += chunk) {
            const long long n = (p.n_images - i0 < chunk) ? p.n_images - i0 : chunk;
            if (small) launch_strip<4>(p, s, i0, n);
            else launch_strip<16>(p, s, i0, n);
            ++*n_kernels;
        }
    } else {
        const long long total = (long long)p.rows * p.row_bytes * p.n_images;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)ctx->sm_count * 32;
        if (blocks > cap) blocks = cap;
        b200blur::blur_generic_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
        ++*n_kernels;
    }
    CU_TRY(cudaGetLastError());
    ctx->launches += *n_kernels;
    return B200BLUR_OK;
}

// tight <-> pitched row re-packing on stream s (see repitch_*_kernel); both pointers 16-byte aligned
void launch_repitch_in(b200blur_ctx *ctx, cudaStream_t s, const void *tight, void *pitched, long long rows, int row_bytes, int pitch)
{
    const long long total = rows * ((row_bytes + 15) / 16);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    b200blur::repitch_in_kernel<<<(unsigned)blocks, 256, 0, s>>>(static_cast<const uint8_t *>(tight), static_cast<uint8_t *>(pitched),
                                                                 rows, row_bytes, pitch);
    ctx->launches++;
}

// `tight_base` is 16-byte aligned; the rows land at byte offset `lo` of it
void launch_repitch_out(b200blur_ctx *ctx, cudaStream_t s, const void *pitched, void *tight_base, long long lo, long long rows,
                        int row_bytes, int pitch)
{
    const long long total = (rows * (long long)row_bytes + 15) / 16 + 1;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    b200blur::repitch_out_kernel<<<(unsigned)blocks, 256, 0, s>>>(static_cast<const uint8_t *>(pitched),
                                                                  static_cast<uint8_t *>(tight_base), lo, rows, row_bytes, pitch);
    ctx->launches++;
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ============================================================================================== C ABI
extern "C" {

const char *b200blur_last_error(void) { return g_last_error.c_str(); }
const char *b200blur_version(void) { return "b200blur 0.1.0 sm_100a"; }

int b200blur_device_count(int *count)
{
    if (!count) return fail(B200BLUR_ERR_INVALID, "count is NULL");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200BLUR_ERR_NO_DEVICE, "%d - no CUDA device: %s", (int)e, cudaGetErrorString(e));
    }
    *count = n;
    if (n == 0) return fail(B200BLUR_ERR_NO_DEVICE, "no CUDA device");
    return B200BLUR_OK;
}

int b200blur_device_name(int device, char *buf, size_t buf_len)
{
    if (!buf || buf_len == 0) return fail(B200BLUR_ERR_INVALID, "buffer is NULL/empty");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    snprintf(buf, buf_len, "%s", prop.name);
    return B200BLUR_OK;
}

int b200blur_device_props(int device, int *sm_count, int *cc, size_t *global_mem_bytes)
{
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc) *cc = prop.major * 10 + prop.minor;
    if (global_mem_bytes) *global_mem_bytes = prop.totalGlobalMem;
    return B200BLUR_OK;
}

int b200blur_ctx_create(int device, int n_queues, b200blur_ctx **out)
{
    if (!out) return fail(B200BLUR_ERR_INVALID, "ctx out-pointer is NULL");
    *out = nullptr;
    int n = 0;
    if (int rc = b200blur_device_count(&n)) return rc;
    if (device < 0 || device >= n) return fail(B200BLUR_ERR_NO_DEVICE, "device %d out of range [0,%d)", device, n);
    if (n_queues <= 0) n_queues = 4;
    if (n_queues > 64) return fail(B200BLUR_ERR_INVALID, "n_queues %d > 64", n_queues);
    CU_TRY(cudaSetDevice(device));
    b200blur_ctx *ctx = new (std::nothrow) b200blur_ctx;
    if (!ctx) return fail(B200BLUR_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(B200BLUR_ERR_CUDA, "%d - cudaGetDeviceProperties: %s", (int)e, cudaGetErrorString(e));
    }
    ctx->sm_count = prop.multiProcessorCount;
    auto env_int = [](const char *name) { const char *v = getenv(name); return v ? atoi(v) : 0; };
    ctx->kernel_variant = env_int("B200BLUR_VARIANT");
    ctx->v2_threads = env_int("B200BLUR_V2_THREADS");
    ctx->v2_seg = env_int("B200BLUR_V2_SEG");
    ctx->v2_cfg = env_int("B200BLUR_V2_CFG");
    ctx->v2_ctas_per_sm = env_int("B200BLUR_V2_CTAS");
    for (int i = 0; i < n_queues; i++) {
        cudaStream_t s;
        e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            for (auto q : ctx->queues) cudaStreamDestroy(q);
            delete ctx;
            return fail(B200BLUR_ERR_CUDA, "%d - cudaStreamCreate: %s", (int)e, cudaGetErrorString(e));
        }
        ctx->queues.push_back(s);
    }
    e = cudaMalloc((void **)&ctx->d_work, sizeof(unsigned long long) * 2 * n_queues);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_work, 0, sizeof(unsigned long long) * 2 * n_queues);
    if (e != cudaSuccess) {
        for (auto q : ctx->queues) cudaStreamDestroy(q);
        delete ctx;
        return fail(B200BLUR_ERR_CUDA, "%d - work counter allocation: %s", (int)e, cudaGetErrorString(e));
    }
    *out = ctx;
    return B200BLUR_OK;
}

static void ring_release(b200blur_ctx *ctx)
{
    for (auto &s : ctx->ring) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.t_in) cudaFree(s.t_in);
        if (s.t_out) cudaFree(s.t_out);
        for (auto &e : s.ev)
            if (e) cudaEventDestroy(e);
    }
    ctx->ring.clear();
    ctx->ring_slot_bytes = 0;
    ctx->ring_tight_bytes = 0;
}

int b200blur_ctx_destroy(b200blur_ctx *ctx)
{
    if (!ctx) return B200BLUR_OK;
    cudaSetDevice(ctx->device);
    for (auto q : ctx->queues) cudaStreamSynchronize(q);
    ring_release(ctx);
    if (ctx->d_work) cudaFree(ctx->d_work);
    if (ctx->scratch_in) cudaFree(ctx->scratch_in);
    if (ctx->scratch_out) cudaFree(ctx->scratch_out);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    for (auto e : ctx->join_events)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->events) {
        if (e.start) cudaEventDestroy(e.start);
        if (e.end) cudaEventDestroy(e.end);
    }
    for (auto q : ctx->queues) cudaStreamDestroy(q);
    delete ctx;
    return B200BLUR_OK;
}

int b200blur_ctx_device(const b200blur_ctx *ctx) { return ctx ? ctx->device : -1; }
int b200blur_ctx_num_queues(const b200blur_ctx *ctx) { return ctx ? (int)ctx->queues.size() : 0; }
void *b200blur_ctx_queue_handle(const b200blur_ctx *ctx, int queue)
{
    if (!ctx || queue < 0 || queue >= (int)ctx->queues.size()) return nullptr;
    return (void *)ctx->queues[queue];
}
int64_t b200blur_ctx_launch_count(const b200blur_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b200blur_set_kernel_variant(b200blur_ctx *ctx, int variant)
{
    if (!ctx) return 0;
    int prev = ctx->kernel_variant;
    ctx->kernel_variant = variant;
    return prev;
}

// ------------------------------------------------------------------------------------------------- memory
int b200blur_dev_alloc(b200blur_ctx *ctx, size_t bytes, void **dptr)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!dptr) return fail(B200BLUR_ERR_INVALID, "dptr is NULL");
    *dptr = nullptr;
    CU_TRY(cudaSetDevice(ctx->device));
    if (bytes == 0) bytes = 16;
    CU_TRY(cudaMalloc(dptr, bytes));
    return B200BLUR_OK;
}

int b200blur_dev_free(b200blur_ctx *ctx, void *dptr)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!dptr) return B200BLUR_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaFree(dptr));
    return B200BLUR_OK;
}

int b200blur_host_alloc(size_t bytes, void **hptr)
{
    if (!hptr) return fail(B200BLUR_ERR_INVALID, "hptr is NULL");
    *hptr = nullptr;
    if (bytes == 0) bytes = 16;
    CU_TRY(cudaHostAlloc(hptr, bytes, cudaHostAllocPortable));
    return B200BLUR_OK;
}

int b200blur_host_free(void *hptr)
{
    if (!hptr) return B200BLUR_OK;
    CU_TRY(cudaFreeHost(hptr));
    return B200BLUR_OK;
}

int b200blur_host_register(void *hptr, size_t bytes)
{
    if (!hptr) return fail(B200BLUR_ERR_INVALID, "hptr is NULL");
    CU_TRY(cudaHostRegister(hptr, bytes, cudaHostRegisterPortable));
    return B200BLUR_OK;
}

int b200blur_host_unregister(void *hptr)
{
    if (!hptr) return B200BLUR_OK;
    CU_TRY(cudaHostUnregister(hptr));
    return B200BLUR_OK;
}

// ------------------------------------------------------------------------------------------------- events
int b200blur_event_ms(b200blur_ctx *ctx, b200blur_event ev, double *ms)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!ms) return fail(B200BLUR_ERR_INVALID, "ms is NULL");
    if (ev < 0 || ev >= (int)ctx->events.size() || !ctx->events[ev].in_use)
        return fail(B200BLUR_ERR_INVALID, "event %d is not live", (int)ev);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaEventSynchronize(ctx->events[ev].end));
    float f = 0.f;
    CU_TRY(cudaEventElapsedTime(&f, ctx->events[ev].start, ctx->events[ev].end));
    *ms = f;
    return B200BLUR_OK;
}

int b200blur_event_release(b200blur_ctx *ctx, b200blur_event ev)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (ev < 0 || ev >= (int)ctx->events.size() || !ctx->events[ev].in_use)
        return fail(B200BLUR_ERR_INVALID, "event %d is not live", (int)ev);
    ctx->events[ev].in_use = false;
    ctx->free_events.push_back(ev);
    return B200BLUR_OK;
}

int b200blur_enqueue_marker(b200blur_ctx *ctx, int queue, b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (!ev) return fail(B200BLUR_ERR_INVALID, "ev is NULL");
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    return event_end(ctx, queue, slot);
}

int b200blur_events_elapsed_ms(b200blur_ctx *ctx, b200blur_event from, b200blur_event to, double *ms)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!ms) return fail(B200BLUR_ERR_INVALID, "ms is NULL");
    for (b200blur_event ev : {from, to})
        if (ev < 0 || ev >= (int)ctx->events.size() || !ctx->events[ev].in_use)
            return fail(B200BLUR_ERR_INVALID, "event %d is not live", (int)ev);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaEventSynchronize(ctx->events[from].end));
    CU_TRY(cudaEventSynchronize(ctx->events[to].end));
    float f = 0.f;
    CU_TRY(cudaEventElapsedTime(&f, ctx->events[from].end, ctx->events[to].end));
    *ms = f;
    return B200BLUR_OK;
}

int b200blur_enqueue_wait(b200blur_ctx *ctx, int queue, b200blur_event ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (ev < 0 || ev >= (int)ctx->events.size() || !ctx->events[ev].in_use)
        return fail(B200BLUR_ERR_INVALID, "event %d is not live", (int)ev);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaStreamWaitEvent(ctx->queues[queue], ctx->events[ev].end, 0));
    return B200BLUR_OK;
}

int b200blur_enqueue_wait_peer(b200blur_ctx *ctx, int queue, b200blur_ctx *src, b200blur_event ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (int rc = ctx_check(src)) return rc;
    if (ev < 0 || ev >= (int)src->events.size() || !src->events[ev].in_use)
        return fail(B200BLUR_ERR_INVALID, "event %d is not live in the source context", (int)ev);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaStreamWaitEvent(ctx->queues[queue], src->events[ev].end, 0));
    return B200BLUR_OK;
}

// ----------------------------------------------------------------------------------------------- transfers
int b200blur_enqueue_write(b200blur_ctx *ctx, int queue, void *dst_dev, const void *src_host, size_t bytes,
                           b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (bytes && (!dst_dev || !src_host)) return fail(B200BLUR_ERR_INVALID, "NULL pointer in enqueue_write");
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    if (bytes) CU_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->queues[queue]));
    return event_end(ctx, queue, slot);
}

int b200blur_enqueue_read(b200blur_ctx *ctx, int queue, void *dst_host, const void *src_dev, size_t bytes,
                          b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (bytes && (!dst_host || !src_dev)) return fail(B200BLUR_ERR_INVALID, "NULL pointer in enqueue_read");
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    if (bytes) CU_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->queues[queue]));
    return event_end(ctx, queue, slot);
}

int b200blur_enqueue_write_2d(b200blur_ctx *ctx, int queue, void *dst_dev, size_t dst_pitch, const void *src_host,
                              size_t src_pitch, size_t row_bytes, size_t rows, b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (row_bytes && rows && (!dst_dev || !src_host)) return fail(B200BLUR_ERR_INVALID, "NULL pointer in enqueue_write_2d");
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    if (row_bytes && rows)
        CU_TRY(cudaMemcpy2DAsync(dst_dev, dst_pitch, src_host, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice,
                                 ctx->queues[queue]));
    return event_end(ctx, queue, slot);
}

int b200blur_enqueue_read_2d(b200blur_ctx *ctx, int queue, void *dst_host, size_t dst_pitch, const void *src_dev,
                             size_t src_pitch, size_t row_bytes, size_t rows, b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (row_bytes && rows && (!dst_host || !src_dev)) return fail(B200BLUR_ERR_INVALID, "NULL pointer in enqueue_read_2d");
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    if (row_bytes && rows)
        CU_TRY(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost,
                                 ctx->queues[queue]));
    return event_end(ctx, queue, slot);
}

int b200blur_finish(b200blur_ctx *ctx, int queue)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaStreamSynchronize(ctx->queues[queue]));
    return B200BLUR_OK;
}

int b200blur_finish_all(b200blur_ctx *ctx)
{
    if (int rc = ctx_check(ctx)) return rc;
    CU_TRY(cudaSetDevice(ctx->device));
    for (auto q : ctx->queues) CU_TRY(cudaStreamSynchronize(q));
    return B200BLUR_OK;
}

// -------------------------------------------------------------------------------------------- kernel launch
int b200blur_launch_rows_pitched(b200blur_launch *l, const void *in, void *out, int width, int in_height, int channels,
                                 int first_row, int n_rows, int64_t n_images, size_t in_image_stride,
                                 size_t out_image_stride, size_t in_row_pitch, size_t out_row_pitch)
{
    if (!l) return fail(B200BLUR_ERR_INVALID, "launch is NULL");
    if (width < 0 || in_height < 0 || channels < 1 || first_row < 0 || n_rows < 0 || n_images < 0 ||
        (long long)first_row + n_rows > in_height)
        return fail(B200BLUR_ERR_INVALID, "bad geometry: width %d in_height %d channels %d rows [%d,%d+%d)", width,
                    in_height, channels, first_row, first_row, n_rows);
    const size_t row_bytes = (size_t)width * channels;
    if ((in_row_pitch && in_row_pitch < row_bytes) || (out_row_pitch && out_row_pitch < row_bytes))
        return fail(B200BLUR_ERR_INVALID, "row pitch smaller than width*channels");
    const size_t pitch = in_row_pitch ? in_row_pitch : row_bytes;
    const uint8_t *base = static_cast<const uint8_t *>(in);
    memset(l, 0, sizeof *l);
    l->in = base ? base + (size_t)first_row * pitch : nullptr;
    l->out = out;
    l->width = width;
    l->channels = channels;
    l->rows = n_rows;
    l->n_images = n_images;
    l->in_image_stride = in_image_stride;
    l->out_image_stride = out_image_stride;
    l->in_row_pitch = in_row_pitch;
    l->out_row_pitch = out_row_pitch;
    if (base && n_rows > 0 && first_row > 0) {
        l->halo_top = base + (size_t)(first_row - 1) * pitch;
        l->halo_top_stride = in_image_stride;
    }
    if (base && n_rows > 0 && first_row + n_rows < in_height) {
        l->halo_bottom = base + (size_t)(first_row + n_rows) * pitch;
        l->halo_bottom_stride = in_image_stride;
    }
    return B200BLUR_OK;
}

int b200blur_launch_rows(b200blur_launch *l, const void *in, void *out, int width, int in_height, int channels,
                         int first_row, int n_rows, int64_t n_images, size_t in_image_stride,
                         size_t out_image_stride)
{
    return b200blur_launch_rows_pitched(l, in, out, width, in_height, channels, first_row, n_rows, n_images,
                                        in_image_stride, out_image_stride, 0, 0);
}

int b200blur_plan_row_edge(int row_bytes, int channels, uint32_t out[10])
{
    if (!out || row_bytes < 1 || channels < 1 || channels > 4 || row_bytes % channels)
        return fail(B200BLUR_ERR_INVALID, "bad row plan request (row_bytes %d channels %d)", row_bytes, channels);
    b200blur::StreamParams sp;
    sp.cpr = (row_bytes + 15) / 16;
    edge_selectors(sp, row_bytes, channels);
    out[0] = (uint32_t)sp.cpr;
    out[1] = (uint32_t)sp.edge_general;
    out[2] = (uint32_t)sp.edge_prev;
    for (int m = 0; m < 6; m++) out[3 + m] = sp.sel_last[m];
    out[9] = sp.sel_prev;
    return B200BLUR_OK;
}

int b200blur_launch_is_vectorised(const b200blur_launch *launch)
{
    if (!launch) return 0;
    return launch_vectorised(launch) ? 1 : 0;
}

int b200blur_enqueue_blur(b200blur_ctx *ctx, int queue, const b200blur_launch *launch, b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (int rc = launch_validate(launch)) return rc;
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    int nk;
    if (int rc = do_launch(ctx, queue, launch, &nk)) return rc;
    return event_end(ctx, queue, slot);
}

// ----------------------------------------------------------------------------------- work distribution (L4)
int b200blur_partition(int64_t n_items, int n_parts, int part, int64_t *begin, int64_t *count)
{
    if (n_items < 0 || n_parts < 1 || part < 0 || part >= n_parts || !begin || !count)
        return fail(B200BLUR_ERR_INVALID, "bad partition request (%lld items, part %d of %d)", (long long)n_items, part,
                    n_parts);
    const int64_t q = n_items / n_parts, r = n_items % n_parts;
    *begin = part * q + (part < r ? part : r);
    *count = q + (part < r ? 1 : 0);
    return B200BLUR_OK;
}

int b200blur_ratio_split_images(int batch_count, float gpu_ratio, int mode, int *n_first, int *n_second)
{
    if (batch_count < 0 || !n_first || !n_second || mode < 0 || mode > 2)
        return fail(B200BLUR_ERR_INVALID, "bad ratio split request");
    int second = 0, first = 0;
    if (mode == 0) {
        second = (int)(batch_count * gpu_ratio);
        first = batch_count - second;
    } else if (mode == 1) {
        first = batch_count;
    } else {
        second = batch_count;
    }
    *n_first = first;
    *n_second = second;
    return B200BLUR_OK;
}

int b200blur_ratio_split_row(int height, float gpu_ratio, int *split_row)
{
    if (height < 2 || !split_row) return fail(B200BLUR_ERR_INVALID, "bad split-row request (height %d)", height);
    int s = (int)(height * (1.0f - gpu_ratio));
    if (s < 1) s = 1;
    if (s > height - 1) s = height - 1;
    *split_row = s;
    return B200BLUR_OK;
}

// ------------------------------------------------------------------------------------------- stream engines
int b200blur_run_resident(b200blur_ctx *ctx, const void *d_in, void *d_out, int width, int height, int channels,
                          int64_t n_images, int batch_size, int coalesce, b200blur_stats *stats)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (batch_size < 1) return fail(B200BLUR_ERR_INVALID, "batch_size %d < 1", batch_size);
    if (width < 0 || height < 0 || channels < 1 || n_images < 0) return fail(B200BLUR_ERR_INVALID, "bad geometry");
    CU_TRY(cudaSetDevice(ctx->device));
    const size_t image_bytes = (size_t)width * height * channels;
    cudaStream_t s = ctx->queues[0];
    const double t0 = now_ms();
    b200blur_event ev_all = -1;
    int slot = -1;
    if (stats)
        if (int rc = event_begin(ctx, 0, &ev_all, &slot)) return rc;
    int64_t launches = 0;
    // Odd widths (width*channels % 16 != 0): tight rows cannot be read 16 bytes at a time, so the stream is re-pitched
    // through a scratch pair with strided device copies, ~128 MB at a time, and blurred on the vectorised path
    // (3 passes over the data instead of the ~20x slower byte-wise generic kernel).
    const size_t row_bytes = (size_t)width * channels;
    if (row_bytes % 16 != 0 && channels <= 4 && row_bytes >= 256 && n_images > 0 && height > 0) {
        const size_t dev_pitch = (row_bytes + 15) / 16 * 16;
        const size_t dev_image_bytes = dev_pitch * (size_t)height;
        int64_t chunk = (int64_t)((128ull << 20) / dev_image_bytes);
        if (chunk < 1) chunk = 1;
        if (chunk > n_images) chunk = n_images;
        const size_t need = dev_image_bytes * (size_t)chunk;
        if (ctx->scratch_bytes < need) {
            if (ctx->scratch_in) cudaFree(ctx->scratch_in);
            if (ctx->scratch_out) cudaFree(ctx->scratch_out);
            ctx->scratch_in = ctx->scratch_out = nullptr;
            ctx->scratch_bytes = 0;
            CU_TRY(cudaMalloc((void **)&ctx->scratch_in, need));
            CU_TRY(cudaMalloc((void **)&ctx->scratch_out, need));
            ctx->scratch_bytes = need;
        }
        for (int64_t i0 = 0; i0 < n_images; i0 += chunk) {
            const int64_t n = (n_images - i0 < chunk) ? n_images - i0 : chunk;
            const uint8_t *src = static_cast<const uint8_t *>(d_in) + (size_t)i0 * image_bytes;
            uint8_t *dst = static_cast<uint8_t *>(d_out) + (size_t)i0 * image_bytes;
            const bool fast_repack = aligned16(d_out);  // the re-pack kernel writes aligned 16-byte words of the tight output
            if (fast_repack) {
                launch_repitch_in(ctx, s, src, ctx->scratch_in, n * (long long)height, (int)row_bytes, (int)dev_pitch);
                launches++;
            } else {
                CU_TRY(cudaMemcpy2DAsync(ctx->scratch_in, dev_pitch, src, row_bytes, row_bytes, (size_t)n * height,
                                         cudaMemcpyDeviceToDevice, s));
            }
            b200blur_launch l;
            if (int rc = b200blur_launch_rows_pitched(&l, ctx->scratch_in, ctx->scratch_out, width, height, channels, 0,
                                                      height, n, dev_image_bytes, dev_image_bytes, dev_pitch, dev_pitch))
                return rc;
            int nk;
            if (int rc = do_launch(ctx, 0, &l, &nk)) return rc;
            launches += nk;
            if (fast_repack) {
                launch_repitch_out(ctx, s, ctx->scratch_out, d_out, (long long)((size_t)i0 * image_bytes), n * (long long)height,
                                   (int)row_bytes, (int)dev_pitch);
                launches++;
            } else {
                CU_TRY(cudaMemcpy2DAsync(dst, row_bytes, ctx->scratch_out, dev_pitch, row_bytes, (size_t)n * height,
                                         cudaMemcpyDeviceToDevice, s));
            }
        }
        if (stats) {
            if (int rc = event_end(ctx, 0, slot)) return rc;
            double ms = 0;
            if (int rc = b200blur_event_ms(ctx, ev_all, &ms)) return rc;
            b200blur_event_release(ctx, ev_all);
            memset(stats, 0, sizeof *stats);
            stats->kernel_ms = ms;
            stats->images = n_images;
            stats->launches = launches;
            stats->wall_ms = now_ms() - t0;
        }
        return B200BLUR_OK;
    }
    const int64_t step = coalesce ? (n_images > 0 ? n_images : 1) : batch_size;
    // One launch per batch (coalesce == 0): the batches are independent, so their launches are spread round-robin
    // over all queues of the context and overlap each other's ramp-up and tail; queue 0 forks and joins the others,
    // so the call still behaves like one in-order operation on queue 0.
    const int nq = (!coalesce && n_images > (int64_t)batch_size) ? (int)ctx->queues.size() : 1;
    cudaEvent_t fork = nullptr;
    if (nq > 1) {
        if (!ctx->fork_event) CU_TRY(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
        if (ctx->join_events.size() < ctx->queues.size()) {
            ctx->join_events.resize(ctx->queues.size(), nullptr);
            for (auto &e : ctx->join_events)
                if (!e) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        fork = ctx->fork_event;
    }
    // Launch-bound regime (many small per-batch launches): the second time the same sequence is requested it is
    // captured into a CUDA graph (fork/join over the queues included) and from then on replayed with one call.
    static const bool env_no_graph = getenv("B200BLUR_NO_GRAPH") != nullptr;
    b200blur_ctx::GraphKey key;
    key.in = d_in; key.out = d_out; key.w = width; key.h = height; key.c = channels; key.batch = batch_size; key.n = n_images;
    const bool graphable = nq > 1 && !env_no_graph && (n_images + step - 1) / step >= 8;
    bool capturing = false, replayed = false;
    if (graphable) {
        if (ctx->graph_exec && ctx->graph_key == key) {
            CU_TRY(cudaGraphLaunch(ctx->graph_exec, s));
            launches = ctx->graph_launches;
            ctx->launches += launches;
            replayed = true;
        } else if (ctx->graph_key == key && ctx->graph_seen >= 1) {
            capturing = true;
        } else {
            if (!(ctx->graph_key == key)) {
                ctx->graph_key = key;
                ctx->graph_seen = 0;
                if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
            }
            ctx->graph_seen++;
        }
    }
    if (!replayed) {
        if (capturing) CU_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        if (nq > 1) {
            CU_TRY(cudaEventRecord(fork, s));
            for (int q = 1; q < nq; q++) CU_TRY(cudaStreamWaitEvent(ctx->queues[q], fork, 0));
        }
        int64_t bi = 0;
        for (int64_t i0 = 0; i0 < n_images; i0 += step, bi++) {
            const int64_t n = (n_images - i0 < step) ? n_images - i0 : step;
            b200blur_launch l;
            if (int rc = b200blur_launch_rows(&l, static_cast<const uint8_t *>(d_in) + (size_t)i0 * image_bytes,
                                              static_cast<uint8_t *>(d_out) + (size_t)i0 * image_bytes, width, height,
                                              channels, 0, height, n, image_bytes, image_bytes))
                return rc;
            if (int rc = launch_validate(&l)) return rc;
            int nk;
            if (int rc = do_launch(ctx, (int)(bi % nq), &l, &nk)) return rc;
            launches += nk;
        }
        for (int q = 1; q < nq; q++) {
            CU_TRY(cudaEventRecord(ctx->join_events[q], ctx->queues[q]));
            CU_TRY(cudaStreamWaitEvent(s, ctx->join_events[q], 0));
        }
        if (capturing) {
            cudaGraph_t graph = nullptr;
            CU_TRY(cudaStreamEndCapture(s, &graph));
            cudaError_t e = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                ctx->graph_exec = nullptr;
                return fail(B200BLUR_ERR_CUDA, "%d - cudaGraphInstantiate: %s", (int)e, cudaGetErrorString(e));
            }
            ctx->graph_launches = launches;
            CU_TRY(cudaGraphLaunch(ctx->graph_exec, s));
        }
    }
    if (stats) {
        if (int rc = event_end(ctx, 0, slot)) return rc;
        double ms = 0;
        if (int rc = b200blur_event_ms(ctx, ev_all, &ms)) return rc;
        b200blur_event_release(ctx, ev_all);
        memset(stats, 0, sizeof *stats);
        stats->kernel_ms = ms;
        stats->images = n_images;
        stats->launches = launches;
        stats->wall_ms = now_ms() - t0;
    }
    return B200BLUR_OK;
}

static int ring_prepare(b200blur_ctx *ctx, size_t slot_bytes, size_t tight_bytes, int n_slots)
{
    if (ctx->ring_slot_bytes >= slot_bytes && ctx->ring_tight_bytes >= tight_bytes && (int)ctx->ring.size() == n_slots)
        return B200BLUR_OK;
    ring_release(ctx);
    ctx->ring.resize(n_slots);
    for (auto &s : ctx->ring) {
        CU_TRY(cudaMalloc((void **)&s.d_in, slot_bytes ? slot_bytes : 16));
        CU_TRY(cudaMalloc((void **)&s.d_out, slot_bytes ? slot_bytes : 16));
        if (tight_bytes) {
            CU_TRY(cudaMalloc((void **)&s.t_in, tight_bytes + 16));
            CU_TRY(cudaMalloc((void **)&s.t_out, tight_bytes + 16));
        }
        for (auto &e : s.ev) CU_TRY(cudaEventCreate(&e));
    }
    ctx->ring_slot_bytes = slot_bytes;
    ctx->ring_tight_bytes = tight_bytes;
    return B200BLUR_OK;
}

int b200blur_run_host(b200blur_ctx *ctx, const void *h_in, void *h_out, int width, int height, int channels,
                      int64_t n_images, int batch_size, b200blur_stats *stats)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (batch_size < 1) return fail(B200BLUR_ERR_INVALID, "batch_size %d < 1", batch_size);
    if (width < 0 || height < 0 || channels < 1 || n_images < 0) return fail(B200BLUR_ERR_INVALID, "bad geometry");
    if (ctx->queues.size() < 3) return fail(B200BLUR_ERR_INVALID, "run_host needs a context with >= 3 queues");
    const size_t image_bytes = (size_t)width * height * channels;
    if (n_images && image_bytes && (!h_in || !h_out)) return fail(B200BLUR_ERR_INVALID, "host pointer is NULL");
    CU_TRY(cudaSetDevice(ctx->device));
    // Transfer granularity.  Batches are independent, so the pipeline moves and launches them in chunks of about
    // 64 MB: several small batches fused, or a large batch cut into pieces.  (At one 8 MB batch per transfer the
    // per-chunk event/dependency gaps cap the link at 33 GB/s each way; at 64 MB it reaches the 45 GB/s this host link
    // sustains with both directions active -- tools/linkbench.py, tools/e2e.py.)  B200BLUR_E2E_CHUNK_MB overrides the
    // target (0 = exactly one batch per chunk, the reference's granularity); B200BLUR_RING overrides the ring depth.
    static const int env_ring = getenv("B200BLUR_RING") ? atoi(getenv("B200BLUR_RING")) : 0;
    static const int env_chunk_mb = getenv("B200BLUR_E2E_CHUNK_MB") ? atoi(getenv("B200BLUR_E2E_CHUNK_MB")) : 64;
    const int n_slots = env_ring > 1 ? env_ring : 4;
    if (env_chunk_mb > 0 && image_bytes > 0) {
        const double target = (double)env_chunk_mb * 1024 * 1024;
        const double batch_bytes = (double)batch_size * image_bytes;
        long long chunk = batch_size;
        if (batch_bytes < target) {
            long long fuse = (long long)(target / batch_bytes + 0.5);
            const long long n_batches = (n_images + batch_size - 1) / batch_size;
            if (fuse > n_batches / 16) fuse = n_batches / 16;   // keep at least ~16 chunks in the pipeline
            if (fuse < 1) fuse = 1;
            chunk = (long long)batch_size * fuse;
        } else if (batch_bytes > 2 * target) {
            long long pieces = (long long)(batch_bytes / target + 0.5);
            chunk = (batch_size + pieces - 1) / pieces;
            if (chunk < 1) chunk = 1;
        }
        if (chunk > 0x7fffffffLL) chunk = 0x7fffffffLL;
        batch_size = (int)chunk;
    }
    // Odd widths (width*channels % 16 != 0): rows travel tight over the host link (linear copies at link speed; the
    // copy engines' strided copies manage only 6.5 GB/s on 750-byte rows) and are re-pitched to a multiple of 16 bytes
    // on the device by two small kernels around the blur, so the vectorised kernel runs on any image width.
    const size_t row_bytes = (size_t)width * channels;
    const bool repitch = row_bytes % 16 != 0 && channels <= 4 && row_bytes >= 256;
    const size_t dev_pitch = repitch ? (row_bytes + 15) / 16 * 16 : row_bytes;
    const size_t dev_image_bytes = dev_pitch * (size_t)height;
    if (int rc = ring_prepare(ctx, dev_image_bytes * (size_t)batch_size, repitch ? image_bytes * (size_t)batch_size : 0, n_slots))
        return rc;
    cudaStream_t q_in = ctx->queues[0], q_k = ctx->queues[1], q_out = ctx->queues[2];
    const double t0 = now_ms();
    double ms_in = 0, ms_k = 0, ms_out = 0;
    int64_t launches = 0;
    const int64_t n_chunks = (n_images + batch_size - 1) / batch_size;

    auto harvest = [&](b200blur_ctx::Slot &s) -> int {
        CU_TRY(cudaEventSynchronize(s.ev[5]));
        float f;
        CU_TRY(cudaEventElapsedTime(&f, s.ev[0], s.ev[1])); ms_in += f;
        CU_TRY(cudaEventElapsedTime(&f, s.ev[2], s.ev[3])); ms_k += f;
        CU_TRY(cudaEventElapsedTime(&f, s.ev[4], s.ev[5])); ms_out += f;
        return B200BLUR_OK;
    };

    for (int64_t ci = 0; ci < n_chunks; ci++) {
        b200blur_ctx::Slot &s = ctx->ring[ci % n_slots];
        if (ci >= n_slots)
            if (int rc = harvest(s)) return rc;  // slot's previous chunk fully drained (also frees d_in/d_out)
        const int64_t i0 = ci * batch_size;
        const int64_t n = (n_images - i0 < batch_size) ? n_images - i0 : batch_size;
        const size_t bytes = (size_t)n * image_bytes;
        const uint8_t *src = static_cast<const uint8_t *>(h_in) + (size_t)i0 * image_bytes;
        uint8_t *dst = static_cast<uint8_t *>(h_out) + (size_t)i0 * image_bytes;
        // H2D
        CU_TRY(cudaEventRecord(s.ev[0], q_in));
        if (bytes) CU_TRY(cudaMemcpyAsync(repitch ? s.t_in : s.d_in, src, bytes, cudaMemcpyHostToDevice, q_in));
        CU_TRY(cudaEventRecord(s.ev[1], q_in));
        // blur
        CU_TRY(cudaStreamWaitEvent(q_k, s.ev[1], 0));
        CU_TRY(cudaEventRecord(s.ev[2], q_k));
        b200blur_launch l;
        if (int rc = b200blur_launch_rows_pitched(&l, s.d_in, s.d_out, width, height, channels, 0, height, n,
                                                  dev_image_bytes, dev_image_bytes, repitch ? dev_pitch : 0,
                                                  repitch ? dev_pitch : 0))
            return rc;
        if (repitch && bytes) {
            launch_repitch_in(ctx, q_k, s.t_in, s.d_in, n * (long long)height, (int)row_bytes, (int)dev_pitch);
            launches++;
        }
        int nk;
        if (int rc = do_launch(ctx, 1, &l, &nk)) return rc;
        launches += nk;
        if (repitch && bytes) {
            launch_repitch_out(ctx, q_k, s.d_out, s.t_out, 0, n * (long long)height, (int)row_bytes, (int)dev_pitch);
            launches++;
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(s.ev[3], q_k));
        // D2H
        CU_TRY(cudaStreamWaitEvent(q_out, s.ev[3], 0));
        CU_TRY(cudaEventRecord(s.ev[4], q_out));
        if (bytes) CU_TRY(cudaMemcpyAsync(dst, repitch ? s.t_out : s.d_out, bytes, cudaMemcpyDeviceToHost, q_out));
        CU_TRY(cudaEventRecord(s.ev[5], q_out));
    }
    const int64_t first_pending = n_chunks > n_slots ? n_chunks - n_slots : 0;
    for (int64_t ci = first_pending; ci < n_chunks; ci++)
        if (int rc = harvest(ctx->ring[ci % n_slots])) return rc;
    CU_TRY(cudaStreamSynchronize(q_out));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->wall_ms = now_ms() - t0;
        stats->h2d_ms = ms_in;
        stats->kernel_ms = ms_k;
        stats->d2h_ms = ms_out;
        stats->images = n_images;
        stats->launches = launches;
        stats->h2d_bytes = (int64_t)(image_bytes * (size_t)n_images);
        stats->d2h_bytes = (int64_t)(image_bytes * (size_t)n_images);
    }
    return B200BLUR_OK;
}

// -------------------------------------------------------------------------------- multi-GPU (Approach 2 bands)
int b200blur_peer_enable(b200blur_ctx *a, b200blur_ctx *b)
{
    if (!a || !b) return fail(B200BLUR_ERR_INVALID, "context is NULL");
    if (a->device == b->device) return B200BLUR_OK;
    int ab = 0, ba = 0;
    CU_TRY(cudaDeviceCanAccessPeer(&ab, a->device, b->device));
    CU_TRY(cudaDeviceCanAccessPeer(&ba, b->device, a->device));
    if (!ab || !ba) return fail(B200BLUR_ERR_PEER, "no peer access between devices %d and %d", a->device, b->device);
    CU_TRY(cudaSetDevice(a->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(B200BLUR_ERR_PEER, "%d - cudaDeviceEnablePeerAccess: %s", (int)e, cudaGetErrorString(e));
    cudaGetLastError();
    CU_TRY(cudaSetDevice(b->device));
    e = cudaDeviceEnablePeerAccess(a->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(B200BLUR_ERR_PEER, "%d - cudaDeviceEnablePeerAccess: %s", (int)e, cudaGetErrorString(e));
    cudaGetLastError();
    return B200BLUR_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == B200BLUR_IPC_HANDLE_BYTES, "IPC handle size");

int b200blur_ipc_export(b200blur_ctx *ctx, void *dptr, unsigned char handle[B200BLUR_IPC_HANDLE_BYTES])
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!dptr || !handle) return fail(B200BLUR_ERR_INVALID, "NULL pointer in ipc_export");
    CU_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, sizeof h);
    return B200BLUR_OK;
}

int b200blur_ipc_open(b200blur_ctx *ctx, const unsigned char handle[B200BLUR_IPC_HANDLE_BYTES], void **dptr)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!dptr || !handle) return fail(B200BLUR_ERR_INVALID, "NULL pointer in ipc_open");
    CU_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200BLUR_OK;
}

int b200blur_ipc_close(b200blur_ctx *ctx, void *dptr)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!dptr) return B200BLUR_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaIpcCloseMemHandle(dptr));
    return B200BLUR_OK;
}

}  // extern "C"
