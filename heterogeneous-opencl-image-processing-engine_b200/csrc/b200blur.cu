// b200blur.cu -- implementation of the C ABI declared in include/b200blur.h.
//
// The thin layer that replaces the OpenCL plumbing of heterogeneous_blur.c / split_image_blur.c
// (clCreateContext / clCreateCommandQueue / clCreateBuffer / clEnqueue{Write,NDRange,Read} / clFinish /
// clGetEventProfilingInfo) with the CUDA runtime, plus the two stream engines that replace
// the batch loop of heterogeneous_blur.c:418-600.  CUDA runtime only: no PyTorch, no OpenCL, no CPU fallback.
#include "b200blur.h"

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
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
    std::atomic<bool> in_use{false};
};

// Capacity of a context's event pool.  The pool is ONE fixed array allocated with the context and slots are only ever
// appended (n_events grows, nothing moves), because b200blur_enqueue_wait_peer lets ANOTHER context's host thread read
// a slot while its owner is creating new ones: a growing std::vector would reallocate under that reader.
constexpr int kMaxEvents = 1 << 15;

}  // namespace

struct b200blur_ctx {
    int device = -1;
    int sm_count = 0;
    std::vector<cudaStream_t> queues;
    std::unique_ptr<EventSlot[]> events;   // kMaxEvents slots, fixed storage (see kMaxEvents)
    std::atomic<int> n_events{0};          // slots created so far; published with release after the slot is complete
    std::vector<int> free_events;          // owner thread only
    int kernel_variant = 0;   // 0 auto, 1 register/shuffle strips, 2 TMA-bulk streamed
    int64_t launches = 0;
    // tuning knobs of the streamed kernel (0 = automatic); set from B200BLUR_V2_* at context creation
    int v2_threads = 0, v2_seg = 0, v2_cfg = 0, v2_ctas_per_sm = 0;
    int tail_div = 4;          // guided tail: last groups are this many times finer (B200BLUR_TAIL_DIV, 1 = off)
    double tail_rounds = 2.0;  // ... about this many fine groups per resident CTA (B200BLUR_TAIL_ROUNDS)
    bool use_pdl = true;       // programmatic dependent launch of the streamed kernel (B200BLUR_NO_PDL disables)
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
    struct b200blur_feed *resident_feed = nullptr;   // descriptor table of b200blur_run_resident(coalesce = 0)
    struct b200blur_feed *batches_feed = nullptr;    // descriptor table of b200blur_enqueue_blur_batches
    int64_t batches_calls = 0;
    const void *resident_in = nullptr; void *resident_out = nullptr;   // ... and what it currently describes
    int64_t resident_n = 0, resident_calls = 0;
    cudaEvent_t fork_event = nullptr;          // fork/join of the per-batch launches of b200blur_run_resident
    std::vector<cudaEvent_t> join_events;
    // CUDA graph of the last per-batch launch sequence (launch-bound loop: hundreds of small kernels)
    struct GraphKey {
        const void *in = nullptr; void *out = nullptr;
        int w = 0, h = 0, c = 0, batch = 0; int64_t n = 0;
        int variant = 0;
        bool operator==(const GraphKey &o) const
        { return in == o.in && out == o.out && w == o.w && h == o.h && c == o.c && batch == o.batch && n == o.n && variant == o.variant; }
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

bool event_live(const b200blur_ctx *ctx, b200blur_event ev)
{
    return ev >= 0 && ev < ctx->n_events.load(std::memory_order_acquire) && ctx->events[ev].in_use.load(std::memory_order_acquire);
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
        idx = ctx->n_events.load(std::memory_order_relaxed);
        if (idx >= kMaxEvents) return fail(B200BLUR_ERR_NOMEM, "event pool exhausted (%d live events): release events you no longer need", idx);
        EventSlot &s = ctx->events[idx];
        CU_TRY(cudaEventCreate(&s.start));
        cudaError_t e = cudaEventCreate(&s.end);
        if (e != cudaSuccess) {
            cudaEventDestroy(s.start);
            s.start = nullptr;
            return fail(B200BLUR_ERR_CUDA, "%d - cudaEventCreate: %s", (int)e, cudaGetErrorString(e));
        }
        ctx->n_events.store(idx + 1, std::memory_order_release);
    }
    cudaError_t e = cudaEventRecord(ctx->events[idx].start, ctx->queues[queue]);
    if (e != cudaSuccess) {
        ctx->free_events.push_back(idx);
        return fail(B200BLUR_ERR_CUDA, "%d - cudaEventRecord: %s", (int)e, cudaGetErrorString(e));
    }
    ctx->events[idx].in_use.store(true, std::memory_order_release);
    *slot_out = idx;
    *ev = idx;
    return B200BLUR_OK;
}

// Gives a slot taken by event_begin back when the command it was meant to time could not be enqueued.
int event_abort(b200blur_ctx *ctx, int slot, b200blur_event *ev, int rc)
{
    if (slot >= 0) {
        ctx->events[slot].in_use.store(false, std::memory_order_release);
        ctx->free_events.push_back(slot);
        if (ev) *ev = -1;
    }
    return rc;
}

int event_end(b200blur_ctx *ctx, int queue, int slot, b200blur_event *ev = nullptr)
{
    if (slot < 0) return B200BLUR_OK;
    cudaError_t e = cudaEventRecord(ctx->events[slot].end, ctx->queues[queue]);
    if (e != cudaSuccess)
        return event_abort(ctx, slot, ev, fail(B200BLUR_ERR_CUDA, "%d - cudaEventRecord: %s", (int)e, cudaGetErrorString(e)));
    return B200BLUR_OK;
}

// CU_TRY for a command enqueued between event_begin and event_end: the event slot goes back to the pool on failure.
#define CU_TRY_EV(expr, slot_, ev_)                                                                    \
    do {                                                                                               \
        cudaError_t e2_ = (expr);                                                                      \
        if (e2_ != cudaSuccess)                                                                        \
            return event_abort(ctx, slot_, ev_,                                                        \
                               fail(B200BLUR_ERR_CUDA, "%d - %s failed: %s", (int)e2_, #expr, cudaGetErrorString(e2_))); \
    } while (0)

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

size_t in_pitch_of(const b200blur_launch *l) { return l->in_row_pitch ? l->in_row_pitch : (size_t)l->width * l->channels; }
size_t out_pitch_of(const b200blur_launch *l) { return l->out_row_pitch ? l->out_row_pitch : (size_t)l->width * l->channels; }

// TIGHT input: rows of `in` (and the halo rows) may have any pitch >= width*channels and any alignment -- the streamed
// kernel copies aligned supersets and re-aligns in shared memory -- as long as the OUTPUT side is 16-byte pitched/aligned.
bool launch_tight_input(const b200blur_launch *l)
{
    if (l->channels < 1 || l->channels > 4) return false;
    const size_t row_bytes = (size_t)l->width * l->channels;
    if (row_bytes < 256 || in_pitch_of(l) > 4096) return false;          // whole-row mode only
    if (out_pitch_of(l) % 16 != 0 || !aligned16(l->out)) return false;
    if (l->n_images > 1 && l->out_image_stride % 16) return false;
    return true;
}

// TIGHT input AND output: tight rows (pitch = width*channels) of any alignment on both sides, rows of 256..4096 bytes.
bool launch_tight_inout(const b200blur_launch *l)
{
    if (l->channels < 1 || l->channels > 4) return false;
    const size_t row_bytes = (size_t)l->width * l->channels;
    if (row_bytes < 256 || in_pitch_of(l) > 4096) return false;
    return out_pitch_of(l) == row_bytes;
}

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
    StreamKernel fn_feed[4];  // by channels-1: FEED mode (per-batch descriptors; tight rows that end on a chunk boundary)
    StreamKernel fn_tight[4]; // by channels-1: TIGHT input rows (row pitch = width*channels, any alignment), pitched output
    StreamKernel fn_tight2[4];  // by channels-1: TIGHT input and output rows (store warps)
};

template <int RB, int NS>
constexpr StreamCfg make_cfg()
{
    return StreamCfg{RB, NS,
                     {b200blur::blur_stream_kernel<1, RB, NS, false>, b200blur::blur_stream_kernel<2, RB, NS, false>,
                      b200blur::blur_stream_kernel<3, RB, NS, false>, b200blur::blur_stream_kernel<4, RB, NS, false>},
                     {b200blur::blur_stream_kernel<1, RB, NS, true>, b200blur::blur_stream_kernel<2, RB, NS, true>,
                      b200blur::blur_stream_kernel<3, RB, NS, true>, b200blur::blur_stream_kernel<4, RB, NS, true>},
                     {b200blur::blur_stream_kernel<1, RB, NS, false, true>, b200blur::blur_stream_kernel<2, RB, NS, false, true>,
                      b200blur::blur_stream_kernel<3, RB, NS, false, true>, b200blur::blur_stream_kernel<4, RB, NS, false, true>},
                     {b200blur::blur_stream_kernel<1, RB, NS, true, false, 1>, b200blur::blur_stream_kernel<2, RB, NS, true, false, 1>,
                      b200blur::blur_stream_kernel<3, RB, NS, true, false, 1>, b200blur::blur_stream_kernel<4, RB, NS, true, false, 1>},
                     {b200blur::blur_stream_kernel<1, RB, NS, true, false, 2>, b200blur::blur_stream_kernel<2, RB, NS, true, false, 2>,
                      b200blur::blur_stream_kernel<3, RB, NS, true, false, 2>, b200blur::blur_stream_kernel<4, RB, NS, true, false, 2>}};
}

// {rows per slot, slots}; index 0 is the default (B200BLUR_V2_CFG selects another for tuning runs)
// (round-1 sweep over {8,4} {8,3} {4,3} {4,4} {4,6} {16,2} {8,2}: all within 3 % once work is handed out dynamically)
const StreamCfg kStreamCfgs[] = {make_cfg<8, 4>(), make_cfg<8, 3>(), make_cfg<8, 6>()};
constexpr int kNumStreamCfgs = sizeof(kStreamCfgs) / sizeof(kStreamCfgs[0]);

// Whether the streamed (variant 2) kernel can run this launch: rows wide enough for bulk copies to pay.
bool stream_eligible(const b200blur::BandParams &p) { return p.row_bytes >= 256; }

// PRMT selectors that apply the right-edge clamp to the {wl, w, wr} window of the chunk holding the end of a row whose
// length is not a multiple of 16 (see StreamParams): window byte idx takes the byte C positions earlier when it is one
// of the C bytes just past the end of the row.  Selector nibbles index the 8 bytes of (previous word, this word).
void edge_selectors(b200blur::StreamParams &sp, int row_bytes, int channels, bool force_general = false)
{
    const int v = row_bytes - (sp.cpr - 1) * 16;  // bytes of the row inside its last chunk, 1..16
    sp.edge_general = (v != 16) || force_general;   // (TIGHT kernels are compiled with the general right-edge code only)
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

// Everything a streamed-kernel launch needs besides the stream: geometry, kernel, block/grid/shared memory.
struct StreamPlan {
    b200blur::StreamParams sp;
    StreamKernel fn = nullptr;
    int block = 0;
    size_t smem = 0;
    long long slots = 0;   // resident CTAs on the whole GPU for this kernel/block/smem
};

// Plans the streamed kernel for band geometry `p` (p.n_images images per launch, or -- feed -- per batch at most).
// `slots_override` > 0: plan for that many resident CTAs without asking the CUDA runtime (host-side introspection).
int plan_stream(b200blur_ctx *ctx, const b200blur::BandParams &p, bool feed, StreamPlan &plan, long long slots_override = 0,
                int tight = 0)   // tight: 0 pitched rows, 1 tight input, 2 tight input and output
{
    b200blur::StreamParams &sp = plan.sp;
    memset(&sp, 0, sizeof sp);
    sp.b = p;
    sp.cpr = (p.row_bytes + 15) / 16;   // live chunks per row; bytes past row_bytes up to the pitch are padding
    edge_selectors(sp, p.row_bytes, p.channels, tight != 0);
    // (tight output stages a second copy of every slot in shared memory: a 3-slot ring keeps 3 CTAs per SM)
    const StreamCfg &cfg = kStreamCfgs[tight == 2 ? 1 : (ctx->v2_cfg >= 0 && ctx->v2_cfg < kNumStreamCfgs) ? ctx->v2_cfg : 0];
    int threads;
    if (p.pitch <= 4096 && sp.cpr <= 256) {
        // whole rows: a CTA step covers `ipc` images side by side; pick the block size that wastes fewest lanes
        sp.cb = sp.cpr;
        sp.ncb = 1;
        const int prefer = ctx->v2_threads > 0 ? ctx->v2_threads : 128;
        int best_t = 0;
        double best_score = -1.0;
        for (int t = 64; t <= 256; t += 32) {
            if (t < sp.cb) continue;
            int ipc = t / sp.cb;
            if (ipc > p.n_images && p.n_images > 0) ipc = (int)p.n_images;   // never more image lanes than images
            const double eff = (double)(ipc * sp.cb) / t;
            const double score = eff - 0.0005 * (t > prefer ? t - prefer : prefer - t);
            if (score > best_score) { best_score = score; best_t = t; }
        }
        threads = best_t;
        sp.ipc = threads / sp.cb;
        if (sp.ipc > p.n_images && p.n_images > 0) sp.ipc = (int)p.n_images;
        // Rows as they lie in memory (RB rows of an image = one bulk copy, padding included) while the padding is small;
        // heavily padded rows (ROI views, pitch >> width*channels) bring only their live chunks, one copy per row.
        sp.margin = (!tight && (long long)p.pitch - sp.cpr * 16 > sp.cpr * 4) ? 16 : 0;
        if (tight && sp.ipc > b200blur::kTightMaxLanes) sp.ipc = b200blur::kTightMaxLanes;
    } else {
        threads = ctx->v2_threads > 0 ? ctx->v2_threads : 128;
        if (threads > 256) threads = 256;
        sp.cb = threads;
        sp.ncb = (sp.cpr + sp.cb - 1) / sp.cb;
        sp.margin = 16;
        sp.ipc = 1;
    }
    // TIGHT: an image lane of a slot = 16 B lead pad + up to three runs of rows, each widened to 16-byte boundaries
    sp.lane_bytes = tight ? (cfg.rb * p.pitch + 15) / 16 * 16 + 128 : 0;
    sp.stage_pitch = tight == 2 ? (p.row_bytes + 15) / 16 * 16 : 0;
    auto smem_for = [&](int ipc) {
        const int sstride = sp.margin ? sp.cb * 16 + 2 * sp.margin : p.pitch;
        const size_t slot = tight ? (size_t)ipc * sp.lane_bytes : (size_t)ipc * cfg.rb * sstride;
        return 16 + (size_t)cfg.ns * slot + 16 + 16 * cfg.ns + sizeof(b200blur::GroupMeta) * cfg.ns +
               (feed ? 24 * b200blur::kFeedDepth : 0) + (tight ? 2 * cfg.ns * b200blur::kTightMaxLanes * cfg.rb : 0) +
               (tight == 2 ? 16 * b200blur::kStageSlots + sizeof(b200blur::StageRec) * b200blur::kStageRecs + 16 +
                                 (size_t)b200blur::kStageSlots * ipc * cfg.rb * sp.stage_pitch + 32
                           : 0);
    };
    while (sp.ipc > 1 && smem_for(sp.ipc) > 200 * 1024) sp.ipc--;   // fewer images side by side rather than no launch
    sp.sstride = sp.margin ? sp.cb * 16 + 2 * sp.margin : p.pitch;   // whole rows land exactly as they lie in memory
    sp.slot_bytes = tight ? sp.ipc * sp.lane_bytes : sp.ipc * cfg.rb * sp.sstride;
    sp.stage_slot_bytes = sp.ipc * cfg.rb * sp.stage_pitch;
    sp.row_recip = p.row_bytes > 1 ? (unsigned)((0x100000000ULL + (unsigned)p.row_bytes - 1) / (unsigned)p.row_bytes) : 0u;
    plan.smem = smem_for(sp.ipc);
    plan.block = threads + (feed ? 64 : tight == 2 ? 32 + 32 * b200blur::kStoreWarps : 32);  // + producer (+ accountant / store warps)
    if (plan.smem > 220 * 1024) return fail(B200BLUR_ERR_INVALID, "streamed kernel needs %zu B of shared memory", plan.smem);
    plan.fn = tight == 2 ? cfg.fn_tight2[p.channels - 1] : tight ? cfg.fn_tight[p.channels - 1]
              : feed ? cfg.fn_feed[p.channels - 1] : sp.edge_general ? cfg.fn_edge[p.channels - 1] : cfg.fn[p.channels - 1];
    int per_sm = slots_override > 0 ? 1 : 0;
    for (auto &ki : ctx->kernel_info)
        if (ki.fn == (const void *)plan.fn && ki.block == plan.block && ki.smem == plan.smem) per_sm = ki.per_sm;
    if (per_sm == 0) {
        // opt in to the largest dynamic shared-memory size once per kernel (any later, smaller request is covered)
        CU_TRY(cudaFuncSetAttribute(plan.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, plan.fn, plan.block, plan.smem));
        if (per_sm < 1) return fail(B200BLUR_ERR_CUDA, "streamed kernel does not fit on an SM");
        ctx->kernel_info.push_back({(const void *)plan.fn, plan.block, plan.smem, per_sm});
    }
    if (ctx->v2_ctas_per_sm > 0 && per_sm > ctx->v2_ctas_per_sm) per_sm = ctx->v2_ctas_per_sm;
    plan.slots = slots_override > 0 ? slots_override : (long long)ctx->sm_count * per_sm;
    sp.img_blocks = (p.n_images + sp.ipc - 1) / sp.ipc;
    // Work unit = `seg` output rows of `ipc` images (or of one column block): ~48 KB in + 48 KB out for whole
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
        if (!feed)
            while (want > 6 && sp.img_blocks * sp.ncb * ((p.rows + want - 1) / want) < ctx->sm_count) want = (want + 1) / 2;
        // The consumers run whole ring slots (RB input rows) fully unrolled; a group of `seg` output rows reads seg + 2
        // input rows, so seg + 2 = a multiple of RB keeps every slot of a full-height group whole.  Among the segment
        // heights near `want`, take the one that processes the fewest input rows (2 halo rows per group) and leaves
        // the fewest partly filled slots.
        long long best = -1, best_cost = 0;
        const long long lo = want * 3 / 4 > 6 ? want * 3 / 4 : 6, hi = want * 4 / 3 + 1;
        for (long long sg = lo; sg <= hi && sg <= p.rows; sg++) {
            const long long full_groups = p.rows / sg, rest = p.rows - full_groups * sg;
            long long rows_in = full_groups * (sg + 2) + (rest ? rest + 2 : 0);
            long long partial = full_groups * ((sg + 2) % cfg.rb ? 1 : 0) + (rest && (rest + 2) % cfg.rb ? 1 : 0);
            const long long cost = rows_in * 16 + partial * 48 + (sg > want ? sg - want : want - sg);
            if (best < 0 || cost < best_cost) { best = sg; best_cost = cost; }
        }
        seg = best > 0 ? (int)best : (int)(want < p.rows ? want : p.rows);
    }
    if (seg > p.rows) seg = p.rows;
    sp.seg = seg;
    sp.nseg = (p.rows + seg - 1) / seg;
    // Guided tail: the image blocks handed out last are cut `tail_div` times finer, about `tail_rounds` groups per
    // resident CTA of them, so the CTAs finish within a fraction of a coarse group of each other.  (Launches shorter
    // than a few rounds of coarse groups are all tail, e.g. a 32-row band of 5000 images on one of 8 GPUs.)
    sp.ib_coarse = sp.img_blocks;
    sp.seg_fine = seg;
    sp.nseg_fine = sp.nseg;
    if (!feed && ctx->tail_div > 1 && ctx->v2_seg <= 0 && sp.img_blocks * sp.nseg * sp.ncb >= 4 * plan.slots) {
        int seg_fine = (seg + ctx->tail_div - 1) / ctx->tail_div;
        if (seg_fine < 6) seg_fine = seg < 6 ? seg : 6;
        if (seg_fine < seg) {
            const long long nseg_fine = (p.rows + seg_fine - 1) / seg_fine;
            seg_fine = (int)((p.rows + nseg_fine - 1) / nseg_fine);
            const long long per_block_fine = nseg_fine * sp.ncb;
            long long ib_fine = (long long)(ctx->tail_rounds * (double)plan.slots / (double)per_block_fine + 0.999);
            if (ib_fine > sp.img_blocks) ib_fine = sp.img_blocks;
            if (ib_fine > 0) {
                sp.ib_coarse = sp.img_blocks - ib_fine;
                sp.seg_fine = seg_fine;
                sp.nseg_fine = (int)((p.rows + seg_fine - 1) / seg_fine);
            }
        }
    }
    sp.g_coarse = sp.ib_coarse * sp.nseg * sp.ncb;
    sp.n_groups = sp.g_coarse + (sp.img_blocks - sp.ib_coarse) * sp.nseg_fine * sp.ncb;
    if (sp.n_groups >= 0x7fffffffLL || (long long)sp.nseg * sp.ncb >= 0x7fffffffLL || (long long)sp.nseg_fine * sp.ncb >= 0x7fffffffLL)
        return fail(B200BLUR_ERR_INVALID, "too many work units for one launch (%lld)", sp.n_groups);
    return B200BLUR_OK;
}

// Launches a planned streamed kernel with programmatic stream serialisation: when the previous command in the stream
// is also one of these kernels, this one's prologue (CTA launch, barrier init) overlaps that one's tail; the kernel's
// griddepcontrol.wait keeps every global access after the previous kernel's completion.
int launch_planned(b200blur_ctx *ctx, const StreamPlan &plan, long long grid, cudaStream_t s)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)plan.block);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->use_pdl ? 1 : 0;
    CU_TRY(cudaLaunchKernelEx(&cfg, plan.fn, plan.sp));
    return B200BLUR_OK;
}

int launch_stream(b200blur_ctx *ctx, const b200blur::BandParams &p, cudaStream_t s, int queue, int tight = 0)
{
    StreamPlan plan;
    if (int rc = plan_stream(ctx, p, false, plan, 0, tight)) return rc;
    plan.sp.work = ctx->d_work + 2 * queue;
    const long long grid = plan.sp.n_groups < plan.slots ? plan.sp.n_groups : plan.slots;
    return launch_planned(ctx, plan, grid, s);
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
    } else if (!vec && launch_tight_input(l) && ctx->kernel_variant != 1) {
        if (int rc = launch_stream(ctx, p, s, queue, 1)) return rc;
        ++*n_kernels;
    } else if (!vec && launch_tight_inout(l) && ctx->kernel_variant != 1 && !getenv("B200BLUR_NO_TIGHT_OUT")) {
        if (int rc = launch_stream(ctx, p, s, queue, 2)) return rc;
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
    const long long per_block_rows = b200blur::kRepitchRows;
    const long long total = per_block_rows * ((row_bytes + 15) / 16);
    long long bx = (total + 255) / 256;
    if (bx > 64) bx = 64;
    const long long by = (rows + per_block_rows - 1) / per_block_rows;
    for (long long y0 = 0; y0 < by; y0 += 65535) {   // grid.y limit
        const long long ny = by - y0 < 65535 ? by - y0 : 65535;
        const long long r0 = y0 * per_block_rows;
        b200blur::repitch_in_kernel<<<dim3((unsigned)bx, (unsigned)ny), 256, 0, s>>>(
            static_cast<const uint8_t *>(tight) + r0 * row_bytes, static_cast<uint8_t *>(pitched) + r0 * pitch, rows - r0, row_bytes, pitch);
        ctx->launches++;
    }
}

// `tight_base` is 16-byte aligned; the rows land at byte offset `lo` of it
void launch_repitch_out(b200blur_ctx *ctx, cudaStream_t s, const void *pitched, void *tight_base, long long lo, long long rows,
                        int row_bytes, int pitch)
{
    const long long per_block_rows = b200blur::kRepitchRows;
    const long long total = (per_block_rows * row_bytes + 15) / 16 + 1;
    long long bx = (total + 255) / 256;
    if (bx > 64) bx = 64;
    const long long by = (rows + per_block_rows - 1) / per_block_rows;
    for (long long y0 = 0; y0 < by; y0 += 65535) {
        const long long ny = by - y0 < 65535 ? by - y0 : 65535;
        const long long r0 = y0 * per_block_rows;
        b200blur::repitch_out_kernel<<<dim3((unsigned)bx, (unsigned)ny), 256, 0, s>>>(
            static_cast<const uint8_t *>(pitched) + r0 * pitch, static_cast<uint8_t *>(tight_base), lo + r0 * row_bytes, rows - r0,
            row_bytes, pitch);
        ctx->launches++;
    }
}

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// A feed: the persistent form of the batch loop.  See include/b200blur.h (b200blur_feed_*) and StreamParams.
struct b200blur_feed {
    b200blur_ctx *ctx = nullptr;
    int width = 0, height = 0, channels = 0, max_batch = 0, cap = 0;
    size_t image_bytes = 0;
    size_t in_stride = 0, out_stride = 0, top_stride = 0, bot_stride = 0;   // image strides shared by all batches
    cudaEvent_t table_copied = nullptr;      // table mode: the last copy of the pinned mirror to the device has run
    cudaEvent_t table_kernel = nullptr;      // table mode: the last kernel reading the device table / counters has finished
    StreamPlan plan;
    b200blur::FeedBatch *d_batches = nullptr, *h_batches = nullptr;   // device ring and its pinned host mirror
    unsigned long long *d_ctl = nullptr;     // [0] tail, [1] closed, [2] watchdog, [3] unused; [4],[5] work counters
    unsigned int *d_count = nullptr;         // per slot: consumer-warp arrivals of the batch in it
    unsigned int *h_done = nullptr;          // cap + 1 words, host-mapped: per slot completion, [cap] = watchdog tripped
    unsigned long long *h_ctl = nullptr;     // pinned ring of control words (sources of the small async copies)
    int h_ctl_pos = 0;
    cudaStream_t kstream = nullptr, cstream = nullptr;
    int64_t submitted = 0, flushed = 0, base = 0;
    bool running = false, failed = false;
    unsigned long long timeout_ns = 5000000000ull;
};

namespace {
constexpr int kFeedCtlRing = 256;

void feed_release(b200blur_feed *f)
{
    if (!f) return;
    if (f->d_batches) cudaFree(f->d_batches);
    if (f->d_ctl) cudaFree(f->d_ctl);
    if (f->d_count) cudaFree(f->d_count);
    if (f->h_batches) cudaFreeHost(f->h_batches);
    if (f->h_done) cudaFreeHost(f->h_done);
    if (f->h_ctl) cudaFreeHost(f->h_ctl);
    if (f->table_copied) cudaEventDestroy(f->table_copied);
    if (f->table_kernel) cudaEventDestroy(f->table_kernel);
    if (f->kstream) cudaStreamDestroy(f->kstream);
    if (f->cstream) cudaStreamDestroy(f->cstream);
    delete f;
}

// Can the feed kernel run this geometry?  (tight rows ending on a 16-byte boundary, wide enough for bulk copies)
bool feed_eligible(int width, int height, int channels)
{
    const size_t row_bytes = (size_t)width * channels;
    return channels >= 1 && channels <= 4 && height >= 1 && row_bytes >= 256 && row_bytes % 16 == 0 && row_bytes <= 0x7fffffffULL;
}

int feed_build(b200blur_ctx *ctx, int width, int height, int channels, int max_batch, int cap, bool own_streams, b200blur_feed **out,
               size_t in_stride = 0, size_t out_stride = 0, size_t top_stride = 0, size_t bot_stride = 0)
{
    *out = nullptr;
    if (!feed_eligible(width, height, channels))
        return fail(B200BLUR_ERR_INVALID, "feed needs channels <= 4 and rows of width*channels >= 256 bytes, a multiple of 16 (got %dx%dx%d)",
                    width, height, channels);
    if (max_batch < 1 || cap < 2) return fail(B200BLUR_ERR_INVALID, "feed needs max_batch >= 1 and capacity >= 2");
    CU_TRY(cudaSetDevice(ctx->device));
    b200blur_feed *f = new (std::nothrow) b200blur_feed;
    if (!f) return fail(B200BLUR_ERR_NOMEM, "out of host memory");
    f->ctx = ctx;
    f->width = width; f->height = height; f->channels = channels; f->max_batch = max_batch; f->cap = cap;
    f->image_bytes = (size_t)width * height * channels;
    if (const char *v = getenv("B200BLUR_FEED_TIMEOUT_MS")) f->timeout_ns = (unsigned long long)atoll(v) * 1000000ull;
    b200blur::BandParams p;
    memset(&p, 0, sizeof p);
    f->in_stride = in_stride ? in_stride : f->image_bytes;
    f->out_stride = out_stride ? out_stride : f->image_bytes;
    f->top_stride = top_stride;
    f->bot_stride = bot_stride;
    p.in_stride = f->in_stride;
    p.out_stride = f->out_stride;
    p.top_stride = top_stride;
    p.bot_stride = bot_stride;
    p.row_bytes = p.pitch = p.out_pitch = width * channels;
    p.rows = height;
    p.width = width;
    p.channels = channels;
    p.n_images = max_batch;
    int rc = plan_stream(ctx, p, true, f->plan);
    auto bail = [&](int code) { feed_release(f); return code; };
    if (rc) return bail(rc);
    b200blur::StreamParams &sp = f->plan.sp;
    const long long gpb = sp.img_blocks * sp.nseg * sp.ncb;
    if (gpb >= (1LL << 24)) return bail(fail(B200BLUR_ERR_INVALID, "batch too large for a feed (%lld groups)", gpb));
#define FEED_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return bail(fail(e_ == cudaErrorMemoryAllocation ? B200BLUR_ERR_NOMEM : B200BLUR_ERR_CUDA, "%d - %s failed: %s", (int)e_, #expr, cudaGetErrorString(e_))); } while (0)
    FEED_TRY(cudaMalloc((void **)&f->d_batches, sizeof(b200blur::FeedBatch) * cap));
    FEED_TRY(cudaMalloc((void **)&f->d_ctl, 8 * 8));
    FEED_TRY(cudaMemset(f->d_ctl, 0, 8 * 8));
    FEED_TRY(cudaMalloc((void **)&f->d_count, sizeof(unsigned int) * cap));
    FEED_TRY(cudaMemset(f->d_count, 0, sizeof(unsigned int) * cap));
    FEED_TRY(cudaHostAlloc((void **)&f->h_batches, sizeof(b200blur::FeedBatch) * cap, cudaHostAllocPortable));
    FEED_TRY(cudaHostAlloc((void **)&f->h_done, sizeof(unsigned int) * (cap + 1), cudaHostAllocPortable | cudaHostAllocMapped));
    memset(f->h_done, 0, sizeof(unsigned int) * (cap + 1));
    FEED_TRY(cudaHostAlloc((void **)&f->h_ctl, 8 * kFeedCtlRing, cudaHostAllocPortable));
    if (own_streams) {
        FEED_TRY(cudaStreamCreateWithFlags(&f->kstream, cudaStreamNonBlocking));
        FEED_TRY(cudaStreamCreateWithFlags(&f->cstream, cudaStreamNonBlocking));
    }
    void *d_done = nullptr;
    FEED_TRY(cudaHostGetDevicePointer(&d_done, f->h_done, 0));
    FEED_TRY(cudaDeviceSynchronize());   // the memsets above are on the legacy stream
#undef FEED_TRY
    sp.batches = f->d_batches;
    sp.feed_ctl = f->d_ctl;
    sp.work = f->d_ctl + 4;
    sp.feed_count = f->d_count;
    sp.feed_done = static_cast<volatile unsigned int *>(d_done);
    sp.feed_cap = cap;
    sp.feed_gpb = (int)gpb;
    sp.feed_target = (unsigned int)gpb * (unsigned int)(f->plan.block / 32 - 2);   // consumer warps per CTA
    sp.feed_timeout_ns = f->timeout_ns;
    *out = f;
    return B200BLUR_OK;
}

// next pinned control word holding `value` (the source of a small async copy must stay untouched until the copy runs)
int feed_ctl_word(b200blur_feed *f, cudaStream_t s, unsigned long long value, unsigned long long **word)
{
    if (f->h_ctl_pos == kFeedCtlRing) {
        CU_TRY(cudaStreamSynchronize(s));
        f->h_ctl_pos = 0;
    }
    f->h_ctl[f->h_ctl_pos] = value;
    *word = f->h_ctl + f->h_ctl_pos++;
    return B200BLUR_OK;
}

// publishes descriptors [flushed, submitted) and the new tail on stream s
int feed_publish(b200blur_feed *f, cudaStream_t s)
{
    if (f->flushed == f->submitted) return B200BLUR_OK;
    int64_t i = f->flushed;
    while (i < f->submitted) {
        const int slot = (int)(i % f->cap);
        int64_t n = f->submitted - i;
        if (n > f->cap - slot) n = f->cap - slot;
        CU_TRY(cudaMemcpyAsync(f->d_batches + slot, f->h_batches + slot, sizeof(b200blur::FeedBatch) * (size_t)n, cudaMemcpyHostToDevice, s));
        i += n;
    }
    unsigned long long *w;
    if (int rc = feed_ctl_word(f, s, (unsigned long long)(f->submitted - f->base), &w)) return rc;
    CU_TRY(cudaMemcpyAsync(f->d_ctl, w, 8, cudaMemcpyHostToDevice, s));
    f->flushed = f->submitted;
    return B200BLUR_OK;
}

bool feed_slot_done(const b200blur_feed *f, int64_t ticket)
{
    const unsigned int seen = *(volatile unsigned int *)(f->h_done + ticket % f->cap);
    return (int)(seen - (unsigned int)(ticket + 1)) >= 0 && seen != 0;
}

int feed_launch(b200blur_feed *f, cudaStream_t s, long long total_groups_or_0)
{
    b200blur_ctx *ctx = f->ctx;
    f->plan.sp.feed_base = (unsigned long long)f->base;
    long long grid = f->plan.slots;
    if (total_groups_or_0 > 0 && total_groups_or_0 < grid) grid = total_groups_or_0;
    if (int rc = launch_planned(ctx, f->plan, grid, s)) return rc;
    ctx->launches++;
    return B200BLUR_OK;
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
    ctx->events.reset(new (std::nothrow) EventSlot[kMaxEvents]);
    if (!ctx->events) {
        delete ctx;
        return fail(B200BLUR_ERR_NOMEM, "out of host memory");
    }
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
    if (getenv("B200BLUR_TAIL_DIV")) ctx->tail_div = env_int("B200BLUR_TAIL_DIV");
    if (getenv("B200BLUR_TAIL_ROUNDS")) ctx->tail_rounds = atof(getenv("B200BLUR_TAIL_ROUNDS"));
    ctx->use_pdl = getenv("B200BLUR_NO_PDL") == nullptr;
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
    if (ctx->resident_feed) feed_release(ctx->resident_feed);
    if (ctx->batches_feed) feed_release(ctx->batches_feed);
    if (ctx->d_work) cudaFree(ctx->d_work);
    if (ctx->scratch_in) cudaFree(ctx->scratch_in);
    if (ctx->scratch_out) cudaFree(ctx->scratch_out);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    for (auto e : ctx->join_events)
        if (e) cudaEventDestroy(e);
    for (int i = 0, n = ctx->n_events.load(); i < n; i++) {
        if (ctx->events[i].start) cudaEventDestroy(ctx->events[i].start);
        if (ctx->events[i].end) cudaEventDestroy(ctx->events[i].end);
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
    if (!event_live(ctx, ev))
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
    if (!event_live(ctx, ev))
        return fail(B200BLUR_ERR_INVALID, "event %d is not live", (int)ev);
    ctx->events[ev].in_use.store(false, std::memory_order_release);
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
    return event_end(ctx, queue, slot, ev);
}

int b200blur_events_elapsed_ms(b200blur_ctx *ctx, b200blur_event from, b200blur_event to, double *ms)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!ms) return fail(B200BLUR_ERR_INVALID, "ms is NULL");
    for (b200blur_event ev : {from, to})
        if (!event_live(ctx, ev))
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
    if (!event_live(ctx, ev))
        return fail(B200BLUR_ERR_INVALID, "event %d is not live", (int)ev);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaStreamWaitEvent(ctx->queues[queue], ctx->events[ev].end, 0));
    return B200BLUR_OK;
}

int b200blur_enqueue_wait_peer(b200blur_ctx *ctx, int queue, b200blur_ctx *src, b200blur_event ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (int rc = ctx_check(src)) return rc;
    if (!event_live(src, ev))
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
    if (bytes) CU_TRY_EV(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->queues[queue]), slot, ev);
    return event_end(ctx, queue, slot, ev);
}

int b200blur_enqueue_read(b200blur_ctx *ctx, int queue, void *dst_host, const void *src_dev, size_t bytes,
                          b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (bytes && (!dst_host || !src_dev)) return fail(B200BLUR_ERR_INVALID, "NULL pointer in enqueue_read");
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    if (bytes) CU_TRY_EV(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->queues[queue]), slot, ev);
    return event_end(ctx, queue, slot, ev);
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
        CU_TRY_EV(cudaMemcpy2DAsync(dst_dev, dst_pitch, src_host, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice,
                                    ctx->queues[queue]), slot, ev);
    return event_end(ctx, queue, slot, ev);
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
        CU_TRY_EV(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost,
                                    ctx->queues[queue]), slot, ev);
    return event_end(ctx, queue, slot, ev);
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

int b200blur_plan_groups(int width, int rows, int channels, int64_t n_images, size_t row_pitch, int resident_ctas, int feed,
                         int64_t out[16])
{
    if (!out || width < 1 || rows < 1 || channels < 1 || channels > 4 || n_images < 1 || resident_ctas < 1)
        return fail(B200BLUR_ERR_INVALID, "bad plan request");
    const size_t row_bytes = (size_t)width * channels;
    if (row_pitch == 0) row_pitch = row_bytes;
    if (row_pitch % 16 || row_pitch < row_bytes || row_pitch > 0x7fffffffULL || row_bytes < 256)
        return fail(B200BLUR_ERR_INVALID, "the streamed kernel needs rows of >= 256 bytes and a row pitch that is a multiple of 16");
    b200blur_ctx tmp;   // knobs at their defaults; no device is touched
    tmp.sm_count = 148;
    b200blur::BandParams p;
    memset(&p, 0, sizeof p);
    p.in_stride = p.out_stride = row_pitch * (size_t)rows;
    p.row_bytes = (int)row_bytes;
    p.pitch = p.out_pitch = (int)row_pitch;
    p.rows = rows;
    p.width = width;
    p.channels = channels;
    p.n_images = n_images;
    StreamPlan plan;
    if (int rc = plan_stream(&tmp, p, feed != 0, plan, resident_ctas, 0)) return rc;
    const b200blur::StreamParams &sp = plan.sp;
    const int64_t v[16] = {sp.cpr, sp.cb, sp.ncb, sp.ipc, sp.seg, sp.nseg, sp.seg_fine, sp.nseg_fine, sp.img_blocks, sp.ib_coarse,
                           sp.g_coarse, sp.n_groups, sp.margin, plan.block, (int64_t)plan.smem, sp.edge_general};
    memcpy(out, v, sizeof v);
    return B200BLUR_OK;
}

int b200blur_launch_is_vectorised(const b200blur_launch *launch)
{
    if (!launch) return 0;
    return (launch_vectorised(launch) || launch_tight_input(launch) || launch_tight_inout(launch)) ? 1 : 0;
}

int b200blur_enqueue_blur(b200blur_ctx *ctx, int queue, const b200blur_launch *launch, b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (int rc = launch_validate(launch)) return rc;
    CU_TRY(cudaSetDevice(ctx->device));
    int slot;
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    int nk;
    if (int rc = do_launch(ctx, queue, launch, &nk)) return event_abort(ctx, slot, ev, rc);
    return event_end(ctx, queue, slot, ev);
}

int b200blur_enqueue_blur_batches(b200blur_ctx *ctx, int queue, const b200blur_launch *launches, int n_launches, b200blur_event *ev)
{
    if (int rc = queue_check(ctx, queue)) return rc;
    if (n_launches < 0 || (n_launches > 0 && !launches)) return fail(B200BLUR_ERR_INVALID, "bad batch list");
    for (int i = 0; i < n_launches; i++)
        if (int rc = launch_validate(&launches[i])) return rc;
    CU_TRY(cudaSetDevice(ctx->device));
    int slot = -1;
    // One kernel launch for all batches when they share one geometry the feed kernel can run: tight 16-byte-multiple rows,
    // aligned pointers, same width / rows / channels / strides; each batch brings its own pointers (halo rows included)
    // and image count.  Anything else is enqueued launch by launch -- same results.
    bool same = n_launches >= 2 && ctx->kernel_variant != 1 && getenv("B200BLUR_NO_FEED") == nullptr;
    int64_t max_n = 0;
    if (same) {
        const b200blur_launch &a = launches[0];
        same = feed_eligible(a.width, a.rows, a.channels) && a.in_row_pitch == 0 && a.out_row_pitch == 0;
        for (int i = 0; same && i < n_launches; i++) {
            const b200blur_launch &l = launches[i];
            same = l.width == a.width && l.rows == a.rows && l.channels == a.channels && l.in_row_pitch == 0 && l.out_row_pitch == 0 &&
                   l.in_image_stride == a.in_image_stride && l.out_image_stride == a.out_image_stride && l.n_images >= 1 &&
                   l.n_images <= 0x7fffffff && launch_vectorised(&l) &&
                   (!l.halo_top || l.halo_top_stride == (a.halo_top ? a.halo_top_stride : l.halo_top_stride)) &&
                   (!l.halo_bottom || l.halo_bottom_stride == (a.halo_bottom ? a.halo_bottom_stride : l.halo_bottom_stride)) &&
                   (l.halo_top != nullptr) == (a.halo_top != nullptr) && (l.halo_bottom != nullptr) == (a.halo_bottom != nullptr);
            if (l.n_images > max_n) max_n = l.n_images;
        }
    }
    if (!same) {
        if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
        for (int i = 0; i < n_launches; i++) {
            int nk;
            if (int rc = do_launch(ctx, queue, &launches[i], &nk)) return event_abort(ctx, slot, ev, rc);
        }
        return event_end(ctx, queue, slot, ev);
    }
    const b200blur_launch &a = launches[0];
    cudaStream_t s = ctx->queues[queue];
    b200blur_feed *f = ctx->batches_feed;
    const size_t top_stride = a.halo_top ? a.halo_top_stride : 0, bot_stride = a.halo_bottom ? a.halo_bottom_stride : 0;
    if (!f || f->width != a.width || f->height != a.rows || f->channels != a.channels || f->max_batch != (int)max_n ||
        f->cap != n_launches || f->in_stride != a.in_image_stride || f->out_stride != a.out_image_stride ||
        f->top_stride != top_stride || f->bot_stride != bot_stride) {
        if (f) {
            for (auto q : ctx->queues) cudaStreamSynchronize(q);   // (the old table may be in use on any queue)
            feed_release(f);
            ctx->batches_feed = nullptr;
        }
        if (int rc = feed_build(ctx, a.width, a.rows, a.channels, (int)max_n, n_launches, false, &f, a.in_image_stride,
                                a.out_image_stride, top_stride, bot_stride))
            return rc;
        ctx->batches_feed = f;
    }
    const long long total_groups = (long long)n_launches * f->plan.sp.feed_gpb;
    if (total_groups >= 0x7fffffffLL) return fail(B200BLUR_ERR_INVALID, "too many work units in one batched enqueue");
    // the pinned mirror may still be the source of the previous call's copy; the device table and its counters may still
    // be in use by the previous call's kernel on ANOTHER queue (on the same queue stream order already covers it)
    if (!f->table_copied) {
        CU_TRY(cudaEventCreateWithFlags(&f->table_copied, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&f->table_kernel, cudaEventDisableTiming));
    } else {
        CU_TRY(cudaEventSynchronize(f->table_copied));
        CU_TRY(cudaStreamWaitEvent(s, f->table_kernel, 0));
    }
    for (int i = 0; i < n_launches; i++) {
        b200blur::FeedBatch &d = f->h_batches[i];
        memset(&d, 0, sizeof d);
        d.in = static_cast<const uint8_t *>(launches[i].in);
        d.out = static_cast<uint8_t *>(launches[i].out);
        d.top = static_cast<const uint8_t *>(launches[i].halo_top);
        d.bot = static_cast<const uint8_t *>(launches[i].halo_bottom);
        d.n_images = (int)launches[i].n_images;
    }
    f->h_ctl[0] = (unsigned long long)n_launches;
    f->h_ctl[1] = 1ull;
    // (the table is built before the event starts: `ev` times the copies of the descriptors and the kernel only)
    if (int rc = event_begin(ctx, queue, ev, &slot)) return rc;
    CU_TRY_EV(cudaMemcpyAsync(f->d_batches, f->h_batches, sizeof(b200blur::FeedBatch) * (size_t)n_launches, cudaMemcpyHostToDevice, s), slot, ev);
    CU_TRY_EV(cudaMemcpyAsync(f->d_ctl, f->h_ctl, 16, cudaMemcpyHostToDevice, s), slot, ev);
    CU_TRY_EV(cudaEventRecord(f->table_copied, s), slot, ev);
    f->base = ctx->batches_calls++ * (int64_t)n_launches;
    if (int rc = feed_launch(f, s, total_groups)) return event_abort(ctx, slot, ev, rc);
    CU_TRY_EV(cudaEventRecord(f->table_kernel, s), slot, ev);
    return event_end(ctx, queue, slot, ev);
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


// ------------------------------------------------------------------------------------------------ feed (L3)
int b200blur_feed_create(b200blur_ctx *ctx, int width, int height, int channels, int max_batch_images, int capacity,
                         b200blur_feed **feed)
{
    if (int rc = ctx_check(ctx)) return rc;
    if (!feed) return fail(B200BLUR_ERR_INVALID, "feed out-pointer is NULL");
    if (capacity <= 0) capacity = 4096;
    return feed_build(ctx, width, height, channels, max_batch_images, capacity, true, feed);
}

int b200blur_feed_start(b200blur_feed *f)
{
    if (!f) return fail(B200BLUR_ERR_INVALID, "feed is NULL");
    if (f->failed) return fail(B200BLUR_ERR_CUDA, "feed is in a failed state (watchdog tripped); destroy it");
    if (f->running) return fail(B200BLUR_ERR_INVALID, "feed is already running");
    CU_TRY(cudaSetDevice(f->ctx->device));
    // batches submitted before start are part of this run: the kernel numbers batches from `base`
    f->base = f->flushed;
    unsigned long long *w;
    if (int rc = feed_ctl_word(f, f->kstream, 0ull, &w)) return rc;
    CU_TRY(cudaMemcpyAsync(f->d_ctl, w, 8, cudaMemcpyHostToDevice, f->kstream));       // tail = 0
    CU_TRY(cudaMemcpyAsync(f->d_ctl + 1, w, 8, cudaMemcpyHostToDevice, f->kstream));   // closed = 0
    CU_TRY(cudaStreamSynchronize(f->kstream));
    if (int rc = feed_launch(f, f->kstream, 0)) return rc;
    f->running = true;
    return B200BLUR_OK;
}

int b200blur_feed_submit(b200blur_feed *f, const void *d_in, void *d_out, int n_images, int64_t *ticket)
{
    if (!f) return fail(B200BLUR_ERR_INVALID, "feed is NULL");
    if (f->failed) return fail(B200BLUR_ERR_CUDA, "feed is in a failed state (watchdog tripped); destroy it");
    if (n_images < 1 || n_images > f->max_batch)
        return fail(B200BLUR_ERR_INVALID, "batch of %d images (feed takes 1..%d)", n_images, f->max_batch);
    if (!d_in || !d_out || !aligned16(d_in) || !aligned16(d_out) || d_in == d_out)
        return fail(B200BLUR_ERR_INVALID, "feed batches need distinct, 16-byte aligned device pointers");
    const int64_t t = f->submitted;
    if (t >= f->cap) {
        // the descriptor slot is free once the batch that used it last has completed
        if (!feed_slot_done(f, t - f->cap)) {
            if (!f->running) return fail(B200BLUR_ERR_INVALID, "feed ring is full (%d batches) and the feed is not running", f->cap);
            if (int rc = b200blur_feed_wait(f, t - f->cap)) return rc;
        }
    }
    b200blur::FeedBatch &b = f->h_batches[t % f->cap];
    memset(&b, 0, sizeof b);
    b.in = static_cast<const uint8_t *>(d_in);
    b.out = static_cast<uint8_t *>(d_out);
    b.n_images = n_images;
    f->submitted = t + 1;
    if (ticket) *ticket = t;
    return B200BLUR_OK;
}

int b200blur_feed_flush(b200blur_feed *f)
{
    if (!f) return fail(B200BLUR_ERR_INVALID, "feed is NULL");
    if (!f->running) return fail(B200BLUR_ERR_INVALID, "feed is not running (b200blur_feed_start first)");
    CU_TRY(cudaSetDevice(f->ctx->device));
    return feed_publish(f, f->cstream);
}

int b200blur_feed_completed(b200blur_feed *f, int64_t ticket, int *done)
{
    if (!f || !done) return fail(B200BLUR_ERR_INVALID, "NULL pointer in feed_completed");
    if (ticket < 0 || ticket >= f->submitted) return fail(B200BLUR_ERR_INVALID, "ticket %lld was never issued", (long long)ticket);
    *done = feed_slot_done(f, ticket) ? 1 : 0;
    return B200BLUR_OK;
}

int b200blur_feed_wait(b200blur_feed *f, int64_t ticket)
{
    if (!f) return fail(B200BLUR_ERR_INVALID, "feed is NULL");
    if (ticket < 0 || ticket >= f->submitted) return fail(B200BLUR_ERR_INVALID, "ticket %lld was never issued", (long long)ticket);
    if (feed_slot_done(f, ticket)) return B200BLUR_OK;
    if (!f->running) return fail(B200BLUR_ERR_INVALID, "feed is not running: batch %lld cannot complete", (long long)ticket);
    if (ticket >= f->flushed)
        if (int rc = b200blur_feed_flush(f)) return rc;
    const double t0 = now_ms();
    const double limit_ms = (double)f->timeout_ns / 1e6 + 2000.0;
    for (unsigned spin = 0;; spin++) {
        if (feed_slot_done(f, ticket)) return B200BLUR_OK;
        if (*(volatile unsigned int *)(f->h_done + f->cap)) {
            f->failed = true;
            return fail(B200BLUR_ERR_CUDA, "feed watchdog tripped: the kernel waited too long for the host and stopped");
        }
        if ((spin & 1023) == 1023) {
            if (cudaStreamQuery(f->kstream) != cudaErrorNotReady) {   // kernel gone (error or stopped) without completing it
                if (feed_slot_done(f, ticket)) return B200BLUR_OK;
                f->failed = true;
                cudaError_t e = cudaGetLastError();
                return fail(B200BLUR_ERR_CUDA, "feed kernel ended before batch %lld completed (%s)", (long long)ticket, cudaGetErrorString(e));
            }
            if (now_ms() - t0 > limit_ms) {
                f->failed = true;
                return fail(B200BLUR_ERR_CUDA, "timed out waiting for batch %lld", (long long)ticket);
            }
        }
    }
}

int b200blur_feed_stop(b200blur_feed *f)
{
    if (!f) return fail(B200BLUR_ERR_INVALID, "feed is NULL");
    if (!f->running) return B200BLUR_OK;
    CU_TRY(cudaSetDevice(f->ctx->device));
    int rc = feed_publish(f, f->cstream);
    unsigned long long *w;
    if (!rc) rc = feed_ctl_word(f, f->cstream, 1ull, &w);
    if (!rc) {
        CU_TRY(cudaMemcpyAsync(f->d_ctl + 1, w, 8, cudaMemcpyHostToDevice, f->cstream));   // closed = 1 (after the final tail)
        CU_TRY(cudaStreamSynchronize(f->cstream));
    }
    cudaError_t e = cudaStreamSynchronize(f->kstream);   // the kernel drains what was published and exits
    f->running = false;
    if (e != cudaSuccess) {
        f->failed = true;
        return fail(B200BLUR_ERR_CUDA, "%d - feed kernel: %s", (int)e, cudaGetErrorString(e));
    }
    if (*(volatile unsigned int *)(f->h_done + f->cap)) {
        f->failed = true;
        return fail(B200BLUR_ERR_CUDA, "feed watchdog tripped: the kernel waited too long for the host and stopped");
    }
    return rc;
}

int b200blur_feed_destroy(b200blur_feed *f)
{
    if (!f) return B200BLUR_OK;
    cudaSetDevice(f->ctx->device);
    if (f->running) b200blur_feed_stop(f);
    feed_release(f);
    return B200BLUR_OK;
}

int64_t b200blur_feed_submitted(const b200blur_feed *f) { return f ? f->submitted : 0; }

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
    // Odd widths (width*channels % 16 != 0).  Rows up to 4 KB need nothing special: the streamed kernel reads the tight
    // rows as they are (aligned-superset bulk copies, re-aligned in shared memory) and its store warps write tight rows
    // (one pass).  With B200BLUR_NO_TIGHT_OUT only the input side is folded: 16-byte-pitched rows go to a scratch buffer
    // that one re-pack kernel turns back into tight rows (2 passes).  Rows wider than 4 KB are re-pitched on the way in
    // as well (3 passes; the byte-wise generic kernel would be ~20x slower).
    const size_t row_bytes = (size_t)width * channels;
    static const bool env_no_tight = getenv("B200BLUR_NO_TIGHT") != nullptr;
    // rows up to 4 KB run in ONE pass: the streamed kernel reads and writes the tight rows itself (TIGHT input and output)
    const bool one_pass = row_bytes <= 4096 && ctx->kernel_variant != 1 && !env_no_tight && !getenv("B200BLUR_NO_TIGHT_OUT");
    if (row_bytes % 16 != 0 && channels <= 4 && row_bytes >= 256 && n_images > 0 && height > 0 && !one_pass) {
        const size_t dev_pitch = (row_bytes + 15) / 16 * 16;
        const size_t dev_image_bytes = dev_pitch * (size_t)height;
        int64_t chunk = (int64_t)((512ull << 20) / dev_image_bytes);   // scratch of up to 512 MB (x2 for wide rows)
        if (chunk < 1) chunk = 1;
        if (chunk > n_images) chunk = n_images;
        const size_t need = dev_image_bytes * (size_t)chunk;
        const bool need_in = !(row_bytes <= 4096 && ctx->kernel_variant != 1 && !env_no_tight);   // TIGHT input needs no scratch_in
        if (ctx->scratch_bytes < need || (need_in && !ctx->scratch_in)) {
            if (ctx->scratch_in) cudaFree(ctx->scratch_in);
            if (ctx->scratch_out) cudaFree(ctx->scratch_out);
            ctx->scratch_in = ctx->scratch_out = nullptr;
            ctx->scratch_bytes = 0;
            if (need_in) CU_TRY(cudaMalloc((void **)&ctx->scratch_in, need));
            CU_TRY(cudaMalloc((void **)&ctx->scratch_out, need));
            ctx->scratch_bytes = need;
        }
        for (int64_t i0 = 0; i0 < n_images; i0 += chunk) {
            const int64_t n = (n_images - i0 < chunk) ? n_images - i0 : chunk;
            const uint8_t *src = static_cast<const uint8_t *>(d_in) + (size_t)i0 * image_bytes;
            uint8_t *dst = static_cast<uint8_t *>(d_out) + (size_t)i0 * image_bytes;
            const bool fast_repack = aligned16(d_out);  // the re-pack kernel writes aligned 16-byte words of the tight output
            const bool tight_in = row_bytes <= 4096 && ctx->kernel_variant != 1 && !env_no_tight;
            if (tight_in) {
                // nothing to do on the way in
            } else if (fast_repack) {
                launch_repitch_in(ctx, s, src, ctx->scratch_in, n * (long long)height, (int)row_bytes, (int)dev_pitch);
                launches++;
            } else {
                CU_TRY(cudaMemcpy2DAsync(ctx->scratch_in, dev_pitch, src, row_bytes, row_bytes, (size_t)n * height,
                                         cudaMemcpyDeviceToDevice, s));
            }
            b200blur_launch l;
            if (int rc = tight_in ? b200blur_launch_rows_pitched(&l, src, ctx->scratch_out, width, height, channels, 0, height, n,
                                                                 image_bytes, dev_image_bytes, 0, dev_pitch)
                                  : b200blur_launch_rows_pitched(&l, ctx->scratch_in, ctx->scratch_out, width, height, channels, 0,
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
    // One work descriptor per batch (coalesce == 0): the batches go through the feed kernel as a descriptor table --
    // one kernel launch per call, groups never span batches, every batch completes on its own (flag in host-mapped
    // memory).  A repeated identical request re-uses the table that is already on the device: no copies at all.
    // coalesce == 2 (or B200BLUR_NO_FEED) keeps the round-1 form, one kernel LAUNCH per batch, for comparison.
    static const bool env_no_feed = getenv("B200BLUR_NO_FEED") != nullptr;
    const int64_t n_batches = (n_images + batch_size - 1) / batch_size;
    if (coalesce == 0 && !env_no_feed && n_batches > 1 && n_batches < (1 << 26) && feed_eligible(width, height, channels) &&
        aligned16(d_in) && aligned16(d_out) && d_in != d_out && ctx->kernel_variant != 1) {
        b200blur_feed *f = ctx->resident_feed;
        if (!f || f->width != width || f->height != height || f->channels != channels || f->max_batch != batch_size ||
            f->cap != n_batches) {   // (table mode: descriptor b lives in ring slot b, so the ring is exactly the table)
            if (f) {
                CU_TRY(cudaStreamSynchronize(s));
                feed_release(f);
                ctx->resident_feed = nullptr;
            }
            if (int rc = feed_build(ctx, width, height, channels, batch_size, (int)(n_batches < 2 ? 2 : n_batches), false, &f))
                return event_abort(ctx, slot, nullptr, rc);
            ctx->resident_feed = f;
            ctx->resident_in = nullptr;
        }
        const long long total_groups = (long long)n_batches * f->plan.sp.feed_gpb;
        if (total_groups < 0x7fffffffLL) {
            if (ctx->resident_in != d_in || ctx->resident_out != d_out || ctx->resident_n != n_images) {
                CU_TRY(cudaStreamSynchronize(s));   // the table (and its pinned mirror) may still be in use
                for (int64_t b = 0; b < n_batches; b++) {
                    b200blur::FeedBatch &d = f->h_batches[b];
                    memset(&d, 0, sizeof d);
                    d.in = static_cast<const uint8_t *>(d_in) + (size_t)b * batch_size * image_bytes;
                    d.out = static_cast<uint8_t *>(d_out) + (size_t)b * batch_size * image_bytes;
                    d.n_images = (int)((n_images - b * batch_size < batch_size) ? n_images - b * batch_size : batch_size);
                }
                f->h_ctl[0] = (unsigned long long)n_batches;   // tail: everything is published before the kernel starts
                f->h_ctl[1] = 1ull;                            // closed
                CU_TRY(cudaMemcpyAsync(f->d_batches, f->h_batches, sizeof(b200blur::FeedBatch) * (size_t)n_batches,
                                       cudaMemcpyHostToDevice, s));
                CU_TRY(cudaMemcpyAsync(f->d_ctl, f->h_ctl, 16, cudaMemcpyHostToDevice, s));
                ctx->resident_in = d_in;
                ctx->resident_out = d_out;
                ctx->resident_n = n_images;
            }
            f->base = ctx->resident_calls++ * n_batches;
            if (int rc = feed_launch(f, s, total_groups)) return event_abort(ctx, slot, nullptr, rc);
            if (stats) {
                if (int rc = event_end(ctx, 0, slot)) return rc;
                double ms = 0;
                if (int rc = b200blur_event_ms(ctx, ev_all, &ms)) return rc;
                b200blur_event_release(ctx, ev_all);
                memset(stats, 0, sizeof *stats);
                stats->kernel_ms = ms;
                stats->images = n_images;
                stats->launches = 1;
                stats->wall_ms = now_ms() - t0;
            }
            return B200BLUR_OK;
        }
    }
    const int64_t step = coalesce == 1 ? (n_images > 0 ? n_images : 1) : batch_size;
    // One launch per batch (coalesce == 0): the batches are independent, so their launches are spread round-robin
    // over all queues of the context and overlap each other's ramp-up and tail; queue 0 forks and joins the others,
    // so the call still behaves like one in-order operation on queue 0.
    const int nq = (coalesce != 1 && n_images > (int64_t)batch_size) ? (int)ctx->queues.size() : 1;
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
    key.variant = ctx->kernel_variant;
    const bool graphable = nq > 1 && !env_no_graph && (n_images + step - 1) / step >= 8;
    bool capturing = false, replayed = false;
    if (graphable) {
        if (ctx->graph_exec && ctx->graph_key == key) {
            CU_TRY(cudaGraphLaunch(ctx->graph_exec, s));
            // the graph's kernels used the work counters of queues 1.. too: later launches on those queues must come
            // after it (the eager path gets this ordering from its join events)
            CU_TRY(cudaEventRecord(fork, s));
            for (int q = 1; q < nq; q++) CU_TRY(cudaStreamWaitEvent(ctx->queues[q], fork, 0));
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

// The end-to-end pipeline of one context.  `shared_next` == nullptr: this context moves every chunk of the stream, in
// order.  Otherwise the chunk indices are TAKEN from *shared_next (b200blur_run_host_multi: several contexts, one host
// thread each, draining one stream); `n_workers` only sizes the chunks so that every worker keeps a full pipeline.
static int run_host_impl(b200blur_ctx *ctx, const void *h_in, void *h_out, int width, int height, int channels,
                         int64_t n_images, int batch_size, b200blur_stats *stats, std::atomic<int64_t> *shared_next,
                         int n_workers)
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
    const int env_ring = getenv("B200BLUR_RING") ? atoi(getenv("B200BLUR_RING")) : 0;
    const int env_chunk_mb = getenv("B200BLUR_E2E_CHUNK_MB") ? atoi(getenv("B200BLUR_E2E_CHUNK_MB")) : 64;
    const int n_slots = env_ring > 1 ? env_ring : 4;
    if (env_chunk_mb > 0 && image_bytes > 0) {
        const double target = (double)env_chunk_mb * 1024 * 1024;
        const double batch_bytes = (double)batch_size * image_bytes;
        long long chunk = batch_size;
        if (batch_bytes < target) {
            long long fuse = (long long)(target / batch_bytes + 0.5);
            const long long n_batches = (n_images + batch_size - 1) / batch_size;
            if (fuse > n_batches / (16 * n_workers)) fuse = n_batches / (16 * n_workers);   // keep at least ~16 chunks in every pipeline
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
    const bool one_pass = row_bytes <= 4096 && ctx->kernel_variant != 1 && !getenv("B200BLUR_NO_TIGHT") && !getenv("B200BLUR_NO_TIGHT_OUT");
    const bool repitch = row_bytes % 16 != 0 && channels <= 4 && row_bytes >= 256 && !one_pass;
    const size_t dev_pitch = repitch ? (row_bytes + 15) / 16 * 16 : row_bytes;
    const size_t dev_image_bytes = dev_pitch * (size_t)height;
    // rows up to 4 KB: the blur reads the tight upload as it is (TIGHT input), only the output is re-packed
    const bool tight_in = repitch && row_bytes <= 4096 && ctx->kernel_variant != 1 && getenv("B200BLUR_NO_TIGHT") == nullptr;
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

    // One chunk's three stages.  `phase`: 0 = all three, 1 = upload + kernel only, 2 = download only.
    auto issue = [&](int64_t ci, int phase, int64_t slot_index) -> int {
        b200blur_ctx::Slot &s = ctx->ring[slot_index % n_slots];
        const int64_t i0 = ci * batch_size;
        const int64_t n = (n_images - i0 < batch_size) ? n_images - i0 : batch_size;
        const size_t bytes = (size_t)n * image_bytes;
        const uint8_t *src = static_cast<const uint8_t *>(h_in) + (size_t)i0 * image_bytes;
        uint8_t *dst = static_cast<uint8_t *>(h_out) + (size_t)i0 * image_bytes;
        if (phase != 2) {
            // H2D
            CU_TRY(cudaEventRecord(s.ev[0], q_in));
            if (bytes) CU_TRY(cudaMemcpyAsync(repitch ? s.t_in : s.d_in, src, bytes, cudaMemcpyHostToDevice, q_in));
            CU_TRY(cudaEventRecord(s.ev[1], q_in));
            // blur
            CU_TRY(cudaStreamWaitEvent(q_k, s.ev[1], 0));
            CU_TRY(cudaEventRecord(s.ev[2], q_k));
            b200blur_launch l;
            if (int rc = tight_in ? b200blur_launch_rows_pitched(&l, s.t_in, s.d_out, width, height, channels, 0, height, n, image_bytes,
                                                                 dev_image_bytes, 0, dev_pitch)
                                  : b200blur_launch_rows_pitched(&l, s.d_in, s.d_out, width, height, channels, 0, height, n,
                                                                 dev_image_bytes, dev_image_bytes, repitch ? dev_pitch : 0,
                                                                 repitch ? dev_pitch : 0))
                return rc;
            if (repitch && !tight_in && bytes) {
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
        }
        if (phase != 1) {
            // D2H
            CU_TRY(cudaStreamWaitEvent(q_out, s.ev[3], 0));
            CU_TRY(cudaEventRecord(s.ev[4], q_out));
            if (bytes) CU_TRY(cudaMemcpyAsync(dst, repitch ? s.t_out : s.d_out, bytes, cudaMemcpyDeviceToHost, q_out));
            CU_TRY(cudaEventRecord(s.ev[5], q_out));
        }
        return B200BLUR_OK;
    };

    // B200BLUR_E2E_PHASED=1: one transfer direction per GPU at a time.  The chunks move in waves of one ring: all uploads
    // (+ kernels) of a wave, then all its downloads, then the next wave's uploads.  A single GPU loses by it (its link is
    // full duplex: 46 GB/s each way at once, 55 one way), but when many GPUs share one host fabric the fabric carries
    // more with fewer flows per link -- measured on an 8-GPU box of this pool (tools/linkbench_multi.py,
    // profiles/r02_linkbench_n8.txt): 64 GB/s each way with all 16 flows at once, 75 GB/s each way with half the GPUs
    // uploading while the other half download.  Processes are not synchronised with each other; their waves interleave.
    const char *env_phased = getenv("B200BLUR_E2E_PHASED");
    const bool phased = env_phased && atoi(env_phased) != 0 && !shared_next;
    int64_t taken = n_chunks, images_done = n_images;   // chunks / images this context moved
    if (shared_next) {
        // take the next chunk only when a ring slot is free for it: a GPU behind a slower link takes fewer
        taken = images_done = 0;
        for (;;) {
            if (taken >= n_slots)
                if (int rc = harvest(ctx->ring[taken % n_slots])) return rc;
            const int64_t ci = shared_next->fetch_add(1, std::memory_order_relaxed);
            if (ci >= n_chunks) break;
            if (int rc = issue(ci, 0, taken)) return rc;
            const int64_t i0 = ci * batch_size;
            images_done += (n_images - i0 < batch_size) ? n_images - i0 : batch_size;
            taken++;
        }
    } else if (!phased) {
        for (int64_t ci = 0; ci < n_chunks; ci++) {
            if (ci >= n_slots)
                if (int rc = harvest(ctx->ring[ci % n_slots])) return rc;  // slot's previous chunk fully drained (also frees d_in/d_out)
            if (int rc = issue(ci, 0, ci)) return rc;
        }
    } else {
        for (int64_t w0 = 0; w0 < n_chunks; w0 += n_slots) {
            const int64_t w1 = (w0 + n_slots < n_chunks) ? w0 + n_slots : n_chunks;
            if (w0 > 0) {
                for (int64_t ci = w0 - n_slots; ci < w0; ci++)
                    if (int rc = harvest(ctx->ring[ci % n_slots])) return rc;   // the previous wave's downloads are done
            }
            for (int64_t ci = w0; ci < w1; ci++)
                if (int rc = issue(ci, 1, ci)) return rc;
            // downloads start when the wave's last upload has finished
            CU_TRY(cudaStreamWaitEvent(q_out, ctx->ring[(w1 - 1) % n_slots].ev[1], 0));
            for (int64_t ci = w0; ci < w1; ci++)
                if (int rc = issue(ci, 2, ci)) return rc;
        }
    }
    // (with a shared counter the slot that was harvested before the failed take must not be harvested twice)
    int64_t first_pending = taken > n_slots ? taken - n_slots : 0;
    if (shared_next && taken >= n_slots) first_pending++;
    if (phased && n_chunks > 0) first_pending = (n_chunks - 1) / n_slots * n_slots;   // earlier waves were harvested
    for (int64_t ci = first_pending; ci < taken; ci++)
        if (int rc = harvest(ctx->ring[ci % n_slots])) return rc;
    CU_TRY(cudaStreamSynchronize(q_out));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->wall_ms = now_ms() - t0;
        stats->h2d_ms = ms_in;
        stats->kernel_ms = ms_k;
        stats->d2h_ms = ms_out;
        stats->images = images_done;
        stats->launches = launches;
        stats->h2d_bytes = (int64_t)(image_bytes * (size_t)images_done);
        stats->d2h_bytes = (int64_t)(image_bytes * (size_t)images_done);
    }
    return B200BLUR_OK;
}

int b200blur_run_host(b200blur_ctx *ctx, const void *h_in, void *h_out, int width, int height, int channels,
                      int64_t n_images, int batch_size, b200blur_stats *stats)
{
    return run_host_impl(ctx, h_in, h_out, width, height, channels, n_images, batch_size, stats, nullptr, 1);
}

int b200blur_run_host_multi(b200blur_ctx *const *ctxs, int n_ctx, const void *h_in, void *h_out, int width, int height,
                            int channels, int64_t n_images, int batch_size, b200blur_stats *stats)
{
    if (!ctxs || n_ctx < 1 || n_ctx > 64) return fail(B200BLUR_ERR_INVALID, "need 1..64 contexts");
    for (int k = 0; k < n_ctx; k++) {
        if (int rc = ctx_check(ctxs[k])) return rc;
        for (int j = 0; j < k; j++)
            if (ctxs[j] == ctxs[k]) return fail(B200BLUR_ERR_INVALID, "context %d is listed twice", k);
    }
    if (n_ctx == 1) return run_host_impl(ctxs[0], h_in, h_out, width, height, channels, n_images, batch_size, stats, nullptr, 1);
    // one host thread per context (the reference's two devices share one thread and one clFinish, A1:538-539); the
    // threads' pipelines take transfer chunks from one counter until the stream is drained
    std::atomic<int64_t> next{0};
    std::vector<int> rcs((size_t)n_ctx, B200BLUR_OK);
    std::vector<std::string> errors((size_t)n_ctx);
    std::vector<b200blur_stats> local((size_t)n_ctx);
    auto work = [&](int k) {
        rcs[k] = run_host_impl(ctxs[k], h_in, h_out, width, height, channels, n_images, batch_size, &local[k], &next, n_ctx);
        if (rcs[k] != B200BLUR_OK) {
            errors[k] = g_last_error;                       // (thread-local: carry it to the caller's thread)
            next.store(INT64_MAX / 2, std::memory_order_relaxed);   // the others stop taking chunks
        }
    };
    std::vector<std::thread> threads;
    for (int k = 1; k < n_ctx; k++) threads.emplace_back(work, k);
    work(0);
    for (auto &t : threads) t.join();
    for (int k = 0; k < n_ctx; k++)
        if (rcs[k] != B200BLUR_OK) return fail(rcs[k], "context %d: %s", k, errors[k].c_str());
    if (stats)
        for (int k = 0; k < n_ctx; k++) stats[k] = local[k];
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
