// blur_kernels.cuh -- hand-written sm_100a device code for the 3x3 Gaussian-blur stencil.
//
// Replaces the reference's one device kernel, gaussian_blur (gaussian_kernel.cl:19-72): one work-item per pixel,
// 27 byte loads + 27 int->float converts + 27 fp32 MACs per pixel, one image per launch.  Here one launch covers a
// row band of every image of a batch, each thread owns a 16-byte column of a strip of rows, and the arithmetic is
// the exact integer form  out = (sum w_int * p) >> 4  (equal to the fp32 form) done
// two pixels-bytes at a time in packed 16-bit lanes:
//
//   flat-byte view: a row is pitch = width*channels bytes; out[b] needs in[b-C], in[b], in[b+C] of three rows.
//   E_k = bytes 0,2 of word k, O_k = bytes 1,3 of word k, each zero-extended into 16-bit lanes;
//   h  = left + 2*centre + right            (<= 1020, horizontal [1 2 1])
//   v  = h_up + 2*h_mid + h_down            (<= 4080, vertical   [1 2 1])
//   out byte = v >> 4: computed as (v << 4) so the byte sits in the high half of each 16-bit lane and one PRMT
//   re-interleaves E and O lanes into the output word.  No intermediate rounding anywhere.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200blur {

struct BandParams {
    const uint8_t *in;        // first band row of image 0
    uint8_t *out;             // first output row of image 0
    const uint8_t *halo_top;  // row above the band (per image), or nullptr -> replicate row 0 (gaussian_kernel.cl:57)
    const uint8_t *halo_bot;  // row below the band (per image), or nullptr -> replicate the last row
    size_t in_stride;         // bytes between consecutive images
    size_t out_stride;
    size_t top_stride;
    size_t bot_stride;
    int pitch;                // bytes between rows of `in` (>= row_bytes; the vectorised kernels need pitch % 16 == 0)
    int out_pitch;            // bytes between rows of `out`
    int row_bytes;            // meaningful bytes per row = width * channels
    int rows;                 // band height (rows computed and stored)
    int width;
    int channels;
    long long n_images;
};

// ----------------------------------------------------------------------------------------------- small helpers
__device__ __forceinline__ uint4 ldg128(const uint8_t *p)
{
    return __ldg(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ uint32_t ldg32(const uint8_t *p)
{
    return __ldg(reinterpret_cast<const uint32_t *>(p));
}
// Streaming store: the output is written once and never re-read by this kernel.
__device__ __forceinline__ void stg128_stream(uint8_t *p, const uint4 &v)
{
    __stcs(reinterpret_cast<uint4 *>(p), v);
}

// (hi:lo) >> 16 as packed lanes: result lane0 = lo.lane1, lane1 = hi.lane0
__device__ __forceinline__ uint32_t lanes_shift(uint32_t lo, uint32_t hi)
{
    return __byte_perm(lo, hi, 0x5432);
}

// 2*a + b as ONE multiply-add on the FMA pipe.  Written as PTX so that ptxas keeps the already-unpacked lanes `a`
// instead of re-deriving 2*a from the raw word with an extra add + mask on the (busier) ALU pipe.
__device__ __forceinline__ uint32_t mad2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// Horizontal [1 2 1] of one 16-byte chunk.  w = the chunk, wl = the word before it, wr = the word after it
// (only the C bytes nearest the chunk matter).  C = bytes per pixel = distance to the horizontal neighbour.
// Results: hE[k] lanes = h of bytes (4k, 4k+2), hO[k] lanes = h of bytes (4k+1, 4k+3).
template <int C>
__device__ __forceinline__ void hpass(const uint4 &w, uint32_t wl, uint32_t wr, uint32_t (&hE)[4], uint32_t (&hO)[4])
{
    const uint32_t W[6] = {wl, w.x, w.y, w.z, w.w, wr};
    uint32_t E[6], O[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        E[k] = W[k] & 0x00FF00FFu;
        O[k] = __byte_perm(W[k], 0u, 0x4341);  // (W >> 8) & 0x00FF00FF in one PRMT
    }
#pragma unroll
    for (int k = 1; k <= 4; k++) {
        uint32_t LE, LO, RE, RO;
        if (C == 3) {         // neighbours 3 bytes away
            LE = O[k - 1];                        // bytes 4k-3, 4k-1
            LO = lanes_shift(E[k - 1], E[k]);     // bytes 4k-2, 4k
            RE = lanes_shift(O[k], O[k + 1]);     // bytes 4k+3, 4k+5
            RO = E[k + 1];                        // bytes 4k+4, 4k+6
        } else if (C == 4) {  // neighbours one word away
            LE = E[k - 1]; LO = O[k - 1]; RE = E[k + 1]; RO = O[k + 1];
        } else if (C == 2) {
            LE = lanes_shift(E[k - 1], E[k]); LO = lanes_shift(O[k - 1], O[k]);
            RE = lanes_shift(E[k], E[k + 1]); RO = lanes_shift(O[k], O[k + 1]);
        } else {              // C == 1
            LE = lanes_shift(O[k - 1], O[k]); LO = E[k];
            RE = O[k];                        RO = lanes_shift(E[k], E[k + 1]);
        }
        hE[k - 1] = mad2(E[k], LE) + RE;
        hO[k - 1] = mad2(O[k], LO) + RO;
    }
}

// Same horizontal pass, results interleaved as h[2k] = E lanes, h[2k+1] = O lanes of word k.
template <int C>
__device__ __forceinline__ void hpass8(const uint4 &w, uint32_t wl, uint32_t wr, uint32_t (&h)[8])
{
    uint32_t hE[4], hO[4];
    hpass<C>(w, wl, wr, hE, hO);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        h[2 * k] = hE[k];
        h[2 * k + 1] = hO[k];
    }
}

// Vertical [1 2 1] + >>4 + re-interleave: one output word from the h lanes of three rows.
__device__ __forceinline__ uint32_t vpass_word(uint32_t upE, uint32_t midE, uint32_t dnE,
                                               uint32_t upO, uint32_t midO, uint32_t dnO)
{
    uint32_t vE = (2u * midE + upE + dnE) << 4;  // <= 4080*16 = 65280 per lane: the output byte is lane bits 15:8
    uint32_t vO = (2u * midO + upO + dnO) << 4;
    return __byte_perm(vE, vO, 0x7351);          // bytes: vE.b1, vO.b1, vE.b3, vO.b3
}

// ------------------------------------------------------------------------- variant 1: register/shuffle stencil
// Each thread owns one 16-byte column chunk of a strip of RS output rows of one image and slides down it:
// RS+2 input rows are loaded once each (LDG.128, prefetched PF rows ahead), the word to the left/right of the
// chunk comes from the neighbouring lane by shuffle (lanes 0/31: one predicated LDG.32), the two previous rows'
// horizontal sums stay in registers, and every output row is one coalesced 16-byte streaming store per thread.
// Grid: x = image, y = blocks of strip*chunk units within an image.
template <int C, int RS>
__global__ void __launch_bounds__(256)
blur_strip_kernel(const BandParams p, int cpr, int n_strips)
{
    const int units = cpr * n_strips;
    const int u_raw = blockIdx.y * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (u_raw - lane >= units) return;  // whole warp past the end
    const bool valid = u_raw < units;
    const int u = valid ? u_raw : units - 1;
    const int strip = u / cpr;
    const int c = u - strip * cpr;
    const long long img = blockIdx.x;

    const int r0 = strip * RS;
    const size_t col = (size_t)c * 16;
    const uint8_t *src = p.in + (size_t)img * p.in_stride + col;
    const uint8_t *top = p.halo_top ? p.halo_top + (size_t)img * p.top_stride + col : src;
    const uint8_t *bot = p.halo_bot ? p.halo_bot + (size_t)img * p.bot_stride + col
                                    : src + (size_t)(p.rows - 1) * p.pitch;
    uint8_t *dst = p.out + (size_t)img * p.out_stride + col;

    const bool first = (c == 0), last = (c == cpr - 1);
    const bool need_l = (lane == 0) && !first;   // left word lives in another warp's chunk
    const bool need_r = (lane == 31) && !last;

    auto row_ptr = [&](int j) -> const uint8_t * {  // j in [-1, ...]
        if (j < 0) return top;
        if (j >= p.rows) return bot;
        return src + (size_t)j * p.pitch;
    };

    constexpr int PF = 2;  // rows in flight per thread beyond the one being consumed
    uint4 q[PF + 1];
    uint32_t ql[PF + 1], qr[PF + 1];
#pragma unroll
    for (int k = 0; k < PF; k++) {
        const uint8_t *rp = row_ptr(r0 - 1 + k);
        q[k] = ldg128(rp);
        ql[k] = need_l ? ldg32(rp - 4) : 0u;
        qr[k] = need_r ? ldg32(rp + 16) : 0u;
    }

    uint32_t h2E[4], h2O[4], h1E[4], h1O[4];  // h of rows j-2 and j-1
#pragma unroll
    for (int k = 0; k < RS + 2; k++) {
        // prefetch row k+PF
        if (k + PF < RS + 2) {
            const uint8_t *rp = row_ptr(r0 - 1 + k + PF);
            q[(k + PF) % (PF + 1)] = ldg128(rp);
            ql[(k + PF) % (PF + 1)] = need_l ? ldg32(rp - 4) : 0u;
            qr[(k + PF) % (PF + 1)] = need_r ? ldg32(rp + 16) : 0u;
        }
        const uint4 w = q[k % (PF + 1)];
        uint32_t wl = __shfl_up_sync(0xffffffffu, w.w, 1);
        uint32_t wr = __shfl_down_sync(0xffffffffu, w.x, 1);
        if (need_l) wl = ql[k % (PF + 1)];
        if (need_r) wr = qr[k % (PF + 1)];
        if (first) wl = w.x << (8 * (4 - C));   // clamp: pixel -1 := pixel 0        (gaussian_kernel.cl:56)
        if (last) wr = w.w >> (8 * (4 - C));    // clamp: pixel width := pixel width-1

        uint32_t hE[4], hO[4];
        hpass<C>(w, wl, wr, hE, hO);
        if (k >= 2) {
            const int r = r0 + k - 2;
            uint4 o;
            o.x = vpass_word(h2E[0], h1E[0], hE[0], h2O[0], h1O[0], hO[0]);
            o.y = vpass_word(h2E[1], h1E[1], hE[1], h2O[1], h1O[1], hO[1]);
            o.z = vpass_word(h2E[2], h1E[2], hE[2], h2O[2], h1O[2], hO[2]);
            o.w = vpass_word(h2E[3], h1E[3], hE[3], h2O[3], h1O[3], hO[3]);
            if (valid && r < p.rows) stg128_stream(dst + (size_t)r * p.out_pitch, o);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h2E[i] = h1E[i]; h2O[i] = h1O[i];
            h1E[i] = hE[i];  h1O[i] = hO[i];
        }
    }
}

// ------------------------------------------------------------------ variant 2: TMA-bulk streamed persistent stencil
// The Blackwell-native form.  Persistent CTAs (a few per SM) each stream "groups" -- a segment of `seg` output rows of
// a column block of `ipc` consecutive images -- through a ring of NS shared-memory slots.  One elected thread feeds
// the ring with 1-D bulk async copies (cp.async.bulk global->shared, completion on an mbarrier; SASS UBLKCP): because
// a small frame's rows are contiguous in memory, RB rows of an image arrive as ONE bulk copy; wide frames use one
// bulk copy per row of a 2 KB column block (+16 B margins).  All threads then read their 16-byte chunk and the two
// neighbouring words from shared memory (no shuffles, no edge lanes), keep the rolling horizontal sums of the two
// previous rows in registers across slots, and emit one coalesced 16-byte store per output row.  In-flight bytes are
// held by the ring (NS-1 slots per CTA), not by registers, and a segment re-reads only 2 halo rows per `seg` rows.
// Halo rows above/below the band come from halo_top/halo_bot -- possibly another GPU's memory (NVLink) -- or are the
// replicated edge row (gaussian_kernel.cl:57).
//
// The same kernel runs in FEED mode: instead of one (in, out, n_images) triple per launch, the groups are spread over
// a ring of per-batch descriptors that the host appends to (b200blur_feed_*), every batch completes individually (a
// flag in host-mapped memory) and the kernel stays resident until the feed is closed.  That is the reference's
// `batch_size` loop (heterogeneous_blur.c:418-539: stage a batch, enqueue it, wait for it) without a kernel launch per
// batch: at 35 images per batch a launch would carry ~2.4 us of HBM work, at 1 image 60 ns.
struct StreamParams {
    BandParams b;
    int cpr;            // 16-byte chunks per row
    int cb;             // chunks per column block (<= blockDim.x)
    int ncb;            // column blocks per row
    int ipc;            // images side by side in one CTA step (ncb == 1 only)
    int seg;            // output rows per group (coarse groups)
    int nseg;           // coarse segments per band
    int margin;         // 0 (whole rows, contiguous copies) or 16 (per-row copies: column blocks, heavily padded rows)
    int sstride;        // shared-memory row stride in bytes = cb*16 + 2*margin, or the row pitch when margin == 0
    int slot_bytes;     // ipc * RB * sstride (TIGHT: ipc * lane_bytes)
    int lane_bytes;     // TIGHT: shared-memory bytes of one image lane of a slot (16 B lead pad + up to 3 runs of rows + slack)
    int stage_pitch;    // TIGHT output: row pitch of the shared-memory output staging (row bytes rounded up to 16)
    int stage_slot_bytes;   // ... bytes of one staging slot = ipc * RB * stage_pitch
    unsigned row_recip;     // ... ceil(2^32 / row_bytes): offset / row_bytes == umulhi(offset, row_recip) for offsets < 2^16
    long long img_blocks;   // ceil(n_images / ipc)
    // Guided tail: image blocks [0, ib_coarse) are cut into `nseg` segments of `seg` rows, the remaining image blocks --
    // the last work handed out -- into `nseg_fine` segments of `seg_fine` rows, so that the CTAs run dry within a
    // fraction of a coarse group's time of each other instead of a whole one.
    long long ib_coarse;
    long long g_coarse;     // ib_coarse * nseg * ncb: first fine group
    int seg_fine, nseg_fine;
    long long n_groups;     // g_coarse + (img_blocks - ib_coarse) * nseg_fine * ncb
    unsigned long long *work;  // work[0] = next group to hand out, work[1] = CTAs finished (both 0 between launches)
    // Right edge of a row that does not end on a chunk boundary (row_bytes % 16 != 0, pitched rows).  The clamp
    // "pixel width := pixel width-1" means window bytes [row_bytes, row_bytes + C) := bytes [row_bytes - C, row_bytes).
    // For the chunk that contains the row end (and the one before it when the end is < 4 bytes into the last chunk)
    // the 24-byte window {wl, w, wr} is rewritten with one PRMT per word; the selectors are computed on the host.
    int edge_general;          // 0: rows end on a chunk boundary (fast path)
    int edge_prev;             // 1: the chunk before the last one needs its wr word patched too
    uint32_t sel_last[6];      // PRMT selectors for window words 0..5 of the last chunk (pairs: previous word, word)
    uint32_t sel_prev;         // PRMT selector for the wr word of the chunk before the last
    // ---- FEED mode (template FEED = true); b.in / b.out / b.n_images are unused, every batch brings its own
    const struct FeedBatch *batches;   // ring of `feed_cap` descriptors in device memory, written by the host's copies
    unsigned long long *feed_ctl;      // [0] = batches published so far (tail), [1] = closed, [2] = watchdog tripped
    unsigned int *feed_count;          // per ring slot: consumer-warp arrivals of the batch in it (0 between batches)
    volatile unsigned int *feed_done;  // per ring slot, HOST-mapped: sequence number (low 32 bits, +1) of the batch completed
    int feed_cap;                      // ring capacity (batches)
    int feed_gpb;                      // group slots per batch = ceil(max_batch / ipc) * nseg * ncb (all coarse)
    unsigned int feed_target;          // arrivals that complete a batch = feed_gpb * consumer warps
    unsigned long long feed_timeout_ns;   // a CTA that waits this long for the host gives up (watchdog; 0 = never wait)
    unsigned long long feed_base;      // sequence number of the batch in descriptor-ring position 0 of this launch
};

// One batch of a feed: `n_images` images starting at `in`, results to `out`, optional halo rows above / below the band
// (Approach 2: rows of the neighbouring bands, possibly in another GPU's memory); strides are the feed's.
struct FeedBatch {
    const uint8_t *in;
    uint8_t *out;
    const uint8_t *top;
    const uint8_t *bot;
    int n_images;
    int pad_[3];
};
static_assert(sizeof(FeedBatch) == 48, "FeedBatch is read as three 16-byte words");

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)   // non-blocking
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// 1-D bulk async copy global -> shared, completion (bytes) signalled on an mbarrier.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4 &v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16 bytes at an arbitrary shared-memory byte address: five aligned words and four funnel shifts
__device__ __forceinline__ uint4 lds_unaligned16(uint32_t a)
{
    const uint32_t a4 = a & ~3u, sh = (a & 3u) * 8u;
    uint32_t x[5];
#pragma unroll
    for (int i = 0; i < 5; i++) x[i] = lds32(a4 + 4u * i);
    return make_uint4(__funnelshift_r(x[0], x[1], sh), __funnelshift_r(x[1], x[2], sh), __funnelshift_r(x[2], x[3], sh),
                      __funnelshift_r(x[3], x[4], sh));
}
// Programmatic dependent launch: the next kernel in the stream may start its prologue while this one drains
// (launch_dependents), and must not touch global memory before the previous kernel has completed and flushed (wait).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
}  // namespace ptx

// What the producer knows about a group.  Groups are ordered image-block major, then segment, then column block, so
// CTAs that run concurrently work on adjacent segments of the same images and the shared halo rows hit in L2.
struct GroupGeom {
    const uint8_t *in;   // image lane 0 of the group: first row of the band
    const uint8_t *top;  // image lane 0: halo row above the band, or nullptr (replicate)
    const uint8_t *bot;
    int n_img;           // images in the group (<= ipc)
    int r0;              // first output row
    int nr;              // output rows
    int x0;              // first byte column of the column block
    int cbe;             // chunks in this column block
    bool left_edge, right_edge;
};

// What the consumers need to know about a group: written by the producer into shared memory next to the ring slot that
// holds the group's first rows (the consumers never decode group indices themselves).
struct __align__(16) GroupMeta {
    uint8_t *out;        // image lane 0: output row r0, byte column x0
    int n_img;           // < 0: no more work
    int nr;
    int cbe;
    int flags;           // bit 0 left edge, bit 1 right edge, bits 8..: see the producer (second-to-last chunk of EDGE rows)
    int feed_slot;       // FEED: descriptor-ring slot of the batch this group belongs to
    unsigned int feed_seq;   // FEED: value to publish in feed_done[feed_slot] when the batch completes
};
static_assert(sizeof(GroupMeta) == 32, "GroupMeta is read as two 16-byte words");

// Segment/column decomposition of group `local` (an index within one image block's groups, or -- FEED -- one batch's).
__device__ __forceinline__ void decode_rows_cols(const StreamParams &sp, unsigned sc, int seg, int rows, GroupGeom &q)
{
    const int si = (int)(sc / (unsigned)sp.ncb);
    const int ci = (int)sc - si * sp.ncb;
    q.r0 = si * seg;
    q.nr = min(seg, rows - q.r0);
    q.x0 = ci * sp.cb * 16;
    q.cbe = min(sp.cb, sp.cpr - ci * sp.cb);
    q.left_edge = (ci == 0);
    q.right_edge = (ci == sp.ncb - 1);
}

__device__ __forceinline__ void decode_group(const StreamParams &sp, long long g64, GroupGeom &q, uint8_t *&out)
{
    // n_groups < 2^31 (checked on the host: a group is at least a few KB), so 32-bit division -- inlined, no call
    unsigned ib, sc;
    int seg;
    if (g64 < sp.g_coarse) {
        const unsigned g = (unsigned)g64, per_block = (unsigned)(sp.nseg * sp.ncb);
        ib = g / per_block;
        sc = g - ib * per_block;
        seg = sp.seg;
    } else {
        const unsigned g = (unsigned)(g64 - sp.g_coarse), per_block = (unsigned)(sp.nseg_fine * sp.ncb);
        ib = g / per_block;
        sc = g - ib * per_block;
        ib += (unsigned)sp.ib_coarse;
        seg = sp.seg_fine;
    }
    decode_rows_cols(sp, sc, seg, sp.b.rows, q);
    const long long img0 = (long long)ib * sp.ipc;
    const long long left = sp.b.n_images - img0;
    q.n_img = left < sp.ipc ? (int)left : sp.ipc;
    q.in = sp.b.in + (size_t)img0 * sp.b.in_stride;
    q.top = sp.b.halo_top ? sp.b.halo_top + (size_t)img0 * sp.b.top_stride : nullptr;
    q.bot = sp.b.halo_bot ? sp.b.halo_bot + (size_t)img0 * sp.b.bot_stride : nullptr;
    out = sp.b.out + (size_t)img0 * sp.b.out_stride + (size_t)q.r0 * sp.b.out_pitch + q.x0;
}

// TIGHT rows (row pitch = width*channels, not a multiple of 16; any pointer alignment): consecutive rows of an image are
// still one contiguous byte range, only not a 16-byte aligned one.  The producer copies the ALIGNED SUPERSET of each run
// of rows (up to 15 bytes more at either end, inside the same 16-byte blocks as the run's first and last byte) and notes,
// for every row of the slot, where its first byte landed (`rowoff`); the consumers read their 24-byte window at that
// byte address with seven aligned 32-bit loads and six funnel shifts.  Nothing else changes: this folds the re-pitch
// pass that used to run before the blur into the blur's own loads.
template <int RB>
__device__ __forceinline__ void stream_issue_slot_tight(const StreamParams &sp, const GroupGeom &q, int slot_in_item,
                                                        uint32_t slot_smem, uint32_t bar, unsigned short *rowoff)
{
    const BandParams &b = sp.b;
    const int k0 = slot_in_item * RB;
    const int k1 = min(k0 + RB, q.nr + 2);
    for (int il = 0; il < q.n_img; il++) {
        const uint8_t *src = q.in + (size_t)il * b.in_stride;
        const uint32_t lane_base = slot_smem + (uint32_t)(il * sp.lane_bytes) + 16u;
        uint32_t cur = 0;   // aligned offset of the next run within the lane
        int k = k0;
        while (k < k1) {
            const int j = q.r0 - 1 + k;
            const uint8_t *rp;
            int run = 1;
            if (j < 0) {
                rp = q.top ? q.top + (size_t)il * b.top_stride : src;
            } else if (j >= b.rows) {
                rp = q.bot ? q.bot + (size_t)il * b.bot_stride : src + (size_t)(b.rows - 1) * b.pitch;
            } else {
                rp = src + (size_t)j * b.pitch;
                run = min(k1 - k, b.rows - j);
            }
            const uint32_t delta = (uint32_t)(reinterpret_cast<uintptr_t>(rp) & 15u);
            // (rows of a run are b.pitch apart: the tight case pitch == row_bytes, or any other unaligned pitch)
            const uint32_t bytes = (delta + (uint32_t)(run - 1) * (uint32_t)b.pitch + (uint32_t)b.row_bytes + 15u) & ~15u;
            for (int i = 0; i < run; i++)
                rowoff[il * RB + (k - k0) + i] = (unsigned short)(cur + delta + (uint32_t)i * (uint32_t)b.pitch);
            ptx::mbar_expect_tx(bar, bytes);
            ptx::bulk_g2s(lane_base + cur, rp - delta, bytes, bar);
            cur += bytes;
            k += run;
        }
    }
    ptx::mbar_arrive(bar);
}

template <int RB>
__device__ __forceinline__ void stream_issue_slot(const StreamParams &sp, const GroupGeom &q, int slot_in_item,
                                                  uint32_t slot_smem, uint32_t bar)
{
    const BandParams &b = sp.b;
    const int k0 = slot_in_item * RB;               // first input-row index of the slot within the item
    const int k1 = min(k0 + RB, q.nr + 2);          // one past the last
    // byte range of a row that this column block needs (margins clipped at the image edges)
    const int lm = q.left_edge ? 0 : sp.margin;
    const int rm = q.right_edge ? 0 : sp.margin;
    const int xb = q.x0 - lm;
    const uint32_t row_bytes = (uint32_t)(q.cbe * 16 + lm + rm);
    const uint32_t dst_col = (uint32_t)(sp.margin - lm);
    for (int il = 0; il < q.n_img; il++) {
        const uint8_t *src = q.in + (size_t)il * b.in_stride;
        const uint32_t dst_img = slot_smem + (uint32_t)(il * RB * sp.sstride);
        int k = k0;
        while (k < k1) {
            const int j = q.r0 - 1 + k;  // input row relative to the band
            const uint8_t *rp;
            int run = 1;
            if (j < 0) {
                rp = q.top ? q.top + (size_t)il * b.top_stride : src;
            } else if (j >= b.rows) {
                rp = q.bot ? q.bot + (size_t)il * b.bot_stride : src + (size_t)(b.rows - 1) * b.pitch;
            } else {
                rp = src + (size_t)j * b.pitch;
                if (sp.margin == 0) run = min(k1 - k, b.rows - j);  // contiguous rows: one copy
            }
            // whole rows: `run` rows as they lie in memory (padding included); a halo row brings only its live chunks
            const uint32_t bytes = (sp.margin != 0) ? row_bytes
                                   : (j < 0 || j >= b.rows) ? (uint32_t)sp.cpr * 16u : (uint32_t)run * (uint32_t)b.pitch;
            const uint32_t dst = dst_img + (uint32_t)((k - k0) * sp.sstride) + dst_col;
            ptx::mbar_expect_tx(bar, bytes);
            ptx::bulk_g2s(dst, rp + xb, bytes, bar);
            k += run;
        }
    }
    ptx::mbar_arrive(bar);
}

// FEED: `n` consumer-warp arrivals for the batch in descriptor slot `slot`; the arrival that completes the batch
// re-arms the slot's counter and publishes the batch's sequence number to the host.  The caller has fenced at GPU scope
// after observing (through a CTA-scope mbarrier) that the consumer warps' stores of the group were issued -- the fence
// is cumulative over them -- and the arrival that completes the batch fences at system scope before the host flag.
__device__ __forceinline__ void feed_count_arrivals(const StreamParams &sp, int slot, unsigned int seq, unsigned int old,
                                                    unsigned int n)
{
    if (old + n == sp.feed_target) {
        sp.feed_count[slot] = 0;
        __threadfence_system();
        sp.feed_done[slot] = seq;
    }
}
__device__ __forceinline__ void feed_arrive(const StreamParams &sp, int slot, unsigned int seq, unsigned int n)
{
    __threadfence();
    feed_count_arrivals(sp, slot, seq, atomicAdd(sp.feed_count + slot, n), n);
}

constexpr int kFeedDepth = 8;   // group records in flight per CTA between producer, consumers and accountant (> NS)

// Warp-specialised: warp 0 is the producer (one elected lane issues the bulk copies and never computes), warps 1..
// are consumers.  full[NS] barriers carry the copies' byte counts; empty[NS] barriers collect one arrival per consumer
// warp, so consumer warps never wait for each other -- only for data.
// Work is handed out dynamically: the producer takes the next group from a global atomic counter and publishes what
// the consumers need to know about it next to the slot (meta[]), so SMs that see more bandwidth simply take more
// groups.  (A static round-robin persistent grid loses ~10 % of HBM bandwidth on B200 -- tools/membench.cu.)
// EDGE = rows that do not end on a chunk boundary (pitched rows); compiled separately so the aligned case pays nothing.
constexpr int kTightMaxLanes = 16;   // image lanes per group in TIGHT mode (rows >= 256 bytes, <= 256 consumer threads)
#ifndef B200BLUR_STORE_WARPS
#define B200BLUR_STORE_WARPS 4
#endif
#ifndef B200BLUR_STAGE_SLOTS
#define B200BLUR_STAGE_SLOTS 2
#endif
constexpr int kStoreWarps = B200BLUR_STORE_WARPS;       // TIGHT output: warps that flush the staged rows to (unaligned) global memory
constexpr int kStageSlots = B200BLUR_STAGE_SLOTS;       // ... output staging slots (one per input ring slot in flight between consumers and store warps)
constexpr int kStageRecs = 8;        // ... slot records in flight between producer and store warps (>= NS + kStageSlots)

// What a store warp needs to know about the output rows of one ring slot (written by the producer).
struct __align__(16) StageRec {
    uint8_t *out0;   // image lane 0: address of the slot's first output row (tight rows, any alignment)
    int n_img;
    int n_rows;      // output rows produced from this slot (0: nothing to flush)
    int stop;        // 1: no more slots
    int pad_[3];
};
static_assert(sizeof(StageRec) == 32, "StageRec is 32 bytes");

// Bytes [first, last) of the 16-byte value v to the 16-byte aligned address p, in pieces aligned to their own size.
__device__ __forceinline__ void store_bytes(uint8_t *p, const uint4 &v, int first, int last)
{
    int pos = first;
    while (pos < last) {
        const int align = pos ? (pos & -pos) : 16;
        int sz = 8;
        while (sz > align || sz > last - pos) sz >>= 1;
        const int wi = pos >> 2;
        const uint32_t w = wi == 0 ? v.x : wi == 1 ? v.y : wi == 2 ? v.z : v.w;
        if (sz == 8) {
            const uint32_t w2 = wi == 0 ? v.y : v.w;
            *reinterpret_cast<uint2 *>(p + pos) = make_uint2(w, w2);
        } else if (sz == 4) {
            *reinterpret_cast<uint32_t *>(p + pos) = w;
        } else if (sz == 2) {
            *reinterpret_cast<unsigned short *>(p + pos) = (unsigned short)(w >> (8 * (pos & 2)));
        } else {
            p[pos] = (uint8_t)(w >> (8 * (pos & 3)));
        }
        pos += sz;
    }
}

// TIGHT output: the store warps write `n_rows` staged rows (shared memory, pitch `spitch`, 16-byte aligned) of one image
// lane to the tight rows starting at the arbitrary byte address `g` -- one aligned 16-byte global word per lane and step,
// words m = first, first + stride, ... of the span.  Every word is built the same way, without divergence: two unaligned
// shared-memory reads -- the row that holds the word's first byte, and the next row read `k` bytes early so that its
// bytes sit at the same positions -- merged with byte masks (k = 16 for the words that lie inside one row, which then
// take everything from the first read).  Only the two words that stick out of the span (their other bytes belong to
// another slot, possibly another CTA) are written in narrower aligned pieces.
__device__ __forceinline__ void flush_rows_tight(uint8_t *g, uint32_t stage_lane, int n_rows, int row_bytes, unsigned row_recip,
                                                 int spitch, int first_word, int word_stride)
{
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
    const int head = (int)(ga & 15u);
    uint8_t *w0 = g - head;
    const int span = n_rows * row_bytes;
    const int n_words = (head + span + 15) >> 4;
    for (int m = first_word; m < n_words; m += word_stride) {
        const int b0 = 16 * m - head;
        const int bb = b0 < 0 ? 0 : b0;                              // first valid byte (b0 itself except for the first word)
        const unsigned r = __umulhi((unsigned)bb, row_recip);        // bb / row_bytes (a slot's span is < 2^16 bytes)
        const int col = b0 - (int)(r * (unsigned)row_bytes);         // may be negative for the first word
        int k = row_bytes - col;                                     // bytes of the word that row r holds
        k = k > 16 ? 16 : k;
        const uint32_t a1 = (uint32_t)((int)(stage_lane + r * (unsigned)spitch) + col);
        const uint32_t a2 = k < 16 ? stage_lane + (r + 1) * (unsigned)spitch - (unsigned)k : a1;
        const uint4 va = ptx::lds_unaligned16(a1), vb = ptx::lds_unaligned16(a2);
        const uint32_t aw[4] = {va.x, va.y, va.z, va.w}, bw[4] = {vb.x, vb.y, vb.z, vb.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int nb = k - 4 * i;
            const uint32_t mk = nb >= 4 ? 0xffffffffu : nb <= 0 ? 0u : (1u << (8 * nb)) - 1u;
            o[i] = (aw[i] & mk) | (bw[i] & ~mk);
        }
        const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
        uint8_t *dst = w0 + (size_t)m * 16;
        if (b0 >= 0 && b0 + 16 <= span) stg128_stream(dst, v);
        else store_bytes(dst, v, b0 < 0 ? -b0 : 0, span - b0 < 16 ? span - b0 : 16);   // first / last word of the span
    }
}

// TIGHT: 0 = 16-byte pitched/aligned rows on both sides; 1 = tight unaligned INPUT rows (re-aligned by the loads),
// pitched output; 2 = tight unaligned input AND output rows: the consumers stage their (row-relative, aligned) output
// chunks in shared memory and kStoreWarps extra warps write them out as aligned 16-byte global words.
template <int C, int RB, int NS, bool EDGE = false, bool FEED = false, int TIGHT = 0>
__global__ void __launch_bounds__((FEED ? 64 : TIGHT == 2 ? 32 + 32 * kStoreWarps : 32) + 256)
blur_stream_kernel(const StreamParams sp)
{
    constexpr bool TOUT = TIGHT == 2;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // layout: [16 B pad][NS slots][16 B pad][NS full barriers][NS empty barriers][NS group records]
    const uint32_t ring = ptx::smem_u32(smem_raw) + 16;
    const uint32_t full = ring + (uint32_t)(NS * sp.slot_bytes) + 16;
    const uint32_t empty = full + 8 * NS;
    GroupMeta *meta = reinterpret_cast<GroupMeta *>(smem_raw + 16 + (size_t)NS * sp.slot_bytes + 16 + 16 * NS);
    // FEED: [kFeedDepth "group done" barriers][kFeedDepth "record free" barriers][kFeedDepth records {slot, seq}]
    const uint32_t gdone = empty + 8 * NS + (uint32_t)sizeof(GroupMeta) * NS;
    const uint32_t gfree = gdone + 8 * kFeedDepth;
    int2 *grec = reinterpret_cast<int2 *>(reinterpret_cast<uint8_t *>(meta + NS) + 16 * kFeedDepth);
    // TIGHT: [NS][kTightMaxLanes][RB] byte offsets of the rows of a slot within their image lane (after the group records)
    unsigned short *rowoff = reinterpret_cast<unsigned short *>(meta + NS);
    // TIGHT output: [kStageSlots "staged" barriers][kStageSlots "flushed" barriers][kStageRecs slot records][16 B pad]
    // [kStageSlots staging slots of ipc x RB rows]
    uint8_t *tout_base = reinterpret_cast<uint8_t *>(rowoff + NS * kTightMaxLanes * RB);
    const uint32_t ofull = ptx::smem_u32(tout_base);
    const uint32_t oempty = ofull + 8 * kStageSlots;
    StageRec *orec = reinterpret_cast<StageRec *>(tout_base + 16 * kStageSlots);
    const uint32_t stage = ofull + 16 * kStageSlots + (uint32_t)sizeof(StageRec) * kStageRecs + 16;
    const int t = threadIdx.x;
    constexpr int LEAD = FEED ? 64 : TOUT ? 32 + 32 * kStoreWarps : 32;   // threads before the consumers
    const int n_cwarps = ((int)blockDim.x - LEAD) >> 5;
    if (t == 0) {
        for (int i = 0; i < NS; i++) {
            ptx::mbar_init(full + 8 * i, 1);
            ptx::mbar_init(empty + 8 * i, n_cwarps);
        }
        if (FEED)
            for (int i = 0; i < kFeedDepth; i++) {
                ptx::mbar_init(gdone + 8 * i, n_cwarps + 1);
                ptx::mbar_init(gfree + 8 * i, 1);
            }
        if (TOUT)
            for (int i = 0; i < kStageSlots; i++) {
                ptx::mbar_init(ofull + 8 * i, n_cwarps);
                ptx::mbar_init(oempty + 8 * i, kStoreWarps);
            }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    // Launched with programmatic stream serialisation: everything above overlapped the previous kernel's tail.
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();

    if (FEED && t >= 32 && t < 64) {
        // ------------------------------------------------------------------ accountant warp (FEED only)
        // Takes the completion of the groups off the consumers' path.  A GPU-scope fence waits for every store the SM
        // has in flight -- microseconds while 12 consumer warps stream -- so a consumer warp that fenced once per group
        // cost 30 % of the kernel's throughput, and one fence per group here would still be slower than the groups
        // arrive.  The accountant therefore takes ALL groups whose consumer warps have arrived, fences once, and then
        // adds each group's arrivals to its batch (the arrival that completes a batch raises the host flag).
        if (t == 32) {
            unsigned ga = 0;
            for (bool stop = false; !stop;) {
                int2 rec[kFeedDepth];
                ptx::mbar_wait(gdone + 8 * (ga % kFeedDepth), (ga / kFeedDepth) & 1);
                int n = 1;
                while (n < kFeedDepth && ptx::mbar_test(gdone + 8 * ((ga + n) % kFeedDepth), ((ga + n) / kFeedDepth) & 1)) n++;
#pragma unroll
                for (int i = 0; i < kFeedDepth; i++)
                    if (i < n) {
                        rec[i] = grec[(ga + i) % kFeedDepth];
                        ptx::mbar_arrive(gfree + 8 * ((ga + i) % kFeedDepth));
                    }
                ga += n;
                __threadfence();
                unsigned int old[kFeedDepth];
#pragma unroll
                for (int i = 0; i < kFeedDepth; i++)
                    if (i < n && rec[i].x >= 0) old[i] = atomicAdd(sp.feed_count + rec[i].x, (unsigned)n_cwarps);
#pragma unroll
                for (int i = 0; i < kFeedDepth; i++)
                    if (i < n) {
                        if (rec[i].x < 0) stop = true;
                        else feed_count_arrivals(sp, rec[i].x, (unsigned int)rec[i].y, old[i], (unsigned)n_cwarps);
                    }
            }
        }
        return;
    }
    if (TOUT && t >= 32 && t < LEAD) {
        // ------------------------------------------------------------------ store warps (TIGHT output only)
        // For every ring slot the consumers finish, flush the rows they staged.
        const int sw = (t >> 5) - 1, lane = t & 31;
        for (unsigned j = 0;; j++) {
            const int os = j % kStageSlots;
            ptx::mbar_wait(ofull + 8 * os, (j / kStageSlots) & 1);
            const StageRec rec = orec[j % kStageRecs];
            if (rec.stop) break;
            if (rec.n_rows > 0)
                for (int il = 0; il < rec.n_img; il++)   // every store warp takes its share of every image lane's words
                    flush_rows_tight(rec.out0 + (size_t)il * sp.b.out_stride,
                                     stage + (uint32_t)(os * sp.stage_slot_bytes + il * RB * sp.stage_pitch), rec.n_rows,
                                     sp.b.row_bytes, sp.row_recip, sp.stage_pitch, sw * 32 + lane, 32 * kStoreWarps);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(oempty + 8 * os);   // the staging slot may be overwritten
        }
        return;
    }
    if (t < 32) {
        // ------------------------------------------------------------------ producer warp
        if (t == 0) {
            unsigned pcount = 0;
            unsigned gp = 0;               // FEED: non-empty groups started by this CTA
            unsigned long long tail = 0;   // FEED: batches known to be published
            unsigned long long b_cur = 0, g_cur = 0;   // FEED: a batch this CTA has reached and its first group slot
            for (;;) {
                const long long g = (long long)atomicAdd(sp.work, 1ull);
                bool done;
                GroupGeom q;
                GroupMeta m;
                m.feed_slot = 0;
                m.feed_seq = 0;
                if (!FEED) {
                    done = g >= sp.n_groups;
                    if (!done) decode_group(sp, g, q, m.out);
                } else {
                    // group slot g belongs to batch g / feed_gpb of this launch; wait until the host has published it
                    // (group slots only grow, so the batch index advances from the last one with a 32-bit division)
                    const unsigned delta = (unsigned)((unsigned long long)g - g_cur);
                    const unsigned db = delta / (unsigned)sp.feed_gpb;
                    const unsigned local = delta - db * (unsigned)sp.feed_gpb;
                    b_cur += db;
                    g_cur += (unsigned long long)db * (unsigned)sp.feed_gpb;
                    const unsigned long long b = b_cur;
                    done = false;
                    if (b >= tail) {
                        const unsigned long long t0 = ptx::globaltimer_ns();
                        for (;;) {
                            tail = __ldcv(sp.feed_ctl);
                            if (b < tail) break;
                            if (__ldcv(sp.feed_ctl + 1)) {          // closed: the tail is final once `closed` is visible
                                tail = __ldcv(sp.feed_ctl);
                                done = b >= tail;
                                break;
                            }
                            if (ptx::globaltimer_ns() - t0 > sp.feed_timeout_ns) {   // host gone: never hang the GPU
                                atomicExch(sp.feed_ctl + 2, 1ull);
                                sp.feed_done[sp.feed_cap] = 1u;
                                __threadfence_system();
                                done = true;
                                break;
                            }
                            __nanosleep(256);
                        }
                        __threadfence();   // descriptors were written before the tail that covers them
                    }
                    if (!done) {
                        const unsigned long long seq = sp.feed_base + b;
                        const int slot = (int)(seq % (unsigned long long)sp.feed_cap);
                        const uint4 *dp = reinterpret_cast<const uint4 *>(sp.batches + slot);
                        const uint4 d0 = __ldcv(dp), d1 = __ldcv(dp + 1), d2 = __ldcv(dp + 2);
                        const uint8_t *bin = reinterpret_cast<const uint8_t *>(((unsigned long long)d0.y << 32) | d0.x);
                        uint8_t *bout = reinterpret_cast<uint8_t *>(((unsigned long long)d0.w << 32) | d0.z);
                        const uint8_t *btop = reinterpret_cast<const uint8_t *>(((unsigned long long)d1.y << 32) | d1.x);
                        const uint8_t *bbot = reinterpret_cast<const uint8_t *>(((unsigned long long)d1.w << 32) | d1.z);
                        const int n_images = (int)d2.x;
                        const unsigned per_block = (unsigned)(sp.nseg * sp.ncb);
                        const unsigned ib = local / per_block;
                        const int img0 = (int)ib * sp.ipc;
                        m.feed_slot = slot;
                        m.feed_seq = (unsigned int)seq + 1u;
                        if (img0 >= n_images) {               // a short batch: this group slot is empty
                            feed_arrive(sp, slot, m.feed_seq, (unsigned)n_cwarps);
                            continue;
                        }
                        decode_rows_cols(sp, local - ib * per_block, sp.seg, sp.b.rows, q);
                        q.n_img = min(sp.ipc, n_images - img0);
                        q.in = bin + (size_t)img0 * sp.b.in_stride;
                        q.top = btop ? btop + (size_t)img0 * sp.b.top_stride : nullptr;
                        q.bot = bbot ? bbot + (size_t)img0 * sp.b.bot_stride : nullptr;
                        m.out = bout + (size_t)img0 * sp.b.out_stride + (size_t)q.r0 * sp.b.out_pitch + q.x0;
                    }
                }
                if (FEED) {
                    // hand the group's record to the accountant (or tell it to stop)
                    const int i = gp % kFeedDepth;
                    ptx::mbar_wait(gfree + 8 * i, ((gp / kFeedDepth) & 1) ^ 1);
                    grec[i] = make_int2(done ? -1 : m.feed_slot, (int)m.feed_seq);
                    for (int a = done ? n_cwarps + 1 : 1; a > 0; a--) ptx::mbar_arrive(gdone + 8 * i);
                    gp++;
                }
                int nslots = 1;
                if (!done) {
                    nslots = (q.nr + 2 + RB - 1) / RB;
                    m.n_img = q.n_img;
                    m.nr = q.nr;
                    m.cbe = q.cbe;
                    const int chunk0 = q.x0 >> 4;
                    // bits 8..: 1 + the index within this column block of the row's second-to-last chunk, when that
                    // chunk's right-neighbour word holds the end of the row (EDGE rows ending < 4 bytes into a chunk)
                    const int pl = sp.cpr - 2 - chunk0;
                    m.flags = (q.left_edge ? 1 : 0) | (q.right_edge ? 2 : 0) |
                              ((EDGE && sp.edge_prev && pl >= 0 && pl < q.cbe) ? (pl + 1) << 8 : 0);
                } else {
                    m.out = nullptr;
                    m.n_img = -1;
                    m.nr = m.cbe = m.flags = 0;
                }
                for (int s = 0; s < nslots; s++, pcount++) {
                    const int buf = pcount % NS;
                    // wait until every consumer warp has released this buffer (passes at once on first use)
                    ptx::mbar_wait(empty + 8 * buf, ((pcount / NS) & 1) ^ 1);
                    if (s == 0) meta[buf] = m;
                    if (TOUT) {
                        // what the store warps will flush for this slot: output row k-2 comes from input row k >= 2
                        StageRec rec;
                        const int k0 = s * RB, k1 = done ? 0 : min(k0 + RB, q.nr + 2), kf = k0 > 2 ? k0 : 2;
                        rec.out0 = done ? nullptr : m.out + (size_t)(kf - 2) * (size_t)sp.b.out_pitch;
                        rec.n_img = done ? 0 : q.n_img;
                        rec.n_rows = k1 > kf ? k1 - kf : 0;
                        rec.stop = done ? 1 : 0;
                        rec.pad_[0] = rec.pad_[1] = rec.pad_[2] = 0;
                        orec[pcount % kStageRecs] = rec;
                    }
                    if (done) ptx::mbar_arrive(full + 8 * buf);   // sentinel slot: no data, tells the consumers to stop
                    else if (TIGHT)
                        stream_issue_slot_tight<RB>(sp, q, s, ring + (uint32_t)(buf * sp.slot_bytes), full + 8 * buf,
                                                    rowoff + buf * (kTightMaxLanes * RB));
                    else stream_issue_slot<RB>(sp, q, s, ring + (uint32_t)(buf * sp.slot_bytes), full + 8 * buf);
                }
                if (done) break;
            }
            // the last CTA to run out of work re-arms the counters for the next launch on this queue
            __threadfence();
            if (atomicAdd(sp.work + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
                sp.work[0] = 0;
                sp.work[1] = 0;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int ct = t - LEAD;
    const int il = ct / sp.cb;           // image lane within the group
    const int c = ct - il * sp.cb;       // chunk within the column block
    const int lane = t & 31;
    unsigned ccount = 0;                 // slots consumed so far
    unsigned gc = 0;                     // FEED: groups finished by this warp
    for (;;) {
        // the first slot of an item carries the group record
        ptx::mbar_wait(full + 8 * (ccount % NS), (ccount / NS) & 1);
        const GroupMeta m = meta[ccount % NS];
        if (m.n_img < 0) {
            if (TOUT) {   // pass the stop record on to the store warps
                ptx::mbar_wait(oempty + 8 * (ccount % kStageSlots), ((ccount / kStageSlots) & 1) ^ 1);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ofull + 8 * (ccount % kStageSlots));
            }
            break;
        }
        const bool active = (il < m.n_img) && (c < m.cbe);
        const bool first = (m.flags & 1) && (c == 0);
        const bool last = (m.flags & 2) && (c == m.cbe - 1);               // the chunk that holds the end of the row
        const bool prev_last = EDGE && ((m.flags >> 8) == c + 1);
        const int nslots = (m.nr + 2 + RB - 1) / RB;
        const int il_c = active ? il : 0, c_c = active ? c : 0;
        const uint32_t lane_off = TIGHT ? (uint32_t)(il_c * sp.lane_bytes + 16 + c_c * 16)
                                        : (uint32_t)(il_c * RB * sp.sstride + sp.margin + c_c * 16);
        // Output row k-2 is produced when input row k of the item arrives: `dst` points two rows early.
        uint8_t *dst = m.out + (size_t)il_c * sp.b.out_stride + c_c * 16 - 2 * (ptrdiff_t)sp.b.out_pitch;
        // Rolling vertical state, pre-scaled by 16: before row k arrives
        //   accA = 16*(h[k-2] + 2*h[k-1])   accB = 16*h[k-1]        (<= 48960 per 16-bit lane)
        uint32_t accA[8], accB[8];
#pragma unroll
        for (int i = 0; i < 8; i++) accA[i] = accB[i] = 0;
        // one input row: horizontal sums, vertical roll, one 16-byte store (if `store`)
        auto row = [&](uint32_t a, uint8_t *out_row, bool store) {
            uint4 w;
            uint32_t wl, wr;
            if (!TIGHT) {
                w = ptx::lds128(a);
                wl = ptx::lds32(a - 4);
                wr = ptx::lds32(a + 16);
            } else {
                // the 24-byte window {wl, w, wr} starts at the arbitrary byte address a - 4
                const uint32_t a4 = (a - 4u) & ~3u, sh = ((a - 4u) & 3u) * 8u;
                uint32_t x[7];
#pragma unroll
                for (int i = 0; i < 7; i++) x[i] = ptx::lds32(a4 + 4u * i);
                wl = __funnelshift_r(x[0], x[1], sh);
                w.x = __funnelshift_r(x[1], x[2], sh);
                w.y = __funnelshift_r(x[2], x[3], sh);
                w.z = __funnelshift_r(x[3], x[4], sh);
                w.w = __funnelshift_r(x[4], x[5], sh);
                wr = __funnelshift_r(x[5], x[6], sh);
            }
            if (first) wl = w.x << (8 * (4 - C));   // clamp: pixel -1 := pixel 0        (gaussian_kernel.cl:56)
            if (!EDGE) {
                if (last) wr = w.w >> (8 * (4 - C));    // clamp: pixel width := pixel width-1
            } else if (last) {                          // row ends inside this chunk: rewrite the window
                const uint32_t n0 = __byte_perm(wl, w.x, sp.sel_last[1]), n1 = __byte_perm(w.x, w.y, sp.sel_last[2]);
                const uint32_t n2 = __byte_perm(w.y, w.z, sp.sel_last[3]), n3 = __byte_perm(w.z, w.w, sp.sel_last[4]);
                wr = __byte_perm(w.w, wr, sp.sel_last[5]);
                w.x = n0; w.y = n1; w.z = n2; w.w = n3;
            } else if (prev_last) {
                wr = __byte_perm(w.w, wr, sp.sel_prev);
            }
            uint32_t h[8];
            hpass8<C>(w, wl, wr, h);
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                v[i] = h[i] * 16u + accA[i];       // 16*(h[k-2] + 2*h[k-1] + h[k])  <= 65280 per lane
                accA[i] = h[i] * 32u + accB[i];
                accB[i] = h[i] << 4;
            }
            uint4 o;
            o.x = __byte_perm(v[0], v[1], 0x7351);
            o.y = __byte_perm(v[2], v[3], 0x7351);
            o.z = __byte_perm(v[4], v[5], 0x7351);
            o.w = __byte_perm(v[6], v[7], 0x7351);
            if (store) {
                if (!TOUT) stg128_stream(out_row, o);
                else ptx::sts128((uint32_t)reinterpret_cast<uintptr_t>(out_row), o);   // out_row = staging address (TIGHT output)
            }
        };
        const int k_end = m.nr + 2;
        int k = 0;  // input-row index within the item
        for (int s = 0; s < nslots; s++, ccount++) {
            const int buf = ccount % NS;
            if (s > 0) ptx::mbar_wait(full + 8 * buf, (ccount / NS) & 1);
            const uint32_t a = ring + (uint32_t)(buf * sp.slot_bytes) + lane_off;
            // TIGHT: where each row of this image lane starts (written by the producer before it released the slot)
            const unsigned short *ro = rowoff + buf * (kTightMaxLanes * RB) + il_c * RB;
            auto row_addr = [&](int r) -> uint32_t {
                return TIGHT ? a + (uint32_t)ro[r] : a + (uint32_t)r * (uint32_t)sp.sstride;
            };
            if (TOUT) {
                // this slot's output goes to staging slot ccount % kStageSlots, once the store warps are done with it;
                // the row produced from input row k of the group lands in staging row k - max(first k of the slot, 2)
                const int os = ccount % kStageSlots;
                ptx::mbar_wait(oempty + 8 * os, ((ccount / kStageSlots) & 1) ^ 1);
                const int kf = k > 2 ? k : 2;
                dst = reinterpret_cast<uint8_t *>((uintptr_t)(stage + (uint32_t)(os * sp.stage_slot_bytes) +
                                                              (uint32_t)((il_c * RB + (k - kf)) * sp.stage_pitch + c_c * 16)));
            }
            // Whole slots (the planner sizes groups so that nearly all are) run fully unrolled with compile-time row
            // offsets and no per-row trip test; the first slot of a group only differs in not storing rows 0 and 1.
            const size_t out_step = TOUT ? (size_t)sp.stage_pitch : (size_t)sp.b.out_pitch;
            if (k_end - k >= RB) {
                if (s == 0) {
#pragma unroll
                    for (int r = 0; r < RB; r++) row(row_addr(r), dst + (size_t)r * out_step, active && r >= 2);
                } else {
#pragma unroll
                    for (int r = 0; r < RB; r++) row(row_addr(r), dst + (size_t)r * out_step, active);
                }
                k += RB;
            } else {
                const int n = k_end - k;                // short last slot of a group (uniform across the CTA)
#pragma unroll 1
                for (int r = 0; r < n; r++) row(row_addr(r), dst + (size_t)r * out_step, active && k + r >= 2);
                k += n;
            }
            dst += (size_t)RB * out_step;
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(empty + 8 * buf);   // this warp is done reading the slot
                if (TOUT) ptx::mbar_arrive(ofull + 8 * (ccount % kStageSlots));   // ... and has staged its output rows
            }
        }
        if (FEED) {
            // this warp's stores of the group are issued: arrive (release, CTA scope) for the accountant
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(gdone + 8 * (gc % kFeedDepth));
            gc++;
        }
    }
}

// ------------------------------------------------------------------------------- re-pitch kernels (odd image widths)
// Tight rows whose length is not a multiple of 16 cannot be read or written 16 bytes at a time, and the copy engines'
// strided (2-D) copies crawl on short rows (6.5 GB/s over the host link for 750-byte rows, 45 GB/s linear).  These two
// kernels move rows between the tight layout and a 16-byte-multiple pitch at memory speed: every thread produces ONE
// aligned 16-byte destination word from five 4-byte-aligned source words and four funnel shifts.
__device__ __forceinline__ uint4 load_unaligned16(const uint8_t *p, const uint8_t *end)
{
    // 16 bytes starting at the arbitrary address p; words at or past `end` (rounded up to 4) read as 0
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t *e = reinterpret_cast<const uint32_t *>((reinterpret_cast<uintptr_t>(end) + 3) & ~(uintptr_t)3);
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; i++) v[i] = (q + i < e) ? __ldg(q + i) : 0u;
    uint4 r;
    r.x = __funnelshift_r(v[0], v[1], sh);
    r.y = __funnelshift_r(v[1], v[2], sh);
    r.z = __funnelshift_r(v[2], v[3], sh);
    r.w = __funnelshift_r(v[3], v[4], sh);
    return r;
}

// tight [rows][row_bytes] -> pitched [rows][pitch].  blockIdx.y = block of kRepitchRows rows (all index arithmetic inside
// a block of rows is 32-bit: a 64-bit division per 16 bytes made these kernels ALU-bound), blockIdx.x/threads = 16-byte
// chunks of those rows.
constexpr int kRepitchRows = 128;
__global__ void __launch_bounds__(256)
repitch_in_kernel(const uint8_t *__restrict__ tight, uint8_t *__restrict__ pitched, long long rows, int row_bytes, int pitch)
{
    const long long row0 = (long long)blockIdx.y * kRepitchRows;
    const int nrows = (int)min((long long)kRepitchRows, rows - row0);
    const unsigned cpr = (unsigned)(row_bytes + 15) / 16;
    const unsigned total = (unsigned)nrows * cpr;
    const uint8_t *end = tight + rows * (long long)row_bytes;
    const uint8_t *src = tight + row0 * (long long)row_bytes;
    uint8_t *dst = pitched + row0 * (long long)pitch;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        const unsigned r = g / cpr, c = g - r * cpr;
        const uint4 v = load_unaligned16(src + (size_t)r * row_bytes + 16 * c, end);
        *reinterpret_cast<uint4 *>(dst + (size_t)r * pitch + 16 * c) = v;
    }
}

// pitched [rows][pitch] -> rows [lo/row_bytes, ...) of a tight buffer; one thread per aligned 16-byte word of the tight
// buffer.  `tight` is the 16-byte aligned base, [lo, lo + rows*row_bytes) the byte range that belongs to these rows: words
// only partly inside a row block's range (its two ends: the neighbouring block writes the other bytes) and words that
// straddle two rows are written byte by byte.
__global__ void __launch_bounds__(256)
repitch_out_kernel(const uint8_t *__restrict__ pitched, uint8_t *__restrict__ tight, long long lo, long long rows, int row_bytes,
                   int pitch)
{
    const long long row0 = (long long)blockIdx.y * kRepitchRows;
    const unsigned nrows = (unsigned)min((long long)kRepitchRows, rows - row0);
    const long long blo = lo + row0 * (long long)row_bytes;          // this row block's byte range of the tight buffer
    const unsigned span = nrows * (unsigned)row_bytes;
    const long long w0 = blo / 16;                                    // first aligned word that holds a byte of the range
    const unsigned head = (unsigned)(blo - w0 * 16);                  // bytes of that word before the range
    const unsigned n_words = (head + span + 15) / 16;
    const uint8_t *src = pitched + row0 * (long long)pitch;
    const uint8_t *end = pitched + rows * (long long)pitch;
    uint8_t *dst = tight + w0 * 16;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < n_words; g += gridDim.x * blockDim.x) {
        const int b0 = (int)(g * 16) - (int)head;                     // offset of the word's first byte within the range
        if (b0 >= 0 && (unsigned)b0 + 16 <= span) {
            const unsigned r = (unsigned)b0 / (unsigned)row_bytes, col = (unsigned)b0 - r * (unsigned)row_bytes;
            if (col + 16 <= (unsigned)row_bytes) {                    // the whole word comes from one row
                *reinterpret_cast<uint4 *>(dst + (size_t)g * 16) = load_unaligned16(src + (size_t)r * pitch + col, end);
                continue;
            }
            // The word straddles rows r and r+1 (one word per row does: a byte loop here would stall every warp).
            // First k bytes = the end of row r, the other 16-k = the start of row r+1, read k bytes early so that they
            // sit at the same byte positions; merge with byte masks.
            const unsigned k = (unsigned)row_bytes - col;             // 1..15
            const uint4 a = load_unaligned16(src + (size_t)r * pitch + col, end);
            const uint4 bq = load_unaligned16(src + (size_t)(r + 1) * pitch - k, end);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int nb = (int)k - 4 * i;                        // bytes of word i that come from row r
                const uint32_t m = nb >= 4 ? 0xffffffffu : nb <= 0 ? 0u : (1u << (8 * nb)) - 1u;
                o[i] = (aw[i] & m) | (bw[i] & ~m);
            }
            *reinterpret_cast<uint4 *>(dst + (size_t)g * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            continue;
        }
        for (int i = 0; i < 16; i++) {
            const int b = b0 + i;
            if (b < 0 || (unsigned)b >= span) continue;
            const unsigned rr = (unsigned)b / (unsigned)row_bytes;
            dst[(size_t)g * 16 + i] = src[(size_t)rr * pitch + ((unsigned)b - rr * (unsigned)row_bytes)];
        }
    }
}

// ------------------------------------------------------------------------------------- generic path (any shape)
// One thread per output byte, grid-stride, exact integer arithmetic.  Used when width*channels is not a multiple
// of 16, channels > 4, or a pointer/stride is not 16-byte aligned.  Same results, no alignment requirements.
__global__ void __launch_bounds__(256)
blur_generic_kernel(const BandParams p)
{
    const long long per_image = (long long)p.rows * p.row_bytes;
    const long long total = per_image * p.n_images;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += step) {
        const long long img = g / per_image;
        const int rem = (int)(g - img * per_image);
        const int r = rem / p.row_bytes;
        const int b = rem - r * p.row_bytes;
        const int x = b / p.channels;
        const int bl = (x > 0) ? b - p.channels : b;
        const int br = (x < p.width - 1) ? b + p.channels : b;
        const uint8_t *src = p.in + (size_t)img * p.in_stride;
        const uint8_t *mid = src + (size_t)r * p.pitch;
        const uint8_t *up = (r > 0) ? mid - p.pitch
                                    : (p.halo_top ? p.halo_top + (size_t)img * p.top_stride : mid);
        const uint8_t *dn = (r < p.rows - 1) ? mid + p.pitch
                                             : (p.halo_bot ? p.halo_bot + (size_t)img * p.bot_stride : mid);
        const int hu = up[bl] + 2 * up[b] + up[br];
        const int hm = mid[bl] + 2 * mid[b] + mid[br];
        const int hd = dn[bl] + 2 * dn[b] + dn[br];
        p.out[(size_t)img * p.out_stride + (size_t)r * p.out_pitch + b] = (uint8_t)((hu + 2 * hm + hd) >> 4);
    }
}

}  // namespace b200blur
