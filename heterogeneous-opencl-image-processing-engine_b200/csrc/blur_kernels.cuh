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
// The Blackwell-native form.  Persistent CTAs (a few per SM) each stream "items" -- a segment of `seg` output rows of
// a column block of `ipc` consecutive images -- through a ring of NS shared-memory slots.  One elected thread feeds
// the ring with 1-D bulk async copies (cp.async.bulk global->shared, completion on an mbarrier; SASS UBLKCP): because
// a small frame's rows are contiguous in memory, RB rows of an image arrive as ONE bulk copy; wide frames use one
// bulk copy per row of a 2 KB column block (+16 B margins).  All threads then read their 16-byte chunk and the two
// neighbouring words from shared memory (no shuffles, no edge lanes), keep the rolling horizontal sums of the two
// previous rows in registers across slots, and emit one coalesced 16-byte store per output row.  In-flight bytes are
// held by the ring (NS-1 slots per CTA), not by registers, and a segment re-reads only 2 halo rows per `seg` rows.
// Halo rows above/below the band come from halo_top/halo_bot -- possibly another GPU's memory (NVLink) -- or are the
// replicated edge row (gaussian_kernel.cl:57).
struct StreamParams {
    BandParams b;
    int cpr;            // 16-byte chunks per row
    int cb;             // chunks per column block (<= blockDim.x)
    int ncb;            // column blocks per row
    int ipc;            // images side by side in one CTA step (ncb == 1 only)
    int seg;            // output rows per item
    int nseg;           // segments per band
    int margin;         // 0 (full-width rows, contiguous copies) or 16 (column blocks, per-row copies)
    int sstride;        // shared-memory row stride in bytes = cb*16 + 2*margin
    int slot_bytes;     // ipc * RB * sstride
    long long img_blocks;   // ceil(n_images / ipc)
    long long n_groups;     // img_blocks * nseg * ncb
    unsigned long long *work;  // work[0] = next group to hand out, work[1] = CTAs finished (both 0 between launches)
    // Right edge of a row that does not end on a chunk boundary (row_bytes % 16 != 0, pitched rows).  The clamp
    // "pixel width := pixel width-1" means window bytes [row_bytes, row_bytes + C) := bytes [row_bytes - C, row_bytes).
    // For the chunk that contains the row end (and the one before it when the end is < 4 bytes into the last chunk)
    // the 24-byte window {wl, w, wr} is rewritten with one PRMT per word; the selectors are computed on the host.
    int edge_general;          // 0: rows end on a chunk boundary (fast path)
    int edge_prev;             // 1: the chunk before the last one needs its wr word patched too
    uint32_t sel_last[6];      // PRMT selectors for window words 0..5 of the last chunk (pairs: previous word, word)
    uint32_t sel_prev;         // PRMT selector for the wr word of the chunk before the last
};

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
}  // namespace ptx

// Decodes group index -> geometry.  Groups are ordered image-block major, then segment, then column block, so CTAs
// that run concurrently work on adjacent segments of the same images and the shared halo rows hit in L2.
struct GroupGeom {
    long long img0;   // first image of the group
    int n_img;        // images in the group (<= ipc)
    int r0;           // first output row
    int nr;           // output rows
    int x0;           // first byte column of the column block
    int chunk0;       // index of its first chunk within the row
    int cbe;          // chunks in this column block
    bool left_edge, right_edge;
};
__device__ __forceinline__ GroupGeom decode_group(const StreamParams &sp, long long g64)
{
    GroupGeom q;
    // n_groups < 2^31 (checked on the host: a group is at least a few KB), so 32-bit division -- inlined, no call
    const unsigned g = (unsigned)g64;
    const unsigned per_block = (unsigned)(sp.nseg * sp.ncb);
    const unsigned ib = g / per_block;
    const int sc = (int)(g - ib * per_block);
    const int si = sc / sp.ncb;
    const int ci = sc - si * sp.ncb;
    q.img0 = (long long)ib * sp.ipc;
    const long long left = sp.b.n_images - q.img0;
    q.n_img = left < sp.ipc ? (int)left : sp.ipc;
    q.r0 = si * sp.seg;
    q.nr = min(sp.seg, sp.b.rows - q.r0);
    q.x0 = ci * sp.cb * 16;
    q.chunk0 = ci * sp.cb;
    q.cbe = min(sp.cb, sp.cpr - ci * sp.cb);
    q.left_edge = (ci == 0);
    q.right_edge = (ci == sp.ncb - 1);
    return q;
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
        const size_t img = (size_t)(q.img0 + il);
        const uint8_t *src = b.in + img * b.in_stride;
        const uint32_t dst_img = slot_smem + (uint32_t)(il * RB * sp.sstride);
        int k = k0;
        while (k < k1) {
            const int j = q.r0 - 1 + k;  // input row relative to the band
            const uint8_t *rp;
            int run = 1;
            if (j < 0) {
                rp = b.halo_top ? b.halo_top + img * b.top_stride : src;
            } else if (j >= b.rows) {
                rp = b.halo_bot ? b.halo_bot + img * b.bot_stride : src + (size_t)(b.rows - 1) * b.pitch;
            } else {
                rp = src + (size_t)j * b.pitch;
                if (sp.margin == 0) run = min(k1 - k, b.rows - j);  // contiguous rows: one copy
            }
            // full-width: `run` whole rows as they lie in memory (padding included); a halo row brings only its live chunks
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

// Warp-specialised: warp 0 is the producer (one elected lane issues the bulk copies and never computes), warps 1..
// are consumers.  full[NS] barriers carry the copies' byte counts; empty[NS] barriers collect one arrival per consumer
// warp, so consumer warps never wait for each other -- only for data.
// Work is handed out dynamically: the producer takes the next group from a global atomic counter and publishes its
// index next to the slot (meta[]), so SMs that see more bandwidth simply take more groups.  (A static round-robin
// persistent grid loses ~10 % of HBM bandwidth on B200 -- tools/membench.cu, profiles/membench_r01.txt.)
// EDGE = rows that do not end on a chunk boundary (pitched rows); compiled separately so the aligned case pays nothing.
template <int C, int RB, int NS, bool EDGE = false>
__global__ void __launch_bounds__(32 + 256)
blur_stream_kernel(const StreamParams sp)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // layout: [16 B pad][NS slots][16 B pad][NS full barriers][NS empty barriers][NS group indices]
    const uint32_t ring = ptx::smem_u32(smem_raw) + 16;
    const uint32_t full = ring + (uint32_t)(NS * sp.slot_bytes) + 16;
    const uint32_t empty = full + 8 * NS;
    volatile long long *meta = reinterpret_cast<volatile long long *>(smem_raw + 16 + (size_t)NS * sp.slot_bytes + 16 + 16 * NS);
    const int t = threadIdx.x;
    const int n_cwarps = (blockDim.x >> 5) - 1;
    if (t == 0) {
        for (int i = 0; i < NS; i++) {
            ptx::mbar_init(full + 8 * i, 1);
            ptx::mbar_init(empty + 8 * i, n_cwarps);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (t < 32) {
        // ------------------------------------------------------------------ producer warp
        if (t == 0) {
            unsigned pcount = 0;
            for (;;) {
                const long long g = (long long)atomicAdd(sp.work, 1ull);
                const bool done = g >= sp.n_groups;
                GroupGeom q;
                int nslots = 1;
                if (!done) {
                    q = decode_group(sp, g);
                    nslots = (q.nr + 2 + RB - 1) / RB;
                }
                for (int s = 0; s < nslots; s++, pcount++) {
                    const int buf = pcount % NS;
                    // wait until every consumer warp has released this buffer (passes at once on first use)
                    ptx::mbar_wait(empty + 8 * buf, ((pcount / NS) & 1) ^ 1);
                    if (s == 0) meta[buf] = done ? -1 : g;
                    if (done) ptx::mbar_arrive(full + 8 * buf);   // sentinel slot: no data, tells the consumers to stop
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
    const int ct = t - 32;
    const int il = ct / sp.cb;           // image lane within the group
    const int c = ct - il * sp.cb;       // chunk within the column block
    const int lane = t & 31;
    unsigned ccount = 0;                 // slots consumed so far
    for (;;) {
        // the first slot of an item carries the group index
        ptx::mbar_wait(full + 8 * (ccount % NS), (ccount / NS) & 1);
        const long long g = meta[ccount % NS];
        if (g < 0) break;
        const GroupGeom q = decode_group(sp, g);
        const bool active = (il < q.n_img) && (c < q.cbe);
        const bool first = q.left_edge && (c == 0);
        const bool last = q.right_edge && (c == q.cbe - 1);               // the chunk that holds the end of the row
        const bool prev_last = EDGE && sp.edge_prev && (q.chunk0 + c == sp.cpr - 2);
        const int nslots = (q.nr + 2 + RB - 1) / RB;
        const int il_c = active ? il : 0, c_c = active ? c : 0;
        const uint32_t lane_off = (uint32_t)(il_c * RB * sp.sstride + sp.margin + c_c * 16);
        // Output row k-2 is produced when input row k of the item arrives; the store pointer starts two rows early
        // and advances every row so the loop body has no branches (stores for k < 2 and k >= nr+2 are predicated off).
        uint8_t *dst = sp.b.out + (size_t)(q.img0 + il_c) * sp.b.out_stride + (size_t)q.r0 * sp.b.out_pitch + q.x0 + c_c * 16 -
                       2 * (ptrdiff_t)sp.b.out_pitch;
        // Rolling vertical state, pre-scaled by 16: before row k arrives
        //   accA = 16*(h[k-2] + 2*h[k-1])   accB = 16*h[k-1]        (<= 48960 per 16-bit lane)
        uint32_t accA[8], accB[8];
#pragma unroll
        for (int i = 0; i < 8; i++) accA[i] = accB[i] = 0;
        const int k_end = q.nr + 2;
        int k = 0;  // input-row index within the item
        for (int s = 0; s < nslots; s++, ccount++) {
            const int buf = ccount % NS;
            if (s > 0) ptx::mbar_wait(full + 8 * buf, (ccount / NS) & 1);
            uint32_t a = ring + (uint32_t)(buf * sp.slot_bytes) + lane_off;
#pragma unroll
            for (int r = 0; r < RB; r++) {
                if (k >= k_end) break;                  // short last slot of an item (uniform across the CTA)
                uint4 w = ptx::lds128(a);
                uint32_t wl = ptx::lds32(a - 4);
                uint32_t wr = ptx::lds32(a + 16);
                a += sp.sstride;
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
                if (active && k >= 2) stg128_stream(dst, o);
                dst += sp.b.out_pitch;
                k++;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty + 8 * buf);   // this warp is done reading the slot
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

// tight [rows][row_bytes] -> pitched [rows][pitch]; grid.x covers chunks of a row, grid.y/z-free: rows flattened in x.
__global__ void __launch_bounds__(256)
repitch_in_kernel(const uint8_t *__restrict__ tight, uint8_t *__restrict__ pitched, long long rows, int row_bytes, int pitch)
{
    const int cpr = (row_bytes + 15) / 16;
    const long long total = rows * cpr;
    const uint8_t *end = tight + rows * (long long)row_bytes;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const long long r = g / cpr;
        const int c = (int)(g - r * cpr);
        const uint4 v = load_unaligned16(tight + r * (long long)row_bytes + 16 * c, end);
        *reinterpret_cast<uint4 *>(pitched + r * (long long)pitch + 16 * c) = v;
    }
}

// pitched [rows][pitch] -> rows [lo/row_bytes, ...) of a tight buffer; one thread per aligned 16-byte word of the tight
// buffer.  `tight` is the 16-byte aligned base, [lo, lo + rows*row_bytes) the byte range that belongs to these rows: words
// only partly inside the range (its two ends) and words that straddle two rows are written byte by byte.
__global__ void __launch_bounds__(256)
repitch_out_kernel(const uint8_t *__restrict__ pitched, uint8_t *__restrict__ tight, long long lo, long long rows, int row_bytes,
                   int pitch)
{
    const long long hi = lo + rows * (long long)row_bytes;
    const long long w0 = lo / 16, w1 = (hi + 15) / 16;
    const uint8_t *end = pitched + rows * (long long)pitch;
    for (long long g = w0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; g < w1; g += (long long)gridDim.x * blockDim.x) {
        const long long b0 = g * 16;
        const long long r = (b0 - lo) / row_bytes;          // row within this range (valid when b0 >= lo)
        const long long col = (b0 - lo) - r * row_bytes;
        if (b0 >= lo && b0 + 16 <= hi && col + 16 <= row_bytes) {   // the whole word comes from one row
            const uint4 v = load_unaligned16(pitched + r * (long long)pitch + col, end);
            *reinterpret_cast<uint4 *>(tight + b0) = v;
        } else {
            for (int i = 0; i < 16; i++) {
                const long long b = b0 + i;
                if (b < lo || b >= hi) continue;
                const long long rr = (b - lo) / row_bytes;
                tight[b] = pitched[rr * (long long)pitch + ((b - lo) - rr * row_bytes)];
            }
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
