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
    int pitch;                // bytes per row = width * channels
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
        hE[k - 1] = 2u * E[k] + LE + RE;
        hO[k - 1] = 2u * O[k] + LO + RO;
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
            if (valid && r < p.rows) stg128_stream(dst + (size_t)r * p.pitch, o);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h2E[i] = h1E[i]; h2O[i] = h1O[i];
            h1E[i] = hE[i];  h1O[i] = hO[i];
        }
    }
}

// ------------------------------------------------------------------------------------- generic path (any shape)
// One thread per output byte, grid-stride, exact integer arithmetic.  Used when width*channels is not a multiple
// of 16, channels > 4, or a pointer/stride is not 16-byte aligned.  Same results, no alignment requirements.
__global__ void __launch_bounds__(256)
blur_generic_kernel(const BandParams p)
{
    const long long per_image = (long long)p.rows * p.pitch;
    const long long total = per_image * p.n_images;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += step) {
        const long long img = g / per_image;
        const int rem = (int)(g - img * per_image);
        const int r = rem / p.pitch;
        const int b = rem - r * p.pitch;
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
        p.out[(size_t)img * p.out_stride + (size_t)r * p.pitch + b] = (uint8_t)((hu + 2 * hm + hd) >> 4);
    }
}

}  // namespace b200blur
