// jpeg_decode.hpp -- dependency-free baseline JPEG reader for the two drop-in CLIs (the ingest
// step right before the hot path).
//
// The reference loads its one source image with CImg<unsigned char>("./image_320x240.jpg") (heterogeneous_blur.c:106),
// i.e. libjpeg with its defaults, and interleaves the planes (:128-135).  This reader restates that decode so the CLIs
// can open the reference's own .jpg files and feed the blur the SAME bytes: sequential (baseline / extended) Huffman
// JPEG, 8-bit, 1 or 3 components, chroma subsampled 4:4:4 / 4:2:2 / 4:2:0, restart intervals -- decoded the way
// libjpeg does by default: the slow-but-accurate integer IDCT (jidctint.c, 13-bit constants, two passes), "fancy"
// triangle-filter chroma upsampling (jdsample.c h2v1 / h2v2 fancy upsample) and the fixed-point YCbCr -> RGB tables of
// jdcolor.c.  tests/test_jpeg_cpu.py checks it byte for byte against libjpeg-turbo (Pillow) on committed fixtures.
// Not supported (clean error): progressive / arithmetic / lossless / 12-bit / CMYK files.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace jpegdec {

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocks_w = 0, blocks_h = 0;   // blocks in the MCU-padded plane
    int width = 0, height = 0;        // true (downsampled) size
    int stride = 0;
    int pred = 0;
    std::vector<uint8_t> plane;
};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    // 9-bit lookahead: value and code length, 0 length = longer code
    uint8_t look_len[512], look_val[512];
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        memset(look_len, 0, sizeof look_len);
        int c = 0;
        k = 0;
        for (int l = 1; l <= 9; l++) {
            for (int i = 0; i < bits[l]; i++, k++, c++) {
                const int first = c << (9 - l);
                for (int j = 0; j < (1 << (9 - l)); j++) {
                    look_len[first + j] = (uint8_t)l;
                    look_val[first + j] = vals[k];
                }
            }
            c <<= 1;
        }
    }
};

class Decoder {
public:
    // Decodes `data` into interleaved rows (RGB, or one channel for greyscale files).  Returns "" or an error message.
    std::string decode(const uint8_t *data, size_t size, int &width, int &height, int &channels, std::vector<uint8_t> &out)
    {
        d_ = data;
        n_ = size;
        pos_ = 0;
        if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return "not a JPEG file (no SOI marker)";
        pos_ = 2;
        bool have_frame = false, done = false;
        while (!done) {
            int m = next_marker();
            if (m < 0) return "unexpected end of file";
            switch (m) {
                case 0xD8: break;
                case 0xD9: done = true; break;
                case 0xC0: case 0xC1: {
                    if (std::string e = read_frame(); !e.empty()) return e;
                    have_frame = true;
                    break;
                }
                case 0xC2: return "progressive JPEG is not supported (baseline Huffman only)";
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    return "unsupported JPEG process (lossless / hierarchical / arithmetic)";
                case 0xC4: if (std::string e = read_dht(); !e.empty()) return e; break;
                case 0xDB: if (std::string e = read_dqt(); !e.empty()) return e; break;
                case 0xDD: {
                    if (pos_ + 4 > n_) return "truncated DRI";
                    restart_interval_ = (d_[pos_ + 2] << 8) | d_[pos_ + 3];
                    pos_ += 4;
                    break;
                }
                case 0xDA: {
                    if (!have_frame) return "SOS before SOF";
                    if (std::string e = read_scan(); !e.empty()) return e;
                    done = scans_done_;
                    break;
                }
                default: {   // APPn, COM, anything else with a length
                    if (pos_ + 2 > n_) return "truncated segment";
                    const size_t len = (d_[pos_] << 8) | d_[pos_ + 1];
                    if (len < 2 || pos_ + len > n_) return "bad segment length";
                    pos_ += len;
                }
            }
        }
        if (!have_frame || !scans_done_) return "no image data";
        width = width_;
        height = height_;
        channels = (int)comps_.size() == 1 ? 1 : 3;
        return finish(out);
    }

private:
    const uint8_t *d_ = nullptr;
    size_t n_ = 0, pos_ = 0;
    int width_ = 0, height_ = 0, hmax_ = 1, vmax_ = 1, mcus_x_ = 0, mcus_y_ = 0;
    int restart_interval_ = 0;
    bool scans_done_ = false;
    int comps_in_scans_ = 0;
    uint16_t quant_[4][64] = {};
    bool quant_ok_[4] = {false, false, false, false};
    HuffTable dc_[4], ac_[4];
    std::vector<Component> comps_;
    // bit reader
    uint32_t bitbuf_ = 0;
    int bitcnt_ = 0;
    bool hit_marker_ = false;

    int next_marker()
    {
        while (pos_ < n_) {
            if (d_[pos_] != 0xFF) { pos_++; continue; }
            while (pos_ < n_ && d_[pos_] == 0xFF) pos_++;
            if (pos_ >= n_) return -1;
            const int m = d_[pos_++];
            if (m != 0) return m;
        }
        return -1;
    }

    std::string read_dqt()
    {
        if (pos_ + 2 > n_) return "truncated DQT";
        size_t len = (d_[pos_] << 8) | d_[pos_ + 1];
        if (pos_ + len > n_) return "truncated DQT";
        size_t p = pos_ + 2;
        const size_t end = pos_ + len;
        static const int zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                   15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        while (p < end) {
            const int pq = d_[p] >> 4, tq = d_[p] & 15;
            p++;
            if (tq > 3) return "bad quantisation table id";
            if (p + (pq ? 128 : 64) > end) return "truncated DQT";
            for (int i = 0; i < 64; i++) {
                quant_[tq][zz[i]] = pq ? (uint16_t)((d_[p] << 8) | d_[p + 1]) : d_[p];
                p += pq ? 2 : 1;
            }
            quant_ok_[tq] = true;
        }
        pos_ = end;
        return "";
    }

    std::string read_dht()
    {
        if (pos_ + 2 > n_) return "truncated DHT";
        size_t len = (d_[pos_] << 8) | d_[pos_ + 1];
        if (pos_ + len > n_) return "truncated DHT";
        size_t p = pos_ + 2;
        const size_t end = pos_ + len;
        while (p < end) {
            const int tc = d_[p] >> 4, th = d_[p] & 15;
            p++;
            if (tc > 1 || th > 3) return "bad Huffman table id";
            HuffTable &t = tc ? ac_[th] : dc_[th];
            if (p + 16 > end) return "truncated DHT";
            int total = 0;
            t.bits[0] = 0;
            for (int i = 1; i <= 16; i++) { t.bits[i] = d_[p++]; total += t.bits[i]; }
            if (total > 256 || p + total > end) return "bad Huffman table";
            memcpy(t.vals, d_ + p, total);
            p += total;
            t.present = true;
            t.build();
        }
        pos_ = end;
        return "";
    }

    std::string read_frame()
    {
        if (pos_ + 8 > n_) return "truncated SOF";
        const size_t len = (d_[pos_] << 8) | d_[pos_ + 1];
        if (pos_ + len > n_) return "truncated SOF";
        if (d_[pos_ + 2] != 8) return "only 8-bit JPEG is supported";
        height_ = (d_[pos_ + 3] << 8) | d_[pos_ + 4];
        width_ = (d_[pos_ + 5] << 8) | d_[pos_ + 6];
        const int nc = d_[pos_ + 7];
        if (width_ <= 0 || height_ <= 0) return "bad image size";
        if (nc != 1 && nc != 3) return "only greyscale and YCbCr JPEG files are supported";
        if (len < (size_t)(8 + 3 * nc)) return "truncated SOF";
        comps_.assign(nc, Component());
        hmax_ = vmax_ = 1;
        for (int i = 0; i < nc; i++) {
            Component &c = comps_[i];
            c.id = d_[pos_ + 8 + 3 * i];
            c.h = d_[pos_ + 9 + 3 * i] >> 4;
            c.v = d_[pos_ + 9 + 3 * i] & 15;
            c.tq = d_[pos_ + 10 + 3 * i];
            if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) return "unsupported sampling factors";
            if (c.h > hmax_) hmax_ = c.h;
            if (c.v > vmax_) vmax_ = c.v;
        }
        if (nc == 1) { comps_[0].h = comps_[0].v = 1; hmax_ = vmax_ = 1; }   // a single component is never subsampled
        if (nc == 3 && (comps_[1].h != 1 || comps_[1].v != 1 || comps_[2].h != 1 || comps_[2].v != 1))
            return "unsupported chroma sampling (need 1x1 chroma with 1x1, 2x1 or 2x2 luma)";
        if (nc == 3 && comps_[0].h == 1 && comps_[0].v == 2) return "unsupported sampling 1x2 (4:4:0)";
        mcus_x_ = (width_ + 8 * hmax_ - 1) / (8 * hmax_);
        mcus_y_ = (height_ + 8 * vmax_ - 1) / (8 * vmax_);
        for (auto &c : comps_) {
            c.blocks_w = mcus_x_ * c.h;
            c.blocks_h = mcus_y_ * c.v;
            c.width = (width_ * c.h + hmax_ - 1) / hmax_;
            c.height = (height_ * c.v + vmax_ - 1) / vmax_;
            c.stride = c.blocks_w * 8;
            c.plane.assign((size_t)c.stride * c.blocks_h * 8, 0);
        }
        pos_ += len;
        return "";
    }

    // ---- entropy-coded segment
    void fill_bits()
    {
        while (bitcnt_ <= 24) {
            int b = 0;
            if (!hit_marker_ && pos_ < n_) {
                b = d_[pos_];
                if (b == 0xFF) {
                    const int b2 = pos_ + 1 < n_ ? d_[pos_ + 1] : 0xD9;
                    if (b2 == 0) pos_ += 2;          // stuffed zero
                    else { hit_marker_ = true; b = 0; }   // a marker: feed zeros (like libjpeg) until the caller handles it
                } else {
                    pos_++;
                }
            } else {
                hit_marker_ = true;
            }
            bitbuf_ |= (uint32_t)b << (24 - bitcnt_);
            bitcnt_ += 8;
        }
    }
    int get_bits(int nbits)
    {
        if (nbits == 0) return 0;
        if (bitcnt_ < nbits) fill_bits();
        const int v = (int)(bitbuf_ >> (32 - nbits));
        bitbuf_ <<= nbits;
        bitcnt_ -= nbits;
        return v;
    }
    int decode_huff(const HuffTable &t)
    {
        if (bitcnt_ < 16) fill_bits();
        const int look = (int)(bitbuf_ >> 23);
        if (t.look_len[look]) {
            bitbuf_ <<= t.look_len[look];
            bitcnt_ -= t.look_len[look];
            return t.look_val[look];
        }
        int code = (int)(bitbuf_ >> 22), l = 10;   // first 10 bits
        for (; l <= 16; l++) {
            if (code <= t.maxcode[l] && t.maxcode[l] >= 0) break;
            code = (int)(bitbuf_ >> (32 - (l + 1)));
        }
        if (l > 16) { bitbuf_ <<= 16; bitcnt_ -= 16; return 0; }   // corrupt data: libjpeg warns and returns 0
        bitbuf_ <<= l;
        bitcnt_ -= l;
        return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    static int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

    bool decode_block(Component &c, int16_t *coef)
    {
        static const int zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                   15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        memset(coef, 0, 64 * sizeof(int16_t));
        const HuffTable &dc = dc_[c.td], &ac = ac_[c.ta];
        int s = decode_huff(dc);
        if (s > 15) return false;
        if (s) c.pred += extend(get_bits(s), s);
        coef[0] = (int16_t)c.pred;
        for (int k = 1; k < 64;) {
            const int rs = decode_huff(ac);
            const int r = rs >> 4;
            s = rs & 15;
            if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) return false;
            coef[zz[k]] = (int16_t)extend(get_bits(s), s);
            k++;
        }
        return true;
    }

    // jidctint.c jpeg_idct_islow: dequantise + 2-pass integer IDCT, 13-bit constants, PASS1_BITS = 2
    static inline int descale(int64_t x, int n) { return (int)((x + ((int64_t)1 << (n - 1))) >> n); }
    static inline uint8_t range_limit(int x)
    {
        const int i = x & 1023;   // the IDCT's view of libjpeg's range-limit table (jdmaster.c prepare_range_limit_table)
        if (i < 128) return (uint8_t)(i + 128);
        if (i < 512) return 255;
        if (i < 896) return 0;
        return (uint8_t)(i - 896);
    }
    static void idct_islow(const int16_t *coef, const uint16_t *q, uint8_t *out, int stride)
    {
        constexpr int CB = 13, P1 = 2;
        constexpr int64_t F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                          F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
        int ws[64];
        for (int col = 0; col < 8; col++) {
            const int16_t *in = coef + col;
            const uint16_t *qq = q + col;
            int *w = ws + col;
            if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0) {
                const int dcval = (int)((unsigned)(in[0] * qq[0]) << P1);
                for (int r = 0; r < 8; r++) w[8 * r] = dcval;
                continue;
            }
            int64_t z2 = in[16] * qq[16], z3 = in[48] * qq[48];
            int64_t z1 = (z2 + z3) * F_0_541;
            int64_t tmp2 = z1 + z3 * (-F_1_847);
            int64_t tmp3 = z1 + z2 * F_0_765;
            z2 = in[0] * qq[0];
            z3 = in[32] * qq[32];
            int64_t tmp0 = (z2 + z3) * ((int64_t)1 << CB);
            int64_t tmp1 = (z2 - z3) * ((int64_t)1 << CB);
            const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = in[56] * qq[56];
            tmp1 = in[40] * qq[40];
            tmp2 = in[24] * qq[24];
            tmp3 = in[8] * qq[8];
            z1 = tmp0 + tmp3;
            z2 = tmp1 + tmp2;
            z3 = tmp0 + tmp2;
            int64_t z4 = tmp1 + tmp3;
            const int64_t z5 = (z3 + z4) * F_1_175;
            tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
            z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            w[0] = descale(tmp10 + tmp3, CB - P1);  w[56] = descale(tmp10 - tmp3, CB - P1);
            w[8] = descale(tmp11 + tmp2, CB - P1);  w[48] = descale(tmp11 - tmp2, CB - P1);
            w[16] = descale(tmp12 + tmp1, CB - P1); w[40] = descale(tmp12 - tmp1, CB - P1);
            w[24] = descale(tmp13 + tmp0, CB - P1); w[32] = descale(tmp13 - tmp0, CB - P1);
        }
        for (int row = 0; row < 8; row++) {
            const int *w = ws + 8 * row;
            uint8_t *o = out + (size_t)row * stride;
            if (w[1] == 0 && w[2] == 0 && w[3] == 0 && w[4] == 0 && w[5] == 0 && w[6] == 0 && w[7] == 0) {
                const uint8_t dcval = range_limit(descale(w[0], P1 + 3));
                for (int c = 0; c < 8; c++) o[c] = dcval;
                continue;
            }
            int64_t z2 = w[2], z3 = w[6];
            int64_t z1 = (z2 + z3) * F_0_541;
            int64_t tmp2 = z1 + z3 * (-F_1_847);
            int64_t tmp3 = z1 + z2 * F_0_765;
            int64_t tmp0 = ((int64_t)w[0] + w[4]) * ((int64_t)1 << CB);
            int64_t tmp1 = ((int64_t)w[0] - w[4]) * ((int64_t)1 << CB);
            const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
            z1 = tmp0 + tmp3;
            z2 = tmp1 + tmp2;
            z3 = tmp0 + tmp2;
            int64_t z4 = tmp1 + tmp3;
            const int64_t z5 = (z3 + z4) * F_1_175;
            tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
            z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            constexpr int S = CB + P1 + 3;
            o[0] = range_limit(descale(tmp10 + tmp3, S)); o[7] = range_limit(descale(tmp10 - tmp3, S));
            o[1] = range_limit(descale(tmp11 + tmp2, S)); o[6] = range_limit(descale(tmp11 - tmp2, S));
            o[2] = range_limit(descale(tmp12 + tmp1, S)); o[5] = range_limit(descale(tmp12 - tmp1, S));
            o[3] = range_limit(descale(tmp13 + tmp0, S)); o[4] = range_limit(descale(tmp13 - tmp0, S));
        }
    }

    std::string read_scan()
    {
        if (pos_ + 3 > n_) return "truncated SOS";
        const size_t len = (d_[pos_] << 8) | d_[pos_ + 1];
        if (pos_ + len > n_) return "truncated SOS";
        const int ns = d_[pos_ + 2];
        if (ns < 1 || ns > (int)comps_.size() || len < (size_t)(6 + 2 * ns)) return "bad SOS";
        std::vector<Component *> sc;
        for (int i = 0; i < ns; i++) {
            const int id = d_[pos_ + 3 + 2 * i], tt = d_[pos_ + 4 + 2 * i];
            Component *c = nullptr;
            for (auto &k : comps_)
                if (k.id == id) c = &k;
            if (!c) return "SOS names an unknown component";
            c->td = tt >> 4;
            c->ta = tt & 15;
            if (c->td > 3 || c->ta > 3 || !dc_[c->td].present || !ac_[c->ta].present) return "missing Huffman table";
            if (!quant_ok_[c->tq]) return "missing quantisation table";
            sc.push_back(c);
        }
        pos_ += len;
        bitbuf_ = 0;
        bitcnt_ = 0;
        hit_marker_ = false;
        for (auto *c : sc) c->pred = 0;
        int16_t coef[64];
        int to_restart = restart_interval_;
        auto restart = [&]() -> bool {   // consume an RSTn marker, reset predictions
            bitbuf_ = 0;
            bitcnt_ = 0;
            hit_marker_ = false;
            while (pos_ + 1 < n_ && !(d_[pos_] == 0xFF && d_[pos_ + 1] >= 0xD0 && d_[pos_ + 1] <= 0xD7)) {
                if (d_[pos_] == 0xFF && d_[pos_ + 1] != 0 && d_[pos_ + 1] != 0xFF) return false;   // some other marker
                pos_++;
            }
            if (pos_ + 1 >= n_) return false;
            pos_ += 2;
            for (auto *c : sc) c->pred = 0;
            return true;
        };
        if (ns == 1) {   // non-interleaved scan: the component's own blocks, only those that hold image samples
            Component &c = *sc[0];
            const int bw = (c.width + 7) / 8, bh = (c.height + 7) / 8;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) {
                    if (restart_interval_ && to_restart == 0) {
                        if (!restart()) return "bad restart marker";
                        to_restart = restart_interval_;
                    }
                    if (!decode_block(c, coef)) return "corrupt entropy-coded data";
                    idct_islow(coef, quant_[c.tq], c.plane.data() + (size_t)by * 8 * c.stride + bx * 8, c.stride);
                    to_restart--;
                }
        } else {
            for (int my = 0; my < mcus_y_; my++)
                for (int mx = 0; mx < mcus_x_; mx++) {
                    if (restart_interval_ && to_restart == 0) {
                        if (!restart()) return "bad restart marker";
                        to_restart = restart_interval_;
                    }
                    for (auto *c : sc)
                        for (int v = 0; v < c->v; v++)
                            for (int h = 0; h < c->h; h++) {
                                if (!decode_block(*c, coef)) return "corrupt entropy-coded data";
                                idct_islow(coef, quant_[c->tq],
                                           c->plane.data() + (size_t)(my * c->v + v) * 8 * c->stride + (mx * c->h + h) * 8, c->stride);
                            }
                    to_restart--;
                }
        }
        comps_in_scans_ += ns;
        if (comps_in_scans_ >= (int)comps_.size()) scans_done_ = true;
        return "";
    }

    // jdsample.c: one output row pair of h2v2 fancy upsampling / one row of h2v1 fancy upsampling
    static void h2v1_fancy_row(const uint8_t *in, int w, uint8_t *out)
    {
        if (w == 1) { out[0] = out[1] = in[0]; return; }
        int v = in[0];
        out[0] = (uint8_t)v;
        out[1] = (uint8_t)((v * 3 + in[1] + 2) >> 2);
        for (int i = 1; i < w - 1; i++) {
            v = in[i] * 3;
            out[2 * i] = (uint8_t)((v + in[i - 1] + 1) >> 2);
            out[2 * i + 1] = (uint8_t)((v + in[i + 1] + 2) >> 2);
        }
        v = in[w - 1];
        out[2 * (w - 1)] = (uint8_t)((v * 3 + in[w - 2] + 1) >> 2);
        out[2 * (w - 1) + 1] = (uint8_t)v;
    }
    static void h2v2_fancy_row(const uint8_t *near0, const uint8_t *far1, int w, uint8_t *out)
    {
        if (w == 1) {
            const int s = near0[0] * 3 + far1[0];
            out[0] = (uint8_t)((s * 4 + 8) >> 4);
            out[1] = (uint8_t)((s * 4 + 7) >> 4);
            return;
        }
        int thiscol = near0[0] * 3 + far1[0], nextcol = near0[1] * 3 + far1[1], lastcol;
        out[0] = (uint8_t)((thiscol * 4 + 8) >> 4);
        out[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
        lastcol = thiscol;
        thiscol = nextcol;
        for (int i = 1; i < w - 1; i++) {
            nextcol = near0[i + 1] * 3 + far1[i + 1];
            out[2 * i] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
            out[2 * i + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
            lastcol = thiscol;
            thiscol = nextcol;
        }
        out[2 * (w - 1)] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
        out[2 * (w - 1) + 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
    }

    std::string finish(std::vector<uint8_t> &out)
    {
        const int W = width_, H = height_;
        if (comps_.size() == 1) {
            out.resize((size_t)W * H);
            for (int y = 0; y < H; y++) memcpy(&out[(size_t)y * W], &comps_[0].plane[(size_t)y * comps_[0].stride], W);
            return "";
        }
        // chroma to full resolution (luma is full resolution in every supported layout)
        const Component &Y = comps_[0];
        std::vector<uint8_t> up[2];
        const bool h2 = hmax_ == 2, v2 = vmax_ == 2;
        const int upw = h2 ? 2 * comps_[1].width : comps_[1].width;
        for (int ci = 0; ci < 2; ci++) {
            const Component &c = comps_[1 + ci];
            up[ci].assign((size_t)(upw + 2) * H + 2 * upw + 4, 0);
            const bool fancy = c.width > 2;   // jdsample.c: narrower components are replicated, not filtered
            for (int y = 0; y < H; y++) {
                uint8_t *o = &up[ci][(size_t)y * (upw + 2)];
                if (!fancy && (h2 || v2)) {
                    const uint8_t *in = &c.plane[(size_t)(v2 ? y >> 1 : y) * c.stride];
                    for (int x = 0; x < upw; x++) o[x] = in[h2 ? x >> 1 : x];
                } else if (!h2 && !v2) {
                    memcpy(o, &c.plane[(size_t)y * c.stride], c.width);
                } else if (h2 && !v2) {
                    h2v1_fancy_row(&c.plane[(size_t)y * c.stride], c.width, o);
                } else {
                    // output rows 2i and 2i+1 come from input row i and the row above / below it (edge rows replicate)
                    const int i = y >> 1;
                    int other = (y & 1) ? i + 1 : i - 1;
                    if (other < 0) other = 0;
                    if (other > c.height - 1) other = c.height - 1;
                    h2v2_fancy_row(&c.plane[(size_t)i * c.stride], &c.plane[(size_t)other * c.stride], c.width, o);
                }
            }
        }
        // jdcolor.c ycc_rgb_convert with its fixed-point tables (SCALEBITS = 16)
        static int cr_r[256], cb_b[256];
        static int32_t cr_g[256], cb_g[256];
        static bool tables = false;
        if (!tables) {
            for (int i = 0; i < 256; i++) {
                const int x = i - 128;
                cr_r[i] = (int)((91881LL * x + 32768) >> 16);      // FIX(1.40200)
                cb_b[i] = (int)((116130LL * x + 32768) >> 16);     // FIX(1.77200)
                cr_g[i] = (int32_t)(-46802LL * x);                 // FIX(0.71414)
                cb_g[i] = (int32_t)(-22554LL * x + 32768);         // FIX(0.34414) + ONE_HALF
            }
            tables = true;
        }
        auto clamp = [](int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
        out.resize((size_t)W * H * 3);
        for (int y = 0; y < H; y++) {
            const uint8_t *yp = &Y.plane[(size_t)y * Y.stride];
            const uint8_t *cbp = &up[0][(size_t)y * (upw + 2)], *crp = &up[1][(size_t)y * (upw + 2)];
            uint8_t *o = &out[(size_t)y * W * 3];
            for (int x = 0; x < W; x++) {
                const int yy = yp[x], cb = cbp[x], cr = crp[x];
                o[3 * x] = clamp(yy + cr_r[cr]);
                o[3 * x + 1] = clamp(yy + (int)((cb_g[cb] + cr_g[cr]) >> 16));
                o[3 * x + 2] = clamp(yy + cb_b[cb]);
            }
        }
        return "";
    }
};

// Reads and decodes a JPEG file.  Returns "" on success, else an error message.
inline std::string load_jpeg(const char *path, int &width, int &height, int &channels, std::vector<uint8_t> &pixels)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return std::string("cannot open ") + path;
    std::vector<uint8_t> data;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) data.insert(data.end(), buf, buf + n);
    fclose(fp);
    Decoder d;
    return d.decode(data.data(), data.size(), width, height, channels, pixels);
}

}  // namespace jpegdec
