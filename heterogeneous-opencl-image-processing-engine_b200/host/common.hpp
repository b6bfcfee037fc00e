// common.hpp -- shared host-side pieces of the two drop-in CLIs (heterogeneous_blur, split_image_blur).
//
// Everything here is plain C++ over the C ABI in include/b200blur.h; there is no CUDA, PyTorch or OpenCL in the host
// programs.  Mirrors the helper layer the reference keeps inline in main(): cl_error (heterogeneous_blur.c:25-30),
// get_time_ms (:32-36), the image load + planar->interleaved step (:106-135) and save_one_image_rgb
// (split_image_blur.c:40-56).
#pragma once
#include <sys/time.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200blur.h"
#include "jpeg_decode.hpp"

// cl_error(): print "<code> - <message>" and exit(-1), the reference's only error path for device calls.
inline void blur_check(int code, const char *what)
{
    if (code != B200BLUR_OK) {
        printf("%d - %s (%s)\n", code, what, b200blur_last_error());
        exit(-1);
    }
}

inline double get_time_ms()
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (tv.tv_sec * 1000.0) + (tv.tv_usec / 1000.0);
}

struct Image {
    int width = 0, height = 0, channels = 3;
    std::vector<unsigned char> data;  // interleaved RGBRGB..., the layout the kernel takes (heterogeneous_blur.c:128-135)
    size_t size() const { return (size_t)width * height * channels; }
};

// Binary PPM (P6, maxval 255) reader: the dependency-free stand-in for CImg + libjpeg.
inline bool load_ppm(const char *path, Image &img)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return false;
    auto next_int = [&](int &v) -> bool {
        int c = fgetc(fp);
        while (c == ' ' || c == '\n' || c == '\r' || c == '\t' || c == '#') {
            if (c == '#') while (c != '\n' && c != EOF) c = fgetc(fp);
            c = fgetc(fp);
        }
        if (c < '0' || c > '9') return false;
        v = 0;
        while (c >= '0' && c <= '9') { v = v * 10 + (c - '0'); c = fgetc(fp); }
        return true;
    };
    char magic[3] = {0, 0, 0};
    if (fread(magic, 1, 2, fp) != 2 || magic[0] != 'P' || magic[1] != '6') { fclose(fp); return false; }
    int w, h, maxval;
    if (!next_int(w) || !next_int(h) || !next_int(maxval) || maxval != 255 || w <= 0 || h <= 0) { fclose(fp); return false; }
    img.width = w; img.height = h; img.channels = 3;
    img.data.resize(img.size());
    const bool ok = fread(img.data.data(), 1, img.size(), fp) == img.size();
    fclose(fp);
    return ok;
}

// save_one_image_rgb() of split_image_blur.c:40-56, as PPM.
inline bool save_ppm(const char *path, const unsigned char *interleaved, int width, int height)
{
    FILE *fp = fopen(path, "wb");
    if (!fp) return false;
    fprintf(fp, "P6\n%d %d\n255\n", width, height);
    const size_t n = (size_t)width * height * 3;
    const bool ok = fwrite(interleaved, 1, n, fp) == n;
    fclose(fp);
    return ok;
}

// Seeded synthetic photo-like image (smooth gradients + texture) for when no input file is given.
inline void make_synthetic(Image &img, int width, int height, uint64_t seed)
{
    img.width = width; img.height = height; img.channels = 3;
    img.data.resize(img.size());
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const int noise = (int)((s >> 33) & 63);
            unsigned char *p = &img.data[((size_t)y * width + x) * 3];
            p[0] = (unsigned char)((x * 255 / (width > 1 ? width - 1 : 1) + noise) & 255);
            p[1] = (unsigned char)((y * 255 / (height > 1 ? height - 1 : 1) + noise) & 255);
            p[2] = (unsigned char)(((x ^ y) * 3 + noise) & 255);
        }
}

// 64-bit FNV-1a over a byte range: lets a run print a checksum of everything it produced.
inline uint64_t fnv1a(const unsigned char *p, size_t n, uint64_t h = 1469598103934665603ull)
{
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// Options that follow the reference's positional arguments.  The positional surface is unchanged; these only expose
// what the reference hard-codes (heterogeneous_blur.c:43-48) so BASELINE.json's other shapes can be run.
struct ExtraOptions {
    std::string input;        // --input file.ppm          (default ./image_320x240.ppm, else synthetic)
    int width = 320, height = 240;  // --width/--height     (synthetic source only)
    int num_images = 5000;    // --images N                (NUM_IMAGES)
    int gpus = 0;             // --gpus G                  (0 = all visible)
    bool resident = false;    // --resident                (keep the stream in HBM: kernels only, no host copies)
    bool quiet = false;       // --quiet                   (no per-batch progress lines)
    std::string save;         // --save out.ppm            (write output image 0)
    bool checksum = false;    // --checksum                (FNV-1a of all outputs, end-to-end mode)
    int repeat = 1;           // --repeat R                (resident mode: passes over the stream)
    bool host_halo = false;   // --host-halo               (Approach 2: upload halo rows from the host like the reference)
    int fill_threads = 0;     // --fill-threads T          (host threads per GPU that replicate the source image into staging; 0 = auto)
    int ring = 4;             // --ring R                  (pinned + device slots per GPU in the end-to-end pipeline, >= 2)
    bool oversubscribe = false;  // --oversubscribe        (allow --gpus G > visible devices: band/shard k runs on device k % visible;
                                 //                          exercises the multi-GPU host logic on a single-GPU box)
    int fuse = 0;                // --fuse N               (batches per transfer/launch; 0 = automatic ~64 MB, 1 = the reference's granularity)
    bool static_split = false;   // --static-split         (Approach 1: fixed even partition per batch instead of work stealing)
    bool stage_once = false;     // --stage-once           (every image of the stream is the same source image, so fill each
                                 //                          pinned ring slot ONCE, before the timer, and re-send it per batch;
                                 //                          the default re-fills per batch inside the timer like the reference)
};

inline int parse_extra(int argc, char **argv, int first, ExtraOptions &o)
{
    for (int i = first; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { printf("Error: %s needs a value\n", name); exit(-1); }
            return argv[++i];
        };
        if (a == "--input") o.input = val("--input");
        else if (a == "--width") o.width = atoi(val("--width"));
        else if (a == "--height") o.height = atoi(val("--height"));
        else if (a == "--images") o.num_images = atoi(val("--images"));
        else if (a == "--gpus") o.gpus = atoi(val("--gpus"));
        else if (a == "--resident") o.resident = true;
        else if (a == "--quiet") o.quiet = true;
        else if (a == "--save") o.save = val("--save");
        else if (a == "--checksum") o.checksum = true;
        else if (a == "--repeat") o.repeat = atoi(val("--repeat"));
        else if (a == "--host-halo") o.host_halo = true;
        else if (a == "--fill-threads") o.fill_threads = atoi(val("--fill-threads"));
        else if (a == "--ring") o.ring = atoi(val("--ring"));
        else if (a == "--oversubscribe") o.oversubscribe = true;
        else if (a == "--static-split") o.static_split = true;
        else if (a == "--stage-once") o.stage_once = true;
        else if (a == "--fuse") o.fuse = atoi(val("--fuse"));
        else { printf("Error: unknown option %s\n", a.c_str()); return -1; }
    }
    if (o.width < 1 || o.height < 1 || o.num_images < 1 || o.repeat < 1 || o.ring < 2 || o.ring > 64) { printf("Error: bad size option\n"); return -1; }
    return 0;
}

// The reference's LOAD ORIGINAL IMAGE section (heterogeneous_blur.c:104-137) without CImg/libjpeg: --input takes a
// .jpg/.jpeg (decoded like libjpeg does by default, host/jpeg_decode.hpp) or a binary .ppm; with no --input the
// reference's hard-coded ./image_320x240.jpg (heterogeneous_blur.c:43) is used if it is in the working directory, then
// ./image_320x240.ppm, then a synthetic image.
inline bool is_jpeg_file(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return false;
    unsigned char m[2] = {0, 0};
    const bool ok = fread(m, 1, 2, fp) == 2 && m[0] == 0xFF && m[1] == 0xD8;
    fclose(fp);
    return ok;
}

inline bool load_image_file(const char *path, Image &img, std::string &err)
{
    if (is_jpeg_file(path)) {
        std::vector<uint8_t> px;
        int w, h, c;
        err = jpegdec::load_jpeg(path, w, h, c, px);
        if (!err.empty()) return false;
        img.width = w; img.height = h; img.channels = c;
        img.data.assign(px.begin(), px.end());
        return true;
    }
    if (load_ppm(path, img)) return true;
    err = "not a readable JPEG or binary PPM (P6, maxval 255) file";
    return false;
}

inline void load_source_image(const ExtraOptions &o, Image &img, std::string &name)
{
    std::string err;
    if (!o.input.empty()) {
        if (!load_image_file(o.input.c_str(), img, err)) { printf("Error: cannot read image file %s (%s)\n", o.input.c_str(), err.c_str()); exit(-1); }
        name = o.input;
    } else if (o.width == 320 && o.height == 240 && load_image_file("./image_320x240.jpg", img, err)) {
        name = "./image_320x240.jpg";
    } else if (o.width == 320 && o.height == 240 && load_ppm("./image_320x240.ppm", img)) {
        name = "./image_320x240.ppm";
    } else {
        make_synthetic(img, o.width, o.height, 2026);
        name = "(synthetic " + std::to_string(o.width) + "x" + std::to_string(o.height) + ")";
    }
}

// The reference's "copy original image to each slot" loop (heterogeneous_blur.c:440-442).  The source image stays in the
// core's cache; the destination is pinned staging memory that the GPU's copy engine reads next and the CPU never reads
// back, so the copy uses non-temporal stores: no read-for-ownership of the destination lines (a third less DRAM traffic
// than memcpy) and no cache pollution.  Spread over a few host threads: one core streams ~10-14 GB/s, a GPU's host link
// takes 45-55 GB/s.
#include <thread>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
inline void copy_streaming(unsigned char *dst, const unsigned char *src, size_t n)
{
    // head: up to the first 16-byte boundary of dst
    size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; i++) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), d);
        src += 64; dst += 64;
    }
    memcpy(dst, src, n - blocks * 64);
}
inline void copy_streaming_done() { _mm_sfence(); }
#else
inline void copy_streaming(unsigned char *dst, const unsigned char *src, size_t n) { memcpy(dst, src, n); }
inline void copy_streaming_done() {}
#endif

// A few persistent host threads per GPU worker that replicate the source image into a staging slot (a thread per
// call would cost ~50 us each to create, 8 of them per 1 ms of copying).
#include <condition_variable>
#include <mutex>
class StagingPool {
public:
    explicit StagingPool(int threads) : n_(threads < 1 ? 1 : threads)
    {
        for (int t = 1; t < n_; t++) pool_.emplace_back([this, t] { loop(t); });
    }
    ~StagingPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &t : pool_) t.join();
    }
    int threads() const { return n_; }
    // dst[i] = src for i in [0, count): `count` copies of `bytes_each` bytes
    void replicate(unsigned char *dst, const unsigned char *src, size_t bytes_each, long long count)
    {
        int use = n_;
        if (count < 2 * use || (double)count * bytes_each < 4e6) use = 1;
        if (use == 1) { work(dst, src, bytes_each, 0, count); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = dst; src_ = src; each_ = bytes_each; count_ = count; use_ = use;
            pending_ = use - 1;
            gen_++;
        }
        cv_.notify_all();
        work(dst, src, bytes_each, 0, count / use);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    static void work(unsigned char *dst, const unsigned char *src, size_t each, long long a, long long b)
    {
        for (long long i = a; i < b; i++) copy_streaming(dst + (size_t)i * each, src, each);
        copy_streaming_done();
    }
    void loop(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            unsigned char *dst; const unsigned char *src; size_t each; long long count; int use;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (quit_) return;
                dst = dst_; src = src_; each = each_; count = count_; use = use_;
            }
            if (t < use) {
                work(dst, src, each, count * t / use, count * (t + 1) / use);
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> pool_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    bool quit_ = false;
    unsigned char *dst_ = nullptr; const unsigned char *src_ = nullptr; size_t each_ = 0; long long count_ = 0;
    int use_ = 1, pending_ = 0;
};

// Staging threads per GPU when --fill-threads is not given: a quarter of the host cores shared out over the GPUs,
// 2..4 each.  More is slower: the staging stores compete with the copy engines for the same host memory (measured on a
// 16-core box, one GPU, 5000 x 320x240: 4 threads 180 k images/s, 8 threads 171 k, 12 threads 170 k).
inline int auto_fill_threads(int gpus)
{
    const unsigned hw = std::thread::hardware_concurrency();
    int t = (int)(hw ? hw : 8) / (4 * (gpus > 0 ? gpus : 1));
    return t < 2 ? 2 : (t > 4 ? 4 : t);
}

struct DeviceTimes {
    double in_ms = 0, kernel_ms = 0, out_ms = 0;
    double fill_ms = 0;  // host time spent replicating the source image into staging (inside the wall-clock window)
    double busy_until_ms = 0;  // wall-clock time at which this GPU's last result was back on the host
    long long images = 0;
    double total() const { return in_ms + kernel_ms + out_ms; }
};
