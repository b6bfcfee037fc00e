// split_image_blur.cpp -- Approach 2 (split-image distribution) on B200s, drop-in for the reference CLI.
//
//   ./split_image_blur [gpu_ratio] [batch_size] [--images N] [--input f.ppm | --width W --height H] [--gpus G]
//                      [--resident [--repeat R]] [--host-halo] [--quiet] [--save out.ppm] [--checksum]
//
// Same positional surface, defaults, warnings and report sections as split_image_blur.c (:59-102, :615-721).
// The reference cuts every image at one row between a CPU and a GPU OpenCL device and hands each device its rows plus
// one halo row from HOST memory (:144-166, :511-517).  Here every image is cut into G row bands, one per B200:
//   * band k owns rows [k*H/G, (k+1)*H/G) of EVERY image (even split replaces split_row = (int)(H*(1-gpu_ratio)));
//   * each GPU uploads only its own rows; the halo row above/below a band is read by the stencil kernel straight from
//     the neighbouring GPU's memory over NVLink (peer-enabled pointers in b200blur_launch.halo_top/halo_bottom) -- the
//     halo exchange is fused into the kernel, there is no separate exchange step and no halo slot;
//   * --host-halo keeps the reference's scheme instead (halo rows uploaded from the host with the band, kernel run with
//     height = rows incl. halo, halo outputs dropped, :401/:414/:526/:537) for machines without peer access;
//   * one host thread per GPU, three queues (H2D / blur / D2H) and a ring of slots per GPU.
#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>

#include "common.hpp"

namespace {

struct Slot {
    unsigned char *h_in = nullptr, *h_out = nullptr;
    void *d_in = nullptr, *d_out = nullptr;
    b200blur_event ev_in = -1, ev_k = -1, ev_out = -1;
    long long count = 0, batch = -1;
    long long staged = 0;  // images whose band rows are already replicated into h_in (--stage-once)
    bool busy = false;
};

struct Worker {
    int gpu = 0;
    b200blur_ctx *ctx = nullptr;
    int row0 = 0, rows = 0;       // owned (= output) rows
    int top = 0, bot = 0;         // 1 if a neighbour band exists above / below
    int in_rows = 0;              // rows per image in the device input buffer (rows, or rows + halos with --host-halo)
    std::vector<Slot> ring;
    DeviceTimes t;
    std::unique_ptr<StagingPool> staging;
    uint64_t checksum = 1469598103934665603ull;
    double resident_ms = 0;
    void *d_res_in = nullptr, *d_res_out = nullptr;
};


// Host-side rendezvous of the per-GPU threads.  With peer halos a band's kernel names its neighbours' upload events
// and a band's next upload names its neighbours' kernel events; the two barriers per batch guarantee those event
// handles are live (already enqueued by their owner, not yet released) when another thread refers to them.
class SpinBarrier {
public:
    explicit SpinBarrier(int n) : n_(n) {}
    void wait()
    {
        const int g = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_acq_rel);
        } else {
            while (gen_.load(std::memory_order_acquire) == g) std::this_thread::yield();
        }
    }
private:
    const int n_;
    std::atomic<int> count_{0}, gen_{0};
};

}  // namespace

int main(int argc, char **argv)
{
    // Configuration (split_image_blur.c:62-70)
    int BATCH_SIZE = 500;
    int local_work_size = 16;
    float gpu_ratio = 0.5f;
    const int HALO = 1;
    ExtraOptions opt;

    int npos = 1;
    while (npos < argc && strncmp(argv[npos], "--", 2) != 0) npos++;
    if (parse_extra(argc, argv, npos, opt) != 0) return -1;
    const int kRing = opt.ring;
    const int NUM_IMAGES = opt.num_images;

    if (npos > 1) {
        gpu_ratio = atof(argv[1]);
        if (gpu_ratio < 0.0f || gpu_ratio > 1.0f) {
            printf("Warning: gpu_ratio must be between 0.0 and 1.0. Using 0.5\n");
            gpu_ratio = 0.5f;
        }
    }
    if (npos > 2) {
        BATCH_SIZE = atoi(argv[2]);
        if (BATCH_SIZE < 1 || BATCH_SIZE > NUM_IMAGES) {
            printf("Warning: BATCH_SIZE must be between 1 and %d. Using 500\n", NUM_IMAGES);
            BATCH_SIZE = 500;
        }
    }
    if (BATCH_SIZE > NUM_IMAGES) BATCH_SIZE = NUM_IMAGES;
    const int NUM_BATCHES = (NUM_IMAGES + BATCH_SIZE - 1) / BATCH_SIZE;

    Image img;
    std::string input_name;
    load_source_image(opt, img, input_name);

    printf("========== SPLIT-IMAGE CONFIGURATION ==========\n");
    printf("Input file: %s\n", input_name.c_str());
    printf("Number of images in stream: %d\n", NUM_IMAGES);
    printf("Batch size: %d images\n", BATCH_SIZE);
    printf("Number of batches: %d\n", NUM_BATCHES);
    printf("Work-group size: %dx%d\n", local_work_size, local_work_size);
    printf("GPU ratio: %.1f%% (rows to GPU)\n", gpu_ratio * 100);
    printf("Halo size: %d row(s)\n", HALO);
    printf("================================================\n\n");

    const int width = img.width, height = img.height, channels = img.channels;
    printf("Original image loaded: %dx%d, %d channels\n", width, height, channels);
    const size_t image_size = img.size();
    const size_t pitch = (size_t)width * channels;
    printf("Size of one image: %zu bytes (%.2f KB)\n", image_size, image_size / 1024.0);
    const unsigned char *original_image = img.data.data();
    printf("Original image converted to interleaved format\n\n");

    // ======================== DEVICE DISCOVERY ========================
    int n_dev = 0;
    if (b200blur_device_count(&n_dev) != B200BLUR_OK || n_dev == 0) {
        printf("Error: Could not find a CUDA device (%s)\n", b200blur_last_error());
        return -1;
    }
    int G = n_dev;
    if (opt.gpus > 0) G = opt.oversubscribe ? opt.gpus : std::min(opt.gpus, n_dev);
    if (G > height) G = height;  // a band needs at least one row
    if (opt.fill_threads <= 0) opt.fill_threads = auto_fill_threads(G);

    // ======================== CALCULATE SPLIT DIMENSIONS (split_image_blur.c:142-173) ========================
    if (height >= 2) {
        int ref_split = 0;
        b200blur_ratio_split_row(height, gpu_ratio, &ref_split);
        printf("Split configuration:\n");
        printf("  Reference two-device split for ratio %.3f: split row %d (kept for CLI compatibility)\n", gpu_ratio, ref_split);
    }
    printf("  Row bands over %d GPU(s), %d halo row(s) per interior edge (%s):\n", G, HALO,
           opt.host_halo ? "uploaded from host" : "read from the neighbour GPU over NVLink");
    std::vector<Worker> workers(G);
    for (int k = 0; k < G; k++) {
        Worker &w = workers[k];
        int64_t b, c;
        b200blur_partition(height, G, k, &b, &c);
        w.gpu = k;
        w.row0 = (int)b;
        w.rows = (int)c;
        w.top = k > 0 ? 1 : 0;
        w.bot = k < G - 1 ? 1 : 0;
        w.in_rows = opt.host_halo ? w.rows + w.top + w.bot : w.rows;
        printf("  GPU %d: rows %d-%d, %d input rows%s, %d output rows (%.2f KB in, %.2f KB out per image)\n", k, w.row0,
               w.row0 + w.rows - 1, w.in_rows, opt.host_halo ? " (inc. halo)" : " + peer halo", w.rows,
               w.in_rows * pitch / 1024.0, w.rows * pitch / 1024.0);
    }
    printf("\n");

    printf("Platform 0: NVIDIA CUDA (%s)\n", b200blur_version());
    for (int k = 0; k < G; k++) {
        char dname[256];
        blur_check(b200blur_device_name(k % n_dev, dname, sizeof dname), "Failed to get device name");
        printf("GPU device %d: %s\n", k, dname);
        blur_check(b200blur_ctx_create(k % n_dev, 3, &workers[k].ctx), "Failed to create context");
    }
    if (!opt.host_halo)
        for (int k = 0; k + 1 < G; k++)
            if (b200blur_peer_enable(workers[k].ctx, workers[k + 1].ctx) != B200BLUR_OK) {
                printf("Error: no peer access between GPU %d and GPU %d (%s); re-run with --host-halo\n", k, k + 1,
                       b200blur_last_error());
                return -1;
            }
    printf("\nKernel objects created (precompiled sm_100a)\n\n");

    // ======================== DEVICE BUFFER ALLOCATION (:359-386) ========================
    printf("Allocating device buffers...\n");
    // Batches are independent: `fuse` consecutive batches travel and launch together (~64 MB per transfer per GPU).
    long long max_band_rows = 0;
    for (auto &w : workers) max_band_rows = std::max<long long>(max_band_rows, w.in_rows);
    long long fuse = (long long)((64.0 * 1024 * 1024) / ((double)BATCH_SIZE * max_band_rows * pitch) + 0.5);
    fuse = std::max(1LL, std::min<long long>(fuse, std::max(1, NUM_BATCHES / 16)));  // keep >= 16 pipeline steps per GPU
    if (opt.fuse > 0) fuse = opt.fuse;
    const long long per_dev_images = opt.resident ? NUM_IMAGES : BATCH_SIZE * fuse;
    for (auto &w : workers) {
        const size_t in_bytes = (size_t)per_dev_images * w.in_rows * pitch, out_bytes = (size_t)per_dev_images * w.rows * pitch;
        if (opt.resident) {
            blur_check(b200blur_dev_alloc(w.ctx, in_bytes, &w.d_res_in), "Failed to create input buffer");
            blur_check(b200blur_dev_alloc(w.ctx, out_bytes, &w.d_res_out), "Failed to create output buffer");
        } else {
            w.ring.resize(kRing);
            for (auto &s : w.ring) {
                blur_check(b200blur_host_alloc(in_bytes, (void **)&s.h_in), "Failed to allocate pinned input");
                blur_check(b200blur_host_alloc(out_bytes, (void **)&s.h_out), "Failed to allocate pinned output");
                blur_check(b200blur_dev_alloc(w.ctx, in_bytes, &s.d_in), "Failed to create input buffer");
                blur_check(b200blur_dev_alloc(w.ctx, out_bytes, &s.d_out), "Failed to create output buffer");
            }
        }
    }
    for (auto &w : workers) w.staging.reset(new StagingPool(opt.fill_threads));
    // Like the reference, whose program build (clBuildProgram, :257-353) happens before its timer starts: load the
    // kernel image and wake the copy paths with one tiny write -> blur -> read per GPU (clamped edges, no halo).
    if (!opt.resident)
        for (auto &w : workers) {
            Slot &s = w.ring[0];
            const size_t bytes = (size_t)w.in_rows * pitch;
            memcpy(s.h_in, original_image, bytes);
            b200blur_launch l;
            blur_check(b200blur_enqueue_write(w.ctx, 0, s.d_in, s.h_in, bytes, NULL), "GPU write failed");
            blur_check(b200blur_launch_rows(&l, s.d_in, s.d_out, width, w.in_rows, channels, 0, w.rows, 1, bytes, (size_t)w.rows * pitch),
                       "Failed to set kernel args");
            blur_check(b200blur_enqueue_blur(w.ctx, 0, &l, NULL), "GPU kernel launch failed");
            blur_check(b200blur_enqueue_read(w.ctx, 0, s.h_out, s.d_out, (size_t)w.rows * pitch, NULL), "GPU read failed");
            blur_check(b200blur_finish_all(w.ctx), "finish failed");
        }
    printf("Device buffers allocated\n\n");
    printf("Starting batch processing of %d images in %d batches...\n\n", NUM_IMAGES, NUM_BATCHES);
    std::vector<unsigned char> first_output;
    if (!opt.save.empty()) first_output.resize(image_size);

    // Builds the launch for `count` images whose band data starts at d_in (slot or resident buffer) on worker w.
    auto make_launch = [&](Worker &w, void *d_in, void *d_out, long long count, int slot_index, b200blur_launch &l) {
        const size_t in_stride = (size_t)w.in_rows * pitch, out_stride = (size_t)w.rows * pitch;
        if (opt.host_halo) {
            // the reference's scheme: kernel height = rows incl. halo, keep rows [top, top + rows)  (:401, :414, :526, :537)
            blur_check(b200blur_launch_rows(&l, d_in, d_out, width, w.in_rows, channels, w.top, w.rows, count, in_stride, out_stride),
                       "Failed to set kernel args");
            return;
        }
        blur_check(b200blur_launch_rows(&l, d_in, d_out, width, w.rows, channels, 0, w.rows, count, in_stride, out_stride),
                   "Failed to set kernel args");
        if (w.top) {  // last row of the band above, in GPU k-1's memory
            Worker &n = workers[w.gpu - 1];
            const unsigned char *base = (const unsigned char *)(opt.resident ? n.d_res_in : n.ring[slot_index].d_in);
            l.halo_top = base + (size_t)(n.rows - 1) * pitch;
            l.halo_top_stride = (size_t)n.in_rows * pitch;
        }
        if (w.bot) {  // first row of the band below, in GPU k+1's memory
            Worker &n = workers[w.gpu + 1];
            l.halo_bottom = opt.resident ? n.d_res_in : n.ring[slot_index].d_in;
            l.halo_bottom_stride = (size_t)n.in_rows * pitch;
        }
    };

    auto fill_rows = [&](Worker &w, unsigned char *dst, long long count) {  // replicate this band's rows (:469-480)
        const int first_row = opt.host_halo ? w.row0 - w.top : w.row0;
        const size_t bytes = (size_t)w.in_rows * pitch;
        const double tf = get_time_ms();
        w.staging->replicate(dst, original_image + (size_t)first_row * pitch, bytes, count);
        w.t.fill_ms += get_time_ms() - tf;
    };

    auto harvest = [&](Worker &w, Slot &s) {
        double ms;
        blur_check(b200blur_event_ms(w.ctx, s.ev_out, &ms), "Failed to read transfer-out time");
        w.t.out_ms += ms;
        blur_check(b200blur_event_ms(w.ctx, s.ev_in, &ms), "Failed to read transfer-in time");
        w.t.in_ms += ms;
        blur_check(b200blur_event_ms(w.ctx, s.ev_k, &ms), "Failed to read kernel time");
        w.t.kernel_ms += ms;
        b200blur_event_release(w.ctx, s.ev_in);
        b200blur_event_release(w.ctx, s.ev_k);
        b200blur_event_release(w.ctx, s.ev_out);
        if (opt.checksum) w.checksum = fnv1a(s.h_out, (size_t)s.count * w.rows * pitch, w.checksum);
        if (!opt.save.empty() && s.batch == 0) memcpy(first_output.data() + (size_t)w.row0 * pitch, s.h_out, (size_t)w.rows * pitch);
        s.busy = false;
    };

    std::atomic<int> resident_ready{0};
    SpinBarrier rendezvous(G);
    auto run_worker = [&](Worker &w) {
        const int k = w.gpu;
        if (opt.resident) {
            // bands of all images resident in HBM; halo rows read from the neighbours' resident buffers
            void *h;
            const size_t band_bytes = (size_t)w.in_rows * pitch;
            // staging buffer of at most ~256 MB (large frames: a few images at a time)
            const long long stage = std::max<long long>(1, std::min<long long>(std::min<long long>(NUM_IMAGES, 256), (256ll << 20) / (long long)band_bytes));
            blur_check(b200blur_host_alloc(stage * band_bytes, &h), "Failed to allocate pinned staging");
            fill_rows(w, (unsigned char *)h, stage);
            for (long long i = 0; i < NUM_IMAGES; i += stage) {
                const long long n = std::min(stage, (long long)NUM_IMAGES - i);
                blur_check(b200blur_enqueue_write(w.ctx, 0, (unsigned char *)w.d_res_in + i * band_bytes, h, n * band_bytes, NULL),
                           "GPU write failed");
            }
            blur_check(b200blur_finish(w.ctx, 0), "finish failed");
            resident_ready.fetch_add(1);
            while (resident_ready.load() < G) std::this_thread::yield();  // neighbours' bands are in place
            b200blur_launch l;
            make_launch(w, w.d_res_in, w.d_res_out, NUM_IMAGES, 0, l);
            blur_check(b200blur_enqueue_blur(w.ctx, 1, &l, NULL), "warm-up failed");
            blur_check(b200blur_finish(w.ctx, 1), "finish failed");
            {
                // the `repeat` passes are independent batches: one launch with a descriptor per pass (the tail of one pass
                // overlaps the start of the next; with small bands a launch ramp and drain per pass would cost 15-20 %)
                std::vector<b200blur_launch> passes((size_t)opt.repeat, l);
                blur_check(b200blur_enqueue_blur_batches(w.ctx, 1, passes.data(), opt.repeat, NULL), "warm-up failed");
                blur_check(b200blur_finish(w.ctx, 1), "finish failed");
                b200blur_event ev;
                blur_check(b200blur_enqueue_blur_batches(w.ctx, 1, passes.data(), opt.repeat, &ev), "GPU kernel launch failed");
                double ms;
                blur_check(b200blur_event_ms(w.ctx, ev, &ms), "Failed to read kernel time");
                b200blur_event_release(w.ctx, ev);
                w.resident_ms += ms;
            }
            w.t.kernel_ms = w.resident_ms;
            w.t.images = (long long)NUM_IMAGES * opt.repeat;
            if (!opt.save.empty()) {
                blur_check(b200blur_enqueue_read(w.ctx, 0, h, w.d_res_out, (size_t)w.rows * pitch, NULL), "GPU read failed");
                blur_check(b200blur_finish(w.ctx, 0), "finish failed");
                memcpy(first_output.data() + (size_t)w.row0 * pitch, h, (size_t)w.rows * pitch);
            }
            b200blur_host_free(h);
            return;
        }
        const long long n_super = (NUM_BATCHES + fuse - 1) / fuse;
        for (long long batch = 0; batch < n_super; batch++) {   // one iteration = `fuse` batches of the reference loop
            const long long batch_start = batch * fuse * BATCH_SIZE;
            long long batch_count = fuse * BATCH_SIZE;
            if (batch_start + batch_count > NUM_IMAGES) batch_count = NUM_IMAGES - batch_start;
            if (k == 0 && !opt.quiet) {
                for (long long b = batch * fuse; b < std::min<long long>(NUM_BATCHES, (batch + 1) * fuse); b++) {
                    printf("=== Processing Batch %lld/%d ===\n", b + 1, NUM_BATCHES);
                    printf("  Processing %lld images (each split into %d row bands)\n",
                           std::min<long long>(BATCH_SIZE, NUM_IMAGES - b * BATCH_SIZE), G);
                }
            }
            const int si = (int)(batch % kRing);
            Slot &s = w.ring[si];
            const bool peer = !opt.host_halo && G > 1;
            // phase 1: this slot's device input is about to be overwritten -- the neighbours' kernels of batch - kRing
            // read their halo rows from it, so the upload queue waits for them (their ev_k is still live here)
            if (peer && s.busy) {
                if (w.top) blur_check(b200blur_enqueue_wait_peer(w.ctx, 0, workers[k - 1].ctx, workers[k - 1].ring[si].ev_k), "peer wait failed");
                if (w.bot) blur_check(b200blur_enqueue_wait_peer(w.ctx, 0, workers[k + 1].ctx, workers[k + 1].ring[si].ev_k), "peer wait failed");
            }
            if (peer) rendezvous.wait();
            // phase 2: recycle the slot, stage this batch's rows, upload
            if (s.busy) harvest(w, s);
            s.count = batch_count;
            s.batch = batch;
            if (s.staged < batch_count) fill_rows(w, s.h_in, batch_count);
            const size_t in_bytes = (size_t)batch_count * w.in_rows * pitch, out_bytes = (size_t)batch_count * w.rows * pitch;
            blur_check(b200blur_enqueue_write(w.ctx, 0, s.d_in, s.h_in, in_bytes, &s.ev_in), "GPU write failed");
            if (peer) rendezvous.wait();
            // phase 3: kernel after this GPU's upload and -- because it reads their edge rows -- the neighbours' uploads
            blur_check(b200blur_enqueue_wait(w.ctx, 1, s.ev_in), "GPU wait failed");
            if (peer) {
                if (w.top) blur_check(b200blur_enqueue_wait_peer(w.ctx, 1, workers[k - 1].ctx, workers[k - 1].ring[si].ev_in), "peer wait failed");
                if (w.bot) blur_check(b200blur_enqueue_wait_peer(w.ctx, 1, workers[k + 1].ctx, workers[k + 1].ring[si].ev_in), "peer wait failed");
            }
            b200blur_launch l;
            make_launch(w, s.d_in, s.d_out, batch_count, si, l);
            blur_check(b200blur_enqueue_blur(w.ctx, 1, &l, &s.ev_k), "GPU kernel launch failed");
            blur_check(b200blur_enqueue_wait(w.ctx, 2, s.ev_k), "GPU wait failed");
            blur_check(b200blur_enqueue_read(w.ctx, 2, s.h_out, s.d_out, out_bytes, &s.ev_out), "GPU read failed");
            s.busy = true;
            w.t.images += batch_count;
        }
        // drain: neighbours may still need this GPU's last uploads; events stay live until everyone is done
        blur_check(b200blur_finish_all(w.ctx), "finish failed");
    };

    if (opt.stage_once && !opt.resident)  // the stream is one image repeated: fill every slot here, outside the timer
        for (auto &w : workers)
            for (auto &r : w.ring) {
                fill_rows(w, r.h_in, per_dev_images);
                r.staged = per_dev_images;
                w.t.fill_ms = 0;
            }
    const double time_start_total = get_time_ms();
    std::vector<std::thread> threads;
    for (int k = 1; k < G; k++) threads.emplace_back(run_worker, std::ref(workers[k]));
    run_worker(workers[0]);
    for (auto &t : threads) t.join();
    if (!opt.resident)
        for (auto &w : workers)
            for (auto &s : w.ring)
                if (s.busy) harvest(w, s);
    const double time_end_total = get_time_ms();
    double time_total_processing = time_end_total - time_start_total;
    if (opt.resident) {
        time_total_processing = 0;
        for (auto &w : workers) time_total_processing = std::max(time_total_processing, w.resident_ms);
    }
    const long long passes = opt.resident ? opt.repeat : 1;
    printf("All batches finished!\n\n");

    // ======================== PERFORMANCE ANALYSIS (:615-721) ========================
    printf("========== PERFORMANCE RESULTS ==========\n\n");
    // Section numbers: the reference prints 2 = CPU, 3 = GPU, 4..9 (split_image_blur.c:617-721); with G bands the device
    // sections are 2 .. G+1 and the rest follow (G <= 2 keeps the reference's numbers).
    const int sec_base = G >= 2 ? G + 2 : 4;
    printf("1. OVERALL EXECUTION TIME\n");
    if (opt.resident)
        printf("   Device-resident kernel time (max over GPUs, %d pass(es)): %.3f ms\n", opt.repeat, time_total_processing);
    else
        printf("   Total wall-clock time: %.2f ms (%.2f seconds)\n", time_total_processing, time_total_processing / 1000.0);
    printf("   Total images processed: %lld\n\n", (long long)NUM_IMAGES * passes);

    for (int k = 0; k < G; k++) {
        const Worker &w = workers[k];
        const double tot = w.t.total();
        printf("%d. GPU %d DEVICE (processed %lld images - rows %d-%d, %d rows each)\n", 2 + k, k, w.t.images, w.row0,
               w.row0 + w.rows - 1, w.rows);
        printf("   Total GPU time:        %.2f ms\n", tot);
        printf("   - Transfer IN:         %.2f ms (%.1f%%)\n", w.t.in_ms, tot > 0 ? w.t.in_ms / tot * 100 : 0.0);
        printf("   - Kernel execution:    %.2f ms (%.1f%%)\n", w.t.kernel_ms, tot > 0 ? w.t.kernel_ms / tot * 100 : 0.0);
        printf("   - Transfer OUT:        %.2f ms (%.1f%%)\n", w.t.out_ms, tot > 0 ? w.t.out_ms / tot * 100 : 0.0);
        if (!opt.resident && opt.stage_once) printf("   Host staging: once per ring slot, before the timer (--stage-once)\n");
        else if (!opt.resident) printf("   Host staging (replicate band rows, %d thread(s)): %.2f ms\n", opt.fill_threads, w.t.fill_ms);
        printf("\n");
    }
    printf("============================\n");
    if (G > 1) {
        int fast = 0, slow = 0;
        for (int k = 1; k < G; k++) {
            if (workers[k].t.total() < workers[fast].t.total()) fast = k;
            if (workers[k].t.total() > workers[slow].t.total()) slow = k;
        }
        const double tf = workers[fast].t.total(), ts = workers[slow].t.total();
        printf("%d. DEVICE COMPARISON\n", sec_base);
        printf("   GPU %d is %.2fx FASTER than GPU %d\n", fast, tf > 0 ? ts / tf : 1.0, slow);
        printf("   slowest/fastest time ratio: %.2f\n\n", tf > 0 ? ts / tf : 1.0);
        printf("%d. WORKLOAD BALANCE\n", sec_base + 1);
        printf("   Workload imbalance: %.1f%%\n", ts > 0 ? fabs(ts - tf) / ts * 100.0 : 0.0);
        printf("   GPU %d is the BOTTLENECK (%.2f ms slower)\n\n", slow, ts - tf);
        printf("%d. BOTTLENECK IDENTIFICATION\n", sec_base + 2);
        for (int k = 0; k < G; k++) {
            const DeviceTimes &t = workers[k].t;
            printf("   GPU %d bottleneck: ", k);
            if (t.in_ms + t.out_ms > t.kernel_ms) printf("COMMUNICATION (%.1f%% of time)\n", (t.in_ms + t.out_ms) / t.total() * 100);
            else printf("COMPUTATION (%.1f%% of time)\n", t.kernel_ms / t.total() * 100);
        }
    }
    printf("\n");

    printf("%d. THROUGHPUT\n", sec_base + 3);
    const double secs = time_total_processing / 1000.0;
    const double n_done = (double)NUM_IMAGES * passes;
    printf("   Overall throughput: %.2f Megapixels/sec\n", n_done * width * height / secs / 1e6);
    printf("   Images per second: %.2f\n", n_done / secs);
    const double gbs = 2.0 * n_done * image_size / secs / 1e9;
    if (opt.resident)
        printf("   Algorithmic HBM traffic: %.1f GB/s over %d GPU(s) (%.1f%% of 8000 GB/s nominal per GPU)\n", gbs, G, gbs / G / 8000.0 * 100);
    else
        printf("   Host link traffic: %.2f GB/s each way over %d GPU(s)\n", gbs / 2, G);
    if (opt.checksum) {
        uint64_t h = 0;
        for (auto &w : workers) h ^= w.checksum * (uint64_t)(2 * w.gpu + 1);
        printf("   Output checksum (FNV-1a per band, combined): %016llx\n", (unsigned long long)h);
    }
    printf("\n=========================================\n\n");

    printf("%d. SPLIT-IMAGE STATISTICS\n", sec_base + 4);
    for (int k = 0; k < G; k++)
        printf("   GPU %d time per image: %.5f ms (for %d rows)\n", k, workers[k].t.total() / std::max(1LL, workers[k].t.images), workers[k].rows);
    printf("   Combined time per image: %.5f ms\n", time_total_processing / n_done);
    printf("   Current GPU ratio: %.1f%%\n\n", gpu_ratio * 100);

    printf("%d. OPTIMAL RATIO RECOMMENDATION\n", sec_base + 5);
    double inv_sum = 0;
    for (auto &w : workers) inv_sum += (double)w.t.images * w.rows / std::max(1e-9, w.t.total());
    for (int k = 0; k < G; k++) {
        const Worker &w = workers[k];
        const double per_row = w.t.total() / std::max(1.0, (double)w.t.images * w.rows);
        printf("   GPU %d: %.7f ms/row -> recommended share of rows %.1f%%\n", k, per_row,
               ((double)w.t.images * w.rows / std::max(1e-9, w.t.total())) / inv_sum * 100);
    }
    printf("   Run with: ./split_image_blur %.3f   (bands are even by construction; ratio kept for compatibility)\n\n", gpu_ratio);

    if (!opt.save.empty()) {
        if (save_ppm(opt.save.c_str(), first_output.data(), width, height)) printf("Saved example output: %s\n", opt.save.c_str());
        else printf("Error: cannot write %s\n", opt.save.c_str());
    }

    // ======================== CLEANUP (:723-745) ========================
    for (auto &w : workers) {
        for (auto &s : w.ring) {
            b200blur_host_free(s.h_in);
            b200blur_host_free(s.h_out);
            b200blur_dev_free(w.ctx, s.d_in);
            b200blur_dev_free(w.ctx, s.d_out);
        }
        if (w.d_res_in) b200blur_dev_free(w.ctx, w.d_res_in);
        if (w.d_res_out) b200blur_dev_free(w.ctx, w.d_res_out);
    }
    for (auto &w : workers) b200blur_ctx_destroy(w.ctx);
    return 0;
}
