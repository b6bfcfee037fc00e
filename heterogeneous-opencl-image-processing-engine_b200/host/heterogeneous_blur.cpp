// heterogeneous_blur.cpp -- Approach 1 (image-level distribution) on B200s, drop-in for the reference CLI.
//
//   ./heterogeneous_blur [cpu|gpu|both] [gpu_ratio] [batch_size] [--images N] [--input f.ppm | --width W --height H]
//                        [--gpus G] [--resident [--repeat R]] [--quiet] [--save out.ppm] [--checksum]
//
// Same positional surface, defaults, warnings and report sections as heterogeneous_blur.c (:38-100, :609-724).
// What changed underneath (BASELINE.json north_star):
//   * devices   -- the CPU+iGPU pair becomes G B200s: `cpu` and `gpu` run on ONE GPU, `both` on all visible GPUs.
//                  There is no CPU device any more, so `gpu_ratio` is accepted and echoed but the per-batch split
//                  (int)(batch_count*gpu_ratio) of :449-451 is replaced: by default the GPUs TAKE work from a shared
//                  counter (groups of whole batches, ~64 MB each), so a GPU that gets less of the box's shared host
//                  link simply takes fewer -- the measured shares are what the reference's ratio recommendation
//                  (:713-722) would have had to be tuned to by hand; --static-split keeps a fixed even partition of
//                  every batch instead.
//   * plumbing  -- every cl* call is the matching b200blur_* call (include/b200blur.h); the kernel is precompiled.
//   * pipeline  -- per GPU one host thread, three queues (H2D / blur / D2H) and a ring of pinned + device slots, so
//                  the write / kernel / read of different batches overlap instead of running back to back per image
//                  (:502-514).  batch_size keeps its meaning: images staged and launched together.
//   * kernel    -- one launch blurs a whole share of a batch (the reference launches once per image, :507).
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
    long long count = 0, first_image = 0;
    long long staged = 0;  // images of the source already replicated into h_in (--stage-once)
    bool busy = false;
};

struct Worker {
    int gpu = 0;
    b200blur_ctx *ctx = nullptr;
    std::vector<Slot> ring;
    DeviceTimes t;
    std::unique_ptr<StagingPool> staging;
    uint64_t checksum = 1469598103934665603ull;
    double resident_ms = 0;
    long long resident_launches = 0;
};

}  // namespace

int main(int argc, char **argv)
{
    // Configuration (heterogeneous_blur.c:41-48)
    int mode = 0;  // 0 = both, 1 = cpu, 2 = gpu
    int BATCH_SIZE = 500;
    int local_work_size = 16;
    float gpu_ratio = 0.5f;
    ExtraOptions opt;

    int npos = 1;
    while (npos < argc && strncmp(argv[npos], "--", 2) != 0) npos++;  // positional arguments end at the first --flag
    if (parse_extra(argc, argv, npos, opt) != 0) return -1;
    const int kRing = opt.ring;
    const int NUM_IMAGES = opt.num_images;

    if (npos > 1) {
        if (strcmp(argv[1], "cpu") == 0) {
            mode = 1;
            printf("Mode: CPU ONLY\n");
        } else if (strcmp(argv[1], "gpu") == 0) {
            mode = 2;
            printf("Mode: GPU ONLY\n");
        } else if (strcmp(argv[1], "both") == 0) {
            mode = 0;
            printf("Mode: HETEROGENEOUS (CPU + GPU)\n");
        } else {
            printf("Usage: %s [cpu|gpu|both]\n", argv[0]);
            printf("Defaulting to heterogeneous mode.\n");
        }
    } else {
        printf("Mode: HETEROGENEOUS (CPU + GPU) [default]\n");
    }
    if (npos > 2) {
        gpu_ratio = atof(argv[2]);
        if (gpu_ratio < 0.0f || gpu_ratio > 1.0f) {
            printf("Warning: gpu_ratio must be between 0.0 and 1.0. Using 0.5\n");
            gpu_ratio = 0.5f;
        }
    }
    if (npos > 3) {
        BATCH_SIZE = atoi(argv[3]);
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

    if (mode == 0) printf("GPU ratio: %.1f%% GPU, %.1f%% CPU\n", gpu_ratio * 100, (1 - gpu_ratio) * 100);
    printf("========== HETEROGENEOUS CONFIGURATION ==========\n");
    printf("Input file: %s\n", input_name.c_str());
    printf("Number of images in stream: %d\n", NUM_IMAGES);
    printf("Batch size: %d images\n", BATCH_SIZE);
    printf("Number of batches: %d\n", NUM_BATCHES);
    printf("Work-group size: %dx%d\n", local_work_size, local_work_size);
    printf("Execution mode : %d\n", mode);
    printf("================================================\n\n");

    const int width = img.width, height = img.height, channels = img.channels;
    printf("Original image loaded: %dx%d, %d channels\n", width, height, channels);
    const size_t image_size = img.size();
    printf("Size of one image: %zu bytes (%.2f KB)\n", image_size, image_size / 1024.0);
    const unsigned char *original_image = img.data.data();
    printf("Original image converted to interleaved format\n\n");

    // ======================== DEVICE DISCOVERY (heterogeneous_blur.c:140-191) ========================
    int n_dev = 0;
    if (b200blur_device_count(&n_dev) != B200BLUR_OK || n_dev == 0) {
        printf("Error: Could not find a CUDA device (%s)\n", b200blur_last_error());
        return -1;
    }
    printf("Platform 0: NVIDIA CUDA (%s)\n", b200blur_version());
    int G = (mode == 0) ? n_dev : 1;  // cpu / gpu: one device; both: every visible GPU
    if (opt.gpus > 0) G = opt.oversubscribe ? opt.gpus : std::min(opt.gpus, n_dev);
    if (opt.fill_threads <= 0) opt.fill_threads = auto_fill_threads(G);
    for (int k = 0; k < G; k++) {
        char dname[256];
        blur_check(b200blur_device_name(k % n_dev, dname, sizeof dname), "Failed to get device name");
        printf("GPU device %d: %s\n", k, dname);
    }
    if (mode == 0 && npos > 2)
        printf("Note: gpu_ratio %.3f is kept for CLI compatibility; %s\n", gpu_ratio,
               G > 1 && !opt.static_split && !opt.resident ? "the GPUs take groups of batches from a shared counter (work stealing)"
                                                            : "images are partitioned evenly over the GPU(s)");
    printf("\n");

    // ======================== CONTEXTS / QUEUES / BUFFERS (:194-355) ========================
    printf("Kernel objects created (precompiled sm_100a, no gaussian_kernel.cl needed at run time)\n\n");
    printf("Allocating device buffers...\n");
    std::vector<Worker> workers(G);
    const bool stealing = G > 1 && !opt.static_split && !opt.resident;
    long long max_share = 0;
    for (int k = 0; k < G; k++) {
        int64_t b, c;
        b200blur_partition(BATCH_SIZE, G, k, &b, &c);
        max_share = std::max<long long>(max_share, c);
    }
    if (stealing) max_share = BATCH_SIZE;   // a GPU takes whole batches
    // Batches are independent, so `fuse` consecutive batches (a GPU's shares of them) travel and launch together: ~64 MB
    // per transfer keeps the host link at its ceiling (8 MB transfers reach only ~33 of 45 GB/s, tools/e2e.py).
    long long fuse = (long long)((64.0 * 1024 * 1024) / ((double)max_share * image_size) + 0.5);
    // keep >= 16 pipeline steps per GPU (and, when the GPUs take work dynamically, enough steps to balance them)
    fuse = std::max(1LL, std::min<long long>(fuse, std::max(1, NUM_BATCHES / (stealing ? 16 * G : 16))));
    if (opt.fuse > 0) fuse = opt.fuse;
    // ring slot = `fuse` batch shares, or -- for batches far above the ~64 MB transfer size -- an even piece of one share
    long long slot_images = max_share * fuse;
    if (fuse == 1 && (double)max_share * image_size > 96.0 * 1024 * 1024 && opt.fuse <= 0) {
        const long long pieces = (long long)((double)max_share * image_size / (64.0 * 1024 * 1024) + 0.5);
        slot_images = (max_share + pieces - 1) / pieces;
    }
    for (int k = 0; k < G; k++) {
        Worker &w = workers[k];
        w.gpu = k;
        blur_check(b200blur_ctx_create(k % n_dev, 3, &w.ctx), "Failed to create context");
        if (!opt.resident) {
            w.ring.resize(kRing);
            for (auto &s : w.ring) {
                blur_check(b200blur_host_alloc(slot_images * image_size, (void **)&s.h_in), "Failed to allocate pinned input");
                blur_check(b200blur_host_alloc(slot_images * image_size, (void **)&s.h_out), "Failed to allocate pinned output");
                blur_check(b200blur_dev_alloc(w.ctx, slot_images * image_size, &s.d_in), "Failed to create input buffer");
                blur_check(b200blur_dev_alloc(w.ctx, slot_images * image_size, &s.d_out), "Failed to create output buffer");
            }
        }
    }
    // Like the reference, whose program build (clBuildProgram, :257-322) happens before its timer starts: load the
    // kernel image and wake the copy paths with one tiny write -> blur -> read per GPU, and start the staging threads.
    if (!opt.resident)
        for (auto &w : workers) {
            w.staging.reset(new StagingPool(opt.fill_threads));
            Slot &s = w.ring[0];
            memcpy(s.h_in, original_image, image_size);
            b200blur_launch l;
            blur_check(b200blur_enqueue_write(w.ctx, 0, s.d_in, s.h_in, image_size, NULL), "GPU write failed");
            blur_check(b200blur_launch_rows(&l, s.d_in, s.d_out, width, height, channels, 0, height, 1, image_size, image_size),
                       "Failed to set kernel args");
            blur_check(b200blur_enqueue_blur(w.ctx, 0, &l, NULL), "GPU kernel launch failed");
            blur_check(b200blur_enqueue_read(w.ctx, 0, s.h_out, s.d_out, image_size, NULL), "GPU read failed");
            blur_check(b200blur_finish_all(w.ctx), "finish failed");
            if (opt.stage_once)  // the stream is one image repeated: fill every slot here, outside the timer, and re-send
                for (auto &r : w.ring) {
                    w.staging->replicate(r.h_in, original_image, image_size, slot_images);
                    r.staged = slot_images;
                }
        }
    printf("Device buffers allocated\n\n");
    printf("Global work size: %d x %d per image, up to %lld image(s) per launch (%lld batch share(s) fused)\n",
           (width + 15) / 16 * 16, (height + 15) / 16 * 16, slot_images, fuse);
    printf("Local work size: 16 x 16 (reference geometry; the CUDA kernel tiles 16-byte columns x row strips)\n\n");

    printf("Starting batch processing of %d images in %d batches...\n\n", NUM_IMAGES, NUM_BATCHES);
    std::vector<unsigned char> first_output;
    if (!opt.save.empty()) first_output.resize(image_size);

    // ======================== BATCH LOOP (:418-600), one host thread per GPU ========================
    auto harvest = [&](Worker &w, Slot &s) {
        double ms;
        blur_check(b200blur_event_ms(w.ctx, s.ev_out, &ms), "Failed to read transfer-out time");  // waits for the read
        w.t.out_ms += ms;
        blur_check(b200blur_event_ms(w.ctx, s.ev_in, &ms), "Failed to read transfer-in time");
        w.t.in_ms += ms;
        blur_check(b200blur_event_ms(w.ctx, s.ev_k, &ms), "Failed to read kernel time");
        w.t.kernel_ms += ms;
        b200blur_event_release(w.ctx, s.ev_in);
        b200blur_event_release(w.ctx, s.ev_k);
        b200blur_event_release(w.ctx, s.ev_out);
        if (opt.checksum) w.checksum = fnv1a(s.h_out, (size_t)s.count * image_size, w.checksum);
        if (!opt.save.empty() && s.first_image == 0 && s.count > 0) memcpy(first_output.data(), s.h_out, image_size);
        s.busy = false;
    };

    std::atomic<long long> next_group{0};
    std::vector<long long> issued_groups(G, 0);
    auto run_worker = [&](Worker &w) {
        const int k = w.gpu;
        if (opt.resident) {
            // Device-resident mode: this GPU's share of the stream lives in HBM; only kernels are timed.
            int64_t begin, count;
            b200blur_partition(NUM_IMAGES, G, k, &begin, &count);
            void *d_in, *d_out, *h;
            blur_check(b200blur_dev_alloc(w.ctx, count * image_size, &d_in), "Failed to create input buffer");
            blur_check(b200blur_dev_alloc(w.ctx, count * image_size, &d_out), "Failed to create output buffer");
            const int64_t stage = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(count, 256), (256ll << 20) / (int64_t)image_size));
            blur_check(b200blur_host_alloc(stage * image_size, &h), "Failed to allocate pinned staging");
            for (int64_t i = 0; i < stage; i++) memcpy((unsigned char *)h + i * image_size, original_image, image_size);
            for (int64_t i = 0; i < count; i += stage) {
                const int64_t n = std::min(stage, count - i);
                blur_check(b200blur_enqueue_write(w.ctx, 0, (unsigned char *)d_in + i * image_size, h, n * image_size, NULL),
                           "GPU write failed");
            }
            blur_check(b200blur_finish(w.ctx, 0), "finish failed");
            b200blur_stats st;
            blur_check(b200blur_run_resident(w.ctx, d_in, d_out, width, height, channels, count, BATCH_SIZE, 1, &st),
                       "warm-up failed");
            for (int r = 0; r < opt.repeat; r++) {
                blur_check(b200blur_run_resident(w.ctx, d_in, d_out, width, height, channels, count, BATCH_SIZE, 1, &st),
                           "GPU kernel launch failed");
                w.resident_ms += st.kernel_ms;
                w.resident_launches += st.launches;
            }
            w.t.kernel_ms = w.resident_ms;
            w.t.images = count * opt.repeat;
            if (!opt.save.empty() && k == 0) {
                blur_check(b200blur_enqueue_read(w.ctx, 0, h, d_out, image_size, NULL), "GPU read failed");
                blur_check(b200blur_finish(w.ctx, 0), "finish failed");
                memcpy(first_output.data(), h, image_size);
            }
            b200blur_host_free(h);
            b200blur_dev_free(w.ctx, d_in);
            b200blur_dev_free(w.ctx, d_out);
            return;
        }
        long long issued = 0;
        for (;;) {
            // next group of `fuse` batches: taken from the shared counter (work stealing), or every group in turn
            // with this GPU's even share of each batch (--static-split)
            long long batch0;
            if (stealing) {
                batch0 = next_group.fetch_add(1, std::memory_order_relaxed) * fuse;
            } else {
                batch0 = issued_groups[k] * fuse;
                issued_groups[k]++;
            }
            if (batch0 >= NUM_BATCHES) break;
            long long count = 0, first_image = -1;
            for (long long batch = batch0; batch < std::min<long long>(NUM_BATCHES, batch0 + fuse); batch++) {
                const long long batch_start = batch * BATCH_SIZE;
                long long batch_count = BATCH_SIZE;
                if (batch_start + batch_count > NUM_IMAGES) batch_count = NUM_IMAGES - batch_start;  // :423-427
                if ((stealing || k == 0) && !opt.quiet) {
                    if (stealing) {
                        printf("=== Processing Batch %lld/%d === taken by GPU %d (%lld images)\n", batch + 1, NUM_BATCHES, k, batch_count);
                    } else {
                        printf("=== Processing Batch %lld/%d ===\n", batch + 1, NUM_BATCHES);
                        printf("  Batch work distribution:");
                        for (int j = 0; j < G; j++) {
                            int64_t b, c;
                            b200blur_partition(batch_count, G, j, &b, &c);
                            printf(" GPU%d=%lld", j, (long long)c);
                        }
                        printf("\n");
                    }
                }
                int64_t begin = 0, c = batch_count;
                if (!stealing) b200blur_partition(batch_count, G, k, &begin, &c);  // replaces (int)(batch_count * gpu_ratio), :449-451
                if (c > 0 && first_image < 0) first_image = batch_start + begin;
                count += c;
            }
            if (count == 0) continue;
            // a share larger than a ring slot (a big batch_size) moves in slot-sized pieces: images are independent
            for (long long done = 0; done < count;) {
                const long long piece = std::min(slot_images, count - done);
                Slot &s = w.ring[issued % kRing];
                if (s.busy) harvest(w, s);
                s.count = piece;
                s.first_image = first_image + done;
                // CREATE BATCH IMAGE STREAM (:431-442): replicate the source image into this share's staging slots
                if (s.staged < piece) {
                    const double tf = get_time_ms();
                    w.staging->replicate(s.h_in, original_image, image_size, piece);
                    w.t.fill_ms += get_time_ms() - tf;
                    if (opt.stage_once) s.staged = piece;
                }
                const size_t bytes = (size_t)piece * image_size;
                blur_check(b200blur_enqueue_write(w.ctx, 0, s.d_in, s.h_in, bytes, &s.ev_in), "GPU write failed");
                blur_check(b200blur_enqueue_wait(w.ctx, 1, s.ev_in), "GPU wait failed");
                b200blur_launch l;
                blur_check(b200blur_launch_rows(&l, s.d_in, s.d_out, width, height, channels, 0, height, piece, image_size, image_size),
                           "Failed to set kernel args");
                blur_check(b200blur_enqueue_blur(w.ctx, 1, &l, &s.ev_k), "GPU kernel launch failed");
                blur_check(b200blur_enqueue_wait(w.ctx, 2, s.ev_k), "GPU wait failed");
                blur_check(b200blur_enqueue_read(w.ctx, 2, s.h_out, s.d_out, bytes, &s.ev_out), "GPU read failed");
                s.busy = true;
                w.t.images += piece;
                issued++;
                done += piece;
            }
        }
        for (auto &s : w.ring)
            if (s.busy) harvest(w, s);
        blur_check(b200blur_finish_all(w.ctx), "finish failed");  // clFinish (:538-539)
        w.t.busy_until_ms = get_time_ms();
    };

    const double time_start_total = get_time_ms();
    std::vector<std::thread> threads;
    for (int k = 1; k < G; k++) threads.emplace_back(run_worker, std::ref(workers[k]));
    run_worker(workers[0]);
    for (auto &t : threads) t.join();
    const double time_end_total = get_time_ms();
    double time_total_processing = time_end_total - time_start_total;
    long long images_done = 0;
    for (auto &w : workers) images_done += w.t.images;
    if (opt.resident) {
        time_total_processing = 0;
        for (auto &w : workers) time_total_processing = std::max(time_total_processing, w.resident_ms);
    }
    printf("All batches finished!\n\n");

    // ======================== PERFORMANCE ANALYSIS (:609-724) ========================
    printf("========== PERFORMANCE RESULTS ==========\n\n");
    printf("1. OVERALL EXECUTION TIME\n");
    if (opt.resident)
        printf("   Device-resident kernel time (max over GPUs, %d pass(es)): %.3f ms\n", opt.repeat, time_total_processing);
    else
        printf("   Total wall-clock time: %.2f ms (%.2f seconds)\n", time_total_processing, time_total_processing / 1000.0);
    printf("   Total images processed: %lld\n\n", images_done);

    // Section numbers: the reference prints 2 = CPU, 3 = GPU, 4..8 = comparison .. recommendation (:621-722).  With G
    // GPUs the device sections are 2 .. G+1 and the rest follow; G = 2 gives the reference's numbers, a single GPU keeps
    // the reference's "2" (cpu) / "3" (gpu) and its fixed "7. THROUGHPUT".
    const int sec_base = G >= 2 ? G + 2 : 4;   // number of the DEVICE COMPARISON section
    for (int k = 0; k < G; k++) {
        const DeviceTimes &t = workers[k].t;
        if (t.images == 0) continue;
        const double tot = t.total();
        printf("%d. GPU %d DEVICE (processed %lld images, %.1f%% of the stream)\n", G == 1 && mode == 2 ? 3 : 2 + k, k, t.images,
               images_done ? 100.0 * t.images / images_done : 0.0);
        printf("   Total GPU time:        %.2f ms\n", tot);
        printf("   - Transfer IN:         %.2f ms (%.1f%%)\n", t.in_ms, tot > 0 ? t.in_ms / tot * 100 : 0.0);
        printf("   - Kernel execution:    %.2f ms (%.1f%%)\n", t.kernel_ms, tot > 0 ? t.kernel_ms / tot * 100 : 0.0);
        printf("   - Transfer OUT:        %.2f ms (%.1f%%)\n", t.out_ms, tot > 0 ? t.out_ms / tot * 100 : 0.0);
        printf("   Average per image:     %.5f ms\n", tot / t.images);
        if (!opt.resident && opt.stage_once) printf("   Host staging: once per ring slot, before the timer (--stage-once)\n");
        else if (!opt.resident) printf("   Host staging (replicate source image, %d thread(s)): %.2f ms\n", opt.fill_threads, t.fill_ms);
        printf("\n");
    }
    printf("====================\n");

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
            if (t.images == 0) continue;
            printf("   GPU %d bottleneck: ", k);
            if (t.in_ms + t.out_ms > t.kernel_ms)
                printf("COMMUNICATION (%.1f%% of time)\n", (t.in_ms + t.out_ms) / t.total() * 100);
            else
                printf("COMPUTATION (%.1f%% of time)\n", t.kernel_ms / t.total() * 100);
        }
    }
    printf("\n");

    printf("%d. THROUGHPUT\n", sec_base + 3);
    const double secs = time_total_processing / 1000.0;
    printf("   Overall throughput: %.2f Megapixels/sec\n", (double)images_done * width * height / secs / 1e6);
    printf("   Images per second: %.2f\n", images_done / secs);
    const double gbs = 2.0 * images_done * image_size / secs / 1e9;
    if (opt.resident)
        printf("   Algorithmic HBM traffic: %.1f GB/s over %d GPU(s) (%.1f%% of 8000 GB/s nominal per GPU)\n", gbs, G,
               gbs / G / 8000.0 * 100);
    else
        printf("   Host link traffic: %.2f GB/s each way over %d GPU(s)\n", gbs / 2, G);
    if (opt.checksum) {
        uint64_t h = 0;
        for (auto &w : workers) h ^= w.checksum;
        printf("   Output checksum (FNV-1a, xor over GPUs): %016llx\n", (unsigned long long)h);
    }
    printf("\n=========================================\n\n");

    if (G > 1) {
        printf("%d. OPTIMAL RATIO RECOMMENDATION\n", sec_base + 4);
        printf("   Based on measured performance:\n");
        double inv_sum = 0;
        for (int k = 0; k < G; k++)
            if (workers[k].t.images) inv_sum += workers[k].t.images / workers[k].t.total();
        for (int k = 0; k < G; k++) {
            const DeviceTimes &t = workers[k].t;
            if (!t.images) continue;
            printf("   GPU %d: %.5f ms/image -> recommended share %.1f%%, share taken %.1f%%\n", k, t.total() / t.images,
                   (t.images / t.total()) / inv_sum * 100, images_done ? 100.0 * t.images / images_done : 0.0);
        }
        if (!opt.resident) {
            // how evenly the GPUs finished: wall-clock time of each GPU's last result relative to the slowest
            double last = 0, first = 1e300;
            for (auto &w : workers) {
                last = std::max(last, w.t.busy_until_ms);
                first = std::min(first, w.t.busy_until_ms);
            }
            printf("   Finish-time spread over GPUs: %.2f ms (%.1f%% of the run; %s)\n", last - first,
                   time_total_processing > 0 ? (last - first) / time_total_processing * 100 : 0.0,
                   stealing ? "work taken dynamically from a shared counter" : "fixed even partition of every batch");
        }
        printf("   Run with: ./heterogeneous_blur both %.3f   (ratio kept for compatibility; shares are %s)\n\n", gpu_ratio,
               stealing ? "measured, not set" : "even by construction");
    }

    if (!opt.save.empty()) {
        if (save_ppm(opt.save.c_str(), first_output.data(), width, height)) printf("Saved example output: %s\n", opt.save.c_str());
        else printf("Error: cannot write %s\n", opt.save.c_str());
    }

    // ======================== CLEANUP (:725-747) ========================
    for (auto &w : workers) {
        for (auto &s : w.ring) {
            b200blur_host_free(s.h_in);
            b200blur_host_free(s.h_out);
            b200blur_dev_free(w.ctx, s.d_in);
            b200blur_dev_free(w.ctx, s.d_out);
        }
        b200blur_ctx_destroy(w.ctx);
    }
    return 0;
}
