#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the one hot path (3x3 Gaussian blur of an RGB image stream).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Approach 1, 5000 x 320x240 RGB uint8 per GPU, batch_size 35, seeded synthetic
images.  One STEP = one pass of the hot path over the whole 5000-image stream of a GPU.

  value      device-resident images/s, whole job (all ranks): inputs already in HBM, K steps timed with CUDA events on
             the launching queue between barriers, max over ranks.
  e2e        the same stream through the reference-facing C-ABI call b200blur_run_host with HOST (pinned) buffers:
             every step copies all inputs host->device and all results device->host inside the timed region.
  roofline   the stencil kernel's algorithmic bytes (2*W*H*3 per image x images per launch) / its launch duration,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference kernel (oracle/_ref: gaussian_kernel.cl compiled unmodified; else the oracle port) on
             this box's host cores over a bounded sample of the same stream (rank 0, N=1 only).

--impl reference times that CPU implementation alone on the same config and prints the same JSON line.
PyTorch is used only as plumbing here (device memory for the synthetic stream, pinned buffers, torch.distributed
barrier / max-reduce); the blur itself is libb200blur.so called through ctypes.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# BASELINE.json configs[1]
N_IMAGES, HEIGHT, WIDTH, CHANNELS, BATCH = 5000, 240, 320, 3, 35
IMAGE_BYTES = HEIGHT * WIDTH * CHANNELS
ALGO_BYTES_PER_IMAGE = 2 * IMAGE_BYTES  # every input byte read once, every output byte written once
METRIC = "images/sec (3x3 Gaussian blur stream; value = device-resident, e2e = incl. host<->device copies)"
UNIT = "images/s"
WORKLOAD = "A1 image-level: 5000x 320x240 RGB uint8 per GPU, batch_size=35 (BASELINE.json configs[1])"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant override (0 = auto)")
    ap.add_argument("--per-batch", action="store_true", help="resident run launches once per batch (no coalescing)")
    ap.add_argument("--scheme", choices=["image", "split"], default="image",
                    help="image = Approach 1 whole-image shards (default, the contract workload); split = Approach 2 row "
                         "bands of 5000 x 256x256 RGB with halo rows read from the neighbour GPU over NVLink (configs[2])")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- CPU arm helpers
def _cpu_impl():
    """-> (kind, batch_fn, threads): the reference kernel source when oracle/_ref is present, else the oracle port."""
    from oracle import oracle
    oracle.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
    if oracle.have_ref():
        return "reference", oracle.ref_blur_batch, oracle.ref_num_threads()
    return "port", oracle.c_blur_batch, oracle.num_threads()


def _synth_host(seed: int, n: int):
    import numpy as np
    return np.random.default_rng(seed).integers(0, 256, size=(n, HEIGHT, WIDTH, CHANNELS), dtype=np.uint8)


def time_cpu_baseline(target_seconds: float = 12.0):
    """Bounded sample of the stream on the host cores: a pilot sizes the sample to ~target_seconds of CPU work."""
    kind, fn, threads = _cpu_impl()
    pilot = _synth_host(1, 64)
    fn(pilot[:8])
    t = time.perf_counter()
    fn(pilot)
    per_img = (time.perf_counter() - t) / len(pilot)
    n = int(max(64, min(N_IMAGES, target_seconds / per_img)))
    passes = int(max(1, min(50, target_seconds / (per_img * n))))
    x = _synth_host(2, n)
    t = time.perf_counter()
    for _ in range(passes):
        fn(x)
    dt = time.perf_counter() - t
    return {"value": n * passes / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{n} of the {N_IMAGES} synthetic 320x240 RGB images x {passes} pass(es), {dt:.2f} s, "
                      f"{'gaussian_kernel.cl compiled unmodified (oracle/_ref), OpenMP over work-group rows' if kind == 'reference' else 'oracle C port, OpenMP over rows'}"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, fn, threads = _cpu_impl()
    # each step = a bounded sample of the workload; sized from a pilot so the whole run ends within a few minutes
    pilot = _synth_host(1, 32)
    fn(pilot[:8])
    t = time.perf_counter()
    fn(pilot)
    per_img = (time.perf_counter() - t) / len(pilot)
    budget = 90.0 / max(1, args.steps + args.warmup)
    n = int(max(32, min(N_IMAGES, budget / per_img)))
    x = _synth_host(3, n)
    for _ in range(args.warmup):
        fn(x)
    t = time.perf_counter()
    for _ in range(args.steps):
        fn(x)
    dt = time.perf_counter() - t
    value = n * args.steps / dt
    sample = f"{n} of the {N_IMAGES} images per step ({'reference kernel source via oracle/_ref' if kind == 'reference' else 'oracle port'}, {threads} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 accumulate of u8 (reference kernel)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": n, "width": WIDTH, "height": HEIGHT, "channels": CHANNELS,
                   "batch_size": BATCH, "device": "host CPU", "threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            hi = [v for v in sm if v >= 0.5 * max(sm)]  # samples taken under load
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this process to the CPU cores next to its GPU (NVML's ideal CPU affinity) BEFORE the pinned staging buffers
    are allocated, so they are first-touched on the GPU's own NUMA node.  One process per GPU on a two-socket box
    otherwise lands half the ranks' staging memory on the far socket.  Best effort: returns a note for the JSON."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [i for i in range(n_cpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"bound to {len(allowed)} cores near GPU {gpu_index} ({allowed[0]}-{allowed[-1]})"
        return "NVML affinity empty; not bound"
    except Exception as e:  # no NVML, container restrictions, ...
        return f"not bound ({type(e).__name__})"


# ----------------------------------------------------------------------------------------------------- product arm
def run_b200_arm(args) -> None:
    import torch
    import torch.distributed as dist
    import b200blur

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if not os.path.exists(b200blur.lib_path()):
        raise SystemExit("libb200blur.so is not built: the product path has no fallback (run __graft_entry__.build())")
    ctx = b200blur.Context(local_rank, 4)
    if args.variant:
        ctx.set_kernel_variant(args.variant)

    # synthetic stream of this rank: images [rank*5000, (rank+1)*5000) of the whole job (weak scaling)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    d_in = torch.randint(0, 256, (N_IMAGES, HEIGHT, WIDTH, CHANNELS), dtype=torch.uint8, device=dev, generator=g)
    d_out = torch.zeros_like(d_in)
    coalesce = not args.per_batch

    # ---- device-resident: W warm-up steps, then exactly K timed steps between barriers
    for _ in range(args.warmup):
        ctx.run_resident(d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, coalesce, stats=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count
    barrier()
    ev0 = ctx.enqueue_marker(0)
    for _ in range(args.steps):
        ctx.run_resident(d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, coalesce, stats=False)
    ev1 = ctx.enqueue_marker(0)
    ctx.finish(0)
    barrier()
    resident_ms = ctx.elapsed_ms(ev0, ev1)
    resident_launches = ctx.launch_count - launches0
    resident_ms = max_over_ranks(resident_ms)
    ms_per_step = resident_ms / args.steps
    value = world * N_IMAGES * args.steps / (resident_ms * 1e-3)

    # the same pass with the reference's launch granularity (one launch per batch of 35: 143 launches spread over the
    # context's queues, replayed as a CUDA graph) -- reported beside `value`, not instead of it
    per_batch_value = None
    if coalesce:
        for _ in range(3):
            ctx.run_resident(d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, False, stats=False)
        ctx.finish()
        barrier()
        p0 = ctx.enqueue_marker(0)
        for _ in range(args.steps):
            ctx.run_resident(d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, False, stats=False)
        p1 = ctx.enqueue_marker(0)
        ctx.finish()
        barrier()
        pb_ms = max_over_ranks(ctx.elapsed_ms(p0, p1))
        per_batch_value = world * N_IMAGES * args.steps / (pb_ms * 1e-3)

    # same-box practical ceiling: a plain device-to-device copy of the same 1.15 GB buffer (torch's copy kernel)
    copy_gbps = None
    if rank == 0:
        scratch = torch.empty_like(d_in)
        for _ in range(3):
            scratch.copy_(d_in)
        torch.cuda.synchronize()
        best = None
        for _ in range(5):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            scratch.copy_(d_in)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
            best = ms if best is None else min(best, ms)
        copy_gbps = 2.0 * N_IMAGES * IMAGE_BYTES / (best * 1e-3) / 1e9
        del scratch

    # sanity on the timed output (cheap, outside the timed region): a few images against the oracle on rank 0
    parity = None
    if rank == 0:
        try:
            from oracle import oracle
            import numpy as np
            idx = list(range(0, N_IMAGES, N_IMAGES // 32))[:32]
            got = d_out[idx].cpu().numpy()
            want = oracle.c_blur_batch(d_in[idx].cpu().numpy(), integer=True)
            diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
            # north_star: report the max-abs-diff histogram (expected: all mass at 0 -- the arithmetic is exact)
            parity = {"bit_exact": bool((got == want).all()), "max_abs_diff": int(diff.max()),
                      "abs_diff_histogram_0_1_2_3plus": [int((diff == 0).sum()), int((diff == 1).sum()),
                                                         int((diff == 2).sum()), int((diff >= 3).sum())],
                      "sample": f"{len(idx)} of {N_IMAGES} images of the timed output vs the oracle"}
        except Exception as e:  # the checker is optional for the measurement itself
            parity = f"unchecked: {e}"

    # ---- end to end through the C-ABI stream engine with pinned host buffers
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
    h_in = torch.empty((N_IMAGES, HEIGHT, WIDTH, CHANNELS), dtype=torch.uint8).pin_memory()
    h_in.copy_(d_in)
    h_out = torch.empty_like(h_in).pin_memory()
    for _ in range(min(args.warmup, 2)):
        ctx.run_host(h_in, h_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, stats=False)
    barrier()
    launches1 = ctx.launch_count
    e0 = ctx.enqueue_marker(0)
    last = None
    for _ in range(e2e_steps):
        last = ctx.run_host(h_in, h_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, stats=True)
    e1 = ctx.enqueue_marker(2)
    ctx.finish()
    barrier()
    e2e_ms = max_over_ranks(ctx.elapsed_ms(e0, e1))
    e2e_launches = ctx.launch_count - launches1
    e2e_value = world * N_IMAGES * e2e_steps / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0 and isinstance(parity, dict):
        parity["e2e_equals_resident"] = bool((h_out[:64].numpy() == d_out[:64].cpu().numpy()).all())

    # host-link roofline for the e2e number: the same pinned buffers moved both ways at once in 64 MB linear copies on
    # two streams, no kernel (rank 0, single-GPU run only)
    link_gbps = None
    if rank == 0 and world == 1:
        flat_in, flat_out = h_in.view(-1), h_out.view(-1)
        dflat_in, dflat_out = d_in.view(-1), d_out.view(-1)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        chunk = 64 << 20
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            s1.wait_event(t0)
            s2.wait_event(t0)
            for off in range(0, flat_in.numel(), chunk):
                with torch.cuda.stream(s1):
                    dflat_in[off:off + chunk].copy_(flat_in[off:off + chunk], non_blocking=True)
                with torch.cuda.stream(s2):
                    flat_out[off:off + chunk].copy_(dflat_out[off:off + chunk], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
            best = ms if best is None else min(best, ms)
        link_gbps = N_IMAGES * IMAGE_BYTES / (best * 1e-3) / 1e9

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = time_cpu_baseline()
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        peaks, peak_note = None, "fallback 6650 GB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak, peak_note = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy kernel, of measured)"
        except Exception:
            peak = 6650.0
        launches_per_step = resident_launches / args.steps
        images_per_launch = N_IMAGES / launches_per_step
        launch_ms = ms_per_step / launches_per_step
        achieved = ALGO_BYTES_PER_IMAGE * images_per_launch / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 (exact integer arithmetic in packed 16-bit lanes)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": N_IMAGES, "width": WIDTH, "height": HEIGHT,
                       "channels": CHANNELS, "batch_size": BATCH, "parallelism": f"image-shard x{world} (no collective)",
                       "resident_launches_per_step": launches_per_step,
                       "resident_mode": "coalesced batches" if coalesce else "one launch per batch",
                       "value_one_launch_per_batch": per_batch_value,
                       "l2": "inputs larger than L2 (1.15 GB in + 1.15 GB out per step vs 126 MB L2)",
                       "kernel_variant": args.variant, "parity_vs_oracle": parity},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "blur_stream_kernel<3,8,4> (TMA-bulk streamed stencil)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_IMAGE * images_per_launch,
                         "launch_ms": launch_ms, "peak_source": peak_note, "frac_of_nominal_8000": achieved / 8000.0,
                         "same_box_d2d_copy_GBps": copy_gbps},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_IMAGES * IMAGE_BYTES,
                    "d2h_bytes_per_step": N_IMAGES * IMAGE_BYTES, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "host_link_GBps_each_way": N_IMAGES * IMAGE_BYTES * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "stage_ms_last_step": {"h2d": last.h2d_ms, "kernel": last.kernel_ms, "d2h": last.d2h_ms},
                    "host_affinity": numa,
                    "link_roofline": None if link_gbps is None else {
                        "bound": "host link, both directions active (same pinned buffers, 64 MB linear copies, no kernel)",
                        "peak_GBps_each_way": link_gbps,
                        "frac": (N_IMAGES * IMAGE_BYTES * e2e_steps / (e2e_ms * 1e-3) / 1e9) / link_gbps},
                    "api": "b200blur_run_host (pinned host buffers, 3 queues, 4-slot device ring, batches fused into ~64 MB transfer chunks)"},
            "gpu_launches": int(resident_launches + e2e_launches),
            "clocks": clocks,
            "published_reference_context": {"a1_best_images_per_s": 8568, "hardware": "i7-12700 + UHD 770",
                                            "source": "BASELINE.md section 1 (end-to-end wall clock)"},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


class _DevView:
    """Exposes a raw device allocation to torch (as plumbing) through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def run_split_arm(args) -> None:
    """Approach 2 under one process per GPU (BASELINE.json configs[2]): 5000 x 256x256 RGB, every image cut into
    WORLD row bands, band k resident on GPU k, halo rows read inside the stencil kernel from the neighbour GPU's memory
    (CUDA IPC handles exchanged once, NVLink peer loads).  Strong scaling: the stream is fixed, bands shrink with N."""
    import torch
    import torch.distributed as dist
    import b200blur
    from b200blur.sharding import plan_bands

    n, h, w, c = 5000, 256, 256, 3
    P = w * c
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = b200blur.Context(local_rank, 4)
    plans = plan_bands(h, world)
    me = plans[rank]
    g = torch.Generator(device=dev).manual_seed(2002)          # same stream on every rank; each keeps only its band
    stream = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device=dev, generator=g)
    nbytes = n * me.rows * P
    d_band, d_out = ctx.dev_alloc(nbytes), ctx.dev_alloc(nbytes)
    torch.as_tensor(_DevView(d_band, nbytes), device=dev).copy_(stream[:, me.row0:me.row0 + me.rows].reshape(-1))
    sample_idx = [0, 1, n // 2, n - 1]
    sample = stream[sample_idx].cpu().numpy() if rank == 0 else None
    del stream
    torch.cuda.synchronize()
    launch = ctx.launch_rows(d_band, d_out, w, me.rows, c, 0, me.rows, n)
    opened = []
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, ctx.ipc_export(d_band))
        if me.has_top:
            up = plans[rank - 1]
            base = ctx.ipc_open(handles[rank - 1])
            opened.append(base)
            launch.halo_top, launch.halo_top_stride = base + (up.rows - 1) * P, up.rows * P
        if me.has_bottom:
            dn = plans[rank + 1]
            base = ctx.ipc_open(handles[rank + 1])
            opened.append(base)
            launch.halo_bottom, launch.halo_bottom_stride = base, dn.rows * P

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        ctx.enqueue_blur(0, launch)
    ctx.finish(0)
    barrier()
    e0 = ctx.enqueue_marker(0)
    for _ in range(args.steps):
        ctx.enqueue_blur(0, launch)
    e1 = ctx.enqueue_marker(0)
    ctx.finish(0)
    barrier()
    ms = ctx.elapsed_ms(e0, e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    parity = None
    if rank == 0:
        from oracle import oracle
        out = torch.as_tensor(_DevView(d_out, nbytes), device=dev).view(n, me.rows, w, c)[sample_idx].cpu().numpy()
        want = oracle.c_blur_batch(sample, integer=True)[:, me.row0:me.row0 + me.rows]
        parity = bool((out == want).all())
        peak = 6650.0
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            pass
        per_gpu_bytes = 2.0 * n * me.rows * P
        achieved = per_gpu_bytes / (ms / args.steps * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": n * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8 (exact integer arithmetic in packed 16-bit lanes)", "data": "synthetic",
            "config": {"workload": "A2 split-image: 5000x 256x256 RGB, row bands + 1-row halo over NVLink (BASELINE.json configs[2])",
                       "bands": [[p.row0, p.rows] for p in plans], "halo": "peer loads inside the stencil kernel (CUDA IPC)",
                       "parity_vs_oracle_rank0_band": parity},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "blur_stream_kernel<3,8,4>", "per": "GPU (rank 0's band)"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(ctx.launch_count)}), flush=True)
    barrier()
    for b in opened:
        ctx.ipc_close(b)
    ctx.dev_free(d_band)
    ctx.dev_free(d_out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.scheme == "split":
        run_split_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
