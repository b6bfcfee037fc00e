#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the one hot path (3x3 Gaussian blur of an RGB image stream).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload of `value` (BASELINE.json configs[1]): Approach 1, 5000 x 320x240 RGB uint8 per GPU, batch_size 35, seeded
synthetic images.  One STEP = one pass of the hot path over the whole 5000-image stream of a GPU.

  value      device-resident images/s, whole job (all ranks): inputs already in HBM, K steps timed with CUDA events on
             the launching queue between barriers, max over ranks.  Batches are fused into one launch (`config.
             resident_mode`); the same pass with ONE WORK DESCRIPTOR PER BATCH of 35 (the reference's batch_size
             granularity, heterogeneous_blur.c:418-539) is `per_batch.value`.
  e2e        the same stream through the reference-facing C-ABI call b200blur_run_host with HOST (pinned) buffers:
             every step copies all inputs host->device and all results device->host inside the timed region.
             e2e.link_roofline = the same pinned buffers moved both ways by ALL ranks at once with no kernel.
  roofline   the stencil kernel's algorithmic bytes (2*W*H*3 per image x images per launch) / its launch duration,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  sustained  >= 2 s of back-to-back resident passes with clocks and power sampled (the burst `value` is a ~7 ms region).
  configs    (N=1) the other BASELINE.json configs, device-resident, each after a bit-exact sample check:
             c0 5000x256x256, c3 batch-size sweep on 50k x 256x256, c4 8192x8192 frames.
  a2_split   (N>1) Approach 2: configs[2] 5000 x 256x256 cut into N row bands, halo rows read over NVLink inside the
             stencil kernel, every rank's band checked against the oracle AND against the whole-image kernel;
             a2_large = the same on a sample of configs[4] (8192x8192 frames).
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
METRIC = ("images/sec (3x3 Gaussian blur stream; value = device-resident with the batches of 35 fused into one launch, "
          "per_batch = one work descriptor per batch, e2e = incl. host<->device copies)")
UNIT = "images/s"
WORKLOAD = "A1 image-level: 5000x 320x240 RGB uint8 per GPU, batch_size=35 (BASELINE.json configs[1]); batches fused for `value`"
DTYPE = "u8 (exact integer arithmetic in packed 16-bit lanes)"


def workload_config(n_gpus: int) -> dict:
    """The `config` object: identical in the product arm and the --impl reference arm (only workload-defining keys)."""
    return {"workload": WORKLOAD, "images_per_gpu": N_IMAGES, "width": WIDTH, "height": HEIGHT, "channels": CHANNELS,
            "batch_size": BATCH, "parallelism": f"image-shard x{n_gpus} (no collective)",
            "l2": "inputs larger than L2 (1.15 GB in + 1.15 GB out per step vs 126 MB L2)"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 5))")
    ap.add_argument("--e2e-phased", type=int, default=-1, help="end-to-end pipeline with one transfer direction per GPU at a time: 1/0 (default 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs / a2_split / sustained (value, e2e, roofline only)")
    ap.add_argument("--sustained-seconds", type=float, default=2.5)
    ap.add_argument("--variant", type=int, default=0, help="kernel variant override (0 = auto)")
    ap.add_argument("--per-batch", action="store_true", help="`value` uses one work descriptor per batch (no coalescing)")
    ap.add_argument("--scheme", choices=["image", "split"], default="image",
                    help="image = the contract line (Approach 1 `value` + the extras above); split = ONLY Approach 2 "
                         "row bands (configs[2]) as its own line")
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
    sample = (f"{n} of the {N_IMAGES} images per step "
              f"({'reference kernel source via oracle/_ref' if kind == 'reference' else 'oracle port'}, {threads} threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 accumulate of u8 (reference kernel)", "data": "synthetic",
        "config": workload_config(args.gpus),
        "detail": {"images_per_step": n, "device": "host CPU", "threads": threads},
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
        sm, power, reasons, mx = [], [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            hi = [v for v in sm if v >= 0.5 * max(sm)]  # samples taken under load
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       sm_mhz_min=min(hi), power_w_max=max(power) if power else None)
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


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy kernel, of measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent (of fallback)"


class _DevView:
    """Exposes a raw device allocation to torch (as plumbing) through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class Env:
    """Per-process plumbing: rank/world, device, barrier and max-reduce over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x: float, op: str = "max") -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                    "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_resident(env, ctx, d_in, d_out, w, h, c, n, batch, coalesce, steps, warmup):
    """W warm-up passes, then exactly K timed passes between barriers, CUDA events on the launching queue, max over ranks.
    -> (ms_per_step, launches_per_step)"""
    for _ in range(warmup):
        ctx.run_resident(d_in, d_out, w, h, c, n, batch, coalesce, stats=False)
    ctx.finish()
    env.barrier()
    l0 = ctx.launch_count
    e0 = ctx.enqueue_marker(0)
    for _ in range(steps):
        ctx.run_resident(d_in, d_out, w, h, c, n, batch, coalesce, stats=False)
    e1 = ctx.enqueue_marker(0)
    ctx.finish()
    env.barrier()
    ms = env.reduce(ctx.elapsed_ms(e0, e1))
    return ms / steps, (ctx.launch_count - l0) / steps


def oracle_sample_check(d_in, d_out, n, k=16):
    """A few images of a timed output against the oracle -> dict (bit_exact, max-abs-diff histogram)."""
    from oracle import oracle
    import numpy as np
    idx = sorted(set(list(range(0, n, max(1, n // k)))[:k] + [n - 1]))
    got = d_out[idx].cpu().numpy()
    want = oracle.c_blur_batch(d_in[idx].cpu().numpy(), integer=True)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    # north_star: report the max-abs-diff histogram (expected: all mass at 0 -- the arithmetic is exact)
    return {"bit_exact": bool((got == want).all()), "max_abs_diff": int(diff.max()),
            "abs_diff_histogram_0_1_2_3plus": [int((diff == 0).sum()), int((diff == 1).sum()), int((diff == 2).sum()),
                                               int((diff >= 3).sum())],
            "sample": f"{len(idx)} of {n} images of the timed output vs the oracle"}


# ------------------------------------------------------------------------------------- host-link roofline (all ranks)
def measure_link(env, ctx, h_in, h_out, d_in, d_out, nbytes, reps=3, chunk=64 << 20, only=None):
    """Every rank moves its pinned input host->device on queue 0 and its output device->host on queue 2 AT THE SAME TIME
    as every other rank, in 64 MB linear copies, no kernel: the ceiling the end-to-end pipeline can reach on this box
    with N GPUs active.  -> best-of-reps (max over ranks per rep) GB/s each way per GPU."""
    best = None
    for _ in range(reps):
        env.barrier()
        e0 = ctx.enqueue_marker(0)
        ctx.enqueue_wait(2, e0)
        for off in range(0, nbytes, chunk):
            m = min(chunk, nbytes - off)
            if only != "d2h":
                ctx.enqueue_write(0, d_in + off, h_in + off, m)
            if only != "h2d":
                ctx.enqueue_read(2, h_out + off, d_out + off, m)
        e2 = ctx.enqueue_marker(2)
        ctx.enqueue_wait(0, e2)
        e1 = ctx.enqueue_marker(0)
        ctx.finish()
        ms = ctx.elapsed_ms(e0, e1)
        ctx._lib.b200blur_event_release(ctx.handle, e2)
        ms = env.reduce(ms)
        best = ms if best is None else min(best, ms)
    return nbytes / (best * 1e-3) / 1e9


# ------------------------------------------------------------------------------------- Approach 2 (row bands, N > 1)
def run_a2(env, ctx, n, h, w, steps, warmup, label, oracle_images=4):
    """Every image cut into WORLD row bands, band k resident on GPU k, halo rows read inside the stencil kernel from the
    neighbour GPU's memory (CUDA IPC handles exchanged once, NVLink peer loads).  Strong scaling: the stream is fixed,
    bands shrink with N.  Parity on EVERY rank: (1) a sample of its band against the oracle's whole-image result,
    (2) its band of ALL images against this GPU's own whole-image kernel output; AND-reduced over ranks."""
    torch = env.torch
    from b200blur.sharding import plan_bands
    c = 3
    P = w * c
    plans = plan_bands(h, env.world)
    me = plans[env.rank]
    g = torch.Generator(device=env.dev).manual_seed(2002)          # same stream on every rank; each keeps only its band
    stream = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device=env.dev, generator=g)
    nbytes = n * me.rows * P
    d_band, d_out = ctx.dev_alloc(nbytes), ctx.dev_alloc(nbytes)
    torch.as_tensor(_DevView(d_band, nbytes), device=env.dev).copy_(stream[:, me.row0:me.row0 + me.rows].reshape(-1))
    torch.cuda.synchronize()
    launch = ctx.launch_rows(d_band, d_out, w, me.rows, c, 0, me.rows, n)
    opened = []
    handles = env.gather(ctx.ipc_export(d_band))
    if me.has_top:
        up = plans[env.rank - 1]
        base = ctx.ipc_open(handles[env.rank - 1])
        opened.append(base)
        launch.halo_top, launch.halo_top_stride = base + (up.rows - 1) * P, up.rows * P
    if me.has_bottom:
        dn = plans[env.rank + 1]
        base = ctx.ipc_open(handles[env.rank + 1])
        opened.append(base)
        launch.halo_bottom, launch.halo_bottom_stride = base, dn.rows * P

    for _ in range(warmup):
        ctx.enqueue_blur(0, launch)
    ctx.finish(0)
    env.barrier()
    l0 = ctx.launch_count
    e0 = ctx.enqueue_marker(0)
    for _ in range(steps):
        ctx.enqueue_blur(0, launch)
    e1 = ctx.enqueue_marker(0)
    ctx.finish(0)
    env.barrier()
    ms = env.reduce(ctx.elapsed_ms(e0, e1)) / steps
    launches = ctx.launch_count - l0
    # the same K passes as K batches of ONE launch (b200blur_enqueue_blur_batches: per-batch descriptors, the tail of one
    # pass overlaps the start of the next instead of a launch ramp and drain each)
    ctx.enqueue_blur_batches(0, [launch] * steps)
    ctx.finish(0)
    env.barrier()
    l1 = ctx.launch_count
    f0 = ctx.enqueue_marker(0)
    ctx.enqueue_blur_batches(0, [launch] * steps)
    f1 = ctx.enqueue_marker(0)
    ctx.finish(0)
    env.barrier()
    ms_batched = env.reduce(ctx.elapsed_ms(f0, f1)) / steps
    launches += ctx.launch_count - l1

    # ---- parity, every rank, its own band (rows next to both halo rows included)
    out = torch.as_tensor(_DevView(d_out, nbytes), device=env.dev).view(n, me.rows, w, c)
    whole = torch.empty_like(stream)
    ctx.run_resident(stream, whole, w, h, c, n, max(1, min(n, 35)), True, stats=False)   # whole-image kernel, this GPU
    ctx.finish()
    same_as_whole = bool(torch.equal(out, whole[:, me.row0:me.row0 + me.rows]))
    from oracle import oracle
    idx = sorted(set([0, n // 3, (2 * n) // 3, n - 1][:max(1, oracle_images)]))
    want = oracle.c_blur_batch(stream[idx].cpu().numpy(), integer=True)
    whole_ok = bool((whole[idx].cpu().numpy() == want).all())
    band_ok = bool((out[idx].cpu().numpy() == want[:, me.row0:me.row0 + me.rows]).all())
    ok = same_as_whole and whole_ok and band_ok
    per_rank = env.gather({"rank": env.rank, "rows": [me.row0, me.rows], "band_vs_oracle_sample": band_ok,
                           "band_vs_whole_image_kernel_all_images": same_as_whole, "whole_vs_oracle_sample": whole_ok})
    all_ok = env.reduce(1.0 if ok else 0.0, "min") == 1.0
    del stream, whole, out
    env.barrier()
    for b in opened:
        ctx.ipc_close(b)
    ctx.dev_free(d_band)
    ctx.dev_free(d_out)
    peak, _ = load_peak()
    max_rows = max(p.rows for p in plans)
    per_gpu = 2.0 * n * max_rows * P / (ms * 1e-3) / 1e9
    halo = sum((int(p.has_top) + int(p.has_bottom)) * n * P for p in plans)
    return {"workload": label, "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "scaling": "strong",
            "bands": [[p.row0, p.rows] for p in plans], "per_gpu_GBps": per_gpu, "frac": per_gpu / peak,
            "frac_of_nominal_8000": per_gpu / 8000.0, "aggregate_GBps": 2.0 * n * h * P / (ms * 1e-3) / 1e9,
            "halo": "peer loads inside the stencil kernel (cp.async.bulk from CUDA-IPC mapped neighbour memory over NVLink)",
            "halo_bytes_per_step": halo, "parity_all_bands": all_ok, "parity_per_rank": per_rank,
            "passes_as_batches_of_one_launch": {"value": n / (ms_batched * 1e-3), "ms_per_step": ms_batched,
                                                "per_gpu_GBps": 2.0 * n * max_rows * P / (ms_batched * 1e-3) / 1e9,
                                                "frac": 2.0 * n * max_rows * P / (ms_batched * 1e-3) / 1e9 / peak,
                                                "api": "b200blur_enqueue_blur_batches (the parity checks above ran on this output)"},
            "oracle_sample_images": len(idx), "gpu_launches": int(launches)}


# ----------------------------------------------------------------------------------------------------- product arm
def run_b200_arm(args) -> None:
    import b200blur

    env = Env()
    torch = env.torch
    rank, world, dev = env.rank, env.world, env.dev
    numa = bind_to_gpu_numa_node(env.local_rank)
    if not os.path.exists(b200blur.lib_path()):
        raise SystemExit("libb200blur.so is not built: the product path has no fallback (run __graft_entry__.build())")
    ctx = b200blur.Context(env.local_rank, 4)
    if args.variant:
        ctx.set_kernel_variant(args.variant)
    peak, peak_note = load_peak()
    total_launches = 0

    # synthetic stream of this rank: images [rank*5000, (rank+1)*5000) of the whole job (weak scaling)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    d_in = torch.randint(0, 256, (N_IMAGES, HEIGHT, WIDTH, CHANNELS), dtype=torch.uint8, device=dev, generator=g)
    d_out = torch.zeros_like(d_in)
    coalesce = not args.per_batch

    # ---- device-resident `value`: W warm-up steps, then exactly K timed steps between barriers
    sampler = ClockSampler(env.local_rank)
    ctx.run_resident(d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, coalesce, stats=False)
    env.barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_per_step, launches_per_step = timed_resident(env, ctx, d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH,
                                                    coalesce, args.steps, args.warmup)
    value = world * N_IMAGES / (ms_per_step * 1e-3)
    total_launches += int(launches_per_step * args.steps)
    parity = oracle_sample_check(d_in, d_out, N_IMAGES, 32) if rank == 0 else None

    # ---- the same pass with the reference's batch granularity: one work descriptor per batch of 35
    per_batch = None
    if coalesce:
        d_out.zero_()
        pb_ms, pb_launches = timed_resident(env, ctx, d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, False,
                                            args.steps, args.warmup)
        total_launches += int(pb_launches * args.steps)
        per_batch = {"value": world * N_IMAGES / (pb_ms * 1e-3), "unit": UNIT, "ms_per_step": pb_ms,
                     "kernel_launches_per_step": pb_launches, "batches_per_step": -(-N_IMAGES // BATCH),
                     "frac_of_coalesced": ms_per_step / pb_ms,
                     "GBps": ALGO_BYTES_PER_IMAGE * N_IMAGES / (pb_ms * 1e-3) / 1e9,
                     "parity_vs_oracle": oracle_sample_check(d_in, d_out, N_IMAGES, 16) if rank == 0 else None}

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

    # ---- end to end through the C-ABI stream engine with pinned host buffers
    # --e2e-phased 1 runs the pipeline with one transfer direction per GPU at a time (B200BLUR_E2E_PHASED, see
    # b200blur_run_host).  Off by default: with unsynchronised ranks it measured SLOWER on this pool's 8-GPU box (275.8 k
    # vs 287.6 k images/s at N=8, 199.0 k vs 222.7 k at N=4; profiles/r02_e2e_phased.md) although the fabric itself carries
    # 16 % more when half the GPUs only upload and half only download (e2e.link_roofline.split_directions).
    phased = bool(args.e2e_phased) if args.e2e_phased >= 0 else False
    os.environ["B200BLUR_E2E_PHASED"] = "1" if phased else "0"
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
    h_in = torch.empty((N_IMAGES, HEIGHT, WIDTH, CHANNELS), dtype=torch.uint8).pin_memory()
    h_in.copy_(d_in)
    h_out = torch.empty_like(h_in).pin_memory()
    d_check = d_out.clone() if rank == 0 else None
    for _ in range(min(args.warmup, 2)):
        ctx.run_host(h_in, h_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, stats=False)
    env.barrier()
    launches1 = ctx.launch_count
    e0 = ctx.enqueue_marker(0)
    last = None
    for _ in range(e2e_steps):
        last = ctx.run_host(h_in, h_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, stats=True)
    e1 = ctx.enqueue_marker(2)
    ctx.finish()
    env.barrier()
    e2e_ms = env.reduce(ctx.elapsed_ms(e0, e1))
    e2e_launches = ctx.launch_count - launches1
    total_launches += int(e2e_launches)
    e2e_value = world * N_IMAGES * e2e_steps / (e2e_ms * 1e-3)
    e2e_gbps = N_IMAGES * IMAGE_BYTES * e2e_steps / (e2e_ms * 1e-3) / 1e9
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and isinstance(parity, dict):
        parity["e2e_equals_resident"] = bool((h_out[:64].numpy() == d_check[:64].cpu().numpy()).all())
    # per-rank e2e times show whether the GPUs get even shares of the shared host fabric
    my_e2e = env.gather(round(N_IMAGES * e2e_steps / max(last.wall_ms, 1e-9) * 1e3)) if world > 1 else None

    # host-link roofline for the e2e number at THIS N: all ranks, both directions at once, same pinned buffers, no kernel
    link_gbps = measure_link(env, ctx, h_in.data_ptr(), h_out.data_ptr(), d_in.data_ptr(), d_out.data_ptr(),
                             N_IMAGES * IMAGE_BYTES)
    # ... and with half the GPUs uploading while the other half download (what the phased pipeline approaches)
    link_split = None
    if world >= 2:
        link_split = measure_link(env, ctx, h_in.data_ptr(), h_out.data_ptr(), d_in.data_ptr(), d_out.data_ptr(),
                                  N_IMAGES * IMAGE_BYTES, only="h2d" if rank % 2 == 0 else "d2h")
    del h_in, h_out, d_check

    # ---- sustained: >= ~2 s of back-to-back passes, clocks and power sampled during the region
    sustained = None
    if not args.no_extras and args.sustained_seconds > 0:
        s2 = ClockSampler(env.local_rank)
        n_pass = int(max(args.steps, args.sustained_seconds * 1e3 / ms_per_step))
        env.barrier()
        if rank == 0:
            s2.start()
        s_ms, _ = timed_resident(env, ctx, d_in, d_out, WIDTH, HEIGHT, CHANNELS, N_IMAGES, BATCH, True, n_pass, 1)
        total_launches += n_pass
        sustained = {"value": world * N_IMAGES / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms, "steps": n_pass,
                     "seconds": s_ms * n_pass * 1e-3, "GBps_per_gpu": ALGO_BYTES_PER_IMAGE * N_IMAGES / (s_ms * 1e-3) / 1e9,
                     "vs_burst": ms_per_step / s_ms, "clocks": s2.stop() if rank == 0 else None}
    del d_in, d_out
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs, device-resident (single GPU line only)
    configs = None
    if world == 1 and not args.no_extras:
        try:
            configs = run_other_configs(env, ctx, peak)
            total_launches += configs.pop("_launches")
        except Exception as e:  # the headline numbers above stand on their own; say what happened instead of losing the line
            configs = {"error": f"{type(e).__name__}: {e}"}

    # ---- Approach 2 over NVLink (multi-GPU lines only): configs[2], then a sample of configs[4]
    a2_split = a2_large = None
    if world > 1 and not args.no_extras:
        try:
            a2_split = run_a2(env, ctx, 5000, 256, 256, max(args.steps, 50), max(args.warmup, 5),   # (a step is only 40-160 us)
                              "A2 split-image: 5000x 256x256 RGB, row bands + 1-row halo over NVLink (BASELINE.json configs[2])")
            torch.cuda.empty_cache()
            a2_large = run_a2(env, ctx, 32, 8192, 8192, max(3, args.steps // 2), 2,
                              "A2 large frames: 32 of the 1000x 8192x8192 RGB frames, row bands over NVLink (BASELINE.json configs[4])",
                              oracle_images=1)
            total_launches += a2_split["gpu_launches"] + a2_large["gpu_launches"]
        except Exception as e:  # e.g. no peer access between the GPUs of this box: the same on every rank, so no rank is left
            err = {"error": f"{type(e).__name__}: {e}"}      # waiting in a barrier; the A1 line above is still printed
            a2_split, a2_large = a2_split or err, a2_large or err

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = time_cpu_baseline()
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        images_per_launch = N_IMAGES / launches_per_step
        launch_ms = ms_per_step / launches_per_step
        achieved = ALGO_BYTES_PER_IMAGE * images_per_launch / (launch_ms * 1e-3) / 1e9
        traffic, traffic_source = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_source = tj.get("dram_bytes_per_launch"), "static ncu capture: " + tj.get("source", "profiles/traffic.json")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": workload_config(world),
            "detail": {"resident_launches_per_step": launches_per_step,
                       "resident_mode": "batches fused into one launch" if coalesce else "one work descriptor per batch",
                       "kernel_variant": args.variant, "parity_vs_oracle": parity},
            "per_batch": per_batch,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_source,
                         "kernel": "blur_stream_kernel<3,8,4> (TMA-bulk streamed stencil)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_IMAGE * images_per_launch,
                         "launch_ms": launch_ms, "peak_source": peak_note, "frac_of_nominal_8000": achieved / 8000.0,
                         "same_box_d2d_copy_GBps": copy_gbps},
            "sustained": sustained,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_IMAGES * IMAGE_BYTES,
                    "d2h_bytes_per_step": N_IMAGES * IMAGE_BYTES, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "host_link_GBps_each_way": e2e_gbps,
                    "stage_ms_last_step": {"h2d": last.h2d_ms, "kernel": last.kernel_ms, "d2h": last.d2h_ms},
                    "host_affinity": numa, "per_rank_images_per_s_last_step": my_e2e,
                    "link_roofline": {
                        "bound": f"host link with all {world} GPU(s) copying both directions at once (same pinned buffers, 64 MB linear copies, no kernel; max over ranks)",
                        "peak_GBps_each_way_per_gpu": link_gbps, "aggregate_GBps_each_way": link_gbps * world,
                        "frac": e2e_gbps / link_gbps,
                        "split_directions": None if link_split is None else {
                            "bound": "even ranks upload only, odd ranks download only (one direction per GPU)",
                            "aggregate_GBps_each_way": link_split * (world // 2 if world > 1 else 1),
                            "note": "GB/s of a direction = bytes moved by the ranks of that direction / time of the slowest rank"}},
                    "pipeline_mode": "one direction per GPU at a time (B200BLUR_E2E_PHASED=1)" if phased else "uploads and downloads overlap on every GPU",
                    "api": "b200blur_run_host (pinned host buffers, 3 queues, 4-slot device ring, batches fused into ~64 MB transfer chunks)"},
            "configs": configs, "a2_split": a2_split, "a2_large": a2_large,
            "gpu_launches": int(total_launches),
            "clocks": clocks,
            "published_reference_context": {"a1_best_images_per_s": 8568, "hardware": "i7-12700 + UHD 770",
                                            "source": "BASELINE.md section 1 (end-to-end wall clock)"},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    env.close()


def run_other_configs(env, ctx, peak):
    """configs[0], [3], [4] of BASELINE.json, device-resident on one GPU; every entry after a bit-exact sample check."""
    torch = env.torch
    out, launches = {}, 0

    def entry(n, h, w, ms, ok, extra=None):
        gb = 2.0 * n * h * w * 3 / (ms * 1e-3) / 1e9
        e = {"images": n, "shape": [h, w, 3], "ms": ms, "images_per_s": n / (ms * 1e-3), "GBps": gb, "frac": gb / peak,
             "frac_of_nominal_8000": gb / 8000.0, "bit_exact_sample": ok}
        if extra:
            e.update(extra)
        return e

    def stream(n, h, w, seed):
        g = torch.Generator(device=env.dev).manual_seed(seed)
        x = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=env.dev, generator=g)
        return x, torch.zeros_like(x)

    # configs[0]: 5000 x 256x256, one device (north_star's ">= 80 % of HBM peak for batched 256x256")
    x, y = stream(5000, 256, 256, 11)
    ms, lp = timed_resident(env, ctx, x, y, 256, 256, 3, 5000, 35, True, 10, 3)
    launches += int(lp * 10)
    out["c0_5000x256x256"] = entry(5000, 256, 256, ms, oracle_sample_check(x, y, 5000, 8)["bit_exact"])
    del x, y
    torch.cuda.empty_cache()

    # configs[3]: batch-size sweep on 50,000 x 256x256, fused vs one work descriptor per batch
    n = 50000
    x, y = stream(n, 256, 256, 12)
    sweep = {}
    for batch in (1, 35, 1200):
        ms_c, lp = timed_resident(env, ctx, x, y, 256, 256, 3, n, batch, True, 3, 2)
        launches += int(lp * 3)
        ok_c = oracle_sample_check(x, y, n, 6)["bit_exact"]
        y.zero_()
        ms_p, lpp = timed_resident(env, ctx, x, y, 256, 256, 3, n, batch, False, 3, 2)
        launches += int(lpp * 3)
        ok_p = oracle_sample_check(x, y, n, 6)["bit_exact"]
        sweep[f"batch_{batch}"] = {"coalesced": entry(n, 256, 256, ms_c, ok_c),
                                   "per_batch": entry(n, 256, 256, ms_p, ok_p, {"kernel_launches_per_step": lpp,
                                                                                "batches_per_step": -(-n // batch),
                                                                                "frac_of_coalesced": ms_c / ms_p})}
    out["c3_sweep_50000x256x256"] = sweep
    del x, y
    torch.cuda.empty_cache()

    # configs[4] on one GPU: whole 8192x8192 frames
    x, y = stream(4, 8192, 8192, 13)
    ms, lp = timed_resident(env, ctx, x, y, 8192, 8192, 3, 4, 4, True, 10, 3)
    launches += int(lp * 10)
    from oracle import oracle
    ok = bool((y[1:2].cpu().numpy() == oracle.c_blur_batch(x[1:2].cpu().numpy(), integer=True)).all())
    out["c4_4x8192x8192"] = entry(4, 8192, 8192, ms, ok)
    del x, y
    torch.cuda.empty_cache()
    out["_launches"] = launches
    return out


def run_split_only(args) -> None:
    """--scheme split: Approach 2 alone as its own line (the contract line embeds the same measurement as a2_split)."""
    import b200blur
    env = Env()
    ctx = b200blur.Context(env.local_rank, 4)
    r = run_a2(env, ctx, 5000, int(os.environ.get("A2_HEIGHT", "256")), 256, args.steps, args.warmup,
               "A2 split-image: 5000x 256x256 RGB, row bands + 1-row halo over NVLink (BASELINE.json configs[2])")
    if env.rank == 0:
        peak, note = load_peak()
        print(json.dumps({
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": {"workload": r["workload"], "bands": r["bands"]},
            "roofline": {"bound": "hbm", "achieved": r["per_gpu_GBps"], "peak": peak, "unit": "GB/s", "frac": r["frac"],
                         "traffic": None, "kernel": "blur_stream_kernel<3,8,4>", "per": "GPU (largest band)", "peak_source": note},
            "a2_split": r, "cpu_baseline": None, "e2e": None, "gpu_launches": r["gpu_launches"]}), flush=True)
    ctx.close()
    env.close()


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.scheme == "split":
        run_split_only(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
