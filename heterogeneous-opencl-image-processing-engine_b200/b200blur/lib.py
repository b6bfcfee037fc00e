"""ctypes binding of include/b200blur.h.  One Python method per C entry point; no compute happens in Python."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_void_p, create_string_buffer)

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.environ.get("B200BLUR_LIB") or os.path.join(_PKG_DIR, "libb200blur.so")   # (B200BLUR_LIB: A/B builds of the library)

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_PEER = 0, -1, -2, -3, -4, -5
IPC_HANDLE_BYTES = 64


class BlurError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b200blur error {code}: {message}")
        self.code = code


class Launch(Structure):
    """struct b200blur_launch"""
    _fields_ = [("in_", c_void_p), ("out", c_void_p), ("width", c_int32), ("channels", c_int32),
                ("rows", c_int32), ("reserved", c_int32), ("n_images", c_int64),
                ("in_image_stride", c_size_t), ("out_image_stride", c_size_t),
                ("halo_top", c_void_p), ("halo_top_stride", c_size_t),
                ("halo_bottom", c_void_p), ("halo_bottom_stride", c_size_t),
                ("in_row_pitch", c_size_t), ("out_row_pitch", c_size_t)]


class Stats(Structure):
    """struct b200blur_stats"""
    _fields_ = [("wall_ms", c_double), ("h2d_ms", c_double), ("kernel_ms", c_double), ("d2h_ms", c_double),
                ("images", c_int64), ("launches", c_int64), ("h2d_bytes", c_int64), ("d2h_bytes", c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# name -> (restype, argtypes); this table is also what the CPU test checks against include/b200blur.h
_SIGNATURES = {
    "b200blur_last_error": (c_char_p, []),
    "b200blur_version": (c_char_p, []),
    "b200blur_device_count": (c_int, [POINTER(c_int)]),
    "b200blur_device_name": (c_int, [c_int, c_char_p, c_size_t]),
    "b200blur_device_props": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "b200blur_ctx_create": (c_int, [c_int, c_int, POINTER(c_void_p)]),
    "b200blur_ctx_destroy": (c_int, [c_void_p]),
    "b200blur_ctx_device": (c_int, [c_void_p]),
    "b200blur_ctx_num_queues": (c_int, [c_void_p]),
    "b200blur_ctx_queue_handle": (c_void_p, [c_void_p, c_int]),
    "b200blur_dev_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "b200blur_dev_free": (c_int, [c_void_p, c_void_p]),
    "b200blur_host_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "b200blur_host_free": (c_int, [c_void_p]),
    "b200blur_host_register": (c_int, [c_void_p, c_size_t]),
    "b200blur_host_unregister": (c_int, [c_void_p]),
    "b200blur_event_ms": (c_int, [c_void_p, c_int32, POINTER(c_double)]),
    "b200blur_event_release": (c_int, [c_void_p, c_int32]),
    "b200blur_enqueue_marker": (c_int, [c_void_p, c_int, POINTER(c_int32)]),
    "b200blur_events_elapsed_ms": (c_int, [c_void_p, c_int32, c_int32, POINTER(c_double)]),
    "b200blur_enqueue_wait": (c_int, [c_void_p, c_int, c_int32]),
    "b200blur_enqueue_wait_peer": (c_int, [c_void_p, c_int, c_void_p, c_int32]),
    "b200blur_enqueue_write": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, POINTER(c_int32)]),
    "b200blur_enqueue_read": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, POINTER(c_int32)]),
    "b200blur_enqueue_write_2d": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t,
                                          POINTER(c_int32)]),
    "b200blur_enqueue_read_2d": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t,
                                         POINTER(c_int32)]),
    "b200blur_finish": (c_int, [c_void_p, c_int]),
    "b200blur_finish_all": (c_int, [c_void_p]),
    "b200blur_launch_rows": (c_int, [POINTER(Launch), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64,
                                     c_size_t, c_size_t]),
    "b200blur_launch_rows_pitched": (c_int, [POINTER(Launch), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64,
                                             c_size_t, c_size_t, c_size_t, c_size_t]),
    "b200blur_enqueue_blur": (c_int, [c_void_p, c_int, POINTER(Launch), POINTER(c_int32)]),
    "b200blur_enqueue_blur_batches": (c_int, [c_void_p, c_int, POINTER(Launch), c_int, POINTER(c_int32)]),
    "b200blur_launch_is_vectorised": (c_int, [POINTER(Launch)]),
    "b200blur_plan_row_edge": (c_int, [c_int, c_int, POINTER(ctypes.c_uint32)]),
    "b200blur_plan_groups": (c_int, [c_int, c_int, c_int, c_int64, c_size_t, c_int, c_int, POINTER(c_int64)]),
    "b200blur_set_kernel_variant": (c_int, [c_void_p, c_int]),
    "b200blur_ctx_launch_count": (c_int64, [c_void_p]),
    "b200blur_partition": (c_int, [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "b200blur_ratio_split_images": (c_int, [c_int, c_float, c_int, POINTER(c_int), POINTER(c_int)]),
    "b200blur_ratio_split_row": (c_int, [c_int, c_float, POINTER(c_int)]),
    "b200blur_run_resident": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int,
                                      POINTER(Stats)]),
    "b200blur_run_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, POINTER(Stats)]),
    "b200blur_run_host_multi": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int,
                                        POINTER(Stats)]),
    "b200blur_feed_create": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "b200blur_feed_destroy": (c_int, [c_void_p]),
    "b200blur_feed_start": (c_int, [c_void_p]),
    "b200blur_feed_submit": (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(c_int64)]),
    "b200blur_feed_flush": (c_int, [c_void_p]),
    "b200blur_feed_wait": (c_int, [c_void_p, c_int64]),
    "b200blur_feed_completed": (c_int, [c_void_p, c_int64, POINTER(c_int)]),
    "b200blur_feed_stop": (c_int, [c_void_p]),
    "b200blur_feed_submitted": (c_int64, [c_void_p]),
    "b200blur_peer_enable": (c_int, [c_void_p, c_void_p]),
    "b200blur_ipc_export": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b200blur_ipc_open": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "b200blur_ipc_close": (c_int, [c_void_p, c_void_p]),
}
DECLARED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib_path() -> str:
    return _LIB


def build(force: bool = False) -> str:
    """Compile libb200blur.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    args = ["make", "-C", _PKG_DIR, "all"]
    if force:
        args.insert(1, "-B")
    out = subprocess.run(args, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("libb200blur build failed:\n" + out.stdout + out.stderr)
    return _LIB


def load():
    """Load the CUDA library.  No fallback: a missing library is an error, not a reason to compute elsewhere."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise BlurError(ERR_NO_DEVICE, f"{_LIB} is not built (run `make -C {_PKG_DIR}` or __graft_entry__.build())")
        lib = ctypes.CDLL(_LIB)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != OK:
        raise BlurError(rc, load().b200blur_last_error().decode(errors="replace"))


def version() -> str:
    return load().b200blur_version().decode()


def device_count() -> int:
    n = c_int(0)
    rc = load().b200blur_device_count(byref(n))
    if rc == ERR_NO_DEVICE:
        return 0
    _check(rc)
    return n.value


def device_name(device: int) -> str:
    buf = create_string_buffer(256)
    _check(load().b200blur_device_name(device, buf, 256))
    return buf.value.decode()


def partition(n_items: int, n_parts: int, part: int):
    """Even contiguous partition -> (begin, count).  Replaces (int)(batch_count*gpu_ratio), heterogeneous_blur.c:449-451."""
    b, c = c_int64(), c_int64()
    _check(load().b200blur_partition(n_items, n_parts, part, byref(b), byref(c)))
    return b.value, c.value


def ratio_split_images(batch_count: int, gpu_ratio: float, mode: int = 0):
    """The reference's own per-batch split (heterogeneous_blur.c:446-458) -> (n_cpu, n_gpu)."""
    a, b = c_int(), c_int()
    _check(load().b200blur_ratio_split_images(batch_count, gpu_ratio, mode, byref(a), byref(b)))
    return a.value, b.value


def ratio_split_row(height: int, gpu_ratio: float) -> int:
    """The reference's split row (split_image_blur.c:144-154)."""
    s = c_int()
    _check(load().b200blur_ratio_split_row(height, gpu_ratio, byref(s)))
    return s.value


def run_host_multi(ctxs, h_in, h_out, width, height, channels, n_images, batch_size):
    """One host stream over several contexts (GPUs) in one call (b200blur_run_host_multi): the contexts' pipelines take
    transfer chunks from a shared counter.  -> list of Stats, one per context."""
    n = len(ctxs)
    handles = (c_void_p * n)(*[c._h for c in ctxs])
    stats = (Stats * n)()
    _check(load().b200blur_run_host_multi(handles, n, _ptr(h_in), _ptr(h_out), width, height, channels, n_images,
                                          batch_size, stats))
    return list(stats)


def plan_row_edge(row_bytes: int, channels: int):
    """Host-side right-edge plan of the vectorised kernel (b200blur_plan_row_edge) -> dict."""
    out = (ctypes.c_uint32 * 10)()
    _check(load().b200blur_plan_row_edge(row_bytes, channels, out))
    return {"chunks": out[0], "edge_general": bool(out[1]), "edge_prev": bool(out[2]),
            "sel_last": [out[3 + m] for m in range(6)], "sel_prev": out[9]}


def plan_groups(width: int, rows: int, channels: int, n_images: int, row_pitch: int = 0, resident_ctas: int = 444,
                feed: bool = False):
    """Host-side work plan of the streamed kernel (b200blur_plan_groups) -> dict; needs no GPU."""
    out = (c_int64 * 16)()
    _check(load().b200blur_plan_groups(width, rows, channels, n_images, row_pitch, resident_ctas, int(feed), out))
    keys = ("cpr", "cb", "ncb", "ipc", "seg", "nseg", "seg_fine", "nseg_fine", "img_blocks", "ib_coarse", "g_coarse", "n_groups",
            "margin", "block", "smem", "edge_general")
    return dict(zip(keys, (int(v) for v in out)))


def _ptr(p) -> int:
    if p is None:
        return None
    if isinstance(p, int):
        return p
    if hasattr(p, "data_ptr"):      # torch tensor
        return p.data_ptr()
    if hasattr(p, "ctypes"):        # numpy array
        return p.ctypes.data
    raise TypeError(f"cannot take the address of {type(p)}")


class Context:
    """One per GPU: cl_context + profiled in-order queues + the (ahead-of-time compiled) kernel."""

    def __init__(self, device: int = 0, n_queues: int = 4):
        self._lib = load()
        h = c_void_p()
        _check(self._lib.b200blur_ctx_create(device, n_queues, byref(h)))
        self._h = h
        self.device = device

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.b200blur_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def launch_count(self) -> int:
        return int(self._lib.b200blur_ctx_launch_count(self._h))

    def queue_handle(self, queue: int) -> int:
        return int(self._lib.b200blur_ctx_queue_handle(self._h, queue) or 0)

    def set_kernel_variant(self, variant: int) -> int:
        return int(self._lib.b200blur_set_kernel_variant(self._h, variant))

    # -- memory
    def dev_alloc(self, nbytes: int) -> int:
        p = c_void_p()
        _check(self._lib.b200blur_dev_alloc(self._h, nbytes, byref(p)))
        return p.value

    def dev_free(self, dptr: int) -> None:
        _check(self._lib.b200blur_dev_free(self._h, dptr))

    @staticmethod
    def host_alloc(nbytes: int) -> int:
        p = c_void_p()
        _check(load().b200blur_host_alloc(nbytes, byref(p)))
        return p.value

    @staticmethod
    def host_free(hptr: int) -> None:
        _check(load().b200blur_host_free(hptr))

    # -- enqueues
    def enqueue_write(self, queue, dst_dev, src_host, nbytes, want_event=False):
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_write(self._h, queue, _ptr(dst_dev), _ptr(src_host), nbytes,
                                                byref(ev) if want_event else None))
        return ev.value if want_event else None

    def enqueue_read(self, queue, dst_host, src_dev, nbytes, want_event=False):
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_read(self._h, queue, _ptr(dst_host), _ptr(src_dev), nbytes,
                                               byref(ev) if want_event else None))
        return ev.value if want_event else None

    def enqueue_write_2d(self, queue, dst_dev, dst_pitch, src_host, src_pitch, row_bytes, rows, want_event=False):
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_write_2d(self._h, queue, _ptr(dst_dev), dst_pitch, _ptr(src_host), src_pitch,
                                                   row_bytes, rows, byref(ev) if want_event else None))
        return ev.value if want_event else None

    def enqueue_read_2d(self, queue, dst_host, dst_pitch, src_dev, src_pitch, row_bytes, rows, want_event=False):
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_read_2d(self._h, queue, _ptr(dst_host), dst_pitch, _ptr(src_dev), src_pitch,
                                                  row_bytes, rows, byref(ev) if want_event else None))
        return ev.value if want_event else None

    def enqueue_blur(self, queue, launch: Launch, want_event=False):
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_blur(self._h, queue, byref(launch), byref(ev) if want_event else None))
        return ev.value if want_event else None

    def enqueue_blur_batches(self, queue, launches, want_event=False):
        """A list of independent launches of one geometry as ONE kernel launch with per-batch descriptors."""
        arr = (Launch * len(launches))(*launches)
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_blur_batches(self._h, queue, arr, len(launches), byref(ev) if want_event else None))
        return ev.value if want_event else None

    def enqueue_marker(self, queue) -> int:
        ev = c_int32(-1)
        _check(self._lib.b200blur_enqueue_marker(self._h, queue, byref(ev)))
        return ev.value

    def elapsed_ms(self, ev_from, ev_to, release=True) -> float:
        ms = c_double()
        _check(self._lib.b200blur_events_elapsed_ms(self._h, ev_from, ev_to, byref(ms)))
        if release:
            _check(self._lib.b200blur_event_release(self._h, ev_from))
            _check(self._lib.b200blur_event_release(self._h, ev_to))
        return ms.value

    def enqueue_wait(self, queue, ev) -> None:
        _check(self._lib.b200blur_enqueue_wait(self._h, queue, ev))

    def enqueue_wait_peer(self, queue, src: "Context", ev) -> None:
        _check(self._lib.b200blur_enqueue_wait_peer(self._h, queue, src._h, ev))

    def finish(self, queue=None) -> None:
        if queue is None:
            _check(self._lib.b200blur_finish_all(self._h))
        else:
            _check(self._lib.b200blur_finish(self._h, queue))

    def event_ms(self, ev, release=True) -> float:
        ms = c_double()
        _check(self._lib.b200blur_event_ms(self._h, ev, byref(ms)))
        if release:
            _check(self._lib.b200blur_event_release(self._h, ev))
        return ms.value

    # -- launch construction
    @staticmethod
    def launch_rows(in_ptr, out_ptr, width, in_height, channels, first_row, n_rows, n_images,
                    in_image_stride=None, out_image_stride=None, in_row_pitch=0, out_row_pitch=0) -> Launch:
        """The reference kernel with height=in_height on `in`, keeping rows [first_row, first_row+n_rows).
        in_row_pitch / out_row_pitch: bytes between rows (0 = tight)."""
        l = Launch()
        if in_image_stride is None:
            in_image_stride = (in_row_pitch or width * channels) * in_height
        if out_image_stride is None:
            out_image_stride = (out_row_pitch or width * channels) * n_rows
        _check(load().b200blur_launch_rows_pitched(byref(l), _ptr(in_ptr), _ptr(out_ptr), width, in_height, channels,
                                                   first_row, n_rows, n_images, in_image_stride, out_image_stride,
                                                   in_row_pitch, out_row_pitch))
        return l

    @staticmethod
    def is_vectorised(launch: Launch) -> bool:
        return bool(load().b200blur_launch_is_vectorised(byref(launch)))

    # -- stream engines
    def run_resident(self, d_in, d_out, width, height, channels, n_images, batch_size, coalesce=True, stats=True):
        """stats=False enqueues asynchronously on queue 0 and returns None (no host sync)."""
        st = Stats() if stats else None
        _check(self._lib.b200blur_run_resident(self._h, _ptr(d_in), _ptr(d_out), width, height, channels, n_images,
                                               batch_size, int(coalesce), byref(st) if stats else None))
        return st

    def run_host(self, h_in, h_out, width, height, channels, n_images, batch_size, stats=True):
        st = Stats() if stats else None
        _check(self._lib.b200blur_run_host(self._h, _ptr(h_in), _ptr(h_out), width, height, channels, n_images,
                                           batch_size, byref(st) if stats else None))
        return st

    def feed(self, width, height, channels, max_batch_images, capacity=0) -> "Feed":
        """The batch loop as a resident kernel fed with per-batch descriptors (b200blur_feed_*)."""
        return Feed(self, width, height, channels, max_batch_images, capacity)

    # -- multi-GPU
    def peer_enable(self, other: "Context") -> None:
        _check(self._lib.b200blur_peer_enable(self._h, other._h))

    def ipc_export(self, dptr: int) -> bytes:
        buf = create_string_buffer(IPC_HANDLE_BYTES)
        _check(self._lib.b200blur_ipc_export(self._h, dptr, buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = c_void_p()
        buf = create_string_buffer(handle, IPC_HANDLE_BYTES)
        _check(self._lib.b200blur_ipc_open(self._h, buf, byref(p)))
        return p.value

    def ipc_close(self, dptr: int) -> None:
        _check(self._lib.b200blur_ipc_close(self._h, dptr))

    # -- convenience used by tests: blur a host numpy batch [N][H][W][C] through explicit write/blur/read enqueues
    def blur_numpy(self, imgs, queue: int = 0):
        import numpy as np
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
        if imgs.ndim == 3:
            return self.blur_numpy(imgs[None], queue)[0]
        n, h, w, c = imgs.shape
        out = np.empty_like(imgs)
        nbytes = imgs.nbytes
        if nbytes == 0:
            return out
        d_in, d_out = self.dev_alloc(nbytes), self.dev_alloc(nbytes)
        try:
            self.enqueue_write(queue, d_in, imgs, nbytes)
            self.enqueue_blur(queue, self.launch_rows(d_in, d_out, w, h, c, 0, h, n))
            self.enqueue_read(queue, out, d_out, nbytes)
            self.finish(queue)
        finally:
            self.dev_free(d_in)
            self.dev_free(d_out)
        return out


class Feed:
    """b200blur_feed: one persistent kernel per GPU that pulls per-batch descriptors from a ring the host appends to
    (the reference's batch loop, heterogeneous_blur.c:418-539, without a kernel launch per batch)."""

    def __init__(self, ctx: Context, width, height, channels, max_batch_images, capacity=0):
        self._lib = load()
        self._ctx = ctx     # keeps the context alive
        h = c_void_p()
        _check(self._lib.b200blur_feed_create(ctx.handle, width, height, channels, max_batch_images, capacity, byref(h)))
        self._h = h

    def start(self) -> None:
        _check(self._lib.b200blur_feed_start(self._h))

    def submit(self, d_in, d_out, n_images) -> int:
        t = c_int64(-1)
        _check(self._lib.b200blur_feed_submit(self._h, _ptr(d_in), _ptr(d_out), n_images, byref(t)))
        return t.value

    def flush(self) -> None:
        _check(self._lib.b200blur_feed_flush(self._h))

    def wait(self, ticket: int) -> None:
        _check(self._lib.b200blur_feed_wait(self._h, ticket))

    def completed(self, ticket: int) -> bool:
        d = c_int(0)
        _check(self._lib.b200blur_feed_completed(self._h, ticket, byref(d)))
        return bool(d.value)

    def stop(self) -> None:
        _check(self._lib.b200blur_feed_stop(self._h))

    @property
    def submitted(self) -> int:
        return int(self._lib.b200blur_feed_submitted(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.b200blur_feed_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
