"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes/numpy front end of the parity checker.  Three independent statements of the reference path:

* ``ref_*``     -- the reference's own ``gaussian_kernel.cl`` compiled unmodified (``oracle/_ref``; present only
                   where ``oracle/Makefile`` found the reference tree, or where the prebuilt ``.so`` travelled).
* ``c_*``       -- the C restatement ``oracle/gaussian_oracle.c`` (always buildable; cites reference file:line).
* ``np_*``      -- a vectorised numpy integer form (``(sum w_int * p) >> 4`` with edge replication).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs import this
module.  The product (``b200blur``) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_long, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libgaussian_ref.so")


def build(quiet: bool = True) -> None:
    """Build liboracle.so (always) and _ref/libgaussian_ref.so (only when /root/reference is present)."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


_c = None
_ref = None


def _lib():
    global _c
    if _c is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        lib = ctypes.CDLL(_ORACLE_SO)
        u8p = c_void_p
        lib.oracle_blur_image_f32.argtypes = [u8p, u8p, c_int, c_int, c_int]
        lib.oracle_blur_image_int.argtypes = [u8p, u8p, c_int, c_int, c_int]
        lib.oracle_blur_batch.argtypes = [u8p, u8p, c_int, c_int, c_int, c_long, c_size_t, c_size_t, c_int]
        lib.oracle_num_threads.restype = c_int
        lib.oracle_a1_batch_split.argtypes = [c_int, c_float, c_int, POINTER(c_int), POINTER(c_int)]
        lib.oracle_a1_totals.argtypes = [c_int, c_int, c_float, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]
        lib.oracle_a2_geometry.argtypes = [c_int, c_float] + [POINTER(c_int)] * 5
        lib.oracle_split_image.argtypes = [u8p, u8p, c_int, c_int, c_int, c_int]
        lib.oracle_split_image.restype = c_int
        lib.oracle_band_split.argtypes = [u8p, u8p, c_int, c_int, c_int, c_int]
        lib.oracle_band_split.restype = c_int
        _c = lib
    return _c


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def _reflib():
    global _ref
    if _ref is None:
        if not have_ref():
            raise FileNotFoundError(_REF_SO + " (built only where the reference tree is present)")
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_gaussian_blur_ndrange.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int]
        lib.ref_gaussian_blur_batch.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_long, c_size_t, c_size_t]
        lib.ref_num_threads.restype = c_int
        _ref = lib
    return _ref


def _check_img(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim != 3:
        raise ValueError("image must be [H][W][C] uint8")
    return img


def _check_batch(imgs: np.ndarray) -> np.ndarray:
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    if imgs.ndim != 4:
        raise ValueError("batch must be [N][H][W][C] uint8")
    return imgs


# ----------------------------------------------------------------------------------------------- kernel forms
def ref_blur(img: np.ndarray) -> np.ndarray:
    """One NDRange launch of the reference's own kernel source on one [H][W][C] image."""
    img = _check_img(img)
    h, w, c = img.shape
    out = np.empty_like(img)
    if img.size:
        _reflib().ref_gaussian_blur_ndrange(img.ctypes.data, out.ctypes.data, w, h, c)
    return out


def ref_blur_batch(imgs: np.ndarray) -> np.ndarray:
    imgs = _check_batch(imgs)
    n, h, w, c = imgs.shape
    out = np.empty_like(imgs)
    if imgs.size:
        _reflib().ref_gaussian_blur_batch(imgs.ctypes.data, out.ctypes.data, w, h, c, n, h * w * c, h * w * c)
    return out


def c_blur(img: np.ndarray, integer: bool = False) -> np.ndarray:
    img = _check_img(img)
    h, w, c = img.shape
    out = np.empty_like(img)
    if img.size:
        fn = _lib().oracle_blur_image_int if integer else _lib().oracle_blur_image_f32
        fn(img.ctypes.data, out.ctypes.data, w, h, c)
    return out


def c_blur_batch(imgs: np.ndarray, integer: bool = False) -> np.ndarray:
    imgs = _check_batch(imgs)
    n, h, w, c = imgs.shape
    out = np.empty_like(imgs)
    if imgs.size:
        _lib().oracle_blur_batch(imgs.ctypes.data, out.ctypes.data, w, h, c, n, h * w * c, h * w * c, int(integer))
    return out


def np_blur(img: np.ndarray) -> np.ndarray:
    """Vectorised integer form on [..., H, W, C]: edge-replicate pad, separable [1,2,1] x [1,2,1]^T, >> 4."""
    a = np.asarray(img, dtype=np.uint8)
    if a.size == 0:
        return a.copy()
    p = np.pad(a.astype(np.int32), [(0, 0)] * (a.ndim - 3) + [(1, 1), (1, 1), (0, 0)], mode="edge")
    hsum = p[..., :, :-2, :] + 2 * p[..., :, 1:-1, :] + p[..., :, 2:, :]
    v = hsum[..., :-2, :, :] + 2 * hsum[..., 1:-1, :, :] + hsum[..., 2:, :, :]
    return (v >> 4).astype(np.uint8)


def num_threads() -> int:
    return int(_lib().oracle_num_threads())


def ref_num_threads() -> int:
    return int(_reflib().ref_num_threads())


def use_all_cores() -> int:
    """Make both CPU implementations use every core this process may run on (torchrun sets OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    _lib().oracle_set_num_threads(n)
    if have_ref():
        _reflib().ref_set_num_threads(n)
    return n


# ------------------------------------------------------------------------------------------- distribution forms
def a1_batch_split(batch_count: int, gpu_ratio: float, mode: int = 0):
    """heterogeneous_blur.c:446-458 -> (n_cpu, n_gpu)."""
    a, b = c_int(), c_int()
    _lib().oracle_a1_batch_split(batch_count, gpu_ratio, mode, a, b)
    return a.value, b.value


def a1_totals(num_images: int, batch_size: int, gpu_ratio: float, mode: int = 0):
    """-> (num_batches, total_cpu, total_gpu) over the whole run (heterogeneous_blur.c:86, :418-461)."""
    nb, tc, tg = c_int(), c_int(), c_int()
    _lib().oracle_a1_totals(num_images, batch_size, gpu_ratio, mode, nb, tc, tg)
    return nb.value, tc.value, tg.value


def a2_geometry(height: int, gpu_ratio: float):
    """split_image_blur.c:144-166 -> dict(split_row, cpu_input_rows, cpu_output_rows, gpu_input_rows, gpu_output_rows)."""
    v = [c_int() for _ in range(5)]
    _lib().oracle_a2_geometry(height, gpu_ratio, *v)
    keys = ("split_row", "cpu_input_rows", "cpu_output_rows", "gpu_input_rows", "gpu_output_rows")
    return dict(zip(keys, (x.value for x in v)))


def split_image(img: np.ndarray, split_row: int) -> np.ndarray:
    """Approach 2 on one image (split_image_blur.c:503-541): two kernel launches with halo, halo outputs dropped."""
    img = _check_img(img)
    h, w, c = img.shape
    out = np.empty_like(img)
    rc = _lib().oracle_split_image(img.ctypes.data, out.ctypes.data, w, h, c, split_row)
    if rc != 0:
        raise MemoryError
    return out


def band_split(img: np.ndarray, n_bands: int) -> np.ndarray:
    img = _check_img(img)
    h, w, c = img.shape
    out = np.empty_like(img)
    rc = _lib().oracle_band_split(img.ctypes.data, out.ctypes.data, w, h, c, n_bands)
    if rc != 0:
        raise ValueError("empty band")
    return out
