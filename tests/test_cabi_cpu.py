"""CPU-side tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200blur.h
declares, and its host logic (partition, ratio arithmetic, launch geometry, error behaviour) is right.  No compute
call is made here -- compute needs a GPU and has no fallback (tests/test_gpu_parity.py)."""
import ctypes
import json
import os
import re

import pytest

import b200blur
from b200blur import lib as L
from b200blur.sharding import band_rows, even_split, image_shard, plan_bands
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200blur.build()


def _header_symbols():
    with open(os.path.join(ROOT, "include", "b200blur.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"B200BLUR_API[^;(]*?\b(b200blur_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 35
    lib = ctypes.CDLL(b200blur.lib_path())
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200blur.h but not exported"
    assert set(syms) == set(b200blur.DECLARED_SYMBOLS), "python binding and header disagree"


def test_library_is_torch_free_and_opencl_free():
    import subprocess
    needed = subprocess.run(["readelf", "-d", b200blur.lib_path()], capture_output=True, text=True).stdout
    libs = re.findall(r"NEEDED.*\[(.*?)\]", needed)
    for bad in ("torch", "OpenCL", "c10", "python"):
        assert not any(bad in x for x in libs), libs


def test_version_and_error_string():
    assert b200blur.version().startswith("b200blur ")
    lib = b200blur.load()
    assert lib.b200blur_partition(-1, 2, 0, None, None) == L.ERR_INVALID
    assert b"bad partition" in lib.b200blur_last_error()


def test_run_host_multi_validates_its_context_list():
    lib = b200blur.load()
    assert lib.b200blur_run_host_multi(None, 2, None, None, 16, 16, 3, 0, 1, None) == L.ERR_INVALID
    arr = (ctypes.c_void_p * 1)(None)
    assert lib.b200blur_run_host_multi(arr, 0, None, None, 16, 16, 3, 0, 1, None) == L.ERR_INVALID
    assert lib.b200blur_run_host_multi(arr, 1, None, None, 16, 16, 3, 0, 1, None) == L.ERR_INVALID   # NULL context
    assert b"context" in lib.b200blur_last_error()


def test_no_device_fails_loudly_not_silently():
    if b200blur.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(b200blur.BlurError) as e:
        b200blur.Context(0)
    assert e.value.code == L.ERR_NO_DEVICE


@pytest.mark.parametrize("n,g", [(5000, 1), (5000, 2), (5000, 8), (35, 8), (30, 4), (3, 8), (0, 4), (50000, 7)])
def test_partition_is_even_contiguous_and_complete(n, g):
    pos = 0
    counts = []
    for k in range(g):
        b, c = b200blur.partition(n, g, k)
        assert (b, c) == even_split(n, g, k) == image_shard(n, g, k)
        assert b == pos
        pos += c
        counts.append(c)
    assert pos == n and max(counts) - min(counts) <= 1
    assert counts == sorted(counts, reverse=True)  # remainder goes to the lowest ranks


def test_ratio_arithmetic_matches_reference_logs(golden_dir):
    with open(os.path.join(golden_dir, "distribution.json")) as f:
        dist = json.load(f)
    for k in dist["a1"]:
        n, b = k["num_images"], k["batch_size"]
        tc = tg = 0
        nb = (n + b - 1) // b
        for i in range(nb):
            cnt = min(b, n - i * b)
            c, g = b200blur.ratio_split_images(cnt, k["gpu_ratio"], k["mode"])
            assert (c, g) == oracle.a1_batch_split(cnt, k["gpu_ratio"], k["mode"])
            tc += c
            tg += g
        assert (nb, tc, tg) == (k["num_batches"], k["total_cpu"], k["total_gpu"]), k["source"]
    for k in dist["a2"]:
        assert b200blur.ratio_split_row(k["height"], k["gpu_ratio"]) == k["split_row"], k["source"]
    for ratio in (0.0, 1.0, 0.999, 0.001):
        assert b200blur.ratio_split_row(240, ratio) == oracle.a2_geometry(240, ratio)["split_row"]
    assert b200blur.ratio_split_images(35, 0.5, 1) == (35, 0)
    assert b200blur.ratio_split_images(35, 0.5, 2) == (0, 35)


def test_launch_rows_geometry_whole_image_and_split_parts():
    W, H, C = 320, 240, 3
    P = W * C
    base_in, base_out = 0x10000000, 0x20000000
    l = b200blur.Context.launch_rows(base_in, base_out, W, H, C, 0, H, 35)
    assert (l.in_, l.out, l.rows, l.n_images) == (base_in, base_out, H, 35)
    assert l.halo_top is None and l.halo_bottom is None
    assert l.in_image_stride == P * H and l.out_image_stride == P * H
    assert b200blur.Context.is_vectorised(l)
    # Approach 2, split_image_blur.c:401/:414/:511-517/:537: top part keeps rows [0,split), bottom part skips its halo row
    split = 39
    top = b200blur.Context.launch_rows(base_in, base_out, W, split + 1, C, 0, split, 5, P * H, P * H)
    assert top.in_ == base_in and top.halo_top is None and top.halo_bottom == base_in + split * P
    assert top.halo_bottom_stride == P * H and top.rows == split
    gpu_in = base_in + (split - 1) * P
    bot = b200blur.Context.launch_rows(gpu_in, base_out + split * P, W, H - split + 1, C, 1, H - split, 5, P * H, P * H)
    assert bot.in_ == base_in + split * P and bot.halo_top == gpu_in and bot.halo_bottom is None
    assert bot.rows == H - split
    with pytest.raises(b200blur.BlurError):
        b200blur.Context.launch_rows(base_in, base_out, W, 10, C, 5, 6, 1)  # rows run past the buffer


def test_vectorised_path_selection():
    mk = b200blur.Context.launch_rows
    assert b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 256, 256, 3, 0, 256, 4))
    assert b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 16, 4, 1, 0, 4, 1))
    assert not b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 17, 33, 3, 0, 33, 1))      # pitch % 16 != 0
    # round 2: tight rows of 256..4096 bytes run on the streamed kernel whatever their length and alignment, on either
    # side (aligned-superset copies re-aligned in shared memory; store warps); short unaligned rows still do not
    assert b200blur.Context.is_vectorised(mk(0x1004, 0x9000, 256, 256, 3, 0, 256, 1))       # unaligned input pointer
    assert b200blur.Context.is_vectorised(mk(0x1000, 0x9004, 256, 256, 3, 0, 256, 1))       # unaligned output pointer
    assert b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 250, 37, 3, 0, 37, 2))         # odd width, tight both sides
    assert b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 250, 37, 3, 0, 37, 2, 250 * 37 * 3, 752 * 37, 0, 752))
    assert not b200blur.Context.is_vectorised(mk(0x1004, 0x9000, 16, 4, 1, 0, 4, 1))        # 16-byte rows, unaligned
    assert not b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 1366, 8, 3, 0, 8, 1))      # 4098-byte odd rows: too wide
    assert not b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 16, 4, 5, 0, 4, 1))        # channels > 4
    assert not b200blur.Context.is_vectorised(mk(0x1000, 0x9000, 16, 4, 3, 0, 4, 2, 16 * 4 * 3 + 4, 192))


@pytest.mark.parametrize("h,g", [(256, 2), (256, 4), (256, 8), (240, 8), (8192, 8), (5, 8), (1, 2)])
def test_band_plan_tiles_the_image_with_one_row_halos(h, g):
    plans = plan_bands(h, g)
    assert sum(p.rows for p in plans) == h
    pos = 0
    for i, p in enumerate(plans):
        assert p.row0 == pos and p.rows >= 1
        assert (p.row0, p.rows) == band_rows(h, g, p.band)
        pos += p.rows
        assert p.has_top == (i > 0) and p.has_bottom == (i < len(plans) - 1)
        assert p.input_rows == p.rows + p.has_top + p.has_bottom
    if h >= g:
        assert len(plans) == g and max(p.rows for p in plans) - min(p.rows for p in plans) <= 1


def test_plain_c_host_compiles_links_and_runs(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (the reference's hosts are C) includes the header with -pedantic
    -Werror, links the library and makes the reference's host-side calls -- known answers from its logs and sources."""
    import subprocess
    from conftest import build_c_client
    out = subprocess.run([build_c_client(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "FAIL" not in out.stdout and out.stdout.rstrip().endswith("done")
    assert "split 10 25" in out.stdout          # heterogeneous_blur.c:449-451 at 0.728: 25 for the GPU, 10 for the CPU
    assert "ok split row 39" in out.stdout      # split_image_blur.c:144 at 0.837
    if b200blur.device_count() == 0:
        assert "devices 0" in out.stdout and "checksum" not in out.stdout
