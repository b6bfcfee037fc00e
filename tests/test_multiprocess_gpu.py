"""GPU test (needs >= 2 GPUs) of the one-process-per-GPU form of Approach 2: each rank owns a row band of every image in
its own GPU's memory, exports it with b200blur_ipc_export, opens its neighbours' bands with b200blur_ipc_open and
launches the stencil with halo pointers into the neighbours' memory (NVLink peer loads inside the kernel).  The
reassembled result must equal the oracle's whole-image blur."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, h, w, out_path):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import b200blur
    from b200blur.sharding import plan_bands
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = 3
    P = w * c
    stream = np.random.default_rng(77).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    plans = plan_bands(h, world)
    me = plans[rank]
    band = np.ascontiguousarray(stream[:, me.row0:me.row0 + me.rows])
    ctx = b200blur.Context(rank, 3)
    d_band, d_out = ctx.dev_alloc(band.nbytes), ctx.dev_alloc(band.nbytes)
    ctx.enqueue_write(0, d_band, band, band.nbytes)
    ctx.finish()
    handles = [None] * world
    dist.all_gather_object(handles, ctx.ipc_export(d_band))   # also orders every rank's upload before any kernel
    launch = ctx.launch_rows(d_band, d_out, w, me.rows, c, 0, me.rows, n)
    opened = []
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
    ctx.enqueue_blur(1, launch)
    got = np.zeros_like(band)
    ctx.finish(1)
    ctx.enqueue_read(2, got, d_out, band.nbytes)
    ctx.finish()
    dist.barrier()          # neighbours have finished reading this rank's band
    for b in opened:
        ctx.ipc_close(b)
    full = np.zeros((n, h, w, c), np.uint8)
    full[:, me.row0:me.row0 + me.rows] = got
    t = torch.from_numpy(full)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    ctx.dev_free(d_band)
    ctx.dev_free(d_out)
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,h,w", [(6, 64, 128), (3, 37, 320)])
def test_row_bands_across_processes_with_ipc_halo_pointers(tmp_path, n, h, w):
    import torch
    import torch.multiprocessing as mp
    from oracle import oracle
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    out_path = str(tmp_path / "bands.npy")
    mp.spawn(_worker, args=(world, _free_port(), n, h, w, out_path), nprocs=world, join=True)
    stream = np.random.default_rng(77).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    assert np.array_equal(np.load(out_path), oracle.c_blur_batch(stream))
