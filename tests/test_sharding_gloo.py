"""world_size-2 (and 4) CPU tests of the multi-GPU host logic over torch.distributed/gloo: the sharding plans that
bench.py and the CLIs use on N GPUs must reassemble to the oracle's whole-stream result.  The band kernel is played
by the oracle here (CPU, test-only); on GPUs the same plans drive libb200blur (tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, h, w):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from b200blur.sharding import image_shard, plan_bands
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = 3
        stream = np.random.default_rng(42).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)  # same on every rank
        want = oracle.c_blur_batch(stream)

        # ---- Approach 1: whole-image shards, no data-path collective; gather only to check
        b, cnt = image_shard(n, world, rank)
        mine = oracle.c_blur_batch(stream[b:b + cnt]) if cnt else np.zeros((0, h, w, c), np.uint8)
        padded = np.zeros((n, h, w, c), np.uint8)
        padded[b:b + cnt] = mine
        t = torch.from_numpy(padded)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)  # shards are disjoint, so the sum is the concatenation
        assert np.array_equal(t.numpy(), want), "Approach 1 shards do not reassemble"

        # ---- Approach 2: row bands; each rank owns its rows only and gets one halo row from each neighbour
        plans = plan_bands(h, world)
        me = next((p for p in plans if p.band == rank), None)
        out = np.zeros((n, h, w, c), np.uint8)
        if me is not None:
            band = np.ascontiguousarray(stream[:, me.row0:me.row0 + me.rows])  # what this GPU holds
            idx = plans.index(me)
            top = torch.zeros((n, w, c), dtype=torch.uint8)
            bot = torch.zeros((n, w, c), dtype=torch.uint8)
            reqs = []
            if me.has_top:      # my first row goes up, their last row comes down
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(band[:, 0])), plans[idx - 1].band))
                reqs.append(dist.irecv(top, plans[idx - 1].band))
            if me.has_bottom:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(band[:, -1])), plans[idx + 1].band))
                reqs.append(dist.irecv(bot, plans[idx + 1].band))
            for r in reqs:
                r.wait()
            parts = ([top.numpy()[:, None]] if me.has_top else []) + [band] + ([bot.numpy()[:, None]] if me.has_bottom else [])
            with_halo = np.ascontiguousarray(np.concatenate(parts, axis=1))
            assert with_halo.shape[1] == me.input_rows
            # the reference's composition: kernel on rows incl. halo, halo outputs dropped (split_image_blur.c:401-541)
            blurred = oracle.c_blur_batch(with_halo)
            lo = 1 if me.has_top else 0
            out[:, me.row0:me.row0 + me.rows] = blurred[:, lo:lo + me.rows]
        t = torch.from_numpy(out)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        assert np.array_equal(t.numpy(), want), "Approach 2 bands do not reassemble"
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,h,w", [(2, 7, 24, 16), (2, 5, 9, 21), (4, 6, 32, 16), (4, 3, 3, 8)])
def test_shards_and_bands_reassemble_over_gloo(world, n, h, w):
    from oracle import oracle
    oracle.build()
    mp.spawn(_worker, args=(world, _free_port(), n, h, w), nprocs=world, join=True)
