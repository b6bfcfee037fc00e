"""CPU tests of the streamed kernel's host-side work plan (b200blur_plan_groups: no GPU needed).

The producer warp of `blur_stream_kernel` decodes a group index into (image block, row segment, column block) from the
numbers this plan holds (csrc/blur_kernels.cuh decode_group / decode_rows_cols); this file restates that decode in Python
and checks, over many geometries, that the groups tile every output byte of every image EXACTLY once -- coarse groups
first, the finer "guided tail" groups after them -- and that the planner's promises hold (whole ring slots, shared-memory
budget, lanes per CTA)."""
import numpy as np
import pytest

import b200blur

RB = 8  # rows per ring slot of the default kernel configuration


def groups(plan, rows, n_images):
    """-> iterator of (img0, n_img, r0, nr, chunk0, n_chunks) in hand-out order, like the device-side decode."""
    for g in range(plan["n_groups"]):
        if g < plan["g_coarse"]:
            per_block, seg, ib0, gg = plan["nseg"] * plan["ncb"], plan["seg"], 0, g
        else:
            per_block, seg, ib0, gg = plan["nseg_fine"] * plan["ncb"], plan["seg_fine"], plan["ib_coarse"], g - plan["g_coarse"]
        ib, sc = divmod(gg, per_block)
        ib += ib0
        si, ci = divmod(sc, plan["ncb"])
        img0 = ib * plan["ipc"]
        r0 = si * seg
        yield (img0, min(plan["ipc"], n_images - img0), r0, min(seg, rows - r0), ci * plan["cb"],
               min(plan["cb"], plan["cpr"] - ci * plan["cb"]))


GEOMETRIES = [  # width, rows, channels, n_images, row_pitch, resident CTAs
    (320, 240, 3, 5000, 0, 444), (256, 256, 3, 5000, 0, 444), (256, 128, 3, 5000, 0, 444), (256, 32, 3, 5000, 0, 444),
    (320, 240, 3, 35, 0, 444), (320, 240, 3, 1, 0, 444), (256, 256, 3, 3, 0, 444), (8192, 8192, 3, 4, 0, 444),
    (8192, 1024, 3, 16, 0, 444), (1366, 77, 3, 9, 4112, 444), (86, 21, 3, 7, 1024, 444), (100, 9, 3, 300, 4096, 444),
    (250, 250, 3, 64, 752, 444), (512, 5, 1, 1000, 0, 444), (128, 300, 2, 77, 0, 296), (64, 64, 4, 2001, 0, 148),
    (1400, 6, 3, 2, 16384, 444), (320, 240, 3, 20000, 0, 444),
]


@pytest.mark.parametrize("w,rows,c,n,pitch,slots", GEOMETRIES)
def test_groups_tile_every_output_exactly_once(w, rows, c, n, pitch, slots):
    plan = b200blur.plan_groups(w, rows, c, n, pitch, slots)
    assert plan["cpr"] == (w * c + 15) // 16 and plan["ncb"] * plan["cb"] >= plan["cpr"] > (plan["ncb"] - 1) * plan["cb"]
    assert plan["img_blocks"] == -(-n // plan["ipc"]) and 0 <= plan["ib_coarse"] <= plan["img_blocks"]
    assert plan["g_coarse"] == plan["ib_coarse"] * plan["nseg"] * plan["ncb"]
    assert plan["n_groups"] == plan["g_coarse"] + (plan["img_blocks"] - plan["ib_coarse"]) * plan["nseg_fine"] * plan["ncb"]
    assert plan["nseg"] == -(-rows // plan["seg"]) and plan["nseg_fine"] == -(-rows // plan["seg_fine"])
    assert plan["smem"] <= 220 * 1024 and 64 <= plan["block"] - 32 <= 256
    assert plan["ipc"] * plan["cb"] <= plan["block"] - 32 or plan["ncb"] > 1     # every chunk of a group has a consumer thread
    assert plan["edge_general"] == int((w * c) % 16 != 0)
    # coverage: count how often each (image, row, chunk) is produced -- sampled images when the stream is long
    sample = sorted(set(range(min(n, 6))) | {n - 1, n // 2} | set(range(max(0, n - 2 * plan["ipc"] - 1), n)))
    index = {img: k for k, img in enumerate(sample)}
    seen = np.zeros((len(sample), rows, plan["cpr"]), np.int32)
    total = 0
    for img0, n_img, r0, nr, ch0, nch in groups(plan, rows, n):
        assert n_img >= 1 and nr >= 1 and nch >= 1
        total += n_img * nr * nch
        for img in range(img0, img0 + n_img):
            if img in index:
                seen[index[img], r0:r0 + nr, ch0:ch0 + nch] += 1
    assert total == n * rows * plan["cpr"]
    assert (seen == 1).all()


@pytest.mark.parametrize("w,rows", [(320, 240), (256, 256), (256, 128), (8192, 8192), (640, 480)])
def test_full_height_groups_fill_whole_ring_slots(w, rows):
    """seg + 2 input rows = a whole number of ring slots, so the consumers' unrolled whole-slot path covers them."""
    plan = b200blur.plan_groups(w, rows, 3, 5000 if w < 4096 else 8)
    assert (plan["seg"] + 2) % RB == 0
    # the guided tail exists for long launches and is finer
    assert plan["ib_coarse"] < plan["img_blocks"] and plan["seg_fine"] < plan["seg"]
    fine_groups = plan["n_groups"] - plan["g_coarse"]
    assert 444 <= fine_groups <= 4 * 444 + plan["nseg_fine"] * plan["ncb"]


def test_short_launches_have_no_fine_tail():
    plan = b200blur.plan_groups(320, 240, 3, 35)
    assert plan["ib_coarse"] == plan["img_blocks"] and plan["n_groups"] == plan["g_coarse"]


@pytest.mark.parametrize("batch", [1, 2, 35, 36, 1200])
def test_feed_plan(batch):
    """FEED form: group slots per batch = image blocks x segments x column blocks, no fine tail, and never more image
    lanes than a batch has images (a batch of 1 runs 64-thread consumer groups, not half-empty 128-thread ones)."""
    plan = b200blur.plan_groups(256, 256, 3, batch, feed=True)
    assert plan["ib_coarse"] == plan["img_blocks"] == -(-batch // plan["ipc"])
    assert plan["ipc"] <= batch and plan["n_groups"] == plan["img_blocks"] * plan["nseg"] * plan["ncb"]
    assert plan["block"] == 64 + 32 * -(-plan["ipc"] * plan["cb"] // 32)    # producer + accountant + consumer warps


def test_plan_rejects_what_the_streamed_kernel_cannot_run():
    for bad in [(16, 16, 3, 10, 0), (320, 240, 5, 10, 0), (320, 240, 3, 10, 1000), (320, 240, 3, 0, 0), (250, 10, 3, 4, 0)]:
        with pytest.raises(b200blur.BlurError):
            b200blur.plan_groups(*bad)


def _drain_with_shared_counter(n_chunks, n_slots, n_workers, rng):
    """The control flow of run_host_impl with a shared chunk counter (csrc/b200blur.cu), restated: every worker harvests
    the ring slot it is about to reuse BEFORE taking the next chunk index, stops at the first index past the stream, and
    finally harvests what is still pending.  Returns per-worker (issued, harvested) lists of local chunk numbers."""
    nxt = 0
    state = [dict(taken=0, issued=[], harvested=[], done=False) for _ in range(n_workers)]
    while not all(s["done"] for s in state):
        s = state[int(rng.integers(0, n_workers))]          # any interleaving of the workers' loop iterations
        if s["done"]:
            continue
        if s["taken"] >= n_slots:
            s["harvested"].append(s["taken"] - n_slots)      # harvest(ring[taken % n_slots]) = local chunk taken - n_slots
        ci, nxt = nxt, nxt + 1
        if ci >= n_chunks:
            first_pending = s["taken"] - n_slots if s["taken"] > n_slots else 0
            if s["taken"] >= n_slots:
                first_pending += 1
            s["harvested"].extend(range(first_pending, s["taken"]))
            s["done"] = True
            continue
        s["issued"].append(ci)
        s["taken"] += 1
    return state


@pytest.mark.parametrize("n_chunks,n_slots,n_workers", [(0, 4, 2), (1, 4, 3), (4, 4, 1), (5, 4, 1), (17, 4, 2), (64, 2, 3),
                                                        (9, 4, 8), (100, 4, 8), (33, 3, 5)])
def test_shared_counter_pipeline_moves_and_harvests_every_chunk_once(n_chunks, n_slots, n_workers):
    """b200blur_run_host_multi's host logic: over random interleavings of the workers, every chunk of the stream is issued
    by exactly one worker, and every worker harvests each of its own chunks exactly once (a slot harvested twice would
    add its stage times twice; one never harvested could still be in flight when the call returns)."""
    import numpy as np
    for seed in range(20):
        state = _drain_with_shared_counter(n_chunks, n_slots, n_workers, np.random.default_rng(seed))
        issued = sorted(c for s in state for c in s["issued"])
        assert issued == list(range(n_chunks))
        for s in state:
            assert sorted(s["harvested"]) == list(range(s["taken"])), (seed, s)
            assert len(s["harvested"]) == len(set(s["harvested"]))
