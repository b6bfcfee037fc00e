"""CPU emulation of the TIGHT-output store path of the streamed kernel (csrc/blur_kernels.cuh: flush_rows_tight,
store_bytes, lds_unaligned16): the consumers stage row-relative, 16-byte aligned output chunks in shared memory with a
16-byte-multiple pitch; store warps then write the tight (unaligned) global rows as ALIGNED 16-byte words -- single-row
words, words straddling two rows (byte-mask merge of two reads) and the two words that stick out of the span (narrower
aligned pieces).  This file restates that address arithmetic in Python and checks, for many (row length, alignment, row
count) combinations, that exactly the span's bytes are written, each once, with the right values, and that every store
is aligned to its own size."""
import numpy as np
import pytest


def lds_unaligned16(stage, a):
    return stage[a:a + 16].copy()          # five aligned words + funnel shifts == 16 bytes at byte address a


def store_bytes(mem, written, p, v, first, last):
    pos = first
    while pos < last:
        align = (pos & -pos) if pos else 16
        sz = 8
        while sz > align or sz > last - pos:
            sz >>= 1
        assert (p + pos) % sz == 0                       # every piece is aligned to its own size
        mem[p + pos:p + pos + sz] = v[pos:pos + sz]
        written[p + pos:p + pos + sz] += 1
        pos += sz


def flush_rows_tight(mem, written, g, stage, stage_lane, n_rows, row_bytes, spitch):
    head = g & 15
    w0 = g - head
    span = n_rows * row_bytes
    n_words = (head + span + 15) >> 4
    recip = (2 ** 32 + row_bytes - 1) // row_bytes
    for m in range(n_words):                              # (lanes take m, m + stride, ...: order does not matter)
        b0 = 16 * m - head
        bb = max(b0, 0)
        r = (bb * recip) >> 32                            # __umulhi(bb, row_recip)
        assert r == bb // row_bytes
        col = b0 - r * row_bytes
        k = min(row_bytes - col, 16)
        a1 = stage_lane + r * spitch + col
        a2 = stage_lane + (r + 1) * spitch - k if k < 16 else a1
        va, vb = lds_unaligned16(stage, a1), lds_unaligned16(stage, a2)
        v = np.concatenate([va[:k], vb[k:]])              # byte-mask merge
        dst = w0 + 16 * m
        if b0 >= 0 and b0 + 16 <= span:
            assert dst % 16 == 0
            mem[dst:dst + 16] = v
            written[dst:dst + 16] += 1
        else:
            store_bytes(mem, written, dst, v, -b0 if b0 < 0 else 0, span - b0 if span - b0 < 16 else 16)


@pytest.mark.parametrize("row_bytes", [256, 257, 258, 263, 300, 750, 1023, 4095])
@pytest.mark.parametrize("n_rows", [1, 2, 6, 8])
def test_flush_writes_exactly_the_span(row_bytes, n_rows):
    rng = np.random.default_rng(row_bytes * 31 + n_rows)
    spitch = (row_bytes + 15) // 16 * 16
    for align in range(16):
        rows = rng.integers(0, 256, size=(n_rows, row_bytes), dtype=np.uint8)
        stage = rng.integers(0, 256, size=64 + (n_rows + 1) * spitch + 64, dtype=np.uint8)   # garbage in the padding
        stage_lane = 64
        for r in range(n_rows):
            stage[stage_lane + r * spitch:stage_lane + r * spitch + row_bytes] = rows[r]
        g = 4096 + align
        mem = np.full(4096 + 16 + n_rows * row_bytes + 4096, 0xEE, np.uint8)
        written = np.zeros(mem.shape, np.int32)
        flush_rows_tight(mem, written, g, stage, stage_lane, n_rows, row_bytes, spitch)
        span = n_rows * row_bytes
        assert np.array_equal(mem[g:g + span], rows.reshape(-1))
        assert (written[g:g + span] == 1).all() and written.sum() == span
        assert (mem[:g] == 0xEE).all() and (mem[g + span:] == 0xEE).all()


def test_slot_records_cover_a_group():
    """The producer's per-slot records: output row k-2 comes from input row k >= 2 of a group of nr output rows; slots
    hold RB input rows.  The records of a group's slots must tile its nr output rows in order."""
    RB = 8
    for nr in range(1, 80):
        nslots = (nr + 2 + RB - 1) // RB
        pos = 0
        for s in range(nslots):
            k0, k1 = s * RB, min(s * RB + RB, nr + 2)
            kf = max(k0, 2)
            n_rows = max(k1 - kf, 0)
            assert kf - 2 == pos or n_rows == 0
            pos += n_rows
        assert pos == nr


# ---------------------------------------------------------------------------------------------- TIGHT input side
def issue_slot_tight(mem, base, rows, row_bytes, pitch, first_row, n_slot_rows, top=None, bot=None, RB=8):
    """The producer's side (stream_issue_slot_tight): one image lane of one ring slot.  Input rows first_row-1 ...
    (row -1 / row `rows` = the halo rows `top` / `bot`, or the replicated edge row) are copied run by run as ALIGNED
    SUPERSETS: the copy starts at the 16-byte boundary below the run's first byte and ends at the one above its last.
    -> (lane bytes, rowoff[]) where rowoff[i] = offset of slot row i's first byte within the lane."""
    lane = np.full(RB * pitch + 16 + 128, 0xCC, np.uint8)
    rowoff = []
    cur = 0
    k = 0
    while k < n_slot_rows:
        j = first_row - 1 + k
        run = 1
        if j < 0:
            rp = top if top is not None else base
        elif j >= rows:
            rp = bot if bot is not None else base + (rows - 1) * pitch
        else:
            rp = base + j * pitch
            run = min(n_slot_rows - k, rows - j)
        delta = rp & 15
        nbytes = (delta + (run - 1) * pitch + row_bytes + 15) & ~15
        assert (rp - delta) % 16 == 0 and nbytes % 16 == 0 and cur % 16 == 0      # what a bulk copy needs
        assert nbytes - (delta + (run - 1) * pitch + row_bytes) < 16 and delta < 16   # never more than 15 extra bytes per end
        lane[16 + cur:16 + cur + nbytes] = mem[rp - delta:rp - delta + nbytes]
        for i in range(run):
            rowoff.append(cur + delta + i * pitch)
        cur += nbytes
        k += run
    assert 16 + cur <= len(lane)
    return lane, rowoff


def window(lane, a):
    """The consumer's 24-byte window {wl, w, wr} at byte address a - 4: seven aligned words, six funnel shifts."""
    a4 = (a - 4) & ~3
    sh = (a - 4) & 3
    x = lane[a4:a4 + 28]
    return x[sh:sh + 24].copy()


@pytest.mark.parametrize("row_bytes,pitch", [(750, 750), (257, 257), (300, 300), (4095, 4095), (750, 777), (256, 301)])
@pytest.mark.parametrize("base_align", [0, 1, 7, 13])
def test_tight_input_rows_land_where_the_consumers_look(row_bytes, pitch, base_align):
    rng = np.random.default_rng(row_bytes + pitch + base_align)
    rows = 21
    base = 4096 + base_align
    mem = rng.integers(0, 256, size=base + rows * pitch + 4096, dtype=np.uint8)
    top_at, bot_at = 64 + 5, 2048 + 11                       # halo rows somewhere else, at their own alignment
    for first_row, n_slot_rows, top, bot in [(0, 8, None, None), (0, 8, top_at, None), (5, 8, None, None), (14, 8, None, bot_at),
                                             (14, 8, None, None), (20, 3, top_at, bot_at), (1, 2, None, None)]:
        lane, rowoff = issue_slot_tight(mem, base, rows, row_bytes, pitch, first_row, n_slot_rows, top, bot)
        assert len(rowoff) == n_slot_rows and max(rowoff) + row_bytes + 20 <= len(lane) - 16 + 16
        for i, off in enumerate(rowoff):
            j = first_row - 1 + i
            if j < 0:
                src = top if top is not None else base
            elif j >= rows:
                src = bot if bot is not None else base + (rows - 1) * pitch
            else:
                src = base + j * pitch
            want = mem[src:src + row_bytes]
            assert np.array_equal(lane[16 + off:16 + off + row_bytes], want)
            for c in (0, 1, (row_bytes - 1) // 16):            # first chunks and the chunk that holds the end of the row
                w = window(lane, 16 + off + 16 * c)
                n = min(16, row_bytes - 16 * c)
                assert np.array_equal(w[4:4 + n], want[16 * c:16 * c + n])
                if c > 0:
                    assert np.array_equal(w[:4], want[16 * c - 4:16 * c])
