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
