"""CPU emulation of the vectorised kernel's arithmetic, checked against the oracle -- no GPU needed.

The CUDA kernels (csrc/blur_kernels.cuh) blur two byte-columns at a time in packed 16-bit lanes with PRMT/IMAD/LOP3; this
test re-states that instruction sequence in Python (byte_perm = PRMT) and drives it with the HOST-side plan the library
computes for the right edge of rows that do not end on a 16-byte chunk boundary (b200blur_plan_row_edge), so both the
lane arithmetic described in DESIGN.md section 4 and the selector logic are pinned on every machine."""
import numpy as np
import pytest

import b200blur
from oracle import oracle

M = 0xFFFFFFFF


def byte_perm(x, y, s):
    b = [(x >> (8 * i)) & 255 for i in range(4)] + [(y >> (8 * i)) & 255 for i in range(4)]
    return sum(b[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def lanes_shift(lo, hi):
    return byte_perm(lo, hi, 0x5432)


def hpass(C, w, wl, wr):
    W = [wl] + list(w) + [wr]
    E = [x & 0x00FF00FF for x in W]
    O = [byte_perm(x, 0, 0x4341) for x in W]
    h = []
    for k in range(1, 5):
        if C == 3:
            LE, LO, RE, RO = O[k - 1], lanes_shift(E[k - 1], E[k]), lanes_shift(O[k], O[k + 1]), E[k + 1]
        elif C == 4:
            LE, LO, RE, RO = E[k - 1], O[k - 1], E[k + 1], O[k + 1]
        elif C == 2:
            LE, LO = lanes_shift(E[k - 1], E[k]), lanes_shift(O[k - 1], O[k])
            RE, RO = lanes_shift(E[k], E[k + 1]), lanes_shift(O[k], O[k + 1])
        else:
            LE, LO, RE, RO = lanes_shift(O[k - 1], O[k]), E[k], O[k], lanes_shift(E[k], E[k + 1])
        h += [(2 * E[k] + LE + RE) & M, (2 * O[k] + LO + RO) & M]
    return h


def emulate(img):
    """One image through the streamed kernel's per-thread program, rows padded to a multiple of 16 bytes with garbage."""
    H, Wd, C = img.shape
    rb = Wd * C
    pitch = (rb + 15) // 16 * 16
    buf = np.random.default_rng(1).integers(0, 256, (H, pitch + 32), dtype=np.uint8)
    buf[:, :rb] = img.reshape(H, rb)
    words = np.ascontiguousarray(buf).view("<u4")
    plan = b200blur.plan_row_edge(rb, C)
    cpr, sl = plan["chunks"], plan["sel_last"]
    assert cpr == (rb + 15) // 16 and plan["edge_general"] == (rb % 16 != 0)
    out = np.zeros((H, pitch), np.uint8)
    for c in range(cpr):
        accA, accB = [0] * 8, [0] * 8      # 16*(h[k-2] + 2h[k-1]) and 16*h[k-1]
        for k, j in enumerate(range(-1, H + 1)):
            jj = min(max(j, 0), H - 1)     # replicated edge rows (gaussian_kernel.cl:57)
            w = [int(x) for x in words[jj, 4 * c:4 * c + 4]]
            wl = int(words[jj, 4 * c - 1]) if c > 0 else (w[0] << (8 * (4 - C))) & M
            wr = int(words[jj, 4 * c + 4])
            last, prev_last = c == cpr - 1, plan["edge_prev"] and c == cpr - 2
            if not plan["edge_general"]:
                if last:
                    wr = w[3] >> (8 * (4 - C))
            elif last:
                n = [byte_perm(a, b, s) for a, b, s in zip([wl] + w[:3], w, sl[1:5])]
                wr = byte_perm(w[3], wr, sl[5])
                w = n
            elif prev_last:
                wr = byte_perm(w[3], wr, plan["sel_prev"])
            h = hpass(C, w, wl, wr)
            v = [(h[i] * 16 + accA[i]) & M for i in range(8)]
            accA = [(h[i] * 32 + accB[i]) & M for i in range(8)]
            accB = [(h[i] << 4) & M for i in range(8)]
            assert all(((x & 0xFFFF) <= 65280) and ((x >> 16) <= 65280) for x in v) or k < 2
            if k >= 2:
                o = [byte_perm(v[2 * i], v[2 * i + 1], 0x7351) for i in range(4)]
                out[k - 2, 16 * c:16 * c + 16] = np.array(o, dtype="<u4").view(np.uint8)
    return out[:, :rb].reshape(H, Wd, C)


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200blur.build()
    oracle.build()


@pytest.mark.parametrize("C", [1, 2, 3, 4])
def test_packed_lane_program_matches_oracle_for_every_row_length(C):
    rng = np.random.default_rng(C)
    for Wd in list(range(1, 24)) + [32, 33, 48, 86, 87, 100]:
        img = rng.integers(0, 256, (4, Wd, C), dtype=np.uint8)
        assert np.array_equal(emulate(img), oracle.c_blur(img)), (C, Wd)
    img = np.full((3, 40, C), 255, np.uint8)     # every lane at its maximum: 4080 << 4 must not carry
    assert np.array_equal(emulate(img), oracle.c_blur(img))


def test_row_edge_plan_known_cases():
    p = b200blur.plan_row_edge(960, 3)           # 320 px RGB: rows end on a chunk boundary -> fast path
    assert p["chunks"] == 60 and not p["edge_general"] and not p["edge_prev"]
    p = b200blur.plan_row_edge(750, 3)           # 250 px RGB: 14 bytes into the last chunk
    assert p["chunks"] == 47 and p["edge_general"] and not p["edge_prev"]
    p = b200blur.plan_row_edge(258, 3)           # 86 px RGB: the last pixel straddles two chunks
    assert p["chunks"] == 17 and p["edge_general"] and p["edge_prev"]
    with pytest.raises(b200blur.BlurError):
        b200blur.plan_row_edge(10, 3)            # not a whole number of pixels
