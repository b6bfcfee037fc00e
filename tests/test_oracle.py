"""CPU tests of the parity checker itself (oracle/): the C restatement, the numpy form and -- where it was built --
the reference's own gaussian_kernel.cl must agree bit for bit, and all must reproduce the committed golden vectors
(tests/golden/, generated from the reference kernel by tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle

SHAPES = [(1, 1, 3), (1, 9, 3), (9, 1, 3), (2, 2, 3), (3, 5, 3), (17, 33, 3), (16, 16, 3), (64, 48, 3), (240, 320, 3),
          (256, 256, 3), (12, 20, 1), (12, 20, 2), (12, 20, 4)]


def synth(seed, n, h, w, c=3):
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)


@pytest.fixture(scope="module", autouse=True)
def _built():
    oracle.build()


def test_known_answers_from_the_kernel_spec():
    """truncation (not rounding) and clamp-to-edge weight folding (gaussian_kernel.cl:36-41,:56-57,:70)."""
    img = np.zeros((5, 5, 3), np.uint8)
    img[2, 2] = 255
    out = oracle.c_blur(img)
    assert out[1:4, 1:4, 0].tolist() == [[15, 31, 15], [31, 63, 31], [15, 31, 15]]
    assert out.sum() == 3 * (4 * 15 + 4 * 31 + 63)
    img = np.zeros((5, 5, 3), np.uint8)
    img[0, 0] = 255
    out = oracle.c_blur(img)
    assert (out[0, 0, 0], out[0, 1, 0], out[1, 0, 0], out[1, 1, 0]) == (143, 47, 47, 15)
    img = np.zeros((5, 5, 3), np.uint8)
    img[0, 2] = 255
    out = oracle.c_blur(img)
    assert out[0, :, 1].tolist() == [0, 47, 95, 47, 0] and out[1, :, 1].tolist() == [0, 15, 31, 15, 0]
    for k in (0, 1, 255):
        assert (oracle.c_blur(np.full((4, 6, 3), k, np.uint8)) == k).all()


@pytest.mark.parametrize("shape", SHAPES)
def test_float_int_numpy_forms_agree(shape):
    h, w, c = shape
    img = synth(h * 1000 + w, 1, h, w, c)[0]
    f = oracle.c_blur(img)
    assert np.array_equal(f, oracle.c_blur(img, integer=True))
    assert np.array_equal(f, oracle.np_blur(img))


@pytest.mark.parametrize("shape", SHAPES)
def test_restatement_matches_reference_kernel_source(shape):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree on this machine)")
    h, w, c = shape
    img = synth(7 + h * w, 1, h, w, c)[0]
    assert np.array_equal(oracle.ref_blur(img), oracle.c_blur(img))


def test_batch_forms_match_single_image_form():
    x = synth(11, 5, 33, 21, 3)
    y = oracle.c_blur_batch(x)
    for i in range(5):
        assert np.array_equal(y[i], oracle.c_blur(x[i]))
    assert np.array_equal(y, oracle.c_blur_batch(x, integer=True))
    assert np.array_equal(y, oracle.np_blur(x))
    if oracle.have_ref():
        assert np.array_equal(y, oracle.ref_blur_batch(x))


def test_golden_vectors(golden_dir):
    vec = np.load(os.path.join(golden_dir, "vectors.npz"))
    names = sorted(k[3:] for k in vec.files if k.startswith("in_"))
    assert len(names) >= 25
    for name in names:
        x, y = vec["in_" + name], vec["out_" + name]
        assert np.array_equal(oracle.c_blur(x), y), name
        assert np.array_equal(oracle.c_blur(x, integer=True), y), name
        assert np.array_equal(oracle.np_blur(x), y), name


def test_golden_checksums(golden_dir):
    with open(os.path.join(golden_dir, "checksums.json")) as f:
        sums = json.load(f)
    for case in sums["cases"]:
        x = synth(case["seed"], case["n"], case["h"], case["w"], case["c"])
        assert hashlib.sha256(x.tobytes()).hexdigest() == case["in_sha256"], "synthetic generator drifted"
        y = oracle.c_blur_batch(x, integer=True)
        assert hashlib.sha256(y.tobytes()).hexdigest() == case["out_sha256"], case


def test_distribution_known_answers_from_reference_logs(golden_dir):
    with open(os.path.join(golden_dir, "distribution.json")) as f:
        dist = json.load(f)
    for k in dist["a1"]:
        nb, tc, tg = oracle.a1_totals(k["num_images"], k["batch_size"], k["gpu_ratio"], k["mode"])
        assert (nb, tc, tg) == (k["num_batches"], k["total_cpu"], k["total_gpu"]), k["source"]
    for k in dist["a2"]:
        g = oracle.a2_geometry(k["height"], k["gpu_ratio"])
        for key in ("split_row", "cpu_input_rows", "cpu_output_rows", "gpu_input_rows", "gpu_output_rows"):
            assert g[key] == k[key], (key, k["source"])
    # modes and clamps: heterogeneous_blur.c:452-458, split_image_blur.c:147-154
    assert oracle.a1_batch_split(35, 0.728, 1) == (35, 0)
    assert oracle.a1_batch_split(35, 0.728, 2) == (0, 35)
    assert oracle.a1_batch_split(30, 0.728, 0) == (9, 21)
    assert oracle.a2_geometry(240, 1.0)["split_row"] == 1
    assert oracle.a2_geometry(240, 0.0)["split_row"] == 239


@pytest.mark.parametrize("h,w", [(240, 320), (256, 256), (9, 5), (2, 7)])
def test_split_image_composition_equals_whole_image(h, w):
    """Approach 2's halo trick is bit-identical to the whole-image kernel."""
    img = synth(h + w, 1, h, w, 3)[0]
    whole = oracle.c_blur(img)
    for split_row in sorted({1, h // 2, h - 1, max(1, h // 3)}):
        assert np.array_equal(oracle.split_image(img, split_row), whole), split_row
    for g in (1, 2, 4, 8):
        if g <= h:
            assert np.array_equal(oracle.band_split(img, g), whole), g
