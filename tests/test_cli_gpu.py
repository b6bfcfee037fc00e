"""GPU tests of the two drop-in CLIs (C++ host programs over the C ABI): their saved output image must equal the oracle's
blur of the input they were given, for Approach 1 and Approach 2, end-to-end and device-resident."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200", "bin")


def write_ppm(path, img):
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(img.tobytes())


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = map(int, f.readline().split())
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module", autouse=True)
def _binaries():
    if not (os.path.exists(os.path.join(BIN, "heterogeneous_blur")) and os.path.exists(os.path.join(BIN, "split_image_blur"))):
        import b200blur
        b200blur.build()  # nvcc + g++ are part of the image on the GPU box too


@pytest.fixture(scope="module")
def photo(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    img = np.random.default_rng(5).integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    path = os.path.join(d, "in.ppm")
    write_ppm(path, img)
    return d, path, img, oracle.c_blur(img)


def run(cmd, cwd):
    out = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize("extra", [[], ["--resident"]], ids=["end-to-end", "resident"])
@pytest.mark.parametrize("mode", ["gpu", "both", "cpu"])
def test_heterogeneous_blur_cli(photo, mode, extra):
    d, path, img, want = photo
    out_path = os.path.join(d, f"a1_{mode}_{len(extra)}.ppm")
    text = run([os.path.join(BIN, "heterogeneous_blur"), mode, "0.728", "35", "--images", "143", "--input", path,
                "--save", out_path, "--quiet", "--checksum"] + extra, d)
    assert "HETEROGENEOUS CONFIGURATION" in text and "7. THROUGHPUT" in text and "Images per second" in text
    assert "Number of batches: 5" in text  # ceil(143/35), heterogeneous_blur.c:86
    assert np.array_equal(read_ppm(out_path), want)


def test_heterogeneous_blur_cli_argument_handling(photo):
    d, path, _, _ = photo
    text = run([os.path.join(BIN, "heterogeneous_blur"), "nonsense", "7", "99999", "--images", "10", "--input", path, "--quiet"], d)
    # heterogeneous_blur.c:62-65, :72-75, :80-83
    assert "Usage:" in text and "Defaulting to heterogeneous mode." in text
    assert "Warning: gpu_ratio must be between 0.0 and 1.0. Using 0.5" in text
    assert "Warning: BATCH_SIZE must be between 1 and 10. Using 500" in text


@pytest.mark.parametrize("extra", [[], ["--resident"], ["--host-halo"]], ids=["p2p", "resident", "host-halo"])
def test_split_image_blur_cli(photo, extra):
    d, path, img, want = photo
    g = min(n_gpus(), 8)
    out_path = os.path.join(d, f"a2_{len(extra)}_{'_'.join(extra)}.ppm")
    text = run([os.path.join(BIN, "split_image_blur"), "0.837", "35", "--images", "80", "--input", path, "--save", out_path,
                "--quiet", "--checksum"] + extra, d)
    assert "SPLIT-IMAGE CONFIGURATION" in text and "split row 39" in text  # split_image_blur.c:144 known answer
    assert f"Row bands over {g} GPU(s)" in text and f"{max(g, 2) + 7}. OPTIMAL RATIO RECOMMENDATION" in text
    assert np.array_equal(read_ppm(out_path), want)


def _field(text, label):
    for line in text.splitlines():
        if label in line:
            return line.split(label, 1)[1].strip()
    raise AssertionError(f"{label!r} not in output")


@pytest.mark.parametrize("extra", [[], ["--static-split"]], ids=["work-stealing", "static-split"])
def test_heterogeneous_blur_cli_four_workers(photo, extra):
    """Approach 1 over 4 GPU workers (on a box with fewer GPUs the workers share devices: --oversubscribe).  By default the
    workers TAKE groups of batches from a shared counter (dynamic scheduling instead of the reference's
    hand-tuned ratio, heterogeneous_blur.c:713-722); every image is processed exactly once either way."""
    d, path, img, want = photo
    out_path = os.path.join(d, f"a1_four_{len(extra)}.ppm")
    text = run([os.path.join(BIN, "heterogeneous_blur"), "both", "0.5", "5", "--images", "1003", "--gpus", "4", "--oversubscribe",
                "--input", path, "--save", out_path, "--quiet"] + extra, d)
    assert "Total images processed: 1003" in text
    assert "6. DEVICE COMPARISON" in text and "9. THROUGHPUT" in text and "10. OPTIMAL RATIO RECOMMENDATION" in text
    shares = [int(line.split("processed")[1].split("images")[0]) for line in text.splitlines() if "DEVICE (processed" in line]
    assert sum(shares) == 1003 and len(shares) >= 1
    if extra:
        assert len(shares) == 4 and max(shares) - min(shares) <= 201   # even shares of every batch of 5 (1 or 2 images)
    assert np.array_equal(read_ppm(out_path), want)


def test_split_image_blur_cli_eight_bands_stress(photo):
    """8 row bands x 200 batches with a ring of only 2 slots: every batch, each band's kernel waits for its neighbours'
    uploads and each upload waits for the neighbours' previous kernels THROUGH THE OTHER THREAD'S CONTEXT
    (b200blur_enqueue_wait_peer) while that thread keeps creating events (round-1 race: the event pool reallocated under
    the reader).  All outputs (checksum) must equal the host-halo scheme's, and image 0 the oracle's."""
    d, path, img, want = photo
    sums = []
    for i, extra in enumerate([[], ["--host-halo"]]):
        out_path = os.path.join(d, f"a2_stress_{i}.ppm")
        text = run([os.path.join(BIN, "split_image_blur"), "0.5", "1", "--images", "200", "--gpus", "8", "--oversubscribe",
                    "--ring", "2", "--fuse", "1", "--input", path, "--save", out_path, "--quiet", "--checksum"] + extra, d)
        assert "Row bands over 8 GPU(s)" in text and "Total images processed: 200" in text
        assert np.array_equal(read_ppm(out_path), want)
        sums.append(_field(text, "Output checksum"))
    assert sums[0] == sums[1]


@pytest.mark.parametrize("prog,args", [("heterogeneous_blur", ["both", "0.5", "7"]), ("split_image_blur", ["0.5", "7"])])
def test_cli_stage_once_gives_the_same_outputs(photo, prog, args):
    """--stage-once fills every pinned ring slot once before the timer and re-sends it per batch (the stream is one image
    repeated); every output byte (FNV-1a over all of them) must equal the default per-batch staging, ragged last batch and
    ring wrap-around included."""
    d, path, img, want = photo
    sums = []
    for i, extra in enumerate([[], ["--stage-once"]]):
        out_path = os.path.join(d, f"{prog}_once_{i}.ppm")
        text = run([os.path.join(BIN, prog)] + args + ["--images", "100", "--gpus", "2", "--oversubscribe", "--ring", "2",
                    "--fuse", "1", "--input", path, "--save", out_path, "--quiet", "--checksum"] + extra, d)
        assert "Total images processed: 100" in text
        assert ("once per ring slot" in text) == bool(extra)
        assert np.array_equal(read_ppm(out_path), want)
        sums.append(_field(text, "Output checksum"))
    assert sums[0] == sums[1]


@pytest.mark.parametrize("name", ["420_q30_96x80.jpg", "420_photo_crop_80x50.jpg", "420_q75_odd_101x67.jpg"])
def test_cli_blurs_a_jpeg_input(tmp_path, name):
    """Ingest -> hot path: the CLI decodes a .jpg itself (host/jpeg_decode.hpp, byte-identical to libjpeg) and its saved
    output equals the oracle's blur of those decoded pixels."""
    src = os.path.join(ROOT, "tests", "golden", "jpeg", name)
    ppm = os.path.join(tmp_path, "decoded.ppm")
    subprocess.run([os.path.join(BIN, "jpeg2ppm"), src, ppm], check=True, capture_output=True)
    want = oracle.c_blur(read_ppm(ppm))
    out_path = os.path.join(tmp_path, "out.ppm")
    text = run([os.path.join(BIN, "heterogeneous_blur"), "gpu", "0.5", "7", "--images", "50", "--input", src, "--save", out_path,
                "--quiet"], str(tmp_path))
    assert "Total images processed: 50" in text
    assert np.array_equal(read_ppm(out_path), want)


@pytest.mark.parametrize("prog,args", [("heterogeneous_blur", ["both", "0.5", "9"]), ("split_image_blur", ["0.6", "9"])])
def test_clis_on_an_odd_width_image(tmp_path, prog, args):
    """A 251x97 image (753-byte rows: neither 16-byte multiples nor aligned from row to row) through both CLIs on three
    workers: the tight-row kernel (unaligned loads, store warps) under Approach 1 shards and under Approach 2 bands whose
    halo rows are read through the neighbours' buffers."""
    img = np.random.default_rng(77).integers(0, 256, size=(97, 251, 3), dtype=np.uint8)
    src, out_path = os.path.join(tmp_path, "odd.ppm"), os.path.join(tmp_path, "odd_out.ppm")
    write_ppm(src, img)
    text = run([os.path.join(BIN, prog)] + args + ["--images", "120", "--gpus", "3", "--oversubscribe", "--input", src, "--save",
                                                   out_path, "--quiet"], str(tmp_path))
    assert "Total images processed: 120" in text
    assert np.array_equal(read_ppm(out_path), oracle.c_blur(img))
