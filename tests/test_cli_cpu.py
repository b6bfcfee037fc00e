"""CPU tests of the CLIs' argument surface: without a GPU the two programs must still parse and echo the reference's
positional arguments, print the reference's warnings, and then fail loudly (exit status 255 = return -1) instead of
falling back to any CPU path (heterogeneous_blur.c:52-100, :181-184; split_image_blur.c:72-102)."""
import os
import subprocess

import pytest

import b200blur

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200blur.build()


def run(args):
    return subprocess.run(args, capture_output=True, text=True, timeout=120)


def no_gpu():
    return b200blur.device_count() == 0


def test_heterogeneous_blur_banner_and_defaults():
    out = run([os.path.join(BIN, "heterogeneous_blur"), "--images", "8"])
    t = out.stdout
    assert "Mode: HETEROGENEOUS (CPU + GPU) [default]" in t
    assert "GPU ratio: 50.0% GPU, 50.0% CPU" in t
    assert "Number of images in stream: 8" in t and "Work-group size: 16x16" in t and "Execution mode : 0" in t
    assert "Original image loaded: 320x240, 3 channels" in t and "Size of one image: 230400 bytes (225.00 KB)" in t
    if no_gpu():
        assert out.returncode == 255 and "Error: Could not find a CUDA device" in t


@pytest.mark.parametrize("mode,line,code", [("cpu", "Mode: CPU ONLY", 1), ("gpu", "Mode: GPU ONLY", 2),
                                            ("both", "Mode: HETEROGENEOUS (CPU + GPU)", 0)])
def test_heterogeneous_blur_modes(mode, line, code):
    t = run([os.path.join(BIN, "heterogeneous_blur"), mode, "0.728", "35", "--images", "70"]).stdout
    assert line in t and f"Execution mode : {code}" in t
    assert "Batch size: 35 images" in t and "Number of batches: 2" in t
    assert ("GPU ratio: 72.8% GPU, 27.2% CPU" in t) == (mode == "both")  # shown for heterogeneous mode only (:90-92)


def test_heterogeneous_blur_bad_arguments_warn_and_default():
    t = run([os.path.join(BIN, "heterogeneous_blur"), "whatever", "1.5", "0"]).stdout
    assert "Usage:" in t and "[cpu|gpu|both]" in t and "Defaulting to heterogeneous mode." in t
    assert "Warning: gpu_ratio must be between 0.0 and 1.0. Using 0.5" in t
    assert "Warning: BATCH_SIZE must be between 1 and 5000. Using 500" in t
    assert "Batch size: 500 images" in t and "Number of batches: 10" in t
    assert run([os.path.join(BIN, "heterogeneous_blur"), "gpu", "--bogus"]).returncode == 255


def test_split_image_blur_banner_and_arguments():
    out = run([os.path.join(BIN, "split_image_blur"), "0.837", "35"])
    t = out.stdout
    assert "SPLIT-IMAGE CONFIGURATION" in t and "GPU ratio: 83.7% (rows to GPU)" in t and "Halo size: 1 row(s)" in t
    assert "Batch size: 35 images" in t and "Number of batches: 143" in t
    if no_gpu():
        assert out.returncode == 255 and "Error: Could not find a CUDA device" in t
    t = run([os.path.join(BIN, "split_image_blur"), "-3", "99999"]).stdout
    assert "Warning: gpu_ratio must be between 0.0 and 1.0. Using 0.5" in t
    assert "Warning: BATCH_SIZE must be between 1 and 5000. Using 500" in t


@pytest.mark.parametrize("prog", ["heterogeneous_blur", "split_image_blur"])
def test_trailing_options_are_parsed_after_the_reference_positionals(prog, tmp_path):
    """The options that follow the reference's positional arguments (they only expose what the reference hard-codes, plus
    the multi-GPU knobs of round 2) are accepted, bad values are refused, and a JPEG --input is decoded before the
    device is looked for."""
    pos = ["both", "0.5", "7"] if prog == "heterogeneous_blur" else ["0.5", "7"]
    exe = os.path.join(BIN, prog)
    jpg = os.path.join(ROOT, "tests", "golden", "jpeg", "420_q90_64x48.jpg")
    ok = run([exe] + pos + ["--images", "21", "--gpus", "3", "--oversubscribe", "--ring", "2", "--fuse", "1", "--fill-threads", "2",
                            "--static-split", "--quiet", "--checksum", "--input", jpg])
    assert "Error: unknown option" not in ok.stdout and "Error: bad size option" not in ok.stdout
    assert "Original image loaded: 64x48, 3 channels" in ok.stdout and "Number of batches: 3" in ok.stdout
    if no_gpu():
        assert ok.returncode == 255 and "Could not find a CUDA device" in ok.stdout
    for bad in (["--ring", "1"], ["--ring", "65"], ["--images", "0"], ["--width", "-5"], ["--repeat", "0"], ["--nonsense"]):
        out = run([exe] + pos + bad)
        assert out.returncode == 255 and "Error:" in out.stdout
    missing = run([exe] + pos + ["--input", os.path.join(str(tmp_path), "nothing.jpg")])
    assert missing.returncode == 255 and "cannot read image file" in missing.stdout
    progressive = run([exe] + pos + ["--input", os.path.join(ROOT, "tests", "golden", "jpeg", "progressive_32x24.jpg")])
    assert progressive.returncode == 255 and "not supported" in progressive.stdout
