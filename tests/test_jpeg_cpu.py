"""CPU tests of the ingest step: host/jpeg_decode.hpp, the dependency-free baseline JPEG reader
that lets the CLIs open the reference's own image_320x240.jpg (heterogeneous_blur.c:43, :106), must decode every
fixture to exactly the bytes libjpeg-turbo decodes (sha256 recorded by tests/golden/make_jpeg_golden.py) -- the blur's
bit-exact parity is only worth something if its input bytes are the reference's too."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import b200blur

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200", "bin")
GOLD = os.path.join(ROOT, "tests", "golden", "jpeg")
with open(os.path.join(GOLD, "expected.json")) as _f:
    EXPECTED = json.load(_f)


@pytest.fixture(scope="module", autouse=True)
def _built():
    b200blur.build()


def decode(path, tmp_path):
    out = os.path.join(tmp_path, "out.ppm")
    r = subprocess.run([os.path.join(BIN, "jpeg2ppm"), path, out], capture_output=True, text=True, timeout=60)
    if r.returncode != 0:
        return None, r.stdout
    data = open(out, "rb").read()
    magic, dims, maxval, pixels = data.split(b"\n", 3)
    w, h = map(int, dims.split())
    c = 1 if magic == b"P5" else 3
    return np.frombuffer(pixels, np.uint8).reshape(h, w, c), r.stdout


@pytest.mark.parametrize("name", sorted(EXPECTED["files"]))
def test_decodes_to_libjpeg_bytes(name, tmp_path):
    want = EXPECTED["files"][name]
    got, msg = decode(os.path.join(GOLD, name), str(tmp_path))
    assert got is not None, msg
    assert got.shape == (want["height"], want["width"], want["channels"])
    assert hashlib.sha256(got.tobytes()).hexdigest() == want["sha256"]
    try:  # where Pillow is importable, also compare pixel by pixel (a clearer failure than a hash)
        from PIL import Image
        im = Image.open(os.path.join(GOLD, name))
        ref = np.array(im if im.mode == "L" else im.convert("RGB")).reshape(got.shape)
        assert np.array_equal(got, ref), f"max abs diff {np.abs(got.astype(int) - ref.astype(int)).max()}"
    except ImportError:
        pass


@pytest.mark.parametrize("name", EXPECTED["unsupported"])
def test_unsupported_files_fail_cleanly(name, tmp_path):
    got, msg = decode(os.path.join(GOLD, name), str(tmp_path))
    assert got is None and "not supported" in msg
    got, msg = decode(os.path.join(GOLD, "expected.json"), str(tmp_path))
    assert got is None and "not a JPEG" in msg


@pytest.mark.skipif(not os.path.exists("/root/reference/image_320x240.jpg"), reason="reference tree not present")
@pytest.mark.parametrize("name", ["image_320x240.jpg", "image_256x256.jpg", "split_output.jpg"])
def test_reference_images_decode_like_libjpeg(name, tmp_path):
    from PIL import Image
    path = os.path.join("/root/reference", name)
    got, msg = decode(path, str(tmp_path))
    assert got is not None, msg
    assert np.array_equal(got, np.array(Image.open(path).convert("RGB")))


def test_cli_opens_a_jpeg(tmp_path):
    """The CLI's LOAD ORIGINAL IMAGE step takes the .jpg directly (and then, with no GPU here, fails like the reference
    does without an OpenCL device, heterogeneous_blur.c:181-184)."""
    name = "420_q90_64x48.jpg"
    out = subprocess.run([os.path.join(BIN, "heterogeneous_blur"), "gpu", "0.5", "4", "--images", "8", "--input",
                          os.path.join(GOLD, name), "--quiet"], capture_output=True, text=True, timeout=120, cwd=str(tmp_path))
    assert "Original image loaded: 64x48, 3 channels" in out.stdout
    assert "Size of one image: 9216 bytes" in out.stdout
