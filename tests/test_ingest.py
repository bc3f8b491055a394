"""Ingest helpers (csrc/ingest.cpp) against OpenCV: cv2.imdecode is the authority for what cv::imread hands the reference's detector
(src/demo.cpp:88-99); the sensor_msgs/Image conversions follow cv_bridge::toCvCopy(msg, "bgr8" / "32FC1") (ros/Node.cpp:165-176).
Host code: no GPU needed (the pinned-buffer test is the exception)."""
import struct
import zlib

import cv2
import numpy as np
import pytest

from partsbaseddetector_b200 import PbdError, ingest
from partsbaseddetector_b200.synth import synth_frame


def _png(img, params=None):
    ok, buf = cv2.imencode(".png", img, params or [])
    assert ok
    return buf.tobytes()


@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (120, 160), (33, 257)])
def test_png_decoding_equals_opencv(shape):
    h, w = shape
    rng = np.random.default_rng(h * w)
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    cases = {
        "bgr8": bgr,
        "grey8": bgr[:, :, 0].copy(),
        "bgra8": np.dstack([bgr, rng.integers(0, 256, (h, w), dtype=np.uint8)]),
        "grey16": rng.integers(0, 65536, (h, w), dtype=np.uint16),
        "bgr16": rng.integers(0, 65536, (h, w, 3), dtype=np.uint16),
        "smooth": cv2.GaussianBlur(np.ascontiguousarray(bgr), (0, 0), 3),      # exercises the Sub / Up / Average / Paeth row filters
    }
    for name, img in cases.items():
        for level in (1, 9):
            data = _png(img, [cv2.IMWRITE_PNG_COMPRESSION, level])
            want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
            got = ingest.imdecode(data)
            assert got.shape == want.shape and np.array_equal(got, want), (name, level)
    d16 = cases["grey16"]
    got = ingest.imdecode_depth(_png(d16), 1.0 / 1000.0)
    assert np.array_equal(got, d16.astype(np.float32) * np.float32(1.0 / 1000.0))          # `depth / 1000.0f`, src/demo.cpp:98


def _chunk(tag, body):
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)


def test_png_palette_and_low_bit_depths():
    """Colour type 3 (palette) and 1/2/4-bit grey, built by hand (cv2 does not write them) and checked against cv2's decoder."""
    rng = np.random.default_rng(4)
    h, w = 9, 21
    for depth in (1, 2, 4, 8):
        idx = rng.integers(0, 1 << depth, (h, w), dtype=np.uint8)
        rows = b""
        for y in range(h):
            bits = "".join(format(int(v), "0%db" % depth) for v in idx[y])
            bits += "0" * (-len(bits) % 8)
            rows += b"\x00" + bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
        pal = rng.integers(0, 256, (1 << depth, 3), dtype=np.uint8)
        for ctype in (3, 0):
            png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
            if ctype == 3:
                png += _chunk(b"PLTE", pal.tobytes())
            png += _chunk(b"IDAT", zlib.compress(rows)) + _chunk(b"IEND", b"")
            want = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_COLOR)
            assert want is not None
            assert np.array_equal(ingest.imdecode(png), want), (depth, ctype)


def test_pnm_and_errors(tmp_path):
    img = synth_frame(3, 48, 64)
    ppm = b"P6\n# comment\n64 48\n255\n" + np.ascontiguousarray(img[:, :, ::-1]).tobytes()
    assert np.array_equal(ingest.imdecode(ppm), img)
    pgm16 = b"P5 64 48 65535\n" + (img[:, :, 0].astype(">u2") * 257).tobytes()
    assert np.array_equal(ingest.imdecode_depth(pgm16, 1.0), img[:, :, 0].astype(np.float32) * 257)
    p = tmp_path / "f.png"
    p.write_bytes(_png(img))
    assert np.array_equal(ingest.imread(str(p)), img)
    for bad, code in ((b"\xff\xd8\xff\xe0jfif", -6), (b"GIF89a", -3), (_png(img)[:200], -3), (b"P6\n64 48\n255\n123", -3)):
        with pytest.raises(PbdError) as e:
            ingest.imdecode(bad)
        assert e.value.code == code, bad[:6]
    corrupt = bytearray(_png(img))
    corrupt[len(corrupt) // 2] ^= 0x55
    with pytest.raises(PbdError):
        ingest.imdecode(bytes(corrupt))


def test_ros_image_encodings():
    rng = np.random.default_rng(8)
    h, w = 17, 29
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(ingest.from_ros_image("bgr8", h, w, bgr.tobytes()), bgr)
    assert np.array_equal(ingest.from_ros_image("rgb8", h, w, bgr[:, :, ::-1].tobytes()), bgr)
    a = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
    assert np.array_equal(ingest.from_ros_image("bgra8", h, w, np.dstack([bgr, a]).tobytes()), bgr)
    assert np.array_equal(ingest.from_ros_image("rgba8", h, w, np.dstack([bgr[:, :, ::-1], a]).tobytes()), bgr)
    mono = bgr[:, :, 1]
    assert np.array_equal(ingest.from_ros_image("mono8", h, w, mono.tobytes()), cv2.cvtColor(np.ascontiguousarray(mono), cv2.COLOR_GRAY2BGR))
    padded = np.zeros((h, w * 3 + 5), np.uint8)                       # row padding: step > width * 3
    padded[:, :w * 3] = bgr.reshape(h, -1)
    assert np.array_equal(ingest.from_ros_image("bgr8", h, w, padded.tobytes(), step=w * 3 + 5), bgr)
    m16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    want = cv2.cvtColor(cv2.convertScaleAbs(m16, alpha=255.0 / 65535.0), cv2.COLOR_GRAY2BGR)    # cv_bridge: convertTo(CV_8U, 255/65535)
    assert np.array_equal(ingest.from_ros_image("mono16", h, w, m16.astype("<u2").tobytes()), want)
    assert np.array_equal(ingest.from_ros_image("mono16", h, w, m16.astype(">u2").tobytes(), is_bigendian=1), want)
    d = rng.standard_normal((h, w)).astype(np.float32)
    assert np.array_equal(ingest.depth_from_ros_image("32FC1", h, w, d.tobytes()), d)
    assert np.array_equal(ingest.depth_from_ros_image("32FC1", h, w, d.astype(">f4").tobytes(), is_bigendian=1), d)
    assert np.array_equal(ingest.depth_from_ros_image("16UC1", h, w, m16.astype("<u2").tobytes()), m16.astype(np.float32))
    with pytest.raises(PbdError) as e:
        ingest.from_ros_image("bayer_rggb8", h, w, mono.tobytes())
    assert e.value.code == -6
    with pytest.raises(PbdError):
        ingest.from_ros_image("bgr8", h, w, bgr.tobytes(), step=w)


@pytest.mark.gpu
def test_decoded_png_through_the_pinned_ring_equals_detect():
    """PNG bytes -> imdecode -> pinned frame ring -> submit / collect_ticket equals detect() on the raw frames."""
    from conftest import golden_model_path
    from partsbaseddetector_b200 import Model, PartsBasedDetector
    frames = np.stack([synth_frame(500 + i, 120, 160) for i in range(4)])
    d = PartsBasedDetector()
    d.distributeModel(Model.load_bin(golden_model_path("Person_26parts")))
    d.set_option("thresh", -1.3)
    ref = d.detect(frames)
    ring = ingest.PinnedFrames(4, 120, 160, 3)
    for i in range(4):
        ring.array[i] = ingest.imdecode(_png(frames[i]))
    got = d.collect_ticket(d.submit(ring.array))
    assert len(got) == len(ref) > 0
    for a, b in zip(got, ref):
        assert a.frame == b.frame and a.score() == b.score() and np.array_equal(a.parts(), b.parts())
    ring.close()
    d.close()
