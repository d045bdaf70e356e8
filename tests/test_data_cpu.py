"""CPU test of the TFRecord reader (`ann3depth_b200/data.py`) against records encoded the way the reference's
writer does (`tools/data_tf_converter.py:41-53`: a `tf.train.Example` with raw float32 `image` / `depth` bytes
stored as value/255 - 0.5 and six int64 dims; framing length:u64 crc:u32 data crc:u32).  The encoder below is an
independent, minimal protobuf writer -- TensorFlow is not installed here."""
import struct

import numpy as np

from ann3depth_b200 import data


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(field, payload):                      # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _bytes_feature(b):
    return _ld(1, _ld(1, b))                  # Feature.bytes_list(1) { value(1): b }


def _int64_feature(v, packed):
    if packed:
        return _ld(3, _ld(1, _varint(v)))     # Int64List.value packed
    return _ld(3, _varint(1 << 3) + _varint(v))


def _example(features, packed):
    entries = b""
    for k, v in features.items():
        feat = _bytes_feature(v) if isinstance(v, bytes) else _int64_feature(v, packed)
        entries += _ld(1, _ld(1, k.encode()) + _ld(2, feat))      # map entry: key(1), value(2)
    return _ld(1, entries)                                         # Example.features(1)


def _write(path, records):
    # masked CRC-32C framing as TFRecordWriter writes it (the CRC itself is pinned to the RFC 3720 vectors in
    # tests/test_summary_cpu.py); the reader verifies it
    from ann3depth_b200.summary import masked_crc32c
    with open(path, "wb") as f:
        for r in records:
            head = struct.pack("<Q", len(r))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + r + struct.pack("<I", masked_crc32c(r)))


def test_tfrecord_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    recs, truth = [], []
    for i, packed in enumerate((True, False, True)):
        img8 = rng.integers(0, 256, size=(4, 6, 3), dtype=np.uint8)
        dep8 = rng.integers(0, 256, size=(2, 3, 1), dtype=np.uint8)
        img = img8.astype(np.float32) / 255. - .5                  # tools/data_tf_converter.py:36-37
        dep = dep8.astype(np.float32) / 255. - .5
        recs.append(_example({"image_height": 4, "image_width": 6, "image_channels": 3, "depth_height": 2,
                              "depth_width": 3, "depth_channels": 1, "image": img.tobytes(), "depth": dep.tobytes()},
                             packed))
        truth.append((img8, dep8))
    path = tmp_path / "train.tfrecords"
    _write(path, recs)
    got = [data.parse_example(r) for r in data.tfrecord_iterator(str(path))]
    assert len(got) == 3
    for e, (img8, dep8) in zip(got, truth):
        assert e["image_height"] == [4] and e["image_width"] == [6] and e["depth_height"] == [2] and e["depth_width"] == [3]
        im = np.frombuffer(e["image"], dtype=np.float32).reshape(4, 6, 3) + 0.5          # src/data.py:84
        dp = np.frombuffer(e["depth"], dtype=np.float32).reshape(2, 3, 1) + 0.5          # src/data.py:85
        assert np.allclose(im, img8 / 255.0, atol=1e-6) and np.allclose(dp, dep8 / 255.0, atol=1e-6)
        assert im.min() >= 0.0 and im.max() <= 1.0


def test_varint_multibyte_lengths(tmp_path):
    """image payloads are > 16 KB in practice: multi-byte varint lengths and 64-bit framing."""
    img = (np.arange(48 * 64 * 3, dtype=np.float32) % 255) / 255. - .5
    rec = _example({"image_height": 48, "image_width": 64, "image_channels": 3, "depth_height": 1, "depth_width": 1,
                    "depth_channels": 1, "image": img.tobytes(), "depth": np.zeros(1, np.float32).tobytes()}, True)
    path = tmp_path / "test.tfrecords"
    _write(path, [rec])
    (e,) = [data.parse_example(r) for r in data.tfrecord_iterator(str(path))]
    assert len(e["image"]) == 48 * 64 * 3 * 4
    assert np.array_equal(np.frombuffer(e["image"], dtype=np.float32), img)


def _dataset(tmp_path, n, hw=(4, 6), dhw=(2, 3)):
    """n records whose image is the constant (i+1)/255 and whose depth is (i+1)/510: a sample identifies its record."""
    import os
    recs = []
    for i in range(n):
        img = np.full((*hw, 3), (i + 1) / 255. - .5, dtype=np.float32)
        dep = np.full((*dhw, 1), (i + 1) / 510. - .5, dtype=np.float32)
        recs.append(_example({"image_height": hw[0], "image_width": hw[1], "image_channels": 3, "depth_height": dhw[0],
                              "depth_width": dhw[1], "depth_channels": 1, "image": img.tobytes(), "depth": dep.tobytes()},
                             i % 2 == 0))
    os.makedirs(tmp_path / "nyu", exist_ok=True)
    _write(tmp_path / "nyu" / "train.tfrecords", recs)
    return str(tmp_path)


def test_threaded_loader_reads_shuffled_records(tmp_path):
    """data.inputs (src/data.py:28-55): worker threads + shuffle pool; every sample is a record of the file (+0.5 applied,
    shapes from the record header), image and depth of a sample belong together, and all records turn up."""
    import torch
    n, B = 7, 4
    inp = data.inputs(_dataset(tmp_path, n), "nyu", B, device="cpu", seed=3, num_threads=2)
    assert tuple(inp.images.shape) == (B, 4, 6, 3) and tuple(inp.depths.shape) == (B, 2, 3, 1)
    assert [tuple(x) for x in data.tfrecord_index(inp.path)][0][0] == 12 and len(data.tfrecord_index(inp.path)) == n
    seen = set()
    for _ in range(12):
        im, dp = inp.next_batch()
        ids = torch.round(im[:, 0, 0, 0] * 255).long()
        assert torch.allclose(im, (ids.float() / 255).view(B, 1, 1, 1).expand_as(im), atol=1e-6)
        assert torch.allclose(dp, (ids.float() / 510).view(B, 1, 1, 1).expand_as(dp), atol=1e-6)
        assert int(ids.min()) >= 1 and int(ids.max()) <= n
        seen.update(ids.tolist())
    inp.close()
    assert seen == set(range(1, n + 1))


def test_threaded_loader_synthetic_host_and_errors(tmp_path):
    import pytest
    inp = data.inputs(str(tmp_path), "none", 2, device="cpu", synthetic="host", image_hw=(8, 8), depth_hw=(5, 7))
    a = inp.images.clone()
    inp.next_batch()
    assert not (a == inp.images).all()
    assert 0.0 <= float(inp.images.min()) and float(inp.images.max()) < 1.0
    assert 0.05 <= float(inp.depths.min()) and float(inp.depths.max()) < 1.0
    inp.close()
    root = _dataset(tmp_path, 3)
    path = tmp_path / "nyu" / "train.tfrecords"
    raw = bytearray(path.read_bytes())
    raw[3] ^= 0x40                                   # corrupt the first length field
    path.write_bytes(bytes(raw))
    with pytest.raises(IOError):
        data.inputs(root, "nyu", 2, device="cpu")
