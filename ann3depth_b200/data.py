"""Input contract of the reference (`src/data.py:28-55,70-86`): `inputs(datadir, dataset, batch_size)`
returns `(images [B,480,640,3], depths [B,h,w,1])`, float32 in [0,1], NHWC -- here as CUDA tensors that
are refilled in place by `next_batch()` (the reference's queue runners do the same behind the graph).

Two sources:
  * synthetic (default when `<datadir>/<dataset>/<split>.tfrecords` is absent): U[0,1) images and
    U[0.05,1) depths, the shapes of BASELINE.json's configs;
  * TFRecords written by the reference's `tools/data_tf_converter.py:41-53`: `tf.train.Example`s with
    raw float32 `image` / `depth` bytes stored as value/255 - 0.5 (read back with +0.5, src/data.py:84-85)
    and six int64 dims.  Parsed with a minimal pure-Python TFRecord + protobuf reader (no TensorFlow).
"""
from __future__ import annotations

import os
import struct

import torch


def _read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_fields(buf):
    """Yield (field_number, wire_type, value) of one protobuf message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def parse_example(record: bytes) -> dict:
    """tf.train.Example -> {name: bytes | [int]} for BytesList / Int64List features."""
    out = {}
    for f, _, features in _parse_fields(record):                 # Example.features = 1
        if f != 1:
            continue
        for f2, _, entry in _parse_fields(features):             # Features.feature (map entry) = 1
            if f2 != 1:
                continue
            key, val = None, None
            for f3, _, v in _parse_fields(entry):
                if f3 == 1:
                    key = bytes(v).decode()
                elif f3 == 2:
                    for f4, _, lst in _parse_fields(v):           # Feature: bytes_list=1, int64_list=3
                        if f4 == 1:
                            val = [bytes(x) for f5, _, x in _parse_fields(lst) if f5 == 1]
                            val = val[0] if len(val) == 1 else val
                        elif f4 == 3:
                            ints = []
                            for f5, wt5, x in _parse_fields(lst):
                                if wt5 == 0:
                                    ints.append(x)
                                else:                              # packed
                                    p = 0
                                    while p < len(x):
                                        iv, p = _read_varint(x, p)
                                        ints.append(iv)
                            val = ints
            out[key] = val
    return out


def tfrecord_iterator(path, verify="length"):
    """Yields the raw records of a TFRecord file: u64 length, masked CRC-32C of the length, data, masked CRC-32C of the
    data (the framing `tools/data_tf_converter.py:41-53` writes through `tf.python_io.TFRecordWriter`).
    verify = "length" (default): the length CRC of every record is checked -- a corrupted length would otherwise send
    the reader to a garbage offset -- and the data CRC of records up to 64 KB; "full": every data CRC too (pure-Python
    CRC-32C, about 1 s per MB); None: nothing.  A truncated record raises instead of ending the iteration silently."""
    from .summary import masked_crc32c
    with open(path, "rb") as f:
        index = 0
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError(f"{path}: truncated header of record {index}")
            (n,) = struct.unpack("<Q", head[:8])
            if verify and struct.unpack("<I", head[8:12])[0] != masked_crc32c(head[:8]):
                raise IOError(f"{path}: corrupted length of record {index}")
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise IOError(f"{path}: truncated record {index} ({len(data)} of {n} bytes)")
            if (verify == "full" or (verify and n <= 65536)) and struct.unpack("<I", tail)[0] != masked_crc32c(data):
                raise IOError(f"{path}: corrupted data of record {index}")
            index += 1
            yield data


class Inputs:
    """`images`, `depths`: persistent CUDA buffers; `next_batch()` refills them in place."""

    def __init__(self, datadir, dataset, batch_size=32, train_or_test="train", device="cuda", seed=0,
                 image_hw=(480, 640), depth_hw=(55, 73)):
        self.B, self.device = batch_size, torch.device(device)
        self.path = os.path.join(datadir, dataset, f"{train_or_test}.tfrecords")      # src/data.py:58-59
        self.gen = torch.Generator().manual_seed(seed)
        self.records = None
        if os.path.exists(self.path):
            self.records = [parse_example(r) for r in tfrecord_iterator(self.path)]
            e = self.records[0]
            image_hw = (e["image_height"][0], e["image_width"][0])
            depth_hw = (e["depth_height"][0], e["depth_width"][0])
        self.image_hw, self.depth_hw = image_hw, depth_hw
        self._h_images = torch.empty(batch_size, *image_hw, 3).pin_memory()
        self._h_depths = torch.empty(batch_size, *depth_hw, 1).pin_memory()
        self.images = torch.empty(batch_size, *image_hw, 3, device=self.device)
        self.depths = torch.empty(batch_size, *depth_hw, 1, device=self.device)
        self.next_batch()

    def next_batch(self):
        if self.records is None:
            self._h_images.copy_(torch.rand(self._h_images.shape, generator=self.gen))
            self._h_depths.copy_(torch.rand(self._h_depths.shape, generator=self.gen) * 0.95 + 0.05)
        else:
            idx = torch.randint(len(self.records), (self.B,), generator=self.gen).tolist()   # shuffle_batch
            for i, j in enumerate(idx):
                e = self.records[j]
                im = torch.frombuffer(bytearray(e["image"]), dtype=torch.float32) + 0.5
                dp = torch.frombuffer(bytearray(e["depth"]), dtype=torch.float32) + 0.5
                self._h_images[i] = im.view(*self.image_hw, 3)
                self._h_depths[i] = dp.view(*self.depth_hw, 1)
        self.images.copy_(self._h_images, non_blocking=True)
        self.depths.copy_(self._h_depths, non_blocking=True)
        return self.images, self.depths


def inputs(datadir, dataset, batch_size=32, train_or_test="train", epochs=None, **kw):
    """Same call as the reference's `data.inputs` (src/data.py:28-29); returns an `Inputs` object whose
    `.images` / `.depths` are the tensors to hand to `models.<model>(images, depths)`."""
    return Inputs(datadir, dataset, batch_size, train_or_test, **kw)
