"""Event-file writer and hooks of ann3depth_b200/summary.py (CPU): the files must be readable by TensorBoard itself."""
import json
import os
import struct

import pytest

from ann3depth_b200 import summary as S


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 test vectors
    assert S.crc32c(b"123456789") == 0xE3069283
    assert S.crc32c(bytes(32)) == 0x8A9136AA
    assert S.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert S.crc32c(bytes(range(32))) == 0x46DD794E
    # incremental == one shot
    assert S.crc32c(b"6789", S.crc32c(b"12345")) == S.crc32c(b"123456789")


def test_tfrecord_frame_layout():
    rec = S.tfrecord_frame(b"abc")
    n, = struct.unpack("<Q", rec[:8])
    assert n == 3 and rec[12:15] == b"abc" and len(rec) == 8 + 4 + 3 + 4
    assert struct.unpack("<I", rec[8:12])[0] == S.masked_crc32c(rec[:8])
    assert struct.unpack("<I", rec[15:])[0] == S.masked_crc32c(b"abc")


def test_tfrecord_frames_are_read_by_the_package_reader(tmp_path):
    from ann3depth_b200 import data
    p = tmp_path / "x.tfrecords"
    p.write_bytes(S.tfrecord_frame(b"first") + S.tfrecord_frame(b"") + S.tfrecord_frame(bytes(range(200))))
    assert list(data.tfrecord_iterator(str(p))) == [b"first", b"", bytes(range(200))]


def test_tfrecord_reader_rejects_corruption(tmp_path):
    from ann3depth_b200 import data
    good = S.tfrecord_frame(b"first") + S.tfrecord_frame(bytes(range(100)))
    cases = {"flipped length bit": bytearray(good), "flipped data bit": bytearray(good), "truncated": bytearray(good[:-3]),
             "half a header": bytearray(good + b"\x01\x02\x03")}
    cases["flipped length bit"][0] ^= 0x10
    cases["flipped data bit"][12 + 2] ^= 0x01
    for name, blob in cases.items():
        p = tmp_path / "bad.tfrecords"
        p.write_bytes(bytes(blob))
        with pytest.raises(IOError):
            list(data.tfrecord_iterator(str(p)))
    p.write_bytes(bytes(cases["flipped data bit"]))
    assert len(list(data.tfrecord_iterator(str(p), verify=None))) == 2        # framing intact, payload unchecked


def test_event_file_is_valid_for_tensorboard(tmp_path):
    loader_mod = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    w = S.EventWriter(str(tmp_path))
    w.add_scalars(150, {"loss/coarse_loss": 22218.68359375, "loss/fine_loss": 0.5}, wall_time=1234.5)
    w.add_scalars(300, {"global_step/sec": 977.25})
    w.add_scalars(2 ** 40, {"big/step": -1.0})
    w.close()
    events = list(loader_mod.EventFileLoader(w.path).Load())          # verifies both CRCs of every record
    assert events[0].file_version == "brain.Event:2"
    got = {}
    for e in events[1:]:
        for v in e.summary.value:
            # TensorBoard's loader upgrades simple_value to a rank-0 float tensor
            val = v.simple_value if v.HasField("simple_value") else v.tensor.float_val[0]
            got[(e.step, v.tag)] = val
    assert got == {(150, "loss/coarse_loss"): 22218.68359375, (150, "loss/fine_loss"): 0.5,
                   (300, "global_step/sec"): 977.25, (2 ** 40, "big/step"): -1.0}
    assert events[1].wall_time == 1234.5


def test_summary_tag_follows_the_reference():
    # src/tfhelper.py:151: '/'.join(name.split('/')[0:2]).split(':')[0]
    assert S.summary_tag("loss/coarse_loss/Mean:0") == "loss/coarse_loss"
    assert S.summary_tag("loss/fine_loss") == "loss/fine_loss"
    assert S.summary_tag("x:0") == "x"


class _FakeOp:
    def __init__(self):
        self.global_step = 0
        self.losses = {"loss/coarse_loss": 2.0, "loss/fine_loss": 3.0}
        self.graph_runs = self.traced_runs = 0

    def run(self, use_graph=True):
        self.global_step += 1
        self.losses = {k: v * 0.5 for k, v in self.losses.items()}
        if use_graph:
            self.graph_runs += 1
        else:
            self.traced_runs += 1


def test_summary_hook_writes_every_n_steps(tmp_path):
    loader_mod = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    op = _FakeOp()
    hook = S.SummaryHook(str(tmp_path), steps=4)
    written = []
    for _ in range(10):
        op.run()
        r = hook.after_run(op, images_per_step=32)
        if r is not None:
            written.append((op.global_step, r))
    hook.writer.close()
    assert [s for s, _ in written] == [4, 8]
    assert written[0][1]["loss/coarse_loss"] == 2.0 * 0.5 ** 4 and "global_step/sec" not in written[0][1]
    assert written[1][1]["images/sec"] == pytest.approx(written[1][1]["global_step/sec"] * 32)
    steps = [e.step for e in loader_mod.EventFileLoader(hook.writer.path).Load() if e.summary.value]
    assert steps == [4, 8]


def test_trace_hook_schedule(tmp_path, monkeypatch):
    """first step after a (re)start and every N-th step are traced (src/tfhelper.py:206,246-249)"""
    class _NoTimeline:
        records = []
        def __init__(self, ctx): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
        def per_op(self): return []
        def chrome_trace(self, **kw): return {"traceEvents": []}
    monkeypatch.setattr(S, "KernelTimeline", _NoTimeline)
    op = _FakeOp()
    op.net = type("N", (), {"ctx": None})()
    hook = S.TraceHook(str(tmp_path), every_step=5)
    traced = []
    for _ in range(12):
        before = op.traced_runs
        hook.run(op)
        if op.traced_runs > before:
            traced.append(op.global_step - 1)
    assert traced == [0, 5, 10]
    assert sorted(os.listdir(tmp_path)) == ["timeline-0.json", "timeline-10.json", "timeline-5.json"]
    assert json.load(open(tmp_path / "timeline-5.json")) == {"traceEvents": []}
