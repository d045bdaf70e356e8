"""Observability hooks of the train loop: TensorBoard scalar summaries and per-kernel step traces.

Counterpart of the reference's `create_summary_hook` (src/tfhelper.py:137-157: scalar summaries of every tensor
in GraphKeys.LOSSES every `--sumfreq` steps, written next to the checkpoints), of the `global_step/sec` scalar
MonitoredTrainingSession's StepCounterHook adds (src/ann3depth.py:113-125), and of `TraceHook`
(src/tfhelper.py:192-249: a FULL_TRACE of the first step after every (re)start and of every N-th step).

No TensorFlow here, so the event file is written directly: TFRecord framing (u64 length, masked CRC-32C of the
length, payload, masked CRC-32C of the payload) around hand-encoded `tensorflow.Event` protobufs

    Event   { double wall_time = 1; int64 step = 2; string file_version = 3; Summary summary = 5; }
    Summary { repeated Value value = 1; }     Value { string tag = 1; float simple_value = 2; }

which TensorBoard reads like any other run (tests/test_summary_cpu.py reads them back with the `tensorboard`
package).  A trace is the CUDA-event timeline of one un-graphed step -- one slice per liba3d call, on the stream it
was launched on -- in Chrome trace-event JSON, the format of TF's `timeline.generate_chrome_trace_format()`.
"""
from __future__ import annotations

import json
import os
import socket
import struct
import time

# ------------------------------------------------------------------------------------------ CRC-32C (Castagnoli)
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """TFRecord's masking: rotate right by 15 and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def tfrecord_frame(payload: bytes) -> bytes:
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload))


# ------------------------------------------------------------------------------------------ protobuf encoding
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field_bytes(num: int, payload: bytes) -> bytes:
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def encode_event(wall_time: float, step: int, scalars=None, file_version=None) -> bytes:
    ev = _varint((1 << 3) | 1) + struct.pack("<d", wall_time)
    if step:
        ev += _varint((2 << 3) | 0) + _varint(int(step))
    if file_version is not None:
        ev += _field_bytes(3, file_version.encode())
    if scalars:
        summary = b""
        for tag, value in scalars.items():
            val = _field_bytes(1, tag.encode()) + _varint((2 << 3) | 5) + struct.pack("<f", float(value))
            summary += _field_bytes(1, val)
        ev += _field_bytes(5, summary)
    return ev


class EventWriter:
    """`tf.summary.FileWriter(logdir)` for scalars: events.out.tfevents.<time>.<host> in logdir."""

    def __init__(self, logdir, suffix=""):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}{suffix}")
        self._f = open(self.path, "ab")
        self._f.write(tfrecord_frame(encode_event(time.time(), 0, file_version="brain.Event:2")))
        self._f.flush()

    def add_scalars(self, step, scalars, wall_time=None):
        self._f.write(tfrecord_frame(encode_event(time.time() if wall_time is None else wall_time, step, scalars)))

    def flush(self):
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.flush()
            self._f.close()


# ------------------------------------------------------------------------------------------ hooks
def summary_tag(name: str) -> str:
    """src/tfhelper.py:151: the first two path components of the tensor name."""
    return "/".join(name.split("/")[0:2]).split(":")[0]


class SummaryHook:
    """Scalar summaries of the op's `losses` (GraphKeys.LOSSES) every `steps` steps + global_step/sec."""

    def __init__(self, ckptdir, steps=150, writer=None):
        self.steps = max(int(steps), 1)
        self.writer = writer or EventWriter(ckptdir)
        self._last_time, self._last_step = time.time(), None

    def after_run(self, op, images_per_step=None):
        step = op.global_step
        if step % self.steps:
            return None
        scalars = {summary_tag(k): float(v) for k, v in op.losses.items()}      # float() synchronises on the loss
        now = time.time()
        if self._last_step is not None and now > self._last_time:
            rate = (step - self._last_step) / (now - self._last_time)
            scalars["global_step/sec"] = rate
            if images_per_step:
                scalars["images/sec"] = rate * images_per_step
        self._last_time, self._last_step = now, step
        self.writer.add_scalars(step, scalars)
        self.writer.flush()
        return scalars


class KernelTimeline:
    """Records one CUDA-event pair per liba3d call made while it is active (`with KernelTimeline(ctx) as t:`).

    The library object is shared by every Context of the process, so only one timeline can be active at a time.
    `chrome_trace()` gives Chrome trace-event JSON (one row per CUDA stream); `per_op()` a list of
    (sequence, entry point, milliseconds)."""

    _SKIP = ("a3d_last_error", "a3d_version", "a3d_launch_count", "a3d_sm_count", "a3d_conv2d_ws_bytes",
             "a3d_pairwise_ws_bytes", "a3d_create", "a3d_destroy", "a3d_comm_unique_id", "a3d_comm_init",
             "a3d_comm_destroy")

    def __init__(self, ctx):
        self.lib = ctx.lib
        self.records = []
        self._wrapped = {}

    def __enter__(self):
        import torch
        from . import _lib as L
        self._origin = torch.cuda.Event(enable_timing=True)
        self._origin.record(torch.cuda.current_stream())
        for name in L.signature_names():
            if name in self._SKIP or not hasattr(self.lib, name):
                continue
            fn = getattr(self.lib, name)
            self._wrapped[name] = fn

            def make(fn, name):
                def call(*a):
                    st = torch.cuda.current_stream()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    rc = fn(*a)
                    e1.record(st)
                    self.records.append((name, st.cuda_stream, e0, e1))
                    return rc
                return call
            setattr(self.lib, name, make(fn, name))
        return self

    def __exit__(self, *exc):
        import torch
        for name, fn in self._wrapped.items():
            setattr(self.lib, name, fn)
        self._wrapped = {}
        torch.cuda.synchronize()
        return False

    def per_op(self):
        return [(i, name, e0.elapsed_time(e1)) for i, (name, _, e0, e1) in enumerate(self.records)]

    def chrome_trace(self, pid=0, label="liba3d"):
        streams = {}
        events = [{"name": "process_name", "ph": "M", "pid": pid, "args": {"name": label}}]
        for name, st, e0, e1 in self.records:
            tid = streams.setdefault(st, len(streams))
            events.append({"name": name, "cat": "kernel", "ph": "X", "pid": pid, "tid": tid,
                           "ts": self._origin.elapsed_time(e0) * 1e3, "dur": max(e0.elapsed_time(e1), 0.0) * 1e3})
        for st, tid in streams.items():
            events.append({"name": "thread_name", "ph": "M", "pid": pid, "tid": tid,
                           "args": {"name": f"cuda stream {st:#x}"}})
        return {"traceEvents": events, "displayTimeUnit": "us"}


class TraceHook:
    """src/tfhelper.py:192-249: trace the first step after every (re)start and every `every_step`-th step."""

    def __init__(self, ckptdir, every_step=50, writer=None):
        self.ckptdir, self.every_step, self._trace, self.writer = ckptdir, max(int(every_step), 1), True, writer

    def wants_trace(self):
        return self._trace

    def run(self, op):
        """One `session.run(model_op)`; traced (un-graphed, timed per liba3d call) when a trace is due."""
        if not self._trace:
            op.run()
        else:
            self._trace = False
            step = op.global_step
            with KernelTimeline(op.net.ctx) as tl:
                op.run(use_graph=False)
            os.makedirs(self.ckptdir, exist_ok=True)
            with open(os.path.join(self.ckptdir, f"timeline-{step}.json"), "w") as f:
                json.dump(tl.chrome_trace(label=f"step {step}"), f)
            if self.writer is not None:
                self.writer.add_scalars(step + 1, {"trace/liba3d_calls": len(tl.records),
                                                   "trace/sum_of_calls_ms": sum(ms for _, _, ms in tl.per_op())})
        if not (op.global_step % self.every_step):          # global_step is already the NEXT step's index
            self._trace = True
