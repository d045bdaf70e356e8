"""The train driver end to end on a GPU (`python -m ann3depth_b200.ann3depth ...`, the counterpart of
`python src/ann3depth.py ...` / `make train`): step limit, periodic + final checkpoints, resume, signal-driven graceful
stop with the reference's exit-code convention (exit code = number of the signal, src/ann3depth.py:129,
src/tfhelper.py:160-189), and the `--timeout` alarm that is armed only for cluster jobs (src/ann3depth.py:107-109)."""
import os
import signal
import subprocess
import sys
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cmd(ckptdir, *extra):
    return [sys.executable, "-u", "-m", "ann3depth_b200.ann3depth", "nyu", "--batchsize", "2", "--ckptdir", str(ckptdir),
            "--datadir", str(ckptdir / "no_data"), "--sumfreq", "1", *extra]


def _run(ckptdir, *extra, timeout=300):
    return subprocess.run(_cmd(ckptdir, *extra), cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def _ckpt(ckptdir, run_id="msdn"):
    return torch.load(os.path.join(ckptdir, run_id, "model.ckpt"), map_location="cpu")


def test_main_step_limit_checkpoint_and_resume(tmp_path):
    r = _run(tmp_path, "--steps", "3")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Session stopped." in r.stdout and "step 3 losses" in r.stdout
    st = _ckpt(tmp_path)
    assert st["global_step"] == 3 and st["format"] == 2
    assert st["adam_t"]["CoarseDense"] == 3 and st["adam_t"]["FineA"] == 0
    # TF variable names and TF layouts (HWIO / [in, out]), not the packed arena
    assert tuple(st["variables"]["coarse/conv/conv2d_0/kernel"].shape) == (11, 11, 3, 96)
    assert tuple(st["variables"]["coarse/dense/dense_1/kernel"].shape) == (4096, 4070)
    assert float(st["adam_m"]["coarse/dense/dense_0/kernel"].abs().max()) > 0          # three steps of momentum
    assert os.path.exists(tmp_path / "msdn" / "timeline-0.json")                       # TraceHook: first step traced
    # restart at the limit: nothing to do, the state survives restore -> save bit for bit
    r2 = _run(tmp_path, "--steps", "3")
    assert r2.returncode == 0 and "Restored checkpoint at global step 3" in r2.stdout
    st2 = _ckpt(tmp_path)
    assert st2["global_step"] == 3 and st2["adam_t"] == st["adam_t"]
    for k in ("variables", "adam_m", "adam_v"):
        for n in st[k]:
            assert torch.equal(st[k][n], st2[k][n]), (k, n)
    # resume and continue to step 5
    r3 = _run(tmp_path, "--steps", "5", "--id", "")
    assert r3.returncode == 0 and "Restored checkpoint at global step 3" in r3.stdout and "step 5 losses" in r3.stdout
    st3 = _ckpt(tmp_path)
    assert st3["global_step"] == 5 and st3["adam_t"]["CoarseDense"] == 5
    # the reference's Adam (beta2 = 1) never moves a weight; its first-moment slots do move
    assert torch.equal(st3["variables"]["coarse/dense/dense_0/kernel"], st["variables"]["coarse/dense/dense_0/kernel"])
    assert not torch.equal(st3["adam_m"]["coarse/dense/dense_0/kernel"], st["adam_m"]["coarse/dense/dense_0/kernel"])


def _wait_for(proc, needle, timeout=240):
    t0, seen = time.time(), ""
    while time.time() - t0 < timeout:
        line = proc.stdout.readline()
        if not line:
            if proc.poll() is not None:
                break
            continue
        seen += line
        if needle in line:
            return seen
    raise AssertionError(f"{needle!r} not seen; output so far:\n{seen[-3000:]}")


def test_signal_stops_gracefully_with_signal_exit_code(tmp_path):
    proc = subprocess.Popen(_cmd(tmp_path, "--steps", "100000000", "--id", "sig"), cwd=ROOT, stdout=subprocess.PIPE,
                            stderr=subprocess.STDOUT, text=True)
    try:
        _wait_for(proc, "step 5 losses")
        proc.send_signal(signal.SIGUSR1)
        out, _ = proc.communicate(timeout=120)
    finally:
        if proc.poll() is None:
            proc.kill()
    assert proc.returncode == signal.SIGUSR1, (proc.returncode, out[-2000:])         # sys.exit(signal_received)
    assert "Session stopped." in out
    st = _ckpt(tmp_path, "msdn_sig")                                                   # final checkpoint was written
    assert st["global_step"] >= 5


def test_timeout_alarm_only_for_cluster_jobs(tmp_path):
    t0 = time.time()
    r = _run(tmp_path, "--steps", "100000000", "--id", "alarm", "--job-name", "worker", "--timeout", "20")
    assert r.returncode == signal.SIGALRM, (r.returncode, r.stdout[-2000:])
    assert "Starting alarm: 20 s timeout." in r.stdout and time.time() - t0 >= 20
    assert _ckpt(tmp_path, "msdn_alarm")["global_step"] > 0
    # a plain local run is not alarmed (src/ann3depth.py:107-109)
    r = _run(tmp_path, "--steps", "2", "--id", "local", "--timeout", "1")
    assert r.returncode == 0 and "Starting alarm" not in r.stdout
