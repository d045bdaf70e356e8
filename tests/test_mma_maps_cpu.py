"""Index algebra of csrc/dense_stream.cu replayed on the CPU (tools/mma_map_check.py): the per-thread addressing of
the mma.sync weight-streaming kernels and of the tiled single-filter convolution against a direct product."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mma_map_check", os.path.join(ROOT, "tools", "mma_map_check.py"))
M = importlib.util.module_from_spec(spec)
spec.loader.exec_module(M)


def test_dense_forward_fragment_map():
    M.check_fwd(M=20, N=70, K=64)          # ragged batch and row count: clamped rows, masked stores


def test_dense_dgrad_fragment_map():
    M.check_dgrad(M=20, N=38, K=64)        # N % 16 != 0: zero-filled dy tail, clamped weight rows


def test_single_filter_conv_tile_map():
    M.check_k1(H=7, W=9, R=5, S=5, pt=2, pl=2, band=3)      # three bands incl. a ragged last one, SAME padding
