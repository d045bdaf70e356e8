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


def test_persistent_kernel_barrier_protocol():
    """tools/persist_protocol_check.py: randomised schedules of the producer / issuer / epilogue roles of
    csrc/tc_persist.cuh under an mbarrier model -- no deadlock, no stage or accumulator overwritten while in use"""
    spec2 = importlib.util.spec_from_file_location("persist_protocol_check",
                                                   os.path.join(ROOT, "tools", "persist_protocol_check.py"))
    P = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(P)
    for nstage in (2, 3, 4):
        for tiles, nkb in ((1, 1), (2, 9), (5, 3), (8, 2)):
            P.simulate(nstage, tiles, nkb, seed=nstage * 31 + tiles)
    for cl, nstage, nkb in ((2, 2, 9), (4, 2, 5), (4, 3, 50), (2, 3, 1)):
        P.simulate_mcast(cl, nstage, nkb, seed=cl * 7 + nkb)
