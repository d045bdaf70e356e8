"""ann3depth_b200 -- B200-native (sm_100a) train/inference path for ann3depth's MSDN and DCNF models.

Python host code (this package) mirrors the reference's plug-in surface (`models.msdn`, `models.dcnf`,
`ann3depth.py`'s loop) and calls hand-written CUDA in `liba3d.so` through the ctypes C-ABI declared in
`include/a3d.h`.  PyTorch is used only for device memory, streams and process-group plumbing.
"""
__version__ = "0.1.0"
