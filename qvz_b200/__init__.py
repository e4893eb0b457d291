"""qvz_b200 -- B200-native compression front end of qvz (k-means, conditional counts, quantize walk).

The product path is the CUDA library qvz_b200/csrc/libqvz_gpu.so behind include/qvz_gpu.h; the
Python in this package is only the ctypes binding and the torch.distributed plumbing around it.
"""
__all__ = ["lib", "hostlib", "dist", "synth"]
