"""dct_b200 -- B200-native 8x8 DCT + quantization hot path behind the reference's C interface.

    dct_b200.build   compiles dct_b200/libdct_cuda.so (nvcc, sm_100a) in-tree
    dct_b200.api     ctypes mirror of include/dct.h, include/quantization.h, include/dct_cuda.h

`api` is imported lazily so that `python -m dct_b200.build` works before the library exists.
"""
__all__ = ["api", "build"]


def __getattr__(name):
    if name in ("api", "build"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
