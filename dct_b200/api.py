"""Host-side mirror of the reference's interface, over libdct_cuda's C ABI (ctypes).

The reference (erkinov-wtf/dct) is compiled C; its public surface for this path is
include/dct.h and include/quantization.h.  This module binds the very symbols a C caller
links against -- same names, same argument meaning -- plus the plane calls of
include/dct_cuda.h, so that the parity tests read like the reference's own tests:

    ctx  = dct_init(8);  qctx = quant_init(8, 50, 0)          # tests/test_entropy.c:283-287
    coeffs = dct_forward(ctx, block)                           # tests/test_entropy.c:311
    q = quantize(qctx, coeffs, variance)                       # tests/test_entropy.c:316

and, for whole planes (the throughput path):

    plan = Plan(ctx, qctx)
    coef = plan.fwd_quant(pixels)                              # numpy (host) or torch cuda tensor
    rec  = plan.dequant_idct(coef, W, H)

There is no fallback of any kind: if libdct_cuda.so is not built, importing this module
raises; if no CUDA device is usable, Plan() and the per-block compute calls fail.
torch is used only to hold device memory and streams (data_ptr / cuda_stream), never to compute.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdct_cuda.so")

NATURAL, ZIGZAG = 0, 1

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is not built: run `python -m dct_b200.build` (there is no CPU fallback)")
_lib = C.CDLL(LIB_PATH)

_PP_D = C.POINTER(C.POINTER(C.c_double))
_PP_I = C.POINTER(C.POINTER(C.c_int))


class DCTContext(C.Structure):  # include/dct.h
    _fields_ = [("block_size", C.c_int), ("dct_matrix", _PP_D), ("transposed_dct", _PP_D)]


class QuantContext(C.Structure):  # include/quantization.h
    _fields_ = [("block_size", C.c_int), ("quality", C.c_int), ("quant_matrix", _PP_D),
                ("dequant_matrix", _PP_D), ("adaptive", C.c_int)]


class Stats(C.Structure):  # dct_cuda_stats
    _fields_ = [("blocks", C.c_uint64), ("replayed_blocks", C.c_uint64), ("near_ties", C.c_uint64),
                ("saturated", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class PlaneDesc(C.Structure):  # dct_cuda_plane
    _fields_ = [("plan", C.c_void_p), ("pixels_in", C.c_void_p), ("pixels_out", C.c_void_p),
                ("pitch", C.c_size_t), ("width", C.c_int), ("height", C.c_int), ("coef", C.c_void_p),
                ("variance", C.c_void_p)]


def _sig(name, restype, *argtypes):
    fn = getattr(_lib, name)
    fn.restype, fn.argtypes = restype, list(argtypes)
    return fn


_P_DCT, _P_Q = C.POINTER(DCTContext), C.POINTER(QuantContext)
_dct_init = _sig("dct_init", _P_DCT, C.c_int)
_dct_free = _sig("dct_free", None, _P_DCT)
_dct_forward = _sig("dct_forward", None, _P_DCT, _PP_D, _PP_D)
_dct_inverse = _sig("dct_inverse", None, _P_DCT, _PP_D, _PP_D)
_create_block = _sig("create_block_from_pixels", _PP_D, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int)
_copy_block = _sig("copy_block_to_coefficients", None, _PP_D, _PP_I, C.c_int)
_quant_init = _sig("quant_init", _P_Q, C.c_int, C.c_int, C.c_int)
_quant_free = _sig("quant_free", None, _P_Q)
_gen_q = _sig("generate_quant_matrix", _PP_D, C.c_int, C.c_int)
_gen_r = _sig("generate_dequant_matrix", _PP_D, _PP_D, C.c_int)
_quantize = _sig("quantize", None, _P_Q, _PP_D, _PP_I, C.c_double)
_dequantize = _sig("dequantize", None, _P_Q, _PP_I, _PP_D, C.c_double)
_variance = _sig("calculate_block_variance", C.c_double, _PP_D, C.c_int)
_adjust = _sig("adjust_matrix_for_block", _PP_D, _P_Q, C.c_double, C.c_int)

_last_error = _sig("dct_cuda_last_error", C.c_char_p)
_device_count = _sig("dct_cuda_device_count", C.c_int)
_plan_create = _sig("dct_cuda_plan_create", C.c_void_p, _P_DCT, _P_Q, C.c_int)
_plan_refresh = _sig("dct_cuda_plan_refresh", C.c_int, C.c_void_p)
_plan_destroy = _sig("dct_cuda_plan_destroy", None, C.c_void_p)
_fwd_dev = _sig("dct_cuda_fwd_quant_u8_dev", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)
_inv_dev = _sig("dct_cuda_dequant_idct_u8_dev", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
_fwd_planes = _sig("dct_cuda_fwd_quant_planes_dev", C.c_int, C.POINTER(PlaneDesc), C.c_int, C.c_int, C.c_void_p)
_inv_planes = _sig("dct_cuda_dequant_idct_planes_dev", C.c_int, C.POINTER(PlaneDesc), C.c_int, C.c_int, C.c_void_p)
_fwd_host = _sig("dct_cuda_fwd_quant_u8", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                 C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Stats))
_inv_host = _sig("dct_cuda_dequant_idct_u8", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                 C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Stats))
_fwd_f32_dev = _sig("dct_cuda_fwd_quant_f32_dev", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                    C.c_void_p, C.c_int, C.c_void_p)
_fwd_f32_host = _sig("dct_cuda_fwd_quant_f32", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                     C.c_void_p, C.c_int, C.POINTER(Stats))
_fwd_host_async = _sig("dct_cuda_fwd_quant_u8_async", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                       C.c_void_p, C.c_int, C.c_void_p)
_inv_host_async = _sig("dct_cuda_dequant_idct_u8_async", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                       C.c_void_p, C.c_void_p, C.c_size_t)
_plan_wait = _sig("dct_cuda_plan_wait", C.c_int, C.c_void_p, C.POINTER(Stats))
_fwd_multi = _sig("dct_cuda_fwd_quant_u8_multi", C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t,
                  C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Stats))
_inv_multi = _sig("dct_cuda_dequant_idct_u8_multi", C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int,
                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Stats))
_stats_fetch = _sig("dct_cuda_stats_fetch", C.c_int, C.c_void_p, C.POINTER(Stats), C.c_void_p)
_skip_replay = _sig("dct_cuda_plan_debug_skip_replay", C.c_int, C.c_void_p, C.c_int)
_profile = _sig("dct_cuda_plan_profile", C.c_int, C.c_void_p, C.c_int)
_profile_fetch = _sig("dct_cuda_profile_fetch", C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                      C.POINTER(C.c_double), C.POINTER(C.c_int))
_rle_count = _sig("dct_cuda_rle_count_dev", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                  C.POINTER(C.c_uint64), C.c_void_p)
_rle_emit = _sig("dct_cuda_rle_emit_dev", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                 C.c_void_p)
_rec_to_block = _sig("dct_cuda_record_to_block", None, C.POINTER(C.c_int16), C.c_int, _PP_I)
_block_to_rec = _sig("dct_cuda_block_to_record", None, _PP_I, C.c_int, C.POINTER(C.c_int16))
_fwd_edge = _sig("dct_cuda_fwd_quant_u8_edge", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                 C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Stats))
_inv_edge = _sig("dct_cuda_dequant_idct_u8_edge", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                 C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Stats))
_pad_edges = _sig("dct_cuda_pad_edges_dev", C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_void_p)


_kernel_launches = _sig("dct_cuda_plan_kernel_launches", C.c_uint64, C.c_void_p)
_fits_i8 = _sig("dct_cuda_plan_records_fit_i8", C.c_int, C.c_void_p)
_fwd_i8 = _sig("dct_cuda_fwd_quant_u8_i8", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
               C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Stats))
_inv_i8 = _sig("dct_cuda_dequant_idct_i8_u8", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
               C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Stats))
_fwd_i8_async = _sig("dct_cuda_fwd_quant_u8_i8_async", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                     C.c_void_p, C.c_int, C.c_void_p)
_inv_i8_async = _sig("dct_cuda_dequant_idct_i8_u8_async", C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                     C.c_void_p, C.c_void_p, C.c_size_t)
_rec8_to_block = _sig("dct_cuda_record8_to_block", None, C.POINTER(C.c_int8), C.c_int, _PP_I)
_peer_share = _sig("dct_cuda_peer_default_share", C.c_float, C.c_void_p, C.c_int, C.c_int)
_fwd_peer = _sig("dct_cuda_fwd_quant_u8_peer", C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                 C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_void_p)
_inv_peer = _sig("dct_cuda_dequant_idct_u8_peer", C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int,
                 C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_float), C.c_void_p)


class Frame420(C.Structure):  # dct_cuda_frame420
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("y_width", C.c_int), ("y_height", C.c_int),
                ("c_width", C.c_int), ("c_height", C.c_int)]


_frame_geometry = _sig("dct_cuda_frame420_geometry", None, C.c_int, C.c_int, C.POINTER(Frame420))
_rgb_to_ycc = _sig("dct_cuda_rgb_to_ycbcr420_dev", C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Frame420),
                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
_ycc_to_rgb = _sig("dct_cuda_ycbcr420_to_rgb_dev", C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                   C.c_size_t, C.POINTER(Frame420), C.c_void_p, C.c_size_t, C.c_void_p)
_encode_rgb = _sig("dct_cuda_encode_rgb420", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Stats))
_decode_rgb = _sig("dct_cuda_decode_rgb420", C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Stats))
_host_alloc = _sig("dct_cuda_host_alloc", C.c_void_p, C.c_size_t)
_host_free = _sig("dct_cuda_host_free", None, C.c_void_p)


class DctCudaError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise DctCudaError(f"libdct_cuda error {rc}: {_last_error().decode(errors='replace')}")


def device_count() -> int:
    return int(_device_count())


# ---- ragged <-> numpy (the reference's blocks are row-pointer arrays, include/utils.h) --------
class _Ragged:
    """A caller-owned ragged n x n array: independent row buffers + a pointer table."""

    def __init__(self, n, ctype, init=None):
        self.n, self.ctype = n, ctype
        self.rows = [(ctype * n)() for _ in range(n)]
        if init is not None:
            a = np.asarray(init)
            for i in range(n):
                for j in range(n):
                    self.rows[i][j] = a[i, j].item()
        self.table = (C.POINTER(ctype) * n)(*[C.cast(r, C.POINTER(ctype)) for r in self.rows])

    @property
    def ptr(self):
        return C.cast(self.table, C.POINTER(C.POINTER(self.ctype)))

    def numpy(self):
        dt = np.float64 if self.ctype is C.c_double else np.int32
        return np.array([[self.rows[i][j] for j in range(self.n)] for i in range(self.n)], dtype=dt)


def _take_ragged(pp, n):
    """Copies a library-returned ragged double array and releases it the way free_array does."""
    out = np.array([[pp[i][j] for j in range(n)] for i in range(n)], dtype=np.float64)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i in range(n):
        libc.free(C.cast(pp[i], C.c_void_p))
    libc.free(C.cast(pp, C.c_void_p))
    return out


def matrix_of(pp, n):
    return np.array([[pp[i][j] for j in range(n)] for i in range(n)], dtype=np.float64)


# ---- include/dct.h ---------------------------------------------------------------------------
def dct_init(block_size: int):
    return _dct_init(int(block_size))


def dct_free(ctx):
    _dct_free(ctx)


def dct_forward(ctx, block):
    n = ctx.contents.block_size
    a, b = _Ragged(n, C.c_double, block), _Ragged(n, C.c_double)
    _dct_forward(ctx, a.ptr, b.ptr)
    return b.numpy()


def dct_inverse(ctx, block):
    n = ctx.contents.block_size
    a, b = _Ragged(n, C.c_double, block), _Ragged(n, C.c_double)
    _dct_inverse(ctx, a.ptr, b.ptr)
    return b.numpy()


def create_block_from_pixels(pixels, width, row_start, col_start, block_size):
    px = np.ascontiguousarray(pixels, dtype=np.uint8)
    pp = _create_block(px.ctypes.data_as(C.POINTER(C.c_ubyte)), width, row_start, col_start, block_size)
    return _take_ragged(pp, block_size)


def copy_block_to_coefficients(block):
    n = np.asarray(block).shape[0]
    a, b = _Ragged(n, C.c_double, block), _Ragged(n, C.c_int)
    _copy_block(a.ptr, b.ptr, n)
    return b.numpy()


# ---- include/quantization.h ------------------------------------------------------------------
def quant_init(block_size: int, quality: int, adaptive: int):
    return _quant_init(int(block_size), int(quality), int(adaptive))


def quant_free(ctx):
    _quant_free(ctx)


def generate_quant_matrix(block_size, quality):
    return _take_ragged(_gen_q(block_size, quality), block_size)


def generate_dequant_matrix(quant_matrix):
    n = np.asarray(quant_matrix).shape[0]
    a = _Ragged(n, C.c_double, quant_matrix)
    return _take_ragged(_gen_r(a.ptr, n), n)


def quantize(ctx, dct_coeffs, block_variance=0.0):
    n = ctx.contents.block_size
    a, b = _Ragged(n, C.c_double, dct_coeffs), _Ragged(n, C.c_int)
    _quantize(ctx, a.ptr, b.ptr, float(block_variance))
    return b.numpy()


def dequantize(ctx, quant_coeffs, block_variance=0.0):
    n = ctx.contents.block_size
    a, b = _Ragged(n, C.c_int, quant_coeffs), _Ragged(n, C.c_double)
    _dequantize(ctx, a.ptr, b.ptr, float(block_variance))
    return b.numpy()


def calculate_block_variance(block):
    n = np.asarray(block).shape[0]
    return float(_variance(_Ragged(n, C.c_double, block).ptr, n))


def adjust_matrix_for_block(ctx, variance, is_quantize):
    return _take_ragged(_adjust(ctx, float(variance), int(is_quantize)), ctx.contents.block_size)


def set_quant_table(qctx, table):
    """Overwrites ctx->quant_matrix (and the matching 1/Q), the way a chroma table gets in."""
    n = qctx.contents.block_size
    t = np.asarray(table, dtype=np.float64)
    for i in range(n):
        for j in range(n):
            qctx.contents.quant_matrix[i][j] = t[i, j]
            qctx.contents.dequant_matrix[i][j] = 1.0 / t[i, j]


def record8_to_block(record8, layout=NATURAL):
    rec = np.ascontiguousarray(record8, dtype=np.int8)
    blk = _Ragged(8, C.c_int)
    _rec8_to_block(rec.ctypes.data_as(C.POINTER(C.c_int8)), layout, blk.ptr)
    return blk.numpy()


def record_to_block(record, layout=NATURAL):
    rec = np.ascontiguousarray(record, dtype=np.int16)
    b = _Ragged(8, C.c_int)
    _rec_to_block(rec.ctypes.data_as(C.POINTER(C.c_int16)), layout, b.ptr)
    return b.numpy()


def block_to_record(block, layout=NATURAL):
    rec = np.zeros(64, dtype=np.int16)
    _block_to_rec(_Ragged(8, C.c_int, block).ptr, layout, rec.ctypes.data_as(C.POINTER(C.c_int16)))
    return rec


# ---- include/dct_cuda.h ------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _stream_ptr(stream, device=None):
    """The CUDA stream a call is queued on: `stream` if given, else torch's current stream OF `device` (the plan's
    or the tensor's GPU -- not of whatever device torch currently has selected)."""
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(getattr(stream, "cuda_stream", stream))


def _on_device(tensor, device, what):
    if tensor is not None and tensor.device.index != device:
        raise DctCudaError(f"{what} lives on cuda:{tensor.device.index}, the plan on cuda:{device}")


class Plan:
    """dct_cuda_plan: a (DCTContext, QuantContext) pair bound to one GPU."""

    def __init__(self, dct_ctx, quant_ctx, device: int = 0):
        self.dct_ctx, self.quant_ctx, self.device = dct_ctx, quant_ctx, device
        self.adaptive = int(quant_ctx.contents.adaptive)
        self.n = int(dct_ctx.contents.block_size)      # records hold n*n coefficients
        self._h = _plan_create(dct_ctx, quant_ctx, device)
        if not self._h:
            raise DctCudaError(_last_error().decode(errors="replace"))

    def close(self):
        if self._h:
            _plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh(self):
        _check(_plan_refresh(self._h))
        self.adaptive = int(self.quant_ctx.contents.adaptive)

    # -- host arrays (numpy): pipelined H2D / kernels / D2H inside the library --------------
    def fwd_quant(self, pixels, layout=NATURAL, coef_out=None, var_out=None, want_stats=False):
        if _is_torch(pixels):
            return self.fwd_quant_dev(pixels, layout, coef_out, var_out)
        assert pixels.dtype == np.uint8 and pixels.ndim == 2
        H, W = pixels.shape
        n = self.n
        nb = (H // n) * (W // n)
        if pixels.size == 0 and H % n == 0 and W % n == 0:   # empty plane: nothing to queue
            out = (np.empty((0, n * n), np.int16), np.empty(0)) if self.adaptive else np.empty((0, n * n), np.int16)
            return (out, Stats().as_dict()) if want_stats else out
        assert pixels.strides[1] == 1
        coef = coef_out if coef_out is not None else np.empty((nb, n * n), dtype=np.int16)
        var = var_out if var_out is not None else (np.empty(nb, dtype=np.float64) if self.adaptive else None)
        st = Stats()
        _check(_fwd_host(self._h, pixels.ctypes.data, pixels.strides[0], W, H, coef.ctypes.data, layout,
                         var.ctypes.data if var is not None else None, C.byref(st)))
        out = (coef, var) if self.adaptive else coef
        return (out, st.as_dict()) if want_stats else out

    def fwd_quant_f32(self, pixels, layout=NATURAL, want_stats=False):
        """Float pixel tiles (numpy float32 on the host, or a torch cuda float32 tensor): block = p - 128."""
        if _is_torch(pixels):
            import torch
            assert pixels.is_cuda and pixels.dtype == torch.float32 and pixels.dim() == 2 and pixels.stride(1) == 1
            H, W = pixels.shape
            coef = torch.empty(((H // 8) * (W // 8), 64), dtype=torch.int16, device=pixels.device)
            _on_device(pixels, self.device, "pixels")
            _check(_fwd_f32_dev(self._h, pixels.data_ptr(), pixels.stride(0) * 4, W, H, coef.data_ptr(), layout,
                                _stream_ptr(None, self.device)))
            return coef
        assert pixels.dtype == np.float32 and pixels.ndim == 2 and pixels.strides[1] == 4
        H, W = pixels.shape
        coef = np.empty(((H // 8) * (W // 8), 64), dtype=np.int16)
        st = Stats()
        _check(_fwd_f32_host(self._h, pixels.ctypes.data, pixels.strides[0], W, H, coef.ctypes.data, layout, C.byref(st)))
        return (coef, st.as_dict()) if want_stats else coef

    def dequant_idct(self, coef, W, H, layout=NATURAL, var=None, pixels_out=None, want_stats=False):
        if _is_torch(coef):
            return self.dequant_idct_dev(coef, W, H, layout, var, pixels_out)
        coef = np.ascontiguousarray(coef, dtype=np.int16)
        px = pixels_out if pixels_out is not None else np.empty((H, W), dtype=np.uint8)
        if var is not None:
            var = np.ascontiguousarray(var, dtype=np.float64)
        st = Stats()
        _check(_inv_host(self._h, coef.ctypes.data, W, H, layout, var.ctypes.data if var is not None else None,
                         px.ctypes.data, px.strides[0], C.byref(st)))
        return (px, st.as_dict()) if want_stats else px

    # -- raw host pointers (pinned torch tensors in bench.py) ---------------------------------
    def fwd_quant_ptr(self, px_ptr, pitch, W, H, coef_ptr, layout=NATURAL, var_ptr=None):
        _check(_fwd_host(self._h, px_ptr, pitch, W, H, coef_ptr, layout, var_ptr, None))

    def dequant_idct_ptr(self, coef_ptr, W, H, px_ptr, pitch, layout=NATURAL, var_ptr=None):
        _check(_inv_host(self._h, coef_ptr, W, H, layout, var_ptr, px_ptr, pitch, None))

    # -- the same, asynchronous: returns at once; wait() blocks until the host buffers hold the result
    def fwd_quant_ptr_async(self, px_ptr, pitch, W, H, coef_ptr, layout=NATURAL, var_ptr=None):
        _check(_fwd_host_async(self._h, px_ptr, pitch, W, H, coef_ptr, layout, var_ptr))

    def dequant_idct_ptr_async(self, coef_ptr, W, H, px_ptr, pitch, layout=NATURAL, var_ptr=None):
        _check(_inv_host_async(self._h, coef_ptr, W, H, layout, var_ptr, px_ptr, pitch))

    # -- int8 records over PCIe (host arrays; only when records_fit_i8) --------------------------
    @property
    def records_fit_i8(self) -> bool:
        return bool(_fits_i8(self._h))

    def fwd_quant_i8(self, pixels, layout=NATURAL, want_stats=False):
        assert pixels.dtype == np.uint8 and pixels.ndim == 2 and pixels.strides[1] == 1
        H, W = pixels.shape
        n = self.n
        nb = (H // n) * (W // n)
        coef = np.empty((nb, n * n), dtype=np.int8)
        var = np.empty(nb, dtype=np.float64) if self.adaptive else None
        st = Stats()
        _check(_fwd_i8(self._h, pixels.ctypes.data, pixels.strides[0], W, H, coef.ctypes.data, layout,
                       var.ctypes.data if var is not None else None, C.byref(st)))
        out = (coef, var) if self.adaptive else coef
        return (out, st.as_dict()) if want_stats else out

    def dequant_idct_i8(self, coef8, W, H, layout=NATURAL, var=None, want_stats=False):
        coef8 = np.ascontiguousarray(coef8, dtype=np.int8)
        px = np.empty((H, W), dtype=np.uint8)
        if var is not None:
            var = np.ascontiguousarray(var, dtype=np.float64)
        st = Stats()
        _check(_inv_i8(self._h, coef8.ctypes.data, W, H, layout, var.ctypes.data if var is not None else None,
                       px.ctypes.data, px.strides[0], C.byref(st)))
        return (px, st.as_dict()) if want_stats else px

    def fwd_quant_i8_ptr_async(self, px_ptr, pitch, W, H, coef_ptr, layout=NATURAL, var_ptr=None):
        _check(_fwd_i8_async(self._h, px_ptr, pitch, W, H, coef_ptr, layout, var_ptr))

    def dequant_idct_i8_ptr_async(self, coef_ptr, W, H, px_ptr, pitch, layout=NATURAL, var_ptr=None):
        _check(_inv_i8_async(self._h, coef_ptr, W, H, layout, var_ptr, px_ptr, pitch))

    def kernel_launches(self) -> int:
        return int(_kernel_launches(self._h))

    def wait(self, want_stats=False):
        st = Stats()
        _check(_plan_wait(self._h, C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    # -- device tensors (torch, cuda): asynchronous on the current / given stream -----------
    def fwd_quant_dev(self, pixels, layout=NATURAL, coef_out=None, var_out=None, stream=None):
        import torch
        assert pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.dim() == 2 and pixels.stride(1) == 1
        H, W = pixels.shape
        nb = (H // self.n) * (W // self.n)
        coef = coef_out if coef_out is not None else torch.empty((nb, self.n * self.n), dtype=torch.int16, device=pixels.device)
        var = var_out
        if self.adaptive and var is None:
            var = torch.empty(nb, dtype=torch.float64, device=pixels.device)
        _on_device(pixels, self.device, "pixels"), _on_device(coef, self.device, "coef_out"), _on_device(var, self.device, "var_out")
        _check(_fwd_dev(self._h, pixels.data_ptr(), pixels.stride(0), W, H, coef.data_ptr(), layout,
                        var.data_ptr() if var is not None else None, _stream_ptr(stream, self.device)))
        return (coef, var) if self.adaptive else coef

    def dequant_idct_dev(self, coef, W, H, layout=NATURAL, var=None, pixels_out=None, stream=None):
        import torch
        assert coef.is_cuda and coef.dtype == torch.int16 and coef.is_contiguous()
        px = pixels_out if pixels_out is not None else torch.empty((H, W), dtype=torch.uint8, device=coef.device)
        _on_device(coef, self.device, "coef"), _on_device(px, self.device, "pixels_out"), _on_device(var, self.device, "var")
        _check(_inv_dev(self._h, coef.data_ptr(), W, H, layout, var.data_ptr() if var is not None else None,
                        px.data_ptr(), px.stride(0), _stream_ptr(stream, self.device)))
        return px

    def debug_skip_replay(self, skip=True):
        _check(_skip_replay(self._h, int(skip)))

    def profile(self, enable=True):
        _check(_profile(self._h, int(enable)))

    def profile_fetch(self):
        """-> {fwd_ms, fwd_launches, inv_ms, inv_launches}: device time of the K1 / K2 launches alone."""
        fm, im, fn, inn = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        _check(_profile_fetch(self._h, C.byref(fm), C.byref(fn), C.byref(im), C.byref(inn)))
        return {"fwd_ms": fm.value, "fwd_launches": fn.value, "inv_ms": im.value, "inv_launches": inn.value}

    def rle_dev(self, coef, layout=NATURAL, stream=None):
        """Run-length symbols of device-resident records (torch int16 [nblocks, 64]).
        -> (offsets uint32-as-int32 tensor [nblocks+1], symbols int32 tensor [total, 2] = (value, run_length))."""
        import torch
        assert coef.is_cuda and coef.dtype == torch.int16 and coef.is_contiguous()
        nb = coef.numel() // 64
        off = torch.empty(nb + 1, dtype=torch.int32, device=coef.device)
        total = C.c_uint64(0)
        _on_device(coef, self.device, "coef")
        _check(_rle_count(self._h, coef.data_ptr(), nb, off.data_ptr(), C.byref(total), _stream_ptr(stream, self.device)))
        sym = torch.empty((int(total.value), 2), dtype=torch.int32, device=coef.device)
        if total.value:
            _check(_rle_emit(self._h, coef.data_ptr(), nb, layout, off.data_ptr(), sym.data_ptr(), _stream_ptr(stream, self.device)))
        return off, sym

    def stats(self, stream=None):
        st = Stats()
        _check(_stats_fetch(self._h, C.byref(st), _stream_ptr(stream, self.device)))
        return st.as_dict()


# ---- planar front / back end (not in the reference; conventions in csrc/planar.cu) -------------------
def frame420_geometry(width: int, height: int) -> Frame420:
    g = Frame420()
    _frame_geometry(int(width), int(height), C.byref(g))
    return g


def rgb_to_ycbcr420_dev(rgb, stream=None):
    """torch cuda uint8 [H, W, 3] -> (y [y_h, y_w], cb, cr [c_h, c_w]) with the edges replicated."""
    import torch
    assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.dim() == 3 and rgb.shape[2] == 3 and rgb.stride(2) == 1 \
        and rgb.stride(1) == 3
    H, W = rgb.shape[:2]
    g = frame420_geometry(W, H)
    y = torch.empty((g.y_height, (g.y_width + 15) // 16 * 16), dtype=torch.uint8, device=rgb.device)[:, :g.y_width]
    cb = torch.empty((g.c_height, g.c_width), dtype=torch.uint8, device=rgb.device)
    cr = torch.empty_like(cb)
    _check(_rgb_to_ycc(rgb.device.index or 0, rgb.data_ptr(), rgb.stride(0), C.byref(g), y.data_ptr(), y.stride(0),
                       cb.data_ptr(), cr.data_ptr(), cb.stride(0), _stream_ptr(stream, rgb.device)))
    return y, cb, cr


def ycbcr420_to_rgb_dev(y, cb, cr, width, height, rgb_out=None, stream=None):
    import torch
    g = frame420_geometry(width, height)
    assert tuple(y.shape) == (g.y_height, g.y_width) and tuple(cb.shape) == (g.c_height, g.c_width) == tuple(cr.shape)
    assert cb.stride(0) == cr.stride(0)
    rgb = rgb_out if rgb_out is not None else _alloc_rgb(height, width, y.device)
    _check(_ycc_to_rgb(y.device.index or 0, y.data_ptr(), y.stride(0), cb.data_ptr(), cr.data_ptr(), cb.stride(0),
                       C.byref(g), rgb.data_ptr(), rgb.stride(0), _stream_ptr(stream, y.device)))
    return rgb


def _alloc_rgb(height, width, device):
    """[H, W, 3] uint8 view whose rows start 16-byte aligned (the vector path of the conversion kernels)."""
    import torch
    pitch = (width * 3 + 15) // 16 * 16
    buf = torch.empty((height, pitch), dtype=torch.uint8, device=device)
    return buf.as_strided((height, width, 3), (pitch, 3, 1))


def pad_edges_dev(plane, width, height, stream=None):
    """Fill plane[:, width:] and plane[height:, :] (torch cuda uint8 [H_pad, W_pad]) by edge replication, in place."""
    _check(_pad_edges(plane.device.index or 0, plane.data_ptr(), plane.stride(0), int(width), int(height),
                      plane.shape[1], plane.shape[0], _stream_ptr(stream, plane.device)))
    return plane


def fwd_quant_edge(plan, pixels, layout=NATURAL, want_stats=False):
    """Host plane of any size: records for ceil(W/n) x ceil(H/n) blocks, edge blocks completed by replication."""
    assert pixels.dtype == np.uint8 and pixels.ndim == 2 and (pixels.size == 0 or pixels.strides[1] == 1)
    H, W = pixels.shape
    n = plan.n
    nb = ((H + n - 1) // n) * ((W + n - 1) // n)
    coef = np.empty((nb, n * n), dtype=np.int16)
    var = np.empty(nb, dtype=np.float64) if plan.adaptive else None
    st = Stats()
    _check(_fwd_edge(plan._h, pixels.ctypes.data, pixels.strides[0], W, H, coef.ctypes.data, layout,
                     var.ctypes.data if var is not None else None, C.byref(st)))
    out = (coef, var) if plan.adaptive else coef
    return (out, st.as_dict()) if want_stats else out


def dequant_idct_edge(plan, coef, W, H, layout=NATURAL, var=None, want_stats=False):
    coef = np.ascontiguousarray(coef, dtype=np.int16)
    px = np.empty((H, W), dtype=np.uint8)
    if var is not None:
        var = np.ascontiguousarray(var, dtype=np.float64)
    st = Stats()
    _check(_inv_edge(plan._h, coef.ctypes.data, W, H, layout, var.ctypes.data if var is not None else None,
                     px.ctypes.data, px.strides[0], C.byref(st)))
    return (px, st.as_dict()) if want_stats else px


def encode_rgb420(luma, chroma, rgb, layout=NATURAL):
    """Host RGB frame [H, W, 3] uint8 -> (coef_y, coef_cb, coef_cr, stats)."""
    assert rgb.dtype == np.uint8 and rgb.ndim == 3 and rgb.shape[2] == 3
    assert rgb.size == 0 or (rgb.strides[2] == 1 and rgb.strides[1] == 3)
    H, W = rgb.shape[:2]
    g = frame420_geometry(W, H)
    ky = np.empty((g.y_width * g.y_height // 64, 64), dtype=np.int16)
    kcb = np.empty((g.c_width * g.c_height // 64, 64), dtype=np.int16)
    kcr = np.empty_like(kcb)
    st = Stats()
    _check(_encode_rgb(luma._h, chroma._h, rgb.ctypes.data, rgb.strides[0], W, H, ky.ctypes.data, kcb.ctypes.data,
                       kcr.ctypes.data, layout, C.byref(st)))
    return ky, kcb, kcr, st.as_dict()


def decode_rgb420(luma, chroma, ky, kcb, kcr, W, H, layout=NATURAL):
    ky, kcb, kcr = (np.ascontiguousarray(k, dtype=np.int16) for k in (ky, kcb, kcr))
    rgb = np.empty((H, W, 3), dtype=np.uint8)
    st = Stats()
    _check(_decode_rgb(luma._h, chroma._h, ky.ctypes.data, kcb.ctypes.data, kcr.ctypes.data, W, H, layout,
                       rgb.ctypes.data, rgb.strides[0], C.byref(st)))
    return rgb, st.as_dict()


def _shares(share, n):
    if share is None:
        return None
    assert len(share) == n
    return (C.c_float * n)(*[float(v) for v in share])


def fwd_quant_peer(plans, pixels, layout=NATURAL, coef_out=None, var_out=None, share=None, stream=None):
    """torch cuda plane resident on plans[0]'s GPU; plans[1:] are NVLink peers working on that memory in place."""
    import torch
    assert pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.dim() == 2 and pixels.stride(1) == 1
    H, W = pixels.shape
    nb = (H // 8) * (W // 8)
    coef = coef_out if coef_out is not None else torch.empty((nb, 64), dtype=torch.int16, device=pixels.device)
    var = var_out
    if plans[0].adaptive and var is None:
        var = torch.empty(nb, dtype=torch.float64, device=pixels.device)
    hs = (C.c_void_p * len(plans))(*[p._h for p in plans])
    with torch.cuda.device(pixels.device):
        _check(_fwd_peer(hs, len(plans), pixels.data_ptr(), pixels.stride(0), W, H, coef.data_ptr(), layout,
                         var.data_ptr() if var is not None else None, _shares(share, len(plans)), _stream_ptr(stream, pixels.device)))
    return (coef, var) if plans[0].adaptive else coef


def dequant_idct_peer(plans, coef, W, H, layout=NATURAL, var=None, pixels_out=None, share=None, stream=None):
    import torch
    assert coef.is_cuda and coef.dtype == torch.int16 and coef.is_contiguous()
    px = pixels_out if pixels_out is not None else torch.empty((H, W), dtype=torch.uint8, device=coef.device)
    hs = (C.c_void_p * len(plans))(*[p._h for p in plans])
    with torch.cuda.device(coef.device):
        _check(_inv_peer(hs, len(plans), coef.data_ptr(), W, H, layout, var.data_ptr() if var is not None else None,
                         px.data_ptr(), px.stride(0), _shares(share, len(plans)), _stream_ptr(stream, coef.device)))
    return px


def fwd_quant_multi(plans, pixels, layout=NATURAL):
    """One host plane over several GPUs (block-row ranges, no inter-GPU traffic)."""
    H, W = pixels.shape
    nb = (H // 8) * (W // 8)
    adaptive = plans[0].adaptive
    coef = np.empty((nb, 64), dtype=np.int16)
    var = np.empty(nb, dtype=np.float64) if adaptive else None
    hs = (C.c_void_p * len(plans))(*[p._h for p in plans])
    st = Stats()
    _check(_fwd_multi(hs, len(plans), pixels.ctypes.data, pixels.strides[0], W, H, coef.ctypes.data, layout,
                      var.ctypes.data if var is not None else None, C.byref(st)))
    return ((coef, var) if adaptive else coef), st.as_dict()


def dequant_idct_multi(plans, coef, W, H, layout=NATURAL, var=None):
    px = np.empty((H, W), dtype=np.uint8)
    hs = (C.c_void_p * len(plans))(*[p._h for p in plans])
    st = Stats()
    _check(_inv_multi(hs, len(plans), coef.ctypes.data, W, H, layout, var.ctypes.data if var is not None else None,
                      px.ctypes.data, px.strides[0], C.byref(st)))
    return px, st.as_dict()


def exported_symbols():
    """Names include/*.h declare; tests check that the library exports every one of them."""
    return ["dct_init", "dct_free", "dct_forward", "dct_inverse", "create_block_from_pixels",
            "copy_block_to_coefficients", "quant_init", "quant_free", "generate_quant_matrix",
            "generate_dequant_matrix", "quantize", "dequantize", "calculate_block_variance",
            "adjust_matrix_for_block", "dct_cuda_last_error", "dct_cuda_device_count", "dct_cuda_plan_create",
            "dct_cuda_plan_refresh", "dct_cuda_plan_destroy", "dct_cuda_plan_device", "dct_cuda_fwd_quant_u8_dev",
            "dct_cuda_dequant_idct_u8_dev", "dct_cuda_fwd_quant_planes_dev", "dct_cuda_dequant_idct_planes_dev",
            "dct_cuda_fwd_quant_u8", "dct_cuda_dequant_idct_u8", "dct_cuda_fwd_quant_u8_async", "dct_cuda_fwd_quant_f32_dev", "dct_cuda_fwd_quant_f32",
            "dct_cuda_dequant_idct_u8_async", "dct_cuda_plan_wait", "dct_cuda_fwd_quant_u8_multi",
            "dct_cuda_dequant_idct_u8_multi", "dct_cuda_stats_fetch", "dct_cuda_plan_profile", "dct_cuda_plan_debug_skip_replay", "dct_cuda_rle_count_dev",
            "dct_cuda_rle_emit_dev",
            "dct_cuda_profile_fetch", "dct_cuda_record_to_block",
            "dct_cuda_block_to_record", "dct_cuda_host_alloc", "dct_cuda_host_free",
            "dct_cuda_fwd_quant_u8_edge", "dct_cuda_dequant_idct_u8_edge", "dct_cuda_pad_edges_dev",
            "dct_cuda_frame420_geometry", "dct_cuda_rgb_to_ycbcr420_dev", "dct_cuda_ycbcr420_to_rgb_dev",
            "dct_cuda_encode_rgb420", "dct_cuda_decode_rgb420", "dct_cuda_peer_default_share",
            "dct_cuda_fwd_quant_u8_peer", "dct_cuda_dequant_idct_u8_peer", "dct_cuda_plan_records_fit_i8", "dct_cuda_plan_kernel_launches",
            "dct_cuda_fwd_quant_u8_i8", "dct_cuda_dequant_idct_i8_u8", "dct_cuda_fwd_quant_u8_i8_async",
            "dct_cuda_dequant_idct_i8_u8_async", "dct_cuda_record8_to_block"]
