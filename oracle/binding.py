"""ctypes binding for the CPU checker libraries (TEST INFRASTRUCTURE ONLY).

`load("oracle")` -> oracle/libdct_oracle.so   (our C restatement, symbols orc_*)
`load("ref")`    -> oracle/_ref/libdct_ref.so (the unmodified reference, symbols ref_*)

Both expose the same flat-array interface, so a test can run the same case through either.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under dct_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NATURAL, ZIGZAG = 0, 1

_P_D = C.POINTER(C.c_double)
_P_I = C.POINTER(C.c_int)
_P_U8 = C.POINTER(C.c_uint8)
_P_I16 = C.POINTER(C.c_int16)
_P_U64 = C.POINTER(C.c_uint64)


def _dp(a):
    return a.ctypes.data_as(_P_D)


def build_oracle(force: bool = False) -> str:
    """Compile oracle/libdct_oracle.so (gcc, < 1 s). Building the checker is not using it."""
    so = os.path.join(HERE, "libdct_oracle.so")
    src = os.path.join(HERE, "dct_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libdct_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref(reference_root: str = "/root/reference") -> str | None:
    """Compile oracle/_ref/* from the reference sources where they lie (only where they exist)."""
    so = os.path.join(HERE, "_ref", "libdct_ref.so")
    if os.path.isdir(os.path.join(reference_root, "src")):
        subprocess.check_call(["make", "-C", HERE, "ref", f"REF={reference_root}"], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


class CpuChecker:
    """Flat-array view of either checker library."""

    def __init__(self, path: str, prefix: str):
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        f = self._f
        f("dct_matrix", None, C.c_int, _P_D)
        f("dct_forward", None, C.c_int, _P_D, _P_D, _P_D)
        f("dct_inverse", None, C.c_int, _P_D, _P_D, _P_D)
        f("quant_table", None, C.c_int, C.c_int, _P_D)
        f("dequant_table", None, C.c_int, _P_D, _P_D)
        f("block_variance", C.c_double, C.c_int, _P_D)
        f("adjust_table", None, C.c_int, _P_D, C.c_double, C.c_int, _P_D)
        f("quantize", None, C.c_int, _P_D, C.c_int, _P_D, _P_I, C.c_double)
        f("dequantize", None, C.c_int, _P_D, _P_D, C.c_int, _P_I, _P_D, C.c_double)
        f("zigzag_order", None, C.c_int, _P_I)
        f("round_to_int", None, C.c_int, _P_D, _P_I)
        f("fwd_quant_plane", C.c_int, _P_U8, C.c_size_t, C.c_int, C.c_int, _P_D, C.c_int, C.c_int, _P_I16,
          _P_D, C.c_int, _P_U64)
        f("dequant_idct_plane", C.c_int, _P_I16, C.c_int, C.c_int, _P_D, _P_D, C.c_int, C.c_int, _P_D,
          _P_U8, C.c_size_t, C.c_int, _P_U64)
        f("fwd_quant_plane_f32", C.c_int, C.POINTER(C.c_float), C.c_size_t, C.c_int, C.c_int, _P_D, C.c_int, C.c_int, _P_I16,
          _P_D, C.c_int, _P_U64)
        f("fwd_quant_plane_n", C.c_int, C.c_int, _P_U8, C.c_size_t, C.c_int, C.c_int, _P_D, C.c_int, C.c_int, _P_I16,
          _P_D, C.c_int, _P_U64)
        f("dequant_idct_plane_n", C.c_int, C.c_int, _P_I16, C.c_int, C.c_int, _P_D, _P_D, C.c_int, C.c_int, _P_D,
          _P_U8, C.c_size_t, C.c_int, _P_U64)
        f("rle_plane", C.c_size_t, _P_I16, C.c_size_t, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int32))
        if prefix == "orc_":
            self.lib.orc_rgb_to_ycbcr420.argtypes = [_P_U8, C.c_size_t, C.c_int, C.c_int, _P_U8, C.c_size_t, C.c_int, C.c_int,
                                                     _P_U8, _P_U8, C.c_size_t, C.c_int, C.c_int]
            self.lib.orc_rgb_to_ycbcr420.restype = None
            self.lib.orc_ycbcr420_to_rgb.argtypes = [_P_U8, C.c_size_t, _P_U8, _P_U8, C.c_size_t, C.c_int, C.c_int, _P_U8,
                                                     C.c_size_t]
            self.lib.orc_ycbcr420_to_rgb.restype = None
            self.lib.orc_pad_edges.argtypes = [_P_U8, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            self.lib.orc_pad_edges.restype = None
            self.lib.orc_fill_xorshift.argtypes = [_P_U8, C.c_size_t, C.c_uint64, C.c_int, C.c_int]
            self.lib.orc_fill_xorshift.restype = None
            self.lib.orc_fnv_i16.argtypes = [_P_I16, C.c_size_t]
            self.lib.orc_fnv_i16.restype = C.c_uint64
            self.lib.orc_fnv_u8_blockorder.argtypes = [_P_U8, C.c_size_t, C.c_int, C.c_int]
            self.lib.orc_fnv_u8_blockorder.restype = C.c_uint64

    def _f(self, name, restype, *argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = restype, list(argtypes)
        setattr(self, "_" + name, fn)

    # ---- block level ----------------------------------------------------------------
    def dct_matrix(self, n=8):
        D = np.zeros((n, n))
        self._dct_matrix(n, _dp(D))
        return D

    def dct_forward(self, blk):
        blk = np.ascontiguousarray(blk, dtype=np.float64)
        n = blk.shape[0]
        D, out = self.dct_matrix(n), np.zeros((n, n))
        self._dct_forward(n, _dp(D), _dp(blk), _dp(out))
        return out

    def dct_inverse(self, blk):
        blk = np.ascontiguousarray(blk, dtype=np.float64)
        n = blk.shape[0]
        D, out = self.dct_matrix(n), np.zeros((n, n))
        self._dct_inverse(n, _dp(D), _dp(blk), _dp(out))
        return out

    def quant_table(self, quality, n=8):
        Q = np.zeros((n, n))
        self._quant_table(n, int(quality), _dp(Q))
        return Q

    def dequant_table(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        R = np.zeros_like(Q)
        self._dequant_table(Q.shape[0], _dp(Q), _dp(R))
        return R

    def block_variance(self, blk):
        blk = np.ascontiguousarray(blk, dtype=np.float64)
        return float(self._block_variance(blk.shape[0], _dp(blk)))

    def adjust_table(self, src, variance, is_quantize):
        src = np.ascontiguousarray(src, dtype=np.float64)
        out = np.zeros_like(src)
        self._adjust_table(src.shape[0], _dp(src), float(variance), int(is_quantize), _dp(out))
        return out

    def quantize(self, Q, coeffs, adaptive=0, variance=0.0):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        q = np.zeros(c.shape, dtype=np.int32)
        self._quantize(c.shape[0], _dp(Q), int(adaptive), _dp(c), q.ctypes.data_as(_P_I), float(variance))
        return q

    def dequantize(self, Q, q, adaptive=0, variance=0.0):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        R = self.dequant_table(Q)
        q = np.ascontiguousarray(q, dtype=np.int32)
        c = np.zeros(q.shape, dtype=np.float64)
        self._dequantize(q.shape[0], _dp(Q), _dp(R), int(adaptive), q.ctypes.data_as(_P_I), _dp(c),
                         float(variance))
        return c

    def zigzag_order(self, n=8):
        o = np.zeros(n * n, dtype=np.int32)
        self._zigzag_order(n, o.ctypes.data_as(_P_I))
        return o

    def round_to_int(self, blk):
        blk = np.ascontiguousarray(blk, dtype=np.float64)
        out = np.zeros(blk.shape, dtype=np.int32)
        self._round_to_int(blk.shape[0], _dp(blk), out.ctypes.data_as(_P_I))
        return out

    # ---- plane level ----------------------------------------------------------------
    def fwd_quant_plane(self, px, Q, adaptive=0, layout=NATURAL, nthreads=1):
        """px: (H, W) uint8 (any row stride) -> (coef int16 [nblocks,64], var float64 [nblocks], near_ties)."""
        assert px.dtype == np.uint8 and px.ndim == 2 and px.strides[1] == 1
        H, W = px.shape
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        nb = (H // 8) * (W // 8)
        coef = np.zeros((nb, 64), dtype=np.int16)
        var = np.zeros(nb, dtype=np.float64)
        ties = C.c_uint64(0)
        rc = self._fwd_quant_plane(px.ctypes.data_as(_P_U8), px.strides[0], W, H, _dp(Q), int(adaptive),
                                   int(layout), coef.ctypes.data_as(_P_I16), _dp(var), int(nthreads),
                                   C.byref(ties))
        if rc != 0:
            raise ValueError(f"{self.prefix}fwd_quant_plane rc={rc}")
        return coef, var, int(ties.value)

    def fwd_quant_plane_f32(self, px, Q, adaptive=0, layout=NATURAL, nthreads=1):
        """px: (H, W) float32 -> like fwd_quant_plane, the block being (double)px - 128.0."""
        px = np.ascontiguousarray(px, dtype=np.float32)
        H, W = px.shape
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        nb = (H // 8) * (W // 8)
        coef = np.zeros((nb, 64), dtype=np.int16)
        var = np.zeros(nb, dtype=np.float64)
        ties = C.c_uint64(0)
        rc = self._fwd_quant_plane_f32(px.ctypes.data_as(C.POINTER(C.c_float)), W, W, H, _dp(Q), int(adaptive), int(layout),
                                       coef.ctypes.data_as(_P_I16), _dp(var), int(nthreads), C.byref(ties))
        if rc != 0:
            raise ValueError(f"{self.prefix}fwd_quant_plane_f32 rc={rc}")
        return coef, var, int(ties.value)

    def dequant_idct_plane(self, coef, W, H, Q, adaptive=0, layout=NATURAL, var=None, nthreads=1):
        coef = np.ascontiguousarray(coef, dtype=np.int16)
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        R = self.dequant_table(Q)
        px = np.zeros((H, W), dtype=np.uint8)
        ties = C.c_uint64(0)
        vptr = _dp(np.ascontiguousarray(var, dtype=np.float64)) if var is not None else None
        rc = self._dequant_idct_plane(coef.ctypes.data_as(_P_I16), W, H, _dp(Q), _dp(R), int(adaptive),
                                      int(layout), vptr, px.ctypes.data_as(_P_U8), W, int(nthreads),
                                      C.byref(ties))
        if rc != 0:
            raise ValueError(f"{self.prefix}dequant_idct_plane rc={rc}")
        return px, int(ties.value)

    def fwd_quant_plane_n(self, n, px, Q, adaptive=0, layout=NATURAL, nthreads=1):
        """Any block size n: px (H, W) uint8 -> (coef int16 [nblocks, n*n], var [nblocks])."""
        px = np.ascontiguousarray(px, dtype=np.uint8)
        H, W = px.shape
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        nb = (H // n) * (W // n)
        coef = np.zeros((nb, n * n), dtype=np.int16)
        var = np.zeros(nb, dtype=np.float64)
        rc = self._fwd_quant_plane_n(n, px.ctypes.data_as(_P_U8), W, W, H, _dp(Q), int(adaptive), int(layout),
                                     coef.ctypes.data_as(_P_I16), _dp(var), int(nthreads), None)
        if rc != 0:
            raise ValueError(f"{self.prefix}fwd_quant_plane_n rc={rc}")
        return coef, var

    def dequant_idct_plane_n(self, n, coef, W, H, Q, adaptive=0, layout=NATURAL, var=None, nthreads=1):
        coef = np.ascontiguousarray(coef, dtype=np.int16)
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        R = self.dequant_table(Q)
        px = np.zeros((H, W), dtype=np.uint8)
        vptr = _dp(np.ascontiguousarray(var, dtype=np.float64)) if var is not None else None
        rc = self._dequant_idct_plane_n(n, coef.ctypes.data_as(_P_I16), W, H, _dp(Q), _dp(R), int(adaptive), int(layout),
                                        vptr, px.ctypes.data_as(_P_U8), W, int(nthreads), None)
        if rc != 0:
            raise ValueError(f"{self.prefix}dequant_idct_plane_n rc={rc}")
        return px

    def rle_plane(self, coef, layout=NATURAL):
        """-> (offsets uint32 [nblocks+1], symbols int32 [total, 2] = (value, run_length))."""
        coef = np.ascontiguousarray(coef, dtype=np.int16).reshape(-1, 64)
        nb = coef.shape[0]
        off = np.zeros(nb + 1, dtype=np.uint32)
        total = self._rle_plane(coef.ctypes.data_as(_P_I16), nb, int(layout), off.ctypes.data_as(C.POINTER(C.c_uint32)), None)
        sym = np.zeros((total, 2), dtype=np.int32)
        self._rle_plane(coef.ctypes.data_as(_P_I16), nb, int(layout), off.ctypes.data_as(C.POINTER(C.c_uint32)),
                        sym.ctypes.data_as(C.POINTER(C.c_int32)))
        return off, sym

    # ---- planar front / back end (oracle only: the reference has none; parity unpinned) ----
    @staticmethod
    def frame420_geometry(W, H):
        up8 = lambda v: (v + 7) // 8 * 8
        return up8(W), up8(H), up8((W + 1) // 2), up8((H + 1) // 2)

    def rgb_to_ycbcr420(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        H, W = rgb.shape[:2]
        yw, yh, cw, ch = self.frame420_geometry(W, H)
        y, cb, cr = np.zeros((yh, yw), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8)
        self.lib.orc_rgb_to_ycbcr420(rgb.ctypes.data_as(_P_U8), W * 3, W, H, y.ctypes.data_as(_P_U8), yw, yw, yh,
                                     cb.ctypes.data_as(_P_U8), cr.ctypes.data_as(_P_U8), cw, cw, ch)
        return y, cb, cr

    def ycbcr420_to_rgb(self, y, cb, cr, W, H):
        y, cb, cr = (np.ascontiguousarray(a, dtype=np.uint8) for a in (y, cb, cr))
        rgb = np.zeros((H, W, 3), np.uint8)
        self.lib.orc_ycbcr420_to_rgb(y.ctypes.data_as(_P_U8), y.shape[1], cb.ctypes.data_as(_P_U8), cr.ctypes.data_as(_P_U8),
                                     cb.shape[1], W, H, rgb.ctypes.data_as(_P_U8), W * 3)
        return rgb

    def pad_edges(self, px, Wp, Hp):
        px = np.ascontiguousarray(px, dtype=np.uint8)
        H, W = px.shape
        out = np.zeros((Hp, Wp), np.uint8)
        out[:H, :W] = px
        self.lib.orc_pad_edges(out.ctypes.data_as(_P_U8), Wp, W, H, Wp, Hp, 1)
        return out

    # ---- oracle-only helpers ----------------------------------------------------------
    def fill_xorshift(self, H, W, seed=0x9E3779B97F4A7C15, dist=0):
        a = np.zeros((H, W), dtype=np.uint8)
        self.lib.orc_fill_xorshift(a.ctypes.data_as(_P_U8), a.size, C.c_uint64(seed), dist, W)
        return a

    def fnv_i16(self, a):
        a = np.ascontiguousarray(a, dtype=np.int16)
        return int(self.lib.orc_fnv_i16(a.ctypes.data_as(_P_I16), a.size))

    def fnv_u8_blockorder(self, px):
        px = np.ascontiguousarray(px, dtype=np.uint8)
        return int(self.lib.orc_fnv_u8_blockorder(px.ctypes.data_as(_P_U8), px.shape[1], px.shape[1], px.shape[0]))


def load(which: str = "oracle") -> CpuChecker:
    if which == "oracle":
        return CpuChecker(build_oracle(), "orc_")
    if which in ("ref", "ref_shipped"):
        name = "libdct_ref.so" if which == "ref" else "libdct_ref_shipped.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        return CpuChecker(path, "ref_")
    raise ValueError(which)


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libdct_ref.so"))
