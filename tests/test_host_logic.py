"""CPU-side checks (no GPU): the C-ABI library loads, exports every symbol include/*.h declares,
its host-side context code reproduces the reference's tables bit for bit, and the compute entry
points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, unhex


@pytest.fixture(scope="module")
def api():
    from dct_b200 import build
    build.build()
    from dct_b200 import api as _api
    return _api


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def declared_functions():
    names = set()
    for h in ("dct.h", "quantization.h", "dct_cuda.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b([a-z_0-9]+)\s*\(", src))
    return {n for n in names if not n.startswith("defined")}


def test_library_exports_every_declared_symbol(api):
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = declared_functions()
    assert len(declared) >= 30
    assert declared <= exported, sorted(declared - exported)
    assert set(api.exported_symbols()) == declared
    # utils.h's four helpers deliberately stay with the reference's untouched utils.c
    assert not ({"alloc_array", "free_array", "alloc_int_array", "free_int_array"} & exported)


def test_no_torch_types_in_the_abi():
    for h in ("dct.h", "quantization.h", "dct_cuda.h", "utils.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        assert "torch" not in src and "at::" not in src and "std::" not in src


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "dct_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in os.path.relpath(dirpath, pkg).split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f
    out = subprocess.run(["ldd", os.path.join(pkg, "libdct_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "dct_ref" not in out


@pytest.mark.parametrize("n", [4, 8, 16])
def test_dct_init_matches_reference_bits(api, golden_blocks, n):
    ctx = api.dct_init(n)
    D = api.matrix_of(ctx.contents.dct_matrix, n)
    Dt = api.matrix_of(ctx.contents.transposed_dct, n)
    api.dct_free(ctx)
    assert np.array_equal(bits(D), bits(unhex(golden_blocks["dct_matrix"][str(n)], (n, n))))
    assert np.array_equal(D.T, Dt)


@pytest.mark.parametrize("n", [4, 8, 16])
def test_quant_init_matches_reference_bits(api, golden_blocks, n):
    for q in (1, 10, 25, 49, 50, 51, 75, 90, 95, 100):
        ctx = api.quant_init(n, q, 0)
        Q = api.matrix_of(ctx.contents.quant_matrix, n)
        R = api.matrix_of(ctx.contents.dequant_matrix, n)
        api.quant_free(ctx)
        want = unhex(golden_blocks["quant_table"][f"{n}:{q}"], (n, n))
        assert np.array_equal(bits(Q), bits(want)), (n, q)
        assert np.array_equal(bits(R), bits(1.0 / want))
        assert np.array_equal(bits(api.generate_quant_matrix(n, q)), bits(want))
    lo, hi = api.quant_init(8, -5, 1), api.quant_init(8, 400, 0)
    assert lo.contents.quality == 1 and hi.contents.quality == 100 and lo.contents.adaptive == 1
    api.quant_free(lo), api.quant_free(hi)
    api.quant_free(None), api.dct_free(None)   # NULL is accepted (src/dct.c:44, src/quantization.c:44)


def test_host_helpers_match_reference(api, golden_blocks):
    kat = golden_blocks["kat"]
    px = np.array(kat["pixels"], dtype=np.uint8)
    blk = api.create_block_from_pixels(px, 8, 0, 0, 8)
    assert np.array_equal(blk, px.reshape(8, 8).astype(np.float64) - 128.0)
    wide = np.arange(16 * 24, dtype=np.uint8).reshape(16, 24)
    sub = api.create_block_from_pixels(wide, 24, 8, 16, 8)
    assert np.array_equal(sub, wide[8:16, 16:24].astype(np.float64) - 128.0)
    assert api.calculate_block_variance(blk) == float.fromhex(kat["variance"])
    coeffs = unhex(kat["coeffs"], (8, 8))
    assert list(api.copy_block_to_coefficients(coeffs).ravel()) == kat["rounded"]
    assert list(api.copy_block_to_coefficients(np.array([[0.5, -0.5], [1.5, -2.5]])).ravel()) == [1, -1, 2, -3]
    var = float.fromhex(kat["variance"])
    for key, case in kat["cases"].items():
        q, adaptive = (int(v) for v in key.split(":"))
        ctx = api.quant_init(8, q, adaptive)
        assert np.array_equal(bits(api.adjust_matrix_for_block(ctx, var, 1)), bits(unhex(case["adjust_q"], (8, 8))))
        assert np.array_equal(bits(api.adjust_matrix_for_block(ctx, var, 0)), bits(unhex(case["adjust_r"], (8, 8))))
        api.quant_free(ctx)
    Q = api.generate_quant_matrix(8, 90)
    assert np.array_equal(bits(api.generate_dequant_matrix(Q)), bits(1.0 / Q))


def test_record_adapters_follow_the_reference_zigzag(api, golden_blocks):
    zz = golden_blocks["zigzag"]["8"]
    nat = np.arange(64).reshape(8, 8)
    rec = api.block_to_record(nat, api.ZIGZAG)
    assert list(rec) == zz                      # == block_to_zigzag of src/entropy.c:158-178
    assert np.array_equal(api.record_to_block(rec, api.ZIGZAG), nat)
    assert np.array_equal(api.record_to_block(np.arange(64), api.NATURAL), nat)


def test_no_cpu_fallback_without_a_device(api):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    d, q = api.dct_init(8), api.quant_init(8, 50, 0)
    with pytest.raises(api.DctCudaError, match="no CUDA device"):
        api.Plan(d, q)
    # the reference-style void calls print to stderr and exit(EXIT_FAILURE), like src/dct.c:9-12
    code = ("import numpy as np; from dct_b200 import api; c = api.dct_init(8); "
            "api.dct_forward(c, np.zeros((8, 8)))")
    r = subprocess.run(["python", "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    api.dct_free(d), api.quant_free(q)


def test_band_tables_are_current():
    """band_tables.h is generated: regenerate and compare, and re-run the empirical bound check."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("derive_bands", os.path.join(ROOT, "tools", "derive_bands.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    tmp = os.path.join(ROOT, "dct_b200", "build", "band_tables.check.h")
    os.makedirs(os.path.dirname(tmp), exist_ok=True)
    m.emit(tmp)
    assert open(tmp).read() == open(os.path.join(ROOT, "dct_b200", "csrc", "band_tables.h")).read()
    m.check(nblocks=2000)
    assert m.check_inverse_sparse(trials=300) <= 1.0
    # the transfer-coefficient bound can only be tighter than the path-by-path one it replaces, and the inverse
    # gains are symmetric under transposition to first order (the flowgraph treats rows and columns alike)
    G, _ = m.inv_tables()
    assert np.allclose(G.reshape(8, 8), G.reshape(8, 8).T, rtol=1e-5)
    o = m.BoundOps(np.full(64, 128.0))
    X = [[o.inp(np.eye(64)[8 * i + j], 0.0, True) for j in range(8)] for i in range(8)]
    Y = m.two_d(o, m.fdct8, X, rows_first=True)
    assert all(o.transfer_error(Y[u][v]) <= Y[u][v].E * (1 + 1e-12) for u in range(8) for v in range(8))


def test_frame420_geometry_is_pure_host_arithmetic():
    """dct_cuda_frame420_geometry needs no GPU: plane sizes rounded up to whole 8x8 blocks (include/dct_cuda.h)."""
    from dct_b200 import api
    for w, h, want in ((3840, 2160, (3840, 2160, 1920, 1080)), (1919, 1081, (1920, 1088, 960, 544)),
                       (1, 1, (8, 8, 8, 8)), (17, 8, (24, 8, 16, 8)), (7680, 4320, (7680, 4320, 3840, 2160))):
        g = api.frame420_geometry(w, h)
        assert (g.width, g.height) == (w, h)
        assert (g.y_width, g.y_height, g.c_width, g.c_height) == want
    assert api._peer_share(None, 8, 1) == 0.0          # no plan, no share


def test_no_cuda_device_means_errors_not_fallbacks():
    """Without a GPU every plan-level entry point reports DCT_CUDA_ENODEV / EINVAL; nothing computes on the CPU."""
    from dct_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    d, q = api.dct_init(8), api.quant_init(8, 50, 0)
    try:
        with pytest.raises(api.DctCudaError, match="no CUDA device"):
            api.Plan(d, q, 0)
    finally:
        api.dct_free(d), api.quant_free(q)
