"""The oracle is pinned here: every function of oracle/dct_oracle.c against the golden vectors
captured from the unmodified reference (tests/golden/, see make_golden.py), bit for bit, and --
where oracle/_ref/libdct_ref.so exists -- against the reference itself on random planes."""
import numpy as np
import pytest

from conftest import unhex
from oracle import binding as B


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("n", [4, 8, 16])
def test_dct_matrix_bit_exact(oracle, golden_blocks, n):
    # reference: src/dct.c:19-30
    want = unhex(golden_blocks["dct_matrix"][str(n)], (n, n))
    assert np.array_equal(bits(oracle.dct_matrix(n)), bits(want))


@pytest.mark.parametrize("n", [4, 8, 16])
def test_quant_tables_bit_exact(oracle, golden_blocks, n):
    # reference: src/quantization.c:51-99, incl. the un-rounded doubles (q90: 16*0.2 != 3.2)
    for q in (1, 10, 25, 49, 50, 51, 75, 90, 95, 100):
        want = unhex(golden_blocks["quant_table"][f"{n}:{q}"], (n, n))
        assert np.array_equal(bits(oracle.quant_table(q, n)), bits(want)), (n, q)


def test_quality_is_clamped(oracle):
    # reference: src/quantization.c:26-31
    assert np.array_equal(oracle.quant_table(0), oracle.quant_table(1))
    assert np.array_equal(oracle.quant_table(1000), oracle.quant_table(100))
    assert np.all(oracle.quant_table(100) == 1.0)


@pytest.mark.parametrize("n", [4, 8, 16])
def test_zigzag_order(oracle, golden_blocks, n):
    # reference: src/entropy.c:158-178 (N=8 is the JPEG order)
    assert list(oracle.zigzag_order(n)) == golden_blocks["zigzag"][str(n)]
    if n == 8:
        assert list(oracle.zigzag_order(8))[:10] == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24]


def test_known_answer_block(oracle, golden_blocks):
    # reference: tests/test_dct.c:33-42 -> dct_forward; tests/test_entropy.c:311-316, :370-373
    kat = golden_blocks["kat"]
    blk = np.array(kat["pixels"], dtype=np.float64).reshape(8, 8) - 128.0
    c = oracle.dct_forward(blk)
    assert np.array_equal(bits(c), bits(unhex(kat["coeffs"], (8, 8))))
    assert abs(c[0, 0] - (-415.375)) < 1e-12 and abs(c[0, 1] - (-30.1857172768)) < 1e-9
    var = oracle.block_variance(blk)
    assert var == float.fromhex(kat["variance"]) and abs(var - 437.5095214844) < 1e-9
    assert np.array_equal(bits(oracle.dct_inverse(c)), bits(unhex(kat["roundtrip"], (8, 8))))
    assert list(oracle.round_to_int(c).ravel()) == kat["rounded"]
    for key, case in kat["cases"].items():
        q, adaptive = (int(v) for v in key.split(":"))
        Q = oracle.quant_table(q)
        qc = oracle.quantize(Q, c, adaptive, var)
        assert list(qc.ravel()) == case["quantized"], key
        dq = oracle.dequantize(Q, qc, adaptive, var)
        assert np.array_equal(bits(dq), bits(unhex(case["dequantized"], (8, 8)))), key
        assert np.array_equal(bits(oracle.dct_inverse(dq)), bits(unhex(case["idct"], (8, 8)))), key
        assert np.array_equal(bits(oracle.adjust_table(Q, var, 1)), bits(unhex(case["adjust_q"], (8, 8))))
        R = oracle.dequant_table(Q)
        assert np.array_equal(bits(oracle.adjust_table(R, var, 0)), bits(unhex(case["adjust_r"], (8, 8))))
    # the textbook result
    assert kat["cases"]["50:0"]["quantized"][:8] == [-26, -3, -6, 2, 2, -1, 0, 0]
    # S2: the non-adaptive dequantize multiplies by 1/Q
    assert unhex(kat["cases"]["50:0"]["dequantized"])[0] == -26 * (1.0 / 16.0)


@pytest.mark.parametrize("n", [4, 16])
def test_generic_block_sizes(oracle, golden_blocks, n):
    # reference: the same triple loops for block_size != 8; custom table src/quantization.c:78-96
    k = golden_blocks["kat"][f"n{n}"]
    b = unhex(k["block"], (n, n))
    c = oracle.dct_forward(b)
    assert np.array_equal(bits(c), bits(unhex(k["coeffs"], (n, n))))
    Q = oracle.quant_table(50, n)
    qc = oracle.quantize(Q, c)
    assert list(qc.ravel()) == k["quantized"]
    assert np.array_equal(bits(oracle.dequantize(Q, qc)), bits(unhex(k["dequantized"], (n, n))))
    assert np.array_equal(bits(oracle.dct_inverse(c)), bits(unhex(k["idct"], (n, n))))


def test_plane_goldens(oracle, golden_planes):
    for key in golden_planes["cases"]:
        name, q, a = key.split("/")
        q, adaptive = int(q[1:]), int(a[1:])
        px = golden_planes[f"{name}/px"]
        H, W = px.shape
        Q = oracle.quant_table(q)
        cn, var, _ = oracle.fwd_quant_plane(px, Q, adaptive, B.NATURAL)
        assert np.array_equal(cn, golden_planes[key + "/coef"]), key
        if key + "/coef_zz" in golden_planes:
            cz, _, _ = oracle.fwd_quant_plane(px, Q, adaptive, B.ZIGZAG, nthreads=3)
            assert np.array_equal(cz, golden_planes[key + "/coef_zz"]), key
            assert np.array_equal(cz, cn[:, oracle.zigzag_order(8)])
            rz, _ = oracle.dequant_idct_plane(cz, W, H, Q, adaptive, B.ZIGZAG, var)
            assert np.array_equal(rz, golden_planes[key + "/rec"]), key
        if adaptive:
            assert np.array_equal(bits(var), bits(golden_planes[key + "/var"])), key
        rec, _ = oracle.dequant_idct_plane(cn, W, H, Q, adaptive, B.NATURAL, var, nthreads=2)
        assert np.array_equal(rec, golden_planes[key + "/rec"]), key


def test_forced_dc_tie_is_counted(oracle, golden_planes):
    # block 3 of the adversarial strip has sum(px-128) = 64 -> DC = 8.0 -> 8/16 = 0.5 exactly
    px = golden_planes["adv/px"]
    _, _, ties = oracle.fwd_quant_plane(px, oracle.quant_table(50))
    assert ties >= 1


def test_a5_image_hashes(oracle, golden_blocks):
    # 512x512 U q50 hashes measured on the reference
    px = oracle.fill_xorshift(512, 512)
    assert [int(v) for v in px.ravel()[:16]] == golden_blocks["a5"]["input_head"]
    Q = oracle.quant_table(50)
    coef, _, ties = oracle.fwd_quant_plane(px, Q, nthreads=4)
    assert f"{oracle.fnv_i16(coef):016x}" == golden_blocks["a5"]["coef_hash"] == "8c0f119ff9a3b113"
    rec, _ = oracle.dequant_idct_plane(coef, 512, 512, Q, nthreads=4)
    assert f"{oracle.fnv_u8_blockorder(rec):016x}" == golden_blocks["a5"]["pixel_hash"] == "6724fe6c8009af02"
    assert ties == 94
    vals, counts = np.unique(rec, return_counts=True)
    assert dict(zip(vals.tolist(), counts.tolist())) == {127: 2014, 128: 257987, 129: 2143}


@pytest.mark.skipif(not B.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
@pytest.mark.parametrize("quality,adaptive,layout", [(50, 0, 0), (90, 0, 1), (10, 1, 0), (75, 1, 1), (100, 0, 0)])
def test_oracle_equals_reference_on_random_planes(oracle, quality, adaptive, layout):
    ref = B.load("ref")
    rng = np.random.default_rng(quality * 10 + adaptive)
    px = rng.integers(0, 256, size=(128, 192), dtype=np.uint8)
    Q = oracle.quant_table(quality)
    assert np.array_equal(bits(Q), bits(ref.quant_table(quality)))
    co, vo, _ = oracle.fwd_quant_plane(px, Q, adaptive, layout, nthreads=2)
    cr, vr, _ = ref.fwd_quant_plane(px, Q, adaptive, layout, nthreads=2)
    assert np.array_equal(co, cr) and np.array_equal(bits(vo), bits(vr))
    po, _ = oracle.dequant_idct_plane(co, 192, 128, Q, adaptive, layout, vo)
    pr, _ = ref.dequant_idct_plane(cr, 192, 128, Q, adaptive, layout, vr)
    assert np.array_equal(po, pr)
    # arbitrary (not K1-produced) coefficients too
    junk = rng.integers(-2000, 2000, size=co.shape).astype(np.int16)
    po, _ = oracle.dequant_idct_plane(junk, 192, 128, Q, adaptive, layout, vo)
    pr, _ = ref.dequant_idct_plane(junk, 192, 128, Q, adaptive, layout, vr)
    assert np.array_equal(po, pr)


def test_rle_symbols_restate_run_length_encode(oracle, golden_planes):
    # reference: src/entropy.c:216-256 (symbols) on top of :158-178 (zigzag)
    rng = np.random.default_rng(8)
    c = rng.integers(-4, 5, size=(300, 64)).astype(np.int16)
    c[rng.random(c.shape) < 0.7] = 0
    c[0] = 0                      # all zero: one closing symbol (0, 64)
    c[1] = 0
    c[1, 63] = 7                  # only the last coefficient: (7, 63)
    c[2] = 3                      # dense: 64 symbols, all runs 0
    for layout in (0, 1):
        off, sym = oracle.rle_plane(c, layout)
        assert off[0] == 0 and off[-1] == len(sym)
        assert sym[off[0]:off[1]].tolist() == [[0, 64]]
        assert sym[off[1]:off[2]].tolist() == [[7, 63]]
        assert sym[off[2]:off[3]].tolist() == [[3, 0]] * 64
        # decoding the symbols gives the zigzag sequence back (run_length_decode, src/entropy.c:333-358)
        zz = oracle.zigzag_order(8)
        for b in (3, 17, 299):
            seq = []
            for v, run in sym[off[b]:off[b + 1]]:
                seq += [0] * run + [v]
            want = c[b] if layout == 1 else c[b][zz]
            assert seq[:64] == list(want) and len(seq) in (64, 65)
    if B.have_ref():
        ref = B.load("ref")
        for layout in (0, 1):
            a, b_ = oracle.rle_plane(c, layout), ref.rle_plane(c, layout)
            assert np.array_equal(a[0], b_[0]) and np.array_equal(a[1], b_[1])
        coef = golden_planes["u64x48/q50/a0/coef"]
        a, b_ = oracle.rle_plane(coef, 0), ref.rle_plane(coef, 0)
        assert np.array_equal(a[0], b_[0]) and np.array_equal(a[1], b_[1])


@pytest.mark.skipif(not B.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_float_pixel_forward_equals_reference(oracle):
    # block filled by hand as tests/test_dct.c:46-50 does, from float pixels (incl. outside [0, 255])
    ref = B.load("ref")
    rng = np.random.default_rng(12)
    px = (rng.random((64, 96)) * 300.0 - 20.0).astype(np.float32)
    for q, adaptive, layout in [(50, 0, 0), (90, 1, 1)]:
        Q = oracle.quant_table(q)
        a = oracle.fwd_quant_plane_f32(px, Q, adaptive, layout, nthreads=2)
        b_ = ref.fwd_quant_plane_f32(px, Q, adaptive, layout)
        assert np.array_equal(a[0], b_[0]) and np.array_equal(bits(a[1]), bits(b_[1]))
    ipx = rng.integers(0, 256, size=(32, 64), dtype=np.uint8)
    Q = oracle.quant_table(50)
    assert np.array_equal(oracle.fwd_quant_plane_f32(ipx.astype(np.float32), Q)[0], oracle.fwd_quant_plane(ipx, Q)[0])


@pytest.mark.skipif(not B.have_ref(), reason="oracle/_ref not built (no /root/reference on this box)")
@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_generic_block_size_planes_equal_reference(oracle, n):
    # the reference's block functions with block_size = n, custom table src/quantization.c:78-96
    ref = B.load("ref")
    rng = np.random.default_rng(n)
    px = rng.integers(0, 256, size=(2 * n, 3 * n), dtype=np.uint8)
    Q = oracle.quant_table(60, n)
    for adaptive, layout in ((0, 0), (1, 1)):
        c1, v1 = oracle.fwd_quant_plane_n(n, px, Q, adaptive, layout, nthreads=2)
        c2, v2 = ref.fwd_quant_plane_n(n, px, Q, adaptive, layout)
        assert np.array_equal(c1, c2) and np.array_equal(bits(v1), bits(v2))
        p1 = oracle.dequant_idct_plane_n(n, c1, 3 * n, 2 * n, Q, adaptive, layout, v1)
        p2 = ref.dequant_idct_plane_n(n, c2, 3 * n, 2 * n, Q, adaptive, layout, v2)
        assert np.array_equal(p1, p2)
    if n == 8:
        c8, _, _ = oracle.fwd_quant_plane(px, Q)
        assert np.array_equal(c8, oracle.fwd_quant_plane_n(8, px, Q)[0])


# ---------------------------------------------------------------------------------------------
# planar front / back end: OUR convention (the reference has none) -- parity unpinned; what CAN be
# anchored are the published JFIF equations and plain numpy restatements.
# ---------------------------------------------------------------------------------------------
def test_colour_oracle_against_the_jfif_equations(oracle):
    rng = np.random.default_rng(601)
    rgb = rng.integers(0, 256, size=(64, 96, 3), dtype=np.uint8)
    rgb[0, :6] = [[255, 0, 0], [255, 0, 0], [0, 255, 0], [0, 255, 0], [0, 0, 255], [0, 0, 255]]
    rgb[1, :6] = rgb[0, :6]
    y, cb, cr = oracle.rgb_to_ycbcr420(rgb)
    assert (y[0, 0], cb[0, 0], cr[0, 0]) == (76, 85, 255)       # published JFIF values of the primaries
    assert (y[0, 2], cb[0, 1], cr[0, 1]) == (150, 44, 21)
    assert (y[0, 4], cb[0, 2], cr[0, 2]) == (29, 255, 107)
    f = rgb.astype(np.float64)
    yf = 0.299 * f[..., 0] + 0.587 * f[..., 1] + 0.114 * f[..., 2]
    assert np.abs(y - yf).max() <= 0.51                         # 0.5 rounding + 16-bit coefficient error
    m = f.reshape(32, 2, 48, 2, 3).mean(axis=(1, 3))             # 2x2 box average, then the JFIF chroma equations
    cbf = -0.168736 * m[..., 0] - 0.331264 * m[..., 1] + 0.5 * m[..., 2] + 128
    crf = 0.5 * m[..., 0] - 0.418688 * m[..., 1] - 0.081312 * m[..., 2] + 128
    assert np.abs(cb - np.clip(cbf, 0, 255)).max() <= 0.51 and np.abs(cr - np.clip(crf, 0, 255)).max() <= 0.51
    back = oracle.ycbcr420_to_rgb(y, cb, cr, 96, 64).astype(np.float64)
    up = lambda c: np.repeat(np.repeat(c.astype(np.float64), 2, 0), 2, 1) - 128
    want = np.stack([y + 1.402 * up(cr), y - 0.344136 * up(cb) - 0.714136 * up(cr), y + 1.772 * up(cb)], axis=-1)
    assert np.abs(back - np.clip(want, 0, 255)).max() <= 1.02   # two roundings (chroma term, then the sum)


def test_colour_oracle_grey_and_ragged_frames(oracle):
    grey = np.repeat(np.arange(256, dtype=np.uint8).reshape(16, 16), 3).reshape(16, 16, 3)
    y, cb, cr = oracle.rgb_to_ycbcr420(grey)
    assert np.array_equal(y, grey[..., 0]) and (cb == 128).all() and (cr == 128).all()
    assert np.array_equal(oracle.ycbcr420_to_rgb(y, cb, cr, 16, 16), grey)
    # ragged frame == the same conversion of the edge-replicated frame
    rng = np.random.default_rng(9)
    rgb = rng.integers(0, 256, size=(13, 21, 3), dtype=np.uint8)
    y, cb, cr = oracle.rgb_to_ycbcr420(rgb)
    assert y.shape == (16, 24) and cb.shape == (8, 16)
    full = np.pad(rgb, ((0, 3), (0, 11), (0, 0)), mode="edge")   # 16 x 32: whole luma AND chroma blocks
    y2, cb2, cr2 = oracle.rgb_to_ycbcr420(full)
    assert np.array_equal(y, y2[:, :24]) and np.array_equal(cb, cb2) and np.array_equal(cr, cr2)


def test_pad_edges_oracle_is_numpy_edge_padding(oracle):
    rng = np.random.default_rng(3)
    for (h, w), (hp, wp) in (((5, 3), (8, 8)), ((1, 1), (8, 16)), ((8, 8), (8, 8)), ((9, 17), (16, 24))):
        px = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        assert np.array_equal(oracle.pad_edges(px, wp, hp), np.pad(px, ((0, hp - h), (0, wp - w)), mode="edge"))
