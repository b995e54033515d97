"""Regenerates tests/golden/* from the UNMODIFIED reference (run in the build container only).

    make -C oracle ref && python tests/golden/make_golden.py

Sources of truth:
  * ref_test_*.stdout  -- stdout of the reference's own three test programs
    (tests/test_dct.c, tests/test_quantization.c, tests/test_entropy.c), built with the
    reference's Justfile flags by oracle/Makefile.  They are the drop-in acceptance goldens.
  * golden_blocks.json -- block-level vectors: DCT matrices, Q tables, the known-answer block
    of tests/test_dct.c:33-42 through dct_forward / quantize / dequantize / dct_inverse at
    several qualities, adaptive on and off, zigzag orders.  Doubles are stored as C99 hex
    strings so the comparison is bit-exact.
  * golden_planes.npz  -- small planes looped through the reference block functions
    (oracle/ref_harness.c): int16 coefficients (NATURAL and ZIGZAG), reconstructed pixels,
    per-block variances, for several qualities and both adaptive settings; plus the FNV
    hashes of the 512x512 case of SURVEY.md Appendix A.5.
/root/reference does not exist on the GPU box, hence the committed fixtures.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

KAT_PIXELS = [  # tests/test_dct.c:33-42
    52, 55, 61, 66, 70, 61, 64, 73, 63, 59, 55, 90, 109, 85, 69, 72,
    62, 59, 68, 113, 144, 104, 66, 73, 63, 58, 71, 122, 154, 106, 70, 69,
    67, 61, 68, 104, 126, 88, 68, 70, 79, 65, 60, 70, 77, 68, 58, 75,
    85, 71, 64, 59, 55, 61, 65, 83, 87, 79, 69, 68, 65, 76, 78, 94]


def hexes(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    ref = B.load("ref")
    orc = B.load("oracle")  # only for the input generator and the hash helpers

    for t in ("test_dct", "test_quantization", "test_entropy"):
        out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", t)], capture_output=True, check=True).stdout
        with open(os.path.join(HERE, f"ref_{t}.stdout"), "wb") as f:
            f.write(out)

    g = {"dct_matrix": {}, "quant_table": {}, "zigzag": {}, "kat": {}}
    for n in (4, 8, 16):
        g["dct_matrix"][str(n)] = hexes(ref.dct_matrix(n))
        g["zigzag"][str(n)] = [int(v) for v in ref.zigzag_order(n)]
        for q in (1, 10, 25, 49, 50, 51, 75, 90, 95, 100):
            g["quant_table"][f"{n}:{q}"] = hexes(ref.quant_table(q, n))

    blk = np.array(KAT_PIXELS, dtype=np.float64).reshape(8, 8) - 128.0
    coeffs = ref.dct_forward(blk)
    var = ref.block_variance(blk)
    g["kat"]["pixels"] = KAT_PIXELS
    g["kat"]["coeffs"] = hexes(coeffs)
    g["kat"]["variance"] = float(var).hex()
    g["kat"]["roundtrip"] = hexes(ref.dct_inverse(coeffs))
    g["kat"]["rounded"] = [int(v) for v in ref.round_to_int(coeffs).ravel()]
    g["kat"]["cases"] = {}
    for q in (10, 50, 75, 90, 95, 100):
        Q = ref.quant_table(q)
        for adaptive in (0, 1):
            qc = ref.quantize(Q, coeffs, adaptive, var)
            dq = ref.dequantize(Q, qc, adaptive, var)
            rec = ref.dct_inverse(dq)
            g["kat"]["cases"][f"{q}:{adaptive}"] = {
                "quantized": [int(v) for v in qc.ravel()],
                "dequantized": hexes(dq),
                "idct": hexes(rec),
                "adjust_q": hexes(ref.adjust_table(Q, var, 1)),
                "adjust_r": hexes(ref.adjust_table(ref.dequant_table(Q), var, 0)),
            }
    # a 4x4 and a 16x16 block through the generic-N path
    rng = np.random.default_rng(1234)
    for n in (4, 16):
        b = rng.integers(0, 256, size=(n, n)).astype(np.float64) - 128.0
        c = ref.dct_forward(b)
        Q = ref.quant_table(50, n)
        qc = ref.quantize(Q, c, 0, 0.0)
        g["kat"][f"n{n}"] = {
            "block": hexes(b), "coeffs": hexes(c), "quantized": [int(v) for v in qc.ravel()],
            "dequantized": hexes(ref.dequantize(Q, qc, 0, 0.0)), "idct": hexes(ref.dct_inverse(c))}

    px512 = orc.fill_xorshift(512, 512)
    Q50 = ref.quant_table(50)
    c512, _, _ = ref.fwd_quant_plane(px512, Q50, 0, B.NATURAL, nthreads=4)
    r512, _ = ref.dequant_idct_plane(c512, 512, 512, Q50, 0, B.NATURAL, nthreads=4)
    g["a5"] = {"coef_hash": f"{orc.fnv_i16(c512):016x}", "pixel_hash": f"{orc.fnv_u8_blockorder(r512):016x}",
               "input_head": [int(v) for v in px512.ravel()[:16]]}
    with open(os.path.join(HERE, "golden_blocks.json"), "w") as f:
        json.dump(g, f, indent=0)

    planes = {}
    cases = []
    for name, (H, W, dist, seed) in {"u64x48": (48, 64, 0, 7), "s40x72": (40, 72, 1, 11),
                                      "u8x8": (8, 8, 0, 3), "u16x264": (16, 264, 0, 5)}.items():
        px = orc.fill_xorshift(H, W, seed=0x9E3779B97F4A7C15 + seed, dist=dist)
        planes[f"{name}/px"] = px
        for q in (10, 50, 90, 100):
            Q = ref.quant_table(q)
            for adaptive in (0, 1):
                cn, var, _ = ref.fwd_quant_plane(px, Q, adaptive, B.NATURAL)
                cz, _, _ = ref.fwd_quant_plane(px, Q, adaptive, B.ZIGZAG)
                rec, _ = ref.dequant_idct_plane(cn, W, H, Q, adaptive, B.NATURAL, var)
                key = f"{name}/q{q}/a{adaptive}"
                planes[key + "/coef"] = cn
                planes[key + "/coef_zz"] = cz
                planes[key + "/rec"] = rec
                if adaptive:
                    planes[key + "/var"] = var
                cases.append(key)
    # adversarial: flat 0, flat 255, checkerboard, forced DC ties
    adv = np.zeros((8, 32), dtype=np.uint8)
    adv[:, 8:16] = 255
    adv[:, 16:24] = (np.indices((8, 8)).sum(0) % 2) * 255
    adv[:, 24:32] = 128
    adv[0, 24] = 192  # sum = 64*128 + 64  =>  DC = 8.0 => c/Q = 0.5 at q50 (exact tie)
    planes["adv/px"] = adv
    for q in (50, 100):
        Q = ref.quant_table(q)
        cn, var, _ = ref.fwd_quant_plane(adv, Q, 0, B.NATURAL)
        planes[f"adv/q{q}/a0/coef"] = cn
        planes[f"adv/q{q}/a0/rec"], _ = ref.dequant_idct_plane(cn, 32, 8, Q, 0, B.NATURAL)
        cases.append(f"adv/q{q}/a0")
    planes["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "golden_planes.npz"), **planes)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
