#!/usr/bin/env python3
"""Generates csrc/m3s_enc_fold_gen.cuh: straight-line code for the encoder's polyphase matrixing
(MP3_Encoder.py:358-368, s_b = sum_j mul(fl[b][j], y_j), every product truncated on its own) in which equal products are
computed ONCE.

The 32 x 64 matrix `fl` (M3S_ENC_FL in m3s_tables_data.h, taken from the reference's object) holds cosines of a
128-point grid, so a column j carries at most 16 distinct magnitudes: 920 distinct (|v|, j) pairs for 2,025 non-zero
entries.  mul(v, y) = (v * y) >> 32 with an arithmetic shift, hence for v > 0

    mul(-v, y) = -mul(v, y) - [ (v * y) mod 2^32 != 0 ]          and   (v * y) mod 2^32 == 0  <=>  tz(v) + tz(y) >= 32  (or y == 0)

so one IMAD.HI per distinct magnitude serves every band that uses +v or -v, as long as the correction bit is known.  The
generated fast path assumes it is 1 for every negative use -- true when the low YBITS bits of y are not all zero (tz(y)
< YBITS = 32 - max tz(v) of the table) -- folds the corrections into the accumulators' start values and
reports `bad` when a y breaks the assumption (silence, mostly); the caller then redoes that slot with the direct form.

The same folding is generated for the MDCT (MP3_Encoder.py:683-701, M3S_ENC_COSL: 488 distinct (|v|, j) pairs of 648).

Layout of the generated functions: one thread owns ALL outputs of its inputs, the coefficient pattern is straight-line code.
m3s_matrix_fold<GUARD>(ld, z, acc, bad, any) also computes its inputs: the windowed value y_j of the thread's time slot from
the eight samples ld(j, k) (window coefficients as immediates), then the products of column j, then the signed adds into
acc[0..31].  m3s_mdct_fold<GUARD>(ld, z, acc, bad, any) reads input j (slot j of the previous granule for j < 18, slot j - 18 of
the current one) as ld(j) and leaves the 18 MDCT lines of the thread's band in acc[].

`z` must be 0 at run time and unknown at compile time (a kernel argument).  Left alone, ptxas folds an ACCUMULATOR into the
multiply's own addend and pays for it by computing a shared product once per use (measured: 2,128 IMAD.HI in the kernel instead
of 1,586).  GUARD = 0 adds z to every product (that add takes the addend slot, so no accumulator can), GUARD = 1 xors it in (one
logic op per product, no register-pair constraints), GUARD = 2 leaves the products unguarded.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mp3-steganography-lib_b200", "csrc")


def table(name):
    src = open(os.path.join(CSRC, "m3s_tables_data.h")).read()
    m = re.search(name + r"\s*\[\d+\]\s*=\s*\{(.*?)\};", src, re.S)
    return [int(x) for x in re.findall(r"-?\d+", m.group(1))]


def tz(v):
    t = 0
    while not (v >> t) & 1:
        t += 1
    return t


def fold(mat, rows, cols, fn, window=None):
    """mat[r][c]: r = output index, c = input index.  Emits one function that walks every column.
    window: the 512 analysis window coefficients -- then input c is itself computed in front of its products,
    y_c = sum_k mul(ld(c, k), window[c + 64 k])  (MP3_Encoder.py:337-356), else it is read as ld(c)."""
    col_mags = []
    for c in range(cols):
        mags = sorted({abs(mat[r][c]) for r in range(rows)} - {0})
        assert all(v <= 0x7FFFFFFF for v in mags)
        col_mags.append(mags)
    ybits = 32 - max(tz(v) for mags in col_mags for v in mags)   # tz(y) < ybits  =>  tz(v) + tz(y) < 32 for every v
    assert ybits >= 16
    out = []
    out.append("#define %s_YMASK 0x%08Xu   // an input whose low %d bits are all zero breaks the fast path's correction term" % (fn.upper(), (1 << ybits) - 1, ybits))
    out.append("// %s: %d outputs x %d inputs, %d distinct products of %d non-zero entries" % (
        fn, rows, cols, sum(len(m) for m in col_mags), sum(1 for r in range(rows) for c in range(cols) if mat[r][c])))
    out.append("template <int GUARD, class Load> M3S_FOLD_FN void %s(const Load &ld, const uint32_t z, uint32_t (&acc)[%d], uint32_t &bad, uint32_t &any)" % (fn, rows))
    out.append("{")
    out.append("    bad = 0u; any = 0u;")
    for r in range(rows):
        out.append("    acc[%d] = %du;" % (r, (-sum(1 for c in range(cols) if mat[r][c] < 0)) & 0xFFFFFFFF))
    for c in range(cols):
        out.append("    {")
        if window is None:
            out.append("        const int32_t y = ld(%d);" % c)
        else:
            out.append("        const int32_t y = (int32_t)(%s);" % " + ".join(
                "(uint32_t)M3S_FOLD_MULHI(ld(%d, %d), %d)" % (c, k, window[c + 64 * k]) for k in range(8)))
        out.append("        bad |= (uint32_t)((y & %s_YMASK) == 0); any |= (uint32_t)y;" % fn.upper())
        for q, v in enumerate(col_mags[c]):
            out.append("        const uint32_t p%d = m3s_fold_guard<GUARD>((uint32_t)M3S_FOLD_MULHI(y, %d), z);" % (q, v))
        for r in range(rows):
            v = mat[r][c]
            if v == 0:
                continue
            out.append("        acc[%d] %s= p%d;" % (r, "+" if v > 0 else "-", col_mags[c].index(abs(v))))
        out.append("    }")
    out.append("}")
    return out


def main():
    fl = table("M3S_ENC_FL")
    assert len(fl) == 2048
    mat_fl = [[fl[b * 64 + j] for j in range(64)] for b in range(32)]
    cos = table("M3S_ENC_COSL")
    assert len(cos) == 648
    mat_cos = [[cos[k * 36 + j] for j in range(36)] for k in range(18)]
    win = table("M3S_ENWINDOW")
    assert len(win) == 512
    lines = [
        "// m3s_enc_fold_gen.cuh -- GENERATED by tools/gen_enc_fold.py from M3S_ENWINDOW / M3S_ENC_FL / M3S_ENC_COSL (m3s_tables_data.h); do not edit.",
        "// Windowing + matrixing and MDCT of the encoder's analysis with equal truncated products computed once; see the generator's docstring.",
        "// tests/test_host_logic.py regenerates this file and compares; tests/model/enc_fold_check.cpp holds it to the direct sums.",
        "#pragma once",
        "#include <stdint.h>",
        "#ifndef M3S_FOLD_FN",
        "#define M3S_FOLD_FN static inline",
        "#endif",
        "#ifndef M3S_FOLD_MULHI",
        "#define M3S_FOLD_MULHI(a, b) ((int32_t)(((int64_t)(a) * (int64_t)(b)) >> 32))",
        "#endif",
        "// how a product is kept from being recomputed per use (see the generator): 0 = z added (folds into the multiply's addend),",
        "// 1 = z xor-ed (one logic op per product), 2 = no guard",
        "template <int GUARD> M3S_FOLD_FN uint32_t m3s_fold_guard(const uint32_t p, const uint32_t z) { return GUARD == 0 ? p + z : GUARD == 1 ? (p ^ z) : p; }",
        "",
    ]
    lines += fold(mat_fl, 32, 64, "m3s_matrix_fold", window=win)
    lines.append("")
    lines += fold(mat_cos, 18, 36, "m3s_mdct_fold")
    path = os.path.join(CSRC, "m3s_enc_fold_gen.cuh")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1 and sys.argv[1] == "--check":
        sys.exit(0 if open(path).read() == text else 1)
    open(path, "w").write(text)
    print("wrote", path, len(lines), "lines")


if __name__ == "__main__":
    main()
