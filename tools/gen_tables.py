#!/usr/bin/env python3
"""Emit the constant-table headers used by the CUDA library and by the oracle.

Runs ONLY in the development container (it imports the read-only reference at
/root/reference, plus a 7-line `bitarray` stand-in, as the *data source* for the
ISO 11172-3 code books / windows and for the reference's derived fixed-point
tables).  Its outputs are committed:

  mp3-steganography-lib_b200/csrc/m3s_tables_data.h   product layout (packed code books, LUT inputs)
  oracle/oracle_tables.h                              oracle layout (left-aligned code/len rows for the
                                                      first-prefix-match search the reference performs)
  tests/golden/tables_digest.json                     sha256 of canonical serialisations taken from the
                                                      reference objects, checked by tests/test_tables.py

Reference sources of each table (file:line under /root/reference/mp3stego):
  decoder/tables.py:8-29 band tables, :33-40 count1 table A, :46 slen, :48 pre_tab, :64-417 hft_*,
  :419-427 big_value_table/linbit/max, :429-514 synth_window;
  encoder/tables.py:34-78 enwindow, :80-250 tNHB/tNl, :271-304 huffman_table, :308-332 MDCT_CA/CS,
  :335-359 subdv_table; encoder/MP3_Encoder.py:419-449 IDX_TO_TRANSFORM_HUF, :528-579 fl/cos_l/steptab/int2idx;
  decoder/util.py:3 H0; decoder/Frame.py:609-611 cs/ca literals.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = "/tmp/refshim"
os.makedirs(SHIM, exist_ok=True)
with open(os.path.join(SHIM, "bitarray.py"), "w") as f:
    f.write("class bitarray(list):\n    def frombytes(self, b):\n        for byte in b:\n"
            "            for n in range(7, -1, -1):\n                self.append((byte >> n) & 1)\n")
sys.path.insert(0, SHIM)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from mp3stego.decoder import tables as dt  # noqa: E402
from mp3stego.decoder import util as du  # noqa: E402
from mp3stego.encoder import tables as et  # noqa: E402
from mp3stego.encoder import MP3_Encoder as me  # noqa: E402


def c_array(name, ctype, values, per_line=12, fmt="{}"):
    out = [f"static const {ctype} {name}[{len(values)}] = {{"]
    for i in range(0, len(values), per_line):
        out.append("    " + ", ".join(fmt.format(v) for v in values[i:i + per_line]) + ",")
    out.append("};")
    return "\n".join(out)


def digest(values):
    h = hashlib.sha256()
    h.update((",".join(str(v) for v in values)).encode())
    return h.hexdigest()


# ---------------------------------------------------------------- code books
# distinct code books: table id -> (dim, codes right-aligned, lens); 16..23 share 16, 24..31 share 24
books = {}
for t in range(34):
    ht = et.huffman_table[t]
    if ht.table is None:
        continue
    books[t] = (ht.x_len, ht.y_len, list(ht.table), list(ht.h_len), ht.lin_bits, ht.lin_max)

# cross-check encoder code books against the decoder's left-aligned lists
for t in range(1, 32):
    dmax = dt.big_value_max[t]
    if t in (4, 14):
        assert dmax == 0 and t not in books
        continue
    xl, yl, codes, lens, lb, lm = books[t]
    assert xl == yl == dmax, (t, xl, dmax)
    assert lb == dt.big_value_linbit[t]
    flat = dt.big_value_table[t]
    for x in range(dmax):
        for y in range(dmax):
            i = 2 * dmax * x + 2 * y
            code_l, ln = flat[i], flat[i + 1]
            assert ln == lens[x * yl + y], (t, x, y)
            assert code_l >> (32 - ln) == codes[x * yl + y], (t, x, y)
# count1 table A (decoder index 8v+4w+2x+y, same code list as encoder t32)
for e in range(16):
    assert dt.quad_table_1.h_len[e] == et.t32l[e]
    assert dt.quad_table_1.h_cod[e] >> (32 - et.t32l[e]) == et.t32HB[e]
    v = dt.quad_table_1.value[e]
    assert 8 * v[0] + 4 * v[1] + 2 * v[2] + v[3] == e

distinct = [1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 24, 32, 33]
packed = []
book_off = [0] * 34
for t in distinct:
    book_off[t] = len(packed)
    xl, yl, codes, lens, lb, lm = books[t]
    for c, l in zip(codes, lens):
        packed.append((c << 8) | l)
for t in range(17, 24):
    book_off[t] = book_off[16]
for t in range(25, 32):
    book_off[t] = book_off[24]
dims = [0] * 34
linbits = [0] * 34
linmax = [0] * 34
for t, (xl, yl, codes, lens, lb, lm) in books.items():
    dims[t] = yl
    linbits[t] = lb
    linmax[t] = lm

sr_order = ["44", "48", "32"]  # header sampling_frequency index order 0,1,2
sfb_long = []
sfb_short = []
sfw_short = []
for s in sr_order:
    sfb_long += list(getattr(dt.band_index_table, "long_" + s))
    sfb_short += list(getattr(dt.band_index_table, "short_" + s))
    sfw_short += list(getattr(dt.band_width_table, "short_" + s))
from mp3stego.encoder import util as eu  # noqa: E402
for i, s in enumerate(sr_order):
    assert list(getattr(dt.band_index_table, "long_" + s)) == eu.scale_fact_band_index[i]

slen = [v for row in dt.slen for v in row]
assert [r[0] for r in dt.slen] == et.slen1_tab and [r[1] for r in dt.slen] == et.slen2_tab
pretab = [int(v) for v in dt.pre_tab] + [0]
synth = [float(v) for v in dt.synth_window]
enwin = [int(v) for v in et.enwindow]
cs = [.8574929257, .8817419973, .9496286491, .9833145925, .9955178161, .9991605582, .9998991952, .9999931551]
ca = [-.5144957554, -.4717319686, -.3133774542, -.1819131996, -.0945741925, -.0409655829, -.0141985686,
      -.0036999747]
# make sure these literals are the ones in Frame.py
src = open("/root/reference/mp3stego/decoder/Frame.py").read()
for v in cs + ca:
    lit = ("%.10f" % abs(v)).lstrip("0")
    assert lit in src, lit
subdv = [v for row in et.subdv_table for v in row]
pair = [0] * 64
for (t, b), n in me.IDX_TO_TRANSFORM_HUF.items():
    pair[t * 2 + b] = n
h0mask = 0
for t in du.H0:
    h0mask |= 1 << t

# derived encoder tables: instantiate the reference initialisers without a WAV file
enc = me.MP3Encoder.__new__(me.MP3Encoder)
enc._MP3Encoder__sub_band = me.Subband()
enc._MP3Encoder__mdct = me.MDCT()
enc._MP3Encoder__l3loop = me.L3Loop()
enc._MP3Encoder__sub_band_initialise()
enc._MP3Encoder__mdct_initialise()
enc._MP3Encoder__loop_initialise()
fl = [int(v) for v in enc._MP3Encoder__sub_band.fl.flatten()]
cosl = [int(v) for v in enc._MP3Encoder__mdct.cos_l.flatten()]
steptab = [float(v) for v in enc._MP3Encoder__l3loop.steptab]
steptabi = [int(v) for v in enc._MP3Encoder__l3loop.steptabi]
int2idx = [int(v) for v in enc._MP3Encoder__l3loop.int2idx]
mdct_ca = [int(getattr(et, "MDCT_CA%d" % i)) for i in range(8)]
mdct_cs = [int(getattr(et, "MDCT_CS%d" % i)) for i in range(8)]
# int2idx is reproducible from correctly-rounded sqrt alone; verify and do not embed it
import math  # noqa: E402
for i in range(10000):
    assert int2idx[i] == int(np.int32(math.sqrt(math.sqrt(float(i)) * float(i)) - 0.0946 + 0.5))
# steptab[i] = 2^((127-i)/4): exact power of two times one of four quarter-root constants
quarter = [steptab[127], steptab[126], steptab[125], steptab[124]]  # 2^0, 2^.25, 2^.5, 2^.75
for i in range(128):
    e = 127 - i
    assert steptab[i] == math.ldexp(quarter[e % 4], e // 4), i
    if steptab[i] * 2 > 0x7fffffff:
        assert steptabi[i] == 0x7fffffff
    else:
        assert steptabi[i] == int(steptab[i] * 2 + 0.5)

digests = {
    "huff_books": {str(t): digest(books[t][2] + books[t][3]) for t in distinct},
    "huff_linbits": digest(linbits[:32]),
    "huff_dim": digest(dims),
    "sfb_long": digest(sfb_long), "sfb_short": digest(sfb_short), "sfw_short": digest(sfw_short),
    "slen": digest(slen), "pretab": digest(pretab[:21]),
    "synth_window": digest(["%.9f" % v for v in synth]),
    "enwindow": digest(enwin), "subdv": digest(subdv), "pair": digest(pair), "h0mask": h0mask,
    "fl": digest(fl), "cos_l": digest(cosl), "steptabi": digest(steptabi),
    "steptab_hex": digest([v.hex() for v in steptab]), "int2idx": digest(int2idx),
    "mdct_ca": digest(mdct_ca), "mdct_cs": digest(mdct_cs),
    "alias_cs": digest(["%.10f" % v for v in cs]), "alias_ca": digest(["%.10f" % v for v in ca]),
}

# ------------------------------------------------------------ product header
H = []
H.append("""// GENERATED by tools/gen_tables.py -- do not edit.
// Constant data for the mp3stego B200 hot path, in this library's own layout.
// Code books are ISO/IEC 11172-3 Table B.7; element-for-element equality with the
// reference's tables is checked by tests/test_tables.py through tests/golden/tables_digest.json.
#pragma once
#include <stdint.h>
""")
H.append("// packed code books: (right-aligned code << 8) | length, row-major [x][y]; books 16..23 and 24..31 are shared")
H.append(c_array("M3S_HUFF_PACKED", "uint32_t", packed, 8, "0x{:x}"))
H.append("// table id (0..31 big_values, 32/33 count1 A/B) -> first element in M3S_HUFF_PACKED")
H.append(c_array("M3S_HUFF_BOOK_OFF", "uint16_t", book_off, 17))
H.append("// table id -> square dimension of the code book (0: tables 0, 4, 14 carry no codes; 32/33: 16 quads)")
H.append(c_array("M3S_HUFF_DIM", "uint8_t", dims, 17))
H.append(c_array("M3S_HUFF_LINBITS", "uint8_t", linbits, 17))
H.append(c_array("M3S_HUFF_LINMAX", "uint16_t", linmax, 17))
H.append("// scalefactor band boundaries, rows in header sampling_frequency index order: 44100, 48000, 32000")
H.append(c_array("M3S_SFB_LONG", "uint16_t", sfb_long, 23))
H.append(c_array("M3S_SFB_SHORT", "uint16_t", sfb_short, 14))
H.append(c_array("M3S_SFW_SHORT", "uint8_t", sfw_short, 12))
H.append(c_array("M3S_SLEN", "uint8_t", slen, 16))
H.append("// pre-emphasis table, padded with one zero (the reference guards sfb >= 21 to 0)")
H.append(c_array("M3S_PRETAB", "uint8_t", pretab, 22))
H.append("// synthesis window D[512] (9-digit literals as the reference carries them)")
H.append(c_array("M3S_SYNTH_WINDOW", "double", synth, 6, "{:.9f}"))
H.append("// decoder alias-reduction literals")
H.append(c_array("M3S_ALIAS_CS", "double", cs, 4, "{:.10f}"))
H.append(c_array("M3S_ALIAS_CA", "double", ca, 4, "{:.10f}"))
H.append("// encoder analysis window (fixed point)")
H.append(c_array("M3S_ENWINDOW", "int32_t", enwin, 10))
H.append("// encoder polyphase matrix fl[32][64] and MDCT matrix cos_l[18][36] (window folded in), Q31")
H.append(c_array("M3S_ENC_FL", "int32_t", fl, 8))
H.append(c_array("M3S_ENC_COSL", "int32_t", cosl, 9))
H.append(c_array("M3S_ENC_ALIAS_CA", "int32_t", mdct_ca, 8))
H.append(c_array("M3S_ENC_ALIAS_CS", "int32_t", mdct_cs, 8))
H.append("// 2^(0/4), 2^(1/4), 2^(2/4), 2^(3/4) as the reference's libm produced them; steptab[i] = ldexp(q[(127-i)%4], (127-i)/4)")
H.append(c_array("M3S_ENC_QUARTER", "double", [v.hex() for v in quarter], 4))
H.append(c_array("M3S_ENC_STEPTABI", "int32_t", steptabi, 8))
H.append("// region split table: big_values band count -> (region0_count, region1_count)")
H.append(c_array("M3S_SUBDV", "uint8_t", subdv, 16))
H.append("// stego pair map [table][payload bit] -> table actually written")
H.append(c_array("M3S_STEGO_PAIR", "uint8_t", pair, 16))
H.append("// bit t set  <=>  table t reveals payload bit '0'")
H.append(f"#define M3S_H0_MASK 0x{h0mask:08x}u")
with open(os.path.join(ROOT, "mp3-steganography-lib_b200/csrc/m3s_tables_data.h"), "w") as f:
    f.write("\n".join(H) + "\n")

# ------------------------------------------------------------- oracle header
O = []
O.append("""/* GENERATED by tools/gen_tables.py -- do not edit.
 * Tables for the CPU oracle, laid out for the reference's own search procedures:
 * big-value code books as left-aligned (code, len) rows in row-major (x, y) order, which is what
 * Frame.__unpack_samples (decoder/Frame.py:491-517) scans for the first prefix match. */
#pragma once
#include <stdint.h>
""")
left_codes = []
left_lens = []
obook_off = [0] * 32
odim = [0] * 32
seen = {}
for t in range(32):
    odim[t] = dt.big_value_max[t] if t != 0 else 0
    if t in (0, 4, 14):
        continue
    key = id(dt.big_value_table[t])
    if key in seen:
        obook_off[t] = seen[key]
        continue
    seen[key] = len(left_codes)
    obook_off[t] = len(left_codes)
    flat = dt.big_value_table[t]
    for i in range(0, len(flat), 2):
        left_codes.append(flat[i])
        left_lens.append(flat[i + 1])
O.append(c_array("ORA_HUFF_CODE_L", "uint32_t", left_codes, 8, "0x{:08x}u"))
O.append(c_array("ORA_HUFF_LEN", "uint8_t", left_lens, 24))
O.append(c_array("ORA_HUFF_OFF", "uint16_t", obook_off, 16))
O.append(c_array("ORA_HUFF_MAX", "uint8_t", odim, 16))
O.append(c_array("ORA_HUFF_LINBITS", "uint8_t", list(dt.big_value_linbit), 16))
O.append(c_array("ORA_QUAD_CODE_L", "uint32_t", list(dt.quad_table_1.h_cod), 8, "0x{:08x}u"))
O.append(c_array("ORA_QUAD_LEN", "uint8_t", list(dt.quad_table_1.h_len), 16))
O.append(c_array("ORA_SFB_LONG", "int", sfb_long, 23))
O.append(c_array("ORA_SFB_SHORT", "int", sfb_short, 14))
O.append(c_array("ORA_SFW_SHORT", "int", sfw_short, 12))
O.append(c_array("ORA_SLEN", "int", slen, 16))
O.append(c_array("ORA_PRETAB", "int", pretab[:21], 21))
O.append(c_array("ORA_SYNTH_WINDOW", "double", synth, 6, "{:.9f}"))
O.append(c_array("ORA_ALIAS_CS", "double", cs, 4, "{:.10f}"))
O.append(c_array("ORA_ALIAS_CA", "double", ca, 4, "{:.10f}"))
# encoder side: ISO-layout code/len per table id (as encoder/tables.py:271-304 indexes them)
enc_codes = []
enc_lens = []
enc_off = [0] * 34
for t in distinct:
    enc_off[t] = len(enc_codes)
    enc_codes += books[t][2]
    enc_lens += books[t][3]
for t in range(17, 24):
    enc_off[t] = enc_off[16]
for t in range(25, 32):
    enc_off[t] = enc_off[24]
O.append(c_array("ORA_ENC_CODE", "uint32_t", enc_codes, 16))
O.append(c_array("ORA_ENC_HLEN", "uint8_t", enc_lens, 24))
O.append(c_array("ORA_ENC_OFF", "uint16_t", enc_off, 17))
O.append(c_array("ORA_ENC_XLEN", "int", [et.huffman_table[t].x_len for t in range(34)], 17))
O.append(c_array("ORA_ENC_YLEN", "int", [et.huffman_table[t].y_len for t in range(34)], 17))
O.append(c_array("ORA_ENC_LINBITS", "int", linbits, 17))
O.append(c_array("ORA_ENC_LINMAX", "int", linmax, 17))
O.append(c_array("ORA_ENWINDOW", "int32_t", enwin, 10))
O.append(c_array("ORA_ENC_FL", "int32_t", fl, 8))
O.append(c_array("ORA_ENC_COSL", "int32_t", cosl, 9))
O.append(c_array("ORA_ENC_CA", "int32_t", mdct_ca, 8))
O.append(c_array("ORA_ENC_CS", "int32_t", mdct_cs, 8))
O.append(c_array("ORA_ENC_STEPTAB", "double", [v.hex() for v in steptab], 4))
O.append(c_array("ORA_ENC_STEPTABI", "int32_t", steptabi, 8))
O.append(c_array("ORA_SUBDV", "int", subdv, 16))
O.append(c_array("ORA_PAIR", "int", pair, 16))
O.append(f"#define ORA_H0_MASK 0x{h0mask:08x}u")
with open(os.path.join(ROOT, "oracle/oracle_tables.h"), "w") as f:
    f.write("\n".join(O) + "\n")

with open(os.path.join(ROOT, "tests/golden/tables_digest.json"), "w") as f:
    json.dump(digests, f, indent=1, sort_keys=True)
print("packed entries", len(packed), "oracle code rows", len(left_codes))
