"""CPU: every constant table the CUDA library holds (exported through m3s_table_export, host-only) hashes to the
digest taken from the reference's own Python objects by tools/gen_tables.py (tests/golden/tables_digest.json)."""
import hashlib
import json
import os
import re

import numpy as np

from conftest import ROOT, golden_path


def digest(values):
    return hashlib.sha256((",".join(str(v) for v in values)).encode()).hexdigest()


def _ids():
    hdr = open(os.path.join(ROOT, "include", "mp3stego_b200.h")).read()
    enum = re.search(r"enum\s*\{\s*(M3S_TAB_HUFF_PACKED.*?)M3S_TAB_COUNT", hdr, re.S).group(1)
    return [t.strip() for t in enum.replace("= 0", "").split(",") if t.strip()]


def _export(L, which, cap=20000):
    buf = np.zeros(cap, np.float64)
    n = L.m3s_table_export(which, buf.ctypes.data, cap)
    assert 0 < n <= cap
    return buf[:n]


def test_tables_match_reference_digests(built):
    from mp3stego_b200 import _lib
    L = _lib.load()
    ids = _ids()
    D = json.load(open(golden_path("tables_digest.json")))
    T = {name: _export(L, i) for i, name in enumerate(ids)}
    ints = lambda name: [int(v) for v in T[name]]  # noqa: E731
    assert digest(ints("M3S_TAB_HUFF_DIM")) == D["huff_dim"]
    assert digest(ints("M3S_TAB_HUFF_LINBITS")[:32]) == D["huff_linbits"]
    packed, off, dim = ints("M3S_TAB_HUFF_PACKED"), ints("M3S_TAB_HUFF_BOOK_OFF"), ints("M3S_TAB_HUFF_DIM")
    for t, want in D["huff_books"].items():
        t = int(t)
        n = dim[t] * dim[t] if t < 32 else 16
        rows = packed[off[t]: off[t] + n]
        assert digest([r >> 8 for r in rows] + [r & 0xFF for r in rows]) == want, f"code book {t}"
    for t in range(17, 24):
        assert off[t] == off[16]
    for t in range(25, 32):
        assert off[t] == off[24]
    assert digest(ints("M3S_TAB_SFB_LONG")) == D["sfb_long"]
    assert digest(ints("M3S_TAB_SFB_SHORT")) == D["sfb_short"]
    assert digest(ints("M3S_TAB_SFW_SHORT")) == D["sfw_short"]
    assert digest(ints("M3S_TAB_SLEN")) == D["slen"]
    assert digest(ints("M3S_TAB_PRETAB")) == D["pretab"]
    assert digest(["%.9f" % v for v in T["M3S_TAB_SYNTH_WINDOW"]]) == D["synth_window"]
    assert digest(ints("M3S_TAB_ENWINDOW")) == D["enwindow"]
    assert digest(ints("M3S_TAB_ENC_FL")) == D["fl"]
    assert digest(ints("M3S_TAB_ENC_COSL")) == D["cos_l"]
    assert digest(ints("M3S_TAB_ENC_STEPTABI")) == D["steptabi"]
    assert digest([float(v).hex() for v in T["M3S_TAB_ENC_STEPTAB"]]) == D["steptab_hex"]
    assert digest(ints("M3S_TAB_ENC_INT2IDX")) == D["int2idx"]
    assert digest(ints("M3S_TAB_ENC_CA")) == D["mdct_ca"]
    assert digest(ints("M3S_TAB_ENC_CS")) == D["mdct_cs"]
    assert digest(ints("M3S_TAB_SUBDV")) == D["subdv"]
    assert digest(ints("M3S_TAB_STEGO_PAIR")) == D["pair"]
    assert digest(["%.10f" % v for v in T["M3S_TAB_ALIAS_CS"]]) == D["alias_cs"]
    assert digest(["%.10f" % v for v in T["M3S_TAB_ALIAS_CA"]]) == D["alias_ca"]
    assert ints("M3S_TAB_H0_MASK")[0] == D["h0mask"]


def test_known_answers_fixed_point(oracle):
    """SURVEY.md A.E10 known answers for the encoder's integer primitives, via the oracle's exported helpers."""
    import ctypes
    L = oracle.lib()
    if not hasattr(L, "ora_fixed_kat"):
        import pytest
        pytest.skip("oracle built without ora_fixed_kat")
    L.ora_fixed_kat.restype = ctypes.c_int
    assert L.ora_fixed_kat() == 0
