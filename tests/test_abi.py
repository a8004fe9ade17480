"""CPU: the C-ABI library loads and exports every symbol include/mp3stego_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mp3stego_b200.h")).read()
    return sorted(set(re.findall(r"M3S_API\s+[\w\s\*]+?\b(m3s_\w+)\s*\(", hdr)))


def test_header_symbols_exported(built):
    from mp3stego_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) == bound, (set(names) ^ bound)


def test_no_device_is_loud(built):
    """Without a GPU the product must fail loudly, never fall back to a CPU path."""
    import torch
    from mp3stego_b200 import _lib
    if torch.cuda.is_available():
        return
    try:
        _lib.Handle(0)
    except _lib.M3SError as e:
        assert "no CPU fallback" in str(e) or "failed" in str(e)
    else:
        raise AssertionError("Handle() succeeded without a CUDA device")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mp3-steganography-lib_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("oracle's", "").replace("the oracle", "") or f.endswith((".cu", ".cuh", ".h")), f
                assert "import oracle" not in src and "from oracle" not in src, f
