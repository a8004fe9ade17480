"""GPU: the ctypes stub of INTEGRATION.md, verbatim, patched into the UNMODIFIED reference (baseline/_ref, installed by
__graft_entry__.build() where /root/reference exists; it travels to the GPU box with the snapshot): the reference's own five
test cases pass, every codec call went through the stub, and the artefacts equal the pure-Python reference's by sha256."""
import json
import os
import subprocess
import sys

import pytest

from conftest import PKG, ROOT, golden_path

pytestmark = pytest.mark.gpu


def test_reference_classes_over_the_stub(built):
    ref = os.path.join(ROOT, "baseline", "_ref", "mp3stego")
    if not os.path.isdir(ref):
        pytest.skip("baseline/_ref is absent (built only where /root/reference exists)")
    try:
        import numba  # noqa: F401  (the reference imports it at module level)
    except ImportError:
        pytest.skip("numba is not installed on this box: the reference cannot be imported")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(PKG, "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "integration_stub_runner.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    out = json.loads(line[7:])
    fac = json.load(open(golden_path("ref_facade.json")))
    assert out["ran"] == 5 and not out["failures"], out["failures"]
    assert out["calls"]["parse_file"] >= 8 and out["calls"]["encode"] >= 6      # the five tests' decodes / encodes all went through the stub
    assert out["decode_returns"] == fac["decode_returns"]
    for k in ("out_wav_sha256", "enc320_sha256", "enc128_sha256", "hid_sha256", "cleared_sha256"):
        assert out[k] == fac[k], k
    assert out["hide_ddd_returns"] == fac["hide_ddd_returns"] and out["reveal_hid"] == fac["reveal_hid"]
    assert out["batch_binding_equals_single"]      # INTEGRATION.md's m3s_decode binding: N files in one call == N single-file calls
