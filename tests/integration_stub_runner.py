"""Runs in a subprocess of tests/test_integration_stub.py: patches the UNMODIFIED reference (baseline/_ref) with the ctypes stub and
the two call-site replacements printed in INTEGRATION.md -- taken VERBATIM from the markdown -- and runs the reference's own test
cases (tests/steganography_test.py:15-60) plus a byte comparison with the artefacts the pure-Python reference produced
(tests/golden/ref_facade.json).  Needs a B200; prints one JSON object."""
import hashlib
import json
import os
import re
import shutil
import sys
import tempfile
import types
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    sys.path.insert(0, REF)
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    stub_src = next(b for b in blocks if "mp3stego/_b200.py" in b.splitlines()[0] and "continued" not in b.splitlines()[0])
    batch_src = next(b for b in blocks if "mp3stego/_b200.py (continued)" in b.splitlines()[0])
    sites_src = next(b for b in blocks if "mp3stego/decoder/MP3_Parser.py" in b.splitlines()[0])
    import numpy as np
    import tqdm
    import mp3stego
    import mp3stego.decoder.MP3_Parser as mp
    import mp3stego.encoder.MP3_Encoder as me
    mp.tqdm = lambda *a, **k: tqdm.tqdm(*a, **{**k, "disable": True})
    me.tqdm = lambda *a, **k: tqdm.tqdm(*a, **{**k, "disable": True})
    stub = types.ModuleType("mp3stego._b200")          # "new file in the reference tree"
    exec(compile(stub_src, "INTEGRATION.md:_b200", "exec"), stub.__dict__)
    sys.modules["mp3stego._b200"] = stub
    mp3stego._b200 = stub
    from scipy.io.wavfile import write
    ns = {"_b200": stub, "np": np, "write": write}
    exec(compile(sites_src, "INTEGRATION.md:call_sites", "exec"), ns)
    calls = {"parse_file": 0, "encode": 0}

    def counted(name, fn):
        def w(self, *a, **k):
            calls[name] += 1
            return fn(self, *a, **k)
        return w

    mp.MP3Parser.parse_file = counted("parse_file", ns["parse_file"])
    mp.MP3Parser.write_to_wav = ns["write_to_wav"]
    mp.MP3Parser.get_bitrate = ns["get_bitrate"]
    me.MP3Encoder.encode = counted("encode", ns["encode"])

    work = tempfile.mkdtemp(prefix="m3s_stub_")
    os.makedirs(os.path.join(work, "tests"))
    shutil.copy(os.path.join(ROOT, "tests", "golden", "test.mp3"), os.path.join(work, "tests", "test.mp3"))
    os.chdir(work)
    from tests import steganography_test as ref_tests     # baseline/_ref/tests: the reference's own test module
    suite = unittest.defaultTestLoader.loadTestsFromModule(ref_tests)
    res = unittest.TextTestRunner(stream=open(os.devnull, "w"), verbosity=0).run(suite)
    out = {"ran": res.testsRun, "failures": [str(f[1])[-400:] for f in res.failures + res.errors], "calls": calls}
    # the same artefacts as SURVEY 8(c), by sha256, through the patched reference classes
    s = mp3stego.Steganography(quiet=True)
    sha = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()   # noqa: E731
    out["decode_returns"] = s.decode_mp3_to_wav("tests/test.mp3", "tests/o.wav")
    out["out_wav_sha256"] = sha("tests/o.wav")
    s.encode_wav_to_mp3("tests/o.wav", "tests/e320.mp3", 320)
    out["enc320_sha256"] = sha("tests/e320.mp3")
    s.encode_wav_to_mp3("tests/o.wav", "tests/e128.mp3", 128)
    out["enc128_sha256"] = sha("tests/e128.mp3")
    out["hide_ddd_returns"] = s.hide_message("tests/test.mp3", "tests/hid.mp3", "ddd")
    out["hid_sha256"] = sha("tests/hid.mp3")
    s.clear_file("tests/hid.mp3", "tests/cleared.mp3")
    out["cleared_sha256"] = sha("tests/cleared.mp3")
    s.reveal_massage("tests/hid.mp3", "tests/r.txt")
    out["reveal_hid"] = open("tests/r.txt").read()
    # the batch binding of INTEGRATION.md (m3s_decode: one pipelined call for N files) against the single-file stub
    exec(compile(batch_src, "INTEGRATION.md:_b200_batch", "exec"), stub.__dict__)
    blobs = [open(p, "rb").read() for p in ("tests/test.mp3", "tests/hid.mp3", "tests/cleared.mp3")]
    many = stub.parse_files(blobs, [0, 0, 0])
    same = True
    for b, m in zip(blobs, many):
        one = stub.parse_file(b, 0)
        same = same and one[0] == m[0] and np.array_equal(one[1], m[1]) and one[2:] == m[2:]
    out["batch_binding_equals_single"] = bool(same)
    os.chdir(ROOT)
    shutil.rmtree(work, ignore_errors=True)
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
