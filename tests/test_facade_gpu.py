"""GPU: the reference's own five end-to-end tests (tests/steganography_test.py:15-60) run against the mirror facade,
plus the stronger byte-level goldens of SURVEY.md 8(c): with the float64 (`exact`) decode instantiation the WAV, the
hidden and the cleared MP3 equal the reference's artefacts by sha256."""
import hashlib
import json
import os
import shutil

import pytest

from conftest import golden_path

pytestmark = pytest.mark.gpu


@pytest.fixture()
def work(tmp_path, built):
    shutil.copy(golden_path("test.mp3"), tmp_path / "test.mp3")
    return tmp_path


def _sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def test_decode(work):
    from mp3stego_b200 import Steganography
    fac = json.load(open(golden_path("ref_facade.json")))
    s = Steganography(quiet=True)
    assert s.decode_mp3_to_wav(str(work / "test.mp3"), str(work / "out.wav")) == 320 == fac["decode_returns"]
    assert os.path.getsize(work / "out.wav") == fac["out_wav_bytes"]
    assert _sha(work / "out.wav") == fac["out_wav_sha256"]


def test_hide_short_and_long(work):
    from mp3stego_b200 import Steganography
    fac = json.load(open(golden_path("ref_facade.json")))
    s = Steganography(quiet=True)
    assert s.hide_message(str(work / "test.mp3"), str(work / "out.mp3"), "ddd") is False
    assert _sha(work / "out.mp3") == fac["hid_sha256"]
    assert not os.path.exists(work / "test.wav")          # the temp WAV next to the input is deleted (steganography.py:159)
    assert s.hide_message(str(work / "test.mp3"), str(work / "long.mp3"), "ddd" * 100) is True
    assert _sha(work / "long.mp3") == fac["hid_long_sha256"]


def test_hide_reveal_clear(work):
    from mp3stego_b200 import Steganography
    fac = json.load(open(golden_path("ref_facade.json")))
    s = Steganography(quiet=True)
    s.hide_message(str(work / "test.mp3"), str(work / "out.mp3"), "ddd")
    s.reveal_massage(str(work / "out.mp3"), str(work / "reveal.txt"))
    assert open(work / "reveal.txt").read() == "ddd"
    s.clear_file(str(work / "out.mp3"), str(work / "cleared.mp3"))
    assert _sha(work / "cleared.mp3") == fac["cleared_sha256"]
    s.reveal_massage(str(work / "cleared.mp3"), str(work / "reveal.txt"))
    assert open(work / "reveal.txt").read() == ""
    s.reveal_massage(str(work / "test.mp3"), str(work / "r0.txt"))
    assert open(work / "r0.txt").read() == fac["reveal_test_mp3"]


def test_encode_wav(work):
    from mp3stego_b200 import Steganography
    fac = json.load(open(golden_path("ref_facade.json")))
    shutil.copy(golden_path("ref_test_out.wav"), work / "in.wav")
    s = Steganography(quiet=True)
    s.encode_wav_to_mp3(str(work / "in.wav"), str(work / "e320.mp3"))
    assert _sha(work / "e320.mp3") == fac["enc320_sha256"]
    s.encode_wav_to_mp3(str(work / "in.wav"), str(work / "e128.mp3"), 128)
    assert _sha(work / "e128.mp3") == fac["enc128_sha256"] and os.path.getsize(work / "e128.mp3") == fac["enc128_bytes"]


def test_reference_encoder_input_contract(work):
    """Mono and ragged-length WAVs raise IndexError in the reference (SURVEY A.E1/A.E2); so does the mirror."""
    import numpy as np
    from mp3stego_b200 import Encoder
    from mp3stego_b200.wavio import write_wav
    write_wav(str(work / "mono.wav"), 44100, np.zeros(2304, np.int16))
    with pytest.raises(IndexError):
        Encoder(str(work / "mono.wav"), str(work / "m.mp3")).encode()
    write_wav(str(work / "ragged.wav"), 44100, np.zeros((1152 + 100, 2), np.int16))
    with pytest.raises(IndexError):
        Encoder(str(work / "ragged.wav"), str(work / "r.mp3")).encode()


def test_exact_decode_equals_reference_int16(handle):
    """The float64 instantiation reproduces the reference's int16 samples exactly on every golden stream."""
    import glob
    import numpy as np
    from conftest import GOLDEN, load_npz
    z = load_npz("ref_test_mp3.npz")
    data = np.frombuffer(open(golden_path("test.mp3"), "rb").read(), np.uint8)
    handle.decode_scan(data, [0, len(data)])
    pcm, _ = handle.decode_run(exact=True)
    assert np.array_equal(pcm.reshape(-1, 2), z["pcm16"])
    for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_synth_*.npz"))):
        zz = np.load(p)
        d = zz["mp3"]
        handle.decode_scan(d, [0, len(d)])
        pcm, _ = handle.decode_run(exact=True)
        assert np.array_equal(pcm.reshape(-1, 2), zz["dec_pcm16"]), p
