"""GPU: the batched file-level API (mp3stego_b200/files.py): paths in -> files out, equal to what the per-file
facade (and therefore the reference, tests/test_facade_gpu.py) writes."""
import hashlib
import json
import os
import shutil

import numpy as np
import pytest

from conftest import golden_path, synth_wav

pytestmark = pytest.mark.gpu


def _sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def test_decode_and_encode_files_match_the_reference_artefacts(handle, tmp_path):
    from mp3stego_b200 import files
    from mp3stego_b200.wavio import write_wav
    fac = json.load(open(golden_path("ref_facade.json")))
    names = ["test.mp3", "ref_test_hid.mp3", "ref_test_cleared.mp3", "stream_mono_crc_48k.mp3", "stream_vbr_32k_pad.mp3"]
    mp3s = []
    for n in names:
        shutil.copy(golden_path(n), tmp_path / n)
        mp3s.append(str(tmp_path / n))
    res = files.decode_files(handle, mp3s, reveal=True, threads=3)
    assert [r["message"] for r in res[:3]] == [fac["reveal_test_mp3"], fac["reveal_hid"], fac["reveal_cleared"]]
    assert res[0]["bitrate"] == 320 and res[0]["n_frames"] == 36 and res[3]["channels"] == 1 and res[3]["sample_rate"] == 48000
    assert _sha(str(tmp_path / "test.wav")) == fac["out_wav_sha256"]            # the reference's own WAV, byte for byte
    for n in names[3:]:                                                        # writer streams: the reference decoder's int16 PCM
        z = np.load(golden_path("ref_" + n[:-4] + ".npz"))
        from scipy.io import wavfile
        sr, pcm = wavfile.read(str(tmp_path / (n[:-4] + ".wav")))
        assert sr == int(z["sampling_rate"]) and np.array_equal(pcm.reshape(z["pcm16"].shape), z["pcm16"])
    # encode: the decoded WAV back to 320 / 128 kbps and hiding 'ddd' -> the reference's bytes
    wav = str(tmp_path / "test.wav")
    outs = [str(tmp_path / "a.mp3"), str(tmp_path / "b.mp3")]
    shutil.copy(wav, tmp_path / "test2.wav")
    long = files.encode_files(handle, [wav, str(tmp_path / "test2.wav")], outs, bitrate=320, messages=["ddd", "ddd" * 100])
    assert long == [False, True]
    assert _sha(outs[0]) == fac["hid_sha256"] and _sha(outs[1]) == fac["hid_long_sha256"]
    files.encode_files(handle, [wav], [str(tmp_path / "c.mp3")], bitrate=128)
    assert _sha(str(tmp_path / "c.mp3")) == fac["enc128_sha256"]
    # a mixed batch: two sample rates in one call
    w48 = str(tmp_path / "w48.wav")
    write_wav(w48, 48000, synth_wav(3, 6, sr=48000))
    files.encode_files(handle, [wav, w48], [str(tmp_path / "d.mp3"), str(tmp_path / "e.mp3")], bitrate=320)
    assert _sha(str(tmp_path / "d.mp3")) == fac["enc320_sha256"]
    back = files.decode_files(handle, [str(tmp_path / "e.mp3")], wav_paths=False)
    assert back[0]["sample_rate"] == 48000 and back[0]["n_frames"] == 6


def test_file_api_failures_follow_the_reference(handle, tmp_path):
    from mp3stego_b200 import files
    with pytest.raises(SystemExit, match="not found"):
        files.decode_files(handle, [str(tmp_path / "nope.mp3")])
    from mp3stego_b200.wavio import write_wav
    bad = str(tmp_path / "ragged.wav")
    write_wav(bad, 44100, synth_wav(1, 2)[:-100])
    with pytest.raises(IndexError):
        files.encode_files(handle, [bad], [str(tmp_path / "x.mp3")])
    with pytest.raises(SystemExit, match="bitrate"):
        files.encode_files(handle, [bad], [str(tmp_path / "x.mp3")], bitrate=100)
