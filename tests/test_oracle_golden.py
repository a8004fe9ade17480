"""CPU: pins the oracle (oracle/mp3stego_oracle.c) against golden vectors produced by the UNMODIFIED Python
reference (tests/golden/make_golden.py, run in the development container).  Integer/byte outputs must be
bit-exact; float64 PCM within 1e-12 (different libm call order is the only slack)."""
import hashlib
import json

import numpy as np
import pytest

from conftest import SYNTH_CASES, golden_path, load_npz


def test_decode_test_mp3(oracle):
    z = load_npz("ref_test_mp3.npz")
    data = open(golden_path("test.mp3"), "rb").read()
    r = oracle.decode(data)
    assert r["n_frames"] == int(z["n_frames"]) == 36
    assert r["bit_rate"] == int(z["bitrate"]) and r["sampling_rate"] == int(z["sampling_rate"])
    assert np.array_equal(r["spectra"], z["spectra"].astype(np.int32))
    assert np.array_equal(r["tables"], z["tables"])
    assert r["bits"] == str(z["bits"])
    assert np.array_equal(r["pcm16"], z["pcm16"])
    for k, f in enumerate(z["pcm64_frames"]):
        got = r["pcm"][f * 1152:(f + 1) * 1152]
        assert np.max(np.abs(got - z["pcm64"][k * 1152:(k + 1) * 1152])) < 1e-12
    for k, f in enumerate(z["xr_frames"]):
        assert np.max(np.abs(r["xr"][f] - z["xr"][k])) <= 1e-12 * max(1.0, np.max(np.abs(z["xr"][k])))


def test_decode_side_info_fields(oracle):
    z = load_npz("ref_test_mp3.npz")
    r = oracle.decode(open(golden_path("test.mp3"), "rb").read())
    n = min(r["side"].shape[-1], z["side"].shape[-1])
    assert np.array_equal(r["side"][..., :n], z["side"].astype(np.int32)[..., :n])


@pytest.mark.parametrize("case", SYNTH_CASES)
def test_encode_synth(oracle, case):
    z = load_npz(f"ref_synth_{case}.npz")
    r = oracle.encode(z["pcm_in"], 44100, int(z["bitrate"]), str(z["hide_bits"]))
    assert r["status"] == 0
    assert r["mp3"] == z["mp3"].tobytes()
    assert r["hide_str_offset"] == int(z["hide_str_offset"])
    assert np.array_equal(r["mdct"][:3], z["mdct"])
    assert np.array_equal(r["ix"], z["ix"].astype(np.int32))
    assert np.array_equal(r["info"][..., :16], z["info"])
    assert np.array_equal(r["scfsi"], z["scfsi"].astype(np.int32))


@pytest.mark.parametrize("case", SYNTH_CASES)
def test_decode_synth(oracle, case):
    z = load_npz(f"ref_synth_{case}.npz")
    r = oracle.decode(z["mp3"].tobytes())
    assert r["n_frames"] == int(z["dec_n_frames"])
    assert np.array_equal(r["spectra"], z["dec_spectra"].astype(np.int32))
    assert np.array_equal(r["tables"], z["dec_tables"])
    assert r["bits"] == str(z["dec_bits"])
    assert np.array_equal(r["pcm16"], z["dec_pcm16"])


def _wav_pcm(path):
    raw = open(path, "rb").read()
    i = raw.find(b"data")
    return np.frombuffer(raw[i + 8:], dtype=np.int16).reshape(-1, 2)


def test_facade_composites(oracle):
    """decode -> encode composites reproduce the reference facade artefacts byte for byte."""
    fac = json.load(open(golden_path("ref_facade.json")))
    data = open(golden_path("test.mp3"), "rb").read()
    assert hashlib.sha256(data).hexdigest() == fac["test_mp3_sha256"]
    d = oracle.decode(data, taps=False)
    wav = _wav_pcm(golden_path("ref_test_out.wav"))
    assert np.array_equal(d["pcm16"], wav)
    assert d["bit_rate"] // 1000 == fac["decode_returns"]
    assert oracle.reveal_parse(d["bits"]) == fac["reveal_test_mp3"]
    for br, key in ((320, "enc320"), (128, "enc128")):
        e = oracle.encode(wav, 44100, br, "", taps=False)
        assert hashlib.sha256(e["mp3"]).hexdigest() == fac[f"{key}_sha256"]
        assert len(e["mp3"]) == fac[f"{key}_bytes"]
    # hide_message = decode + encode(hide) at the decoded bitrate (steganography.py:137-162)
    bits = oracle.str_to_bits("3#ddd")
    e = oracle.encode(wav, 44100, 320, bits, taps=False)
    assert hashlib.sha256(e["mp3"]).hexdigest() == fac["hid_sha256"]
    assert (e["hide_str_offset"] < len(bits) - 1) == fac["hide_ddd_returns"]
    dh = oracle.decode(e["mp3"], taps=False)
    assert oracle.reveal_parse(dh["bits"]) == fac["reveal_hid"] == "ddd"
    # clear_file = decode + plain encode (steganography.py:164-182)
    c = oracle.encode(dh["pcm16"], 44100, 320, "", taps=False)
    assert hashlib.sha256(c["mp3"]).hexdigest() == fac["cleared_sha256"]
    assert oracle.reveal_parse(oracle.decode(c["mp3"], taps=False)["bits"]) == fac["reveal_cleared"] == ""
    long_bits = oracle.str_to_bits("300#" + "ddd" * 100)
    el = oracle.encode(wav, 44100, 320, long_bits, taps=False)
    assert hashlib.sha256(el["mp3"]).hexdigest() == fac["hid_long_sha256"]
    assert (el["hide_str_offset"] < len(long_bits) - 1) == fac["hide_long_returns"]


STREAM_CASES = ["long_alltables", "reservoir", "short_mixed", "ms_stereo", "is_only_bit", "mono_crc_48k", "vbr_32k_pad",
                "loud_wrap"]


@pytest.mark.parametrize("case", STREAM_CASES)
def test_decode_writer_streams(oracle, case):
    """BASELINE configs[3]: streams from the test-bitstream writer (tests/golden/make_streams.py) that use what the
    reference encoder cannot emit -- all Huffman tables, scalefactors + scfsi, short / mixed / start / stop blocks, MS
    stereo (and the ignored intensity bit), the bit reservoir, CRC, mono, 32/48 kHz, VBR, padding, int16 wrap -- against
    the unmodified reference decoder's output."""
    z = load_npz(f"ref_stream_{case}.npz")
    r = oracle.decode(open(golden_path(f"stream_{case}.mp3"), "rb").read())
    assert r["n_frames"] == int(z["n_frames"])
    assert r["bit_rate"] == int(z["bitrate"]) and r["sampling_rate"] == int(z["sampling_rate"])
    assert np.array_equal(r["spectra"], z["spectra"].astype(np.int32))
    k = z["tables"].shape[1]
    assert np.array_equal(r["tables"][:, :k], z["tables"]) and not r["tables"][:, k:].any()
    assert r["bits"] == str(z["bits"])
    assert np.array_equal(r["pcm16"].reshape(z["pcm16"].shape), z["pcm16"])


def _fuzz_cases():
    import os
    return sorted(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fuzz.json"))))


@pytest.mark.parametrize("name", _fuzz_cases())
def test_decode_fuzz_corpus(oracle, name):
    """BASELINE configs[3] as a corpus: 16 writer streams over random combinations of sample rate, CBR / VBR, padding, mono /
    stereo / joint stereo, CRC, bit reservoir, short / mixed / start / stop blocks, scalefactors + scfsi and loud spectra
    (tests/golden/make_streams.py --fuzz), decoded by the unmodified reference: digests of its int16 PCM, integer spectra, table ids,
    and its reveal bits."""
    ref = json.load(open(golden_path("ref_fuzz.json")))[name]
    r = oracle.decode(open(golden_path(name + ".mp3"), "rb").read())
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    assert r["n_frames"] == ref["n_frames"] and r["bit_rate"] == ref["bitrate"] and r["sampling_rate"] == ref["sampling_rate"]
    assert r["bits"] == ref["bits"]
    assert sha(r["tables"][:, :6 * ref["channels"]].astype(np.uint8)) == ref["tables_sha256"]   # the reference lists 6 ids per channel
    assert sha(r["spectra"].astype(np.int16)) == ref["spectra_sha256"]
    assert sha(r["pcm16"].astype(np.int16)) == ref["pcm16_sha256"]


def _edge_cases():
    import os
    return sorted(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_edge.json"))))


def edge_audio_start(blob):
    """ID3v2 skip as Decoder.__init__ computes it (decoder.py:29-33)."""
    if blob[:3] == b"ID3" and not blob[5] & 0x0F:
        size = 0
        for i in range(4):
            size = (size << 7) + blob[6 + i]
        return size + (20 if blob[5] & 0x10 else 10)
    return 0


@pytest.mark.parametrize("name", _edge_cases())
def test_decode_edge_reservoirs(oracle, name):
    """Files whose first frames point into a bit reservoir that is not there (cut out of a longer stream, bare or behind an ID3v2
    tag) and files whose header + side-info length changes under a live reservoir (CRC on / off): the reference assembles the
    bytes physically in front of the frame, or keeps the previous frame's main data (Frame.py:318-363, A.D9).  Goldens by the
    unmodified reference decoder (tests/golden/make_streams.py --edge)."""
    ref = json.load(open(golden_path("ref_edge.json")))[name]
    blob = open(golden_path(name + ".mp3"), "rb").read()
    r = oracle.decode(blob, edge_audio_start(blob))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    assert r["n_frames"] == ref["n_frames"] and r["bit_rate"] == ref["bitrate"] and r["sampling_rate"] == ref["sampling_rate"]
    assert r["bits"] == ref["bits"]
    assert sha(r["tables"][:, :6 * ref["channels"]].astype(np.uint8)) == ref["tables_sha256"]
    assert sha(r["spectra"].astype(np.int16)) == ref["spectra_sha256"]
    assert sha(r["pcm16"].astype(np.int16)) == ref["pcm16_sha256"]
