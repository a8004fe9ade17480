"""GPU: the batch composites (mp3stego_b200/batch.py) against the reference's own artefacts
(sha256 goldens of SURVEY 8c) and against the oracle chained the way the facade chains the reference
(decode -> int16 WAV -> encode at the file's bitrate)."""
import hashlib
import json

import numpy as np
import pytest

from conftest import golden_path, synth_wav

pytestmark = pytest.mark.gpu


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def _blob(name):
    return open(golden_path(name), "rb").read()


def test_reveal_batch_needs_no_decode(handle):
    from mp3stego_b200 import batch
    fac = json.load(open(golden_path("ref_facade.json")))
    l0 = handle.launches
    got = batch.reveal_batch(handle, [_blob("ref_test_hid.mp3"), _blob("test.mp3"), _blob("ref_test_cleared.mp3")])
    assert got == [fac["reveal_hid"], fac["reveal_test_mp3"], fac["reveal_cleared"]]
    assert handle.launches - l0 <= 6          # walk, layout, per-file scan, side info, publish, copy-out: no Huffman / synthesis kernels
    assert batch.reveal_batch(handle, []) == []


def test_hide_and_clear_equal_the_reference_artefacts(handle):
    from mp3stego_b200 import batch
    fac = json.load(open(golden_path("ref_facade.json")))
    src = _blob("test.mp3")
    out, too_long = batch.hide_batch(handle, [src, src], ["ddd", "ddd" * 100])
    assert [_sha(o) for o in out] == [fac["hid_sha256"], fac["hid_long_sha256"]]
    assert too_long == [fac["hide_ddd_returns"], fac["hide_long_returns"]]
    assert batch.reveal_batch(handle, out[:1]) == ["ddd"]
    cleared = batch.clear_batch(handle, [out[0]])
    assert _sha(cleared[0]) == fac["cleared_sha256"]
    assert batch.reveal_batch(handle, cleared) == [""]


def test_mixed_bitrate_batch_vs_oracle_chain(handle, oracle):
    """Files of different bitrates and lengths in one call, one of them with an ID3v2 tag; every output equals
    oracle.encode(oracle.decode(file).pcm16, bitrate of the file, bits)."""
    from mp3stego_b200 import batch
    from mp3stego_b200.steganography import str_to_binary_str
    specs = [(128, 9, 60), (320, 14, 61), (64, 5, 62), (128, 21, 63)]
    blobs = [oracle.encode(synth_wav(seed, n), 44100, br, "", taps=False)["mp3"] for br, n, seed in specs]
    tag = b"ID3\x04\x00\x00" + bytes([0, 0, 0, 40]) + bytes(40)
    blobs[1] = tag + blobs[1]
    msgs = ["hello", "x" * 400, "", "The quick brown fox"]
    out, too_long = batch.hide_batch(handle, blobs, msgs)
    for (br, n, seed), b, m, o, tl in zip(specs, blobs, msgs, out, too_long):
        skip = len(tag) if b.startswith(b"ID3") else 0
        dec = oracle.decode(b, skip, taps=False)
        bits = str_to_binary_str(str(len(m)) + "#" + m)
        ref = oracle.encode(np.asarray(dec["pcm16"], np.int16).reshape(-1, 2), 44100, br, bits, taps=False)
        assert o == ref["mp3"]
        assert tl == (ref["hide_str_offset"] < len(bits) - 1)
    revealed = batch.reveal_batch(handle, out)
    for m, tl, r in zip(msgs, too_long, revealed):
        if not tl:
            assert r == m


def _stereo_writer_corpus(oracle):
    """Every stereo writer / fuzz / edge stream (BASELINE configs[3]: short / mixed / start / stop blocks, MS stereo, scalefactors,
    reservoir, CRC, VBR, 32 / 48 kHz, loud spectra with int16 wrap), with its oracle decode."""
    import glob
    import os
    from conftest import GOLDEN
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "stream_*.mp3")) + glob.glob(os.path.join(GOLDEN, "fuzz_*.mp3")) +
                    glob.glob(os.path.join(GOLDEN, "edge_*.mp3"))):
        blob = open(p, "rb").read()
        dec = oracle.decode(blob, oracle.id3_offset(blob), taps=False)
        if dec["channels"] == 2 and dec["n_frames"] > 0:
            out.append((os.path.basename(p), blob, dec))
    return out


def test_clean_and_hide_on_the_mixed_block_corpus_vs_oracle_chain(handle, oracle):
    """BASELINE configs[3] 'decode / reveal / clean': clear_file and hide_message (steganography.py:137-182) over the stereo writer
    streams in ONE batch call each -- different sample rates and bitrates, so several encode groups -- equal, byte for byte, the
    oracle chained the way the facade chains the reference: oracle.encode(oracle.decode(x).pcm16, bitrate of the LAST frame)."""
    from mp3stego_b200 import batch
    from mp3stego_b200.steganography import str_to_binary_str
    corpus = _stereo_writer_corpus(oracle)
    assert len(corpus) >= 20
    blobs = [b for _, b, _ in corpus]
    cleared = batch.clear_batch(handle, blobs)
    msgs = ["clean me %d" % i if i % 3 else "x" * 300 for i in range(len(corpus))]
    hid, too_long = batch.hide_batch(handle, blobs, msgs)
    for (name, blob, dec), c, hm, tl, m in zip(corpus, cleared, hid, too_long, msgs):
        pcm = np.asarray(dec["pcm16"], np.int16).reshape(-1, 2)
        ref = oracle.encode(pcm, dec["sampling_rate"], dec["bit_rate"] // 1000, "", taps=False)
        assert c == ref["mp3"], name
        bits = str_to_binary_str(str(len(m)) + "#" + m)
        refh = oracle.encode(pcm, dec["sampling_rate"], dec["bit_rate"] // 1000, bits, taps=False)
        assert hm == refh["mp3"], name
        assert tl == (refh["hide_str_offset"] < len(bits) - 1), name
    # reveal of the results = what the oracle reveals from the same bytes (the reference's own hide -> reveal round trip is not
    # an identity on every input: e.g. a 32 kHz file whose stale region addresses carry bits the decoder never reads back, A.E6)
    assert batch.reveal_batch(handle, cleared) == [oracle.reveal_parse(oracle.decode(c, 0, taps=False)["bits"]) for c in cleared]
    assert batch.reveal_batch(handle, hid) == [oracle.reveal_parse(oracle.decode(c, 0, taps=False)["bits"]) for c in hid]
    assert sum(r == m for m, tl, r in zip(msgs, too_long, batch.reveal_batch(handle, hid)) if not tl) >= 5   # and some do round-trip
