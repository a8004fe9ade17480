"""GPU: decode + reveal parity of the CUDA path (through the C ABI) against the oracle and the reference-generated
golden vectors.  Gates (BASELINE.json north_star): integer spectra, table ids and reveal bits bit-exact;
PCM within 1 LSB at 16 bit and 1e-5 abs in float."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, SYNTH_CASES, golden_path, load_npz, synth_wav

pytestmark = pytest.mark.gpu

PCM_TOL_LSB = 1       # int16 samples: |cuda - reference| <= 1
PCM_TOL_FLOAT = 1e-5  # pre-quantisation float samples


def _decode_batch(handle, blobs, audio_start=None, spectra=True, as_float=False):
    data = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
    sc = handle.decode_scan(data, off, audio_start)
    ids, bits = handle.decode_reveal()
    pcm, sp = handle.decode_run(spectra=spectra, as_float=as_float)
    fb = np.concatenate([[0], np.cumsum(sc["n_frames"])])
    eb = np.concatenate([[0], np.cumsum(sc["pcm_rows"] * np.maximum(sc["channels"], 1))])
    out = []
    for i in range(len(blobs)):
        ch = max(int(sc["channels"][i]), 1)
        out.append(dict(n_frames=int(sc["n_frames"][i]), status=int(sc["status"][i]), bitrate=int(sc["bitrate"][i]),
                        sample_rate=int(sc["sample_rate"][i]), channels=int(sc["channels"][i]),
                        ids=ids[fb[i]:fb[i + 1]], bits=bits[i],
                        spectra=None if sp is None else sp[fb[i]:fb[i + 1]].astype(np.int32),
                        pcm=pcm[eb[i]:eb[i + 1]].reshape(-1, ch)))
    return out


def _check(got, ref, pcm16=True):
    assert got["n_frames"] == ref["n_frames"]
    if ref["n_frames"] == 0:
        assert got["pcm"].size == 0 and got["bits"] == ""
        return
    if got["spectra"] is not None:
        assert np.array_equal(got["spectra"], ref["spectra"]), "integer spectra differ"
    assert np.array_equal(got["ids"], ref["tables"]), "table ids differ"
    assert got["bits"] == ref["bits"], "reveal bits differ"
    if pcm16:
        assert got["pcm"].shape == ref["pcm16"].shape
        d = np.abs(got["pcm"].astype(np.int32) - ref["pcm16"].astype(np.int32))
        assert d.max() <= PCM_TOL_LSB, f"PCM off by {d.max()} LSB"
    else:
        assert got["pcm"].shape == ref["pcm"].shape
        # 1e-5 of full scale: streams that run (far) past full scale -- the loud / wrap-around cases, |pcm| up to ~30 -- scale with it
        tol = PCM_TOL_FLOAT * max(1.0, float(np.abs(ref["pcm"]).max()))
        assert np.abs(got["pcm"].astype(np.float64) - ref["pcm"]).max() <= tol


def test_test_mp3_vs_reference_golden(handle):
    """configs[0]: tests/test.mp3 against the values the unmodified reference produced."""
    z = load_npz("ref_test_mp3.npz")
    got = _decode_batch(handle, [open(golden_path("test.mp3"), "rb").read()])[0]
    assert got["n_frames"] == 36 and got["bitrate"] == 320000 and got["sample_rate"] == 44100
    assert np.array_equal(got["spectra"], z["spectra"].astype(np.int32))
    assert np.array_equal(got["ids"], z["tables"])
    assert got["bits"] == str(z["bits"])
    d = np.abs(got["pcm"].astype(np.int32) - z["pcm16"].astype(np.int32))
    assert d.max() <= PCM_TOL_LSB
    gotf = _decode_batch(handle, [open(golden_path("test.mp3"), "rb").read()], spectra=False, as_float=True)[0]
    for k, f in enumerate(z["pcm64_frames"]):
        df = np.abs(gotf["pcm"][f * 1152:(f + 1) * 1152].astype(np.float64) - z["pcm64"][k * 1152:(k + 1) * 1152])
        assert df.max() <= PCM_TOL_FLOAT


@pytest.mark.parametrize("case", SYNTH_CASES)
def test_synth_vs_reference_golden(handle, case):
    z = load_npz(f"ref_synth_{case}.npz")
    got = _decode_batch(handle, [z["mp3"].tobytes()])[0]
    assert got["n_frames"] == int(z["dec_n_frames"])
    assert np.array_equal(got["spectra"], z["dec_spectra"].astype(np.int32))
    assert np.array_equal(got["ids"], z["dec_tables"])
    assert got["bits"] == str(z["dec_bits"])
    assert np.abs(got["pcm"].astype(np.int32) - z["dec_pcm16"].astype(np.int32)).max() <= PCM_TOL_LSB


STREAM_CASES = ["long_alltables", "reservoir", "short_mixed", "ms_stereo", "is_only_bit", "mono_crc_48k", "vbr_32k_pad",
                "loud_wrap"]


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("case", STREAM_CASES)
def test_writer_streams_vs_reference_golden(handle, case, exact):
    """BASELINE configs[3]: short / mixed blocks, MS stereo, scalefactors, scfsi, all tables, reservoir, CRC, mono, VBR
    (tests/golden/make_streams.py) against what the unmodified reference decoder produced.  FP32: <= 1 LSB (modulo the
    reference's int16 wrap, A.D8); float64 instantiation: sample-exact."""
    z = load_npz(f"ref_stream_{case}.npz")
    data = np.frombuffer(open(golden_path(f"stream_{case}.mp3"), "rb").read(), np.uint8)
    sc = handle.decode_scan(data, [0, len(data)])
    ids, bits = handle.decode_reveal()
    pcm, sp = handle.decode_run(spectra=True, exact=exact)
    assert int(sc["n_frames"][0]) == int(z["n_frames"])
    assert int(sc["bitrate"][0]) == int(z["bitrate"]) and int(sc["sample_rate"][0]) == int(z["sampling_rate"])
    ch = int(sc["channels"][0])
    assert np.array_equal(sp.astype(np.int32)[:, :, :ch], z["spectra"].astype(np.int32)[:, :, :ch])
    k = z["tables"].shape[1]
    assert np.array_equal(ids[:, :k], z["tables"])
    assert bits[0] == str(z["bits"])
    got = pcm.reshape(z["pcm16"].shape).astype(np.int32)
    ref = z["pcm16"].astype(np.int32)
    if exact:
        assert np.array_equal(got, ref)
    else:
        d = np.abs(got - ref)
        d = np.minimum(d, 65536 - d)   # a 1-LSB difference across the wrap boundary shows up as 65535
        assert d.max() <= PCM_TOL_LSB


def test_batch_of_all_goldens_vs_oracle(handle, oracle):
    """Every golden MP3 in ONE batch (mixed bitrates / lengths, incl. stream-writer cases) vs the oracle."""
    blobs = [open(golden_path("test.mp3"), "rb").read()]
    blobs += [np.load(p)["mp3"].tobytes() for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_synth_*.npz")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_test_*.mp3")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "stream_*.mp3")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "fuzz_*.mp3")))]   # configs[3] as a corpus
    res = _decode_batch(handle, blobs)
    resf = _decode_batch(handle, blobs, spectra=False, as_float=True)
    for b, g, gf in zip(blobs, res, resf):
        ref = oracle.decode(b)
        _check(g, ref)
        _check(gf, ref, pcm16=False)


def test_long_file_run_boundaries(handle, oracle):
    """Files longer than one CTA run (32 frames): every run boundary re-decodes one warm-up frame;
    the result must equal the sequential oracle everywhere."""
    wav = synth_wav(21, 150)
    blobs = [oracle.encode(wav, 44100, br, "", taps=False)["mp3"] for br in (320, 128, 64)]
    for g, b in zip(_decode_batch(handle, blobs), blobs):
        _check(g, oracle.decode(b))


def test_id3_and_trailing_junk(handle, oracle):
    """ID3v2 skip (decoder.py:29-33) and the bad-sync stop that repeats the last frame once (MP3_Parser.py:68-79)."""
    z = load_npz("ref_synth_s12_320_plain.npz")
    mp3 = z["mp3"].tobytes()
    body = b"\x00" * 57
    size = len(body)
    tag = b"ID3\x03\x00\x00" + bytes([(size >> 21) & 127, (size >> 14) & 127, (size >> 7) & 127, size & 127]) + body
    with_tag = tag + mp3
    with_junk = mp3 + b"TAG" + b"\x00" * 125
    truncated = mp3[:-300]
    blobs = [with_tag, with_junk, truncated, mp3]
    starts = [oracle.id3_offset(b) for b in blobs]
    assert starts[0] == 67
    res = _decode_batch(handle, blobs, audio_start=starts)
    for b, s, g in zip(blobs, starts, res):
        ref = oracle.decode(b, s)
        _check(g, ref)
    assert res[1]["status"] & 4 and res[1]["pcm"].shape[0] == 1152 * (res[1]["n_frames"] + 1)


def test_no_sync_and_empty(handle, oracle):
    z = load_npz("ref_synth_s13_64_hide.npz")
    mp3 = z["mp3"].tobytes()
    blobs = [b"RIFFxxxxWAVE" + b"\x00" * 100, mp3, b"", b"\xff"]
    res = _decode_batch(handle, blobs)
    assert res[0]["n_frames"] == 0 and res[0]["status"] & 1
    assert res[2]["n_frames"] == 0 and res[3]["n_frames"] == 0
    _check(res[1], oracle.decode(mp3))


def test_full_size_properties(handle, oracle):
    """BASELINE configs[1] shape (3-minute 320 kbps stereo files), checked through size-independent properties:
    a long file built by concatenating self-contained clips (reference-encoder frames have main_data_begin = 0)
    must decode, segment by segment, to the PCM of the clips decoded alone -- except the first frame of every
    segment, whose overlap/synthesis history comes from the preceding clip -- and reveal the concatenated bits."""
    clips = [oracle.encode(synth_wav(100 + s, 65), 44100, 320, "", taps=False)["mp3"] for s in range(4)]
    # a 4-byte-flush truncated clip (A.E8) is 0-3 bytes short: pad the last frame so the next header lands right
    full = []
    for c in clips:
        ref = oracle.decode(c)
        last = int(ref["frame_off"][-1])
        fs = (144 * 320000) // 44100 + ((c[last + 2] >> 1) & 1)
        full.append(c + b"\x00" * (last + fs - len(c)))
    order = np.random.default_rng(0).integers(0, 4, size=106)   # 106 * 65 = 6,890 frames = 179.98 s
    big = b"".join(full[i] for i in order)
    got = _decode_batch(handle, [big, full[0]], spectra=False)
    assert got[0]["n_frames"] == 6890
    singles = [oracle.decode(c) for c in full]
    bits = "".join(singles[i]["bits"] for i in order)
    assert got[0]["bits"] == bits
    pos = 0
    for i in order:
        seg = got[0]["pcm"][pos * 1152:(pos + 65) * 1152].astype(np.int32)
        ref = singles[i]["pcm16"].astype(np.int32)
        assert np.abs(seg[1152:] - ref[1152:]).max() <= PCM_TOL_LSB
        pos += 65
    _check(got[1], singles[0])


def test_fuzz_corpus_exact_decode_vs_reference_digests(handle):
    """The float64 instantiation on the 16 fuzz streams (short / mixed blocks, MS, reservoir, CRC, mono, VBR ...): the int16 PCM, the
    integer spectra, the table ids and the reveal bits equal the unmodified reference's, by digest."""
    import hashlib
    import json
    ref = json.load(open(golden_path("ref_fuzz.json")))
    names = sorted(ref)
    blobs = [open(golden_path(n + ".mp3"), "rb").read() for n in names]
    data = np.frombuffer(b"".join(blobs), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
    sc = handle.decode_scan(data, off)
    ids, bits = handle.decode_reveal()
    pcm, sp = handle.decode_run(spectra=True, exact=True)
    fb = np.concatenate([[0], np.cumsum(sc["n_frames"])])
    eb = np.concatenate([[0], np.cumsum(sc["pcm_rows"] * np.maximum(sc["channels"], 1))])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for i, n in enumerate(names):
        r = ref[n]
        assert int(sc["n_frames"][i]) == r["n_frames"] and int(sc["sample_rate"][i]) == r["sampling_rate"], n
        assert bits[i] == r["bits"], n
        assert sha(ids[fb[i]:fb[i + 1], :6 * r["channels"]]) == r["tables_sha256"], n   # the reference lists 6 ids per channel
        assert sha(sp[fb[i]:fb[i + 1]].astype(np.int16)) == r["spectra_sha256"], n
        assert sha(pcm[eb[i]:eb[i + 1]].astype(np.int16)) == r["pcm16_sha256"], n


def _edge_blobs():
    import json
    ref = json.load(open(golden_path("ref_edge.json")))
    names = sorted(ref)
    return names, [open(golden_path(n + ".mp3"), "rb").read() for n in names], ref


def test_edge_reservoirs_exact_vs_reference_digests(handle, oracle):
    """Cut streams (first frames point into a reservoir that is not there, bare or behind an ID3v2 tag) and CRC switches under a
    live reservoir: the reference assembles the bytes physically in front of the frame or keeps the previous frame's main data
    (Frame.py:318-363, A.D9).  Float64 instantiation vs the unmodified reference's digests; FP32 vs the oracle within 1 LSB."""
    import hashlib
    names, blobs, ref = _edge_blobs()
    starts = [oracle.id3_offset(b) for b in blobs]
    data = np.frombuffer(b"".join(blobs), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
    sc = handle.decode_scan(data, off, starts)
    ids, bits = handle.decode_reveal()
    pcm, sp = handle.decode_run(spectra=True, exact=True)
    fb = np.concatenate([[0], np.cumsum(sc["n_frames"])])
    eb = np.concatenate([[0], np.cumsum(sc["pcm_rows"] * np.maximum(sc["channels"], 1))])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for i, n in enumerate(names):
        r = ref[n]
        assert int(sc["n_frames"][i]) == r["n_frames"] and int(sc["sample_rate"][i]) == r["sampling_rate"], n
        assert bits[i] == r["bits"], n
        assert sha(ids[fb[i]:fb[i + 1], :6 * r["channels"]]) == r["tables_sha256"], n
        assert sha(sp[fb[i]:fb[i + 1]].astype(np.int16)) == r["spectra_sha256"], n
        assert sha(pcm[eb[i]:eb[i + 1]].astype(np.int16)) == r["pcm16_sha256"], n
    for b, s, g in zip(blobs, starts, _decode_batch(handle, blobs, audio_start=starts)):
        _check(g, oracle.decode(b, s))


def _all_golden_blobs():
    blobs = [open(golden_path("test.mp3"), "rb").read()]
    blobs += [np.load(p)["mp3"].tobytes() for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_synth_*.npz")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_test_*.mp3")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "stream_*.mp3")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "fuzz_*.mp3")))]
    blobs += [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "edge_*.mp3")))]
    blobs += [b"RIFFxxxxWAVE" + b"\x00" * 100, b"", blobs[1] + b"TAG" + b"\x00" * 125]
    return blobs


def _split_decode(res, strings, as_float=False):
    out = []
    fb = np.concatenate([[0], np.cumsum(res["n_frames"])])
    for i in range(len(res["n_frames"])):
        ch = max(int(res["channels"][i]), 1)
        out.append(dict(n_frames=int(res["n_frames"][i]), status=int(res["status"][i]), bitrate=int(res["bitrate"][i]),
                        sample_rate=int(res["sample_rate"][i]), channels=int(res["channels"][i]), spectra=None,
                        ids=np.asarray(res["table_ids"][12 * fb[i]:12 * fb[i + 1]]).reshape(-1, 12), bits=strings[i],
                        pcm=np.asarray(res["pcm"][res["pcm_off"][i]:res["pcm_off"][i + 1]]).reshape(-1, ch)))
    return out


@pytest.mark.parametrize("wave_bytes", [3000, 40000, 1 << 28])
def test_pipelined_batch_call_vs_oracle(built, oracle, wave_bytes, monkeypatch):
    """m3s_decode -- ONE self-pipelining call for the whole batch (waves of files on copy / scan / compute / copy-out streams) --
    over every golden stream, with wave sizes that put one file per wave, a few files per wave, and everything in one wave:
    host buffers, int16 and float PCM, vs the oracle."""
    from mp3stego_b200 import _lib
    monkeypatch.setenv("M3S_DEC_WAVE_BYTES", str(wave_bytes))
    h = _lib.Handle(0)
    blobs = _all_golden_blobs()
    starts = [oracle.id3_offset(b) for b in blobs]
    data = np.frombuffer(b"".join(blobs), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
    refs = [oracle.decode(b, s) for b, s in zip(blobs, starts)]
    for rep in range(2):   # the second call reuses every workspace
        res = h.decode(data, off, starts)
        for g, ref in zip(_split_decode(res, h.reveal_strings(res)), refs):
            _check(g, ref)
    resf = h.decode(data, off, starts, as_float=True)
    for g, ref in zip(_split_decode(resf, h.reveal_strings(resf)), refs):
        _check(g, ref, pcm16=False)
    h.close()


def test_pipelined_batch_call_device_resident_and_exact(built, oracle, monkeypatch):
    """m3s_decode with device pointers (nothing crosses PCIe), float64 instantiation: sample-exact vs the oracle's int16 PCM; the
    first call under-estimates the frame count of device-resident input on purpose (417 bytes / frame guess vs 64 kbps files),
    which exercises the grow-and-rescan path."""
    import torch
    from mp3stego_b200 import _lib
    monkeypatch.setenv("M3S_DEC_WAVE_BYTES", "30000")
    h = _lib.Handle(0)
    blobs = [np.load(golden_path("ref_synth_s13_64_hide.npz"))["mp3"].tobytes()] * 3 + _all_golden_blobs()
    starts = [oracle.id3_offset(b) for b in blobs]
    data = torch.from_numpy(np.frombuffer(b"".join(blobs), np.uint8).copy()).cuda()
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
    refs = [oracle.decode(b, s) for b, s in zip(blobs, starts)]
    nfr = sum(r["n_frames"] for r in refs) + 8
    pcm = torch.zeros(nfr * 1152 * 2 + 4096, dtype=torch.int16, device="cuda")
    res = h.decode(data, off, starts, pcm=pcm, frames_bound=nfr, exact=True)
    res["pcm"] = res["pcm"].cpu().numpy()
    res["table_ids"] = res["table_ids"].cpu().numpy()
    res["reveal_bits"] = res["reveal_bits"].cpu().numpy()
    for g, ref in zip(_split_decode(res, h.reveal_strings(res)), refs):
        _check(g, ref)
        assert np.array_equal(g["pcm"], ref["pcm16"].reshape(g["pcm"].shape))
    with pytest.raises(_lib.M3SError, match="pcm holds"):
        h.decode(data, off, starts, pcm=pcm[:1152 * 2 * 40], frames_bound=nfr)
    res2 = h.decode(data, off, starts, pcm=pcm, frames_bound=nfr)     # the handle is usable after the capacity error
    assert np.array_equal(res2["n_frames"], res["n_frames"])
    h.close()


def test_decode_run_leaves_gaps_alone(handle, oracle):
    """Host PCM with caller offsets: the bytes between the files are not touched (per-file copies, not one staging-range copy)."""
    mp3 = np.load(golden_path("ref_synth_s12_320_plain.npz"))["mp3"].tobytes()
    data = np.frombuffer(mp3 + mp3, np.uint8)
    sc = handle.decode_scan(data, [0, len(mp3), 2 * len(mp3)])
    n = int(sc["pcm_rows"][0]) * 2
    pcm = np.full(2 * n + 3000, 12345, np.int16)
    handle.decode_run(pcm=pcm, pcm_off=[1000, 1000 + n + 1000])
    ref = oracle.decode(mp3)["pcm16"].reshape(-1).astype(np.int32)
    assert (pcm[:1000] == 12345).all() and (pcm[1000 + n:2000 + n] == 12345).all() and (pcm[2000 + 2 * n:] == 12345).all()
    assert np.abs(pcm[1000:1000 + n].astype(np.int32) - ref).max() <= PCM_TOL_LSB
    assert np.abs(pcm[2000 + n:2000 + 2 * n].astype(np.int32) - ref).max() <= PCM_TOL_LSB


def test_files_the_reference_raises_on_are_flagged(handle):
    """A channel-count change inside one file makes MP3Parser.parse_file raise ValueError (rows of different widths, MP3_Parser.py:83);
    big_values > 288 makes Frame.__unpack_samples raise IndexError (Frame.py:461-478).  The scan flags both (the kernels clamp and
    stay inside their buffers) and the Python mirror re-raises."""
    from mp3stego_b200 import _lib
    from mp3stego_b200.decoder import MP3Parser
    mono = open(golden_path("stream_mono_crc_48k.mp3"), "rb").read()
    stereo = open(golden_path("stream_short_mixed.mp3"), "rb").read()
    for blob in (mono + stereo, stereo + mono):
        data = np.frombuffer(blob, np.uint8)
        sc = handle.decode_scan(data, [0, len(blob)])
        assert int(sc["status"][0]) & _lib.M3S_FILE_CHANNEL_SWITCH and int(sc["n_frames"][0]) == 18
        handle.decode_run()          # memory-safe: completes, output unspecified
        with pytest.raises(ValueError):
            MP3Parser(data, 0, "/dev/null").parse_file()
    bad = bytearray(open(golden_path("stream_long_alltables.mp3"), "rb").read())
    # side info starts at byte 4: main_data_begin 9 + private 3 + scfsi 8 = 20 bits, then part2_3_length 12, then big_values 9 bits
    bitpos = 8 * 4 + 20 + 12
    for k in range(9):            # big_values = 511
        bad[(bitpos + k) >> 3] |= 0x80 >> ((bitpos + k) & 7)
    data = np.frombuffer(bytes(bad), np.uint8)
    sc = handle.decode_scan(data, [0, len(bad)])
    assert int(sc["status"][0]) & _lib.M3S_FILE_BAD_SIDEINFO
    handle.decode_run()
    with pytest.raises(IndexError):
        MP3Parser(data, 0, "/dev/null").parse_file()
    ok = handle.decode_scan(np.frombuffer(stereo, np.uint8), [0, len(stereo)])
    assert not int(ok["status"][0]) & (_lib.M3S_FILE_CHANNEL_SWITCH | _lib.M3S_FILE_BAD_SIDEINFO)


def test_corrupted_inputs_are_memory_safe(built, oracle):
    """Crafted / damaged input must never take the kernels outside their buffers or hang them: golden streams with random byte
    flips, bit flips in headers and side info, truncations and garbage tails go through both decode paths (tools/gpu_sanitize.sh
    runs this test under compute-sanitizer memcheck); afterwards the same handle still decodes a clean file exactly.  Where the
    damage leaves the frame walk intact and the oracle can read the file, the reveal bits must agree too."""
    from mp3stego_b200 import _lib
    h = _lib.Handle(0)
    rng = np.random.default_rng(20251017)
    names = ["test.mp3", "stream_reservoir.mp3", "stream_short_mixed.mp3", "stream_mono_crc_48k.mp3", "stream_vbr_32k_pad.mp3",
             "fuzz_03.mp3", "fuzz_09.mp3", "edge_cut05.mp3", "edge_crc_on.mp3"]
    clean = [open(golden_path(n), "rb").read() for n in names]
    blobs = []
    for rep in range(6):
        for b in clean:
            a = bytearray(b)
            kind = rng.integers(0, 5)
            if kind == 0:                                   # random byte flips anywhere
                for p in rng.integers(0, len(a), size=max(1, len(a) // 200)):
                    a[p] ^= int(rng.integers(1, 256))
            elif kind == 1:                                 # damage inside the first frames' headers + side info
                for p in rng.integers(0, min(len(a), 400), size=12):
                    a[p] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2:                                 # truncation at a random byte
                a = a[: int(rng.integers(1, len(a)))]
            elif kind == 3:                                 # garbage tail (may contain sync-like bytes)
                a += bytes(rng.integers(0, 256, size=int(rng.integers(1, 3000)), dtype=np.uint8)) + b"\xff\xfb\x90\x64" * 3
            else:                                           # side info fields pushed to their limits in every frame it can find
                for p in range(4, len(a) - 40, 417):
                    a[p:p + 8] = b"\xff" * 8
            blobs.append(bytes(a))
    data = np.frombuffer(b"".join(blobs), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
    sc = h.decode_scan(data, off)
    _, bits3 = h.decode_reveal()
    h.decode_run()
    res = h.decode(data, off, frames_bound=int(sc["n_frames"].sum()) + 8)
    assert np.array_equal(res["n_frames"], sc["n_frames"]) and np.array_equal(res["status"], sc["status"])
    assert h.reveal_strings(res) == bits3
    agree = 0
    for b, n, st, bits in zip(blobs, sc["n_frames"], sc["status"], bits3):
        if st & (_lib.M3S_FILE_UNSUPPORTED | _lib.M3S_FILE_CHANNEL_SWITCH | _lib.M3S_FILE_BAD_SIDEINFO | _lib.M3S_FILE_NO_SYNC):
            continue
        try:
            ref = oracle.decode(b, 0, taps=False)
        except Exception:
            continue
        if ref["n_frames"] == n and ref["channels"] in (1, 2):
            assert ref["bits"] == bits
            agree += 1
    assert agree >= 10
    z = load_npz("ref_test_mp3.npz")                        # the handle is still sound
    got = _decode_batch(h, [clean[0]])[0]
    assert np.array_equal(got["spectra"], z["spectra"].astype(np.int32)) and got["bits"] == str(z["bits"])
    h.close()
