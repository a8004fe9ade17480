"""CPU: host-side logic of the Python mirror (no GPU): WAV writer/reader, ID3 skip, reveal parse, message framing."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT, golden_path, load_npz


def test_write_wav_matches_reference_bytes(tmp_path, built):
    """MP3Parser.write_to_wav (scipy.io.wavfile.write) byte layout: the reference's out.wav, by sha256."""
    from mp3stego_b200.wavio import write_wav
    fac = json.load(open(golden_path("ref_facade.json")))
    z = load_npz("ref_test_mp3.npz")
    p = tmp_path / "out.wav"
    write_wav(str(p), 44100, z["pcm16"])
    raw = open(p, "rb").read()
    assert len(raw) == fac["out_wav_bytes"]
    assert hashlib.sha256(raw).hexdigest() == fac["out_wav_sha256"]
    assert raw == open(golden_path("ref_test_out.wav"), "rb").read()


def test_wav_reader_fields_and_errors(tmp_path, built):
    from mp3stego_b200.wavio import WavReader, write_wav
    w = WavReader(golden_path("ref_test_out.wav"), 128)
    assert (w.samplerate, w.num_of_channels, w.num_of_samples, w.bitrate) == (44100, 2, 41472, 128)
    assert w.mpeg_mode == 0 and w.original == 1 and w.copyright == 0 and w.emphasis == 0
    assert len(w.buffer) == 41472 * 2 and w.buffer.dtype == np.int16
    assert w.get_buffer_pos(1) == 1
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wave file at all" * 10)
    with pytest.raises(SystemExit, match="Bad WAVE file."):
        WavReader(str(bad))
    with pytest.raises(SystemExit, match="Unsupported bitrate configuration."):
        WavReader(golden_path("ref_test_out.wav"), 100)
    p = tmp_path / "sr.wav"
    write_wav(str(p), 22050, np.zeros((1152, 2), np.int16))
    with pytest.raises(SystemExit, match="Unsupported sampling frequency."):
        WavReader(str(p))


def test_id3_offset_rule(built, oracle):
    from mp3stego_b200.decoder import id3_offset
    body = bytes(range(200))
    for flags, extra in ((0x00, 10), (0x10, 20), (0x40, 10)):
        tag = b"ID3\x04\x00" + bytes([flags]) + bytes([0, 0, 1, 0x48])
        data = tag + body
        assert id3_offset(data) == 128 + 0x48 + extra == oracle.id3_offset(data)
    assert id3_offset(b"ID3\x04\x00\x01\x00\x00\x01\x48" + body) == 0   # protected low flag bit set: tag ignored
    assert id3_offset(b"\xff\xfb\x90\x00" + body) == 0
    assert id3_offset(b"ID") == 0
    # numpy uint8 input (what Decoder / files.py / batch.py hand in): no wrap-around at 256
    big = np.frombuffer(b"ID3\x03\x00\x00\x00\x00\x07\x04", np.uint8)
    assert id3_offset(big) == 910 == oracle.id3_offset(bytes(big) + bytes(900)) and isinstance(id3_offset(big), int)


def test_reveal_parse_matches_reference_rule(built, oracle):
    from mp3stego_b200.decoder import parse_reveal
    from mp3stego_b200.steganography import str_to_binary_str
    assert str_to_binary_str("3#ddd") == oracle.str_to_bits("3#ddd")
    assert str_to_binary_str("é#") == "1100001110101001" + "00100011"
    cases = ["3#ddd", "3#dddXYZ", "10#short", "#abc", "x#abc", "", "12", "0#"]
    for c in cases:
        bits = str_to_binary_str(c) + "101"   # trailing partial byte is dropped
        assert parse_reveal(bits) == oracle.reveal_parse(bits)
    assert parse_reveal(str_to_binary_str("3#dddXYZ")) == "ddd"
    assert parse_reveal(str_to_binary_str("10#short")) == "short"
    assert parse_reveal(str_to_binary_str("x#abc")) == ""      # non-numeric prefix: length 0 and the prefix itself counts as empty
    z = load_npz("ref_test_mp3.npz")
    assert parse_reveal(str(z["bits"])) == ""


def test_facade_path_checks(tmp_path, built):
    """The facade's sys.exit messages (steganography.py:63-78) fire before any GPU work."""
    from mp3stego_b200 import Steganography
    s = Steganography(quiet=True)
    with pytest.raises(SystemExit, match="not found"):
        s.decode_mp3_to_wav(str(tmp_path / "missing.mp3"))
    f = tmp_path / "a.txt"
    f.write_bytes(b"x")
    with pytest.raises(SystemExit, match="input_file_path must be mp3 file"):
        s.decode_mp3_to_wav(str(f))
    with pytest.raises(SystemExit, match="wav_file_path must be wav file"):
        s.encode_wav_to_mp3(str(f), str(tmp_path / "o.mp3"))
    m = tmp_path / "a.mp3"
    m.write_bytes(open(golden_path("test.mp3"), "rb").read())
    with pytest.raises(SystemExit, match="txt_file_path must be txt file"):
        s.reveal_massage(str(m), str(tmp_path / "o.bin"))


def test_fast_transforms_match_the_direct_formulas(tmp_path):
    """The FP32 hybrid kernel's in-register transforms (csrc/m3s_fast_transforms.cuh: 18-point DCT-IV for the 36-point IMDCT, 32-point
    Lee DCT-II for the 64 x 32 matrixing, with their index / sign maps), compiled for the host, against the reference's direct
    formulas (Frame.py:81-87,119-133) in double."""
    import subprocess
    exe = str(tmp_path / "ftc")
    src = os.path.join(ROOT, "tests", "model", "fast_transforms_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "imdct36 via dct4_18" in out.stdout and "dct2_lee<32>" in out.stdout


def test_enc_fold_generated_code(tmp_path):
    """csrc/m3s_enc_fold_gen.cuh (the analysis kernel's matrixing / MDCT with equal truncated products computed once) is what
    tools/gen_enc_fold.py makes of the committed tables, and -- compiled for the host -- equals the direct sums of individually
    truncated products (MP3_Encoder.py:358-368, :683-701, util.py:121-127) whenever it does not ask for the direct redo."""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_enc_fold.py"), "--check"]).returncode == 0
    exe = str(tmp_path / "efc")
    src = os.path.join(ROOT, "tests", "model", "enc_fold_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "errors 0" in out.stdout, out.stdout


def test_bench_clock_sampler_window(tmp_path):
    """bench.ClockSampler.stop(t0, t1): only the nvidia-smi samples whose timestamps fall inside the timed window count; when none does
    (a timed region shorter than the tool's start-up) every sample of the leg is used instead of reporting nothing; unparsable rows are
    skipped; a throttle reason seen inside the window is reported."""
    import datetime
    import sys
    import time
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench

    class Done:
        def terminate(self): pass
        def wait(self, timeout=None): pass

    def sampler(rows):
        cs = bench.ClockSampler.__new__(bench.ClockSampler)
        cs.f = open(tmp_path / "smi.csv", "w+")
        cs.f.write("".join(rows))
        cs.f.flush()
        cs.p = Done()
        return cs

    now = time.time()
    stamp = lambda dt: datetime.datetime.fromtimestamp(now + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]   # noqa: E731
    rows = ["%d, 1965, %.2f, Not Active, Not Active, Not Active, %s, %s\n" % (1500 if k < 3 else 1965, 150.0 + 60 * k, "Active" if k == 5 else "Not Active", stamp(0.1 * k))
            for k in range(10)] + ["[N/A], garbage\n"]
    out = sampler(rows).stop(now + 0.33, now + 0.72)    # +- 50 ms of slack: the samples at 0.3 .. 0.7 s
    assert out["samples"] == 5 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    out = sampler(rows).stop(now + 5.0, now + 6.0)          # nothing inside: the whole leg, and the record says so
    assert out["samples"] == 10 and "no sample" in out["window"] and out["sm_mhz"] == 1965.0
    out = sampler([]).stop(now, now + 1.0)
    assert out["sm_mhz"] is None and out["reasons"] == []


def test_host_affinity_helpers(tmp_path, monkeypatch):
    """hostaffinity: cpulist parsing, PCI bus-id normalisation, and the no-op on single-node hosts (the GPU pool's VMs)."""
    from mp3stego_b200 import hostaffinity as ha
    assert ha._parse_cpulist("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11]
    assert ha._parse_cpulist("") == []
    assert ha.gpu_numa_node("00000000:ZZ:00.0") == -1            # unknown device: the platform "does not say"
    info = ha.bind_to_device("0000:00:00.0")
    assert set(info) >= {"numa_nodes", "gpu_node", "bound", "cpus"}
    if info["numa_nodes"] <= 1:
        assert info["bound"] is False and "nothing to bind" in info["note"]
    # two nodes, device without a node: ranks are spread round-robin and the process is bound to CPUs it is allowed to use
    cpus = sorted(os.sched_getaffinity(0))
    half = max(1, len(cpus) // 2)
    monkeypatch.setattr(ha, "numa_nodes", lambda: {0: cpus[:half], 1: cpus[half:] or cpus[:half]})
    monkeypatch.setattr(ha, "gpu_numa_node", lambda bdf: -1)
    try:
        info = ha.bind_to_device("0000:00:00.0", local_rank=1, local_world=2)
        assert info["bound"] and info["node"] == 1 and os.sched_getaffinity(0) == set(cpus[half:] or cpus[:half])
    finally:
        os.sched_setaffinity(0, cpus)
