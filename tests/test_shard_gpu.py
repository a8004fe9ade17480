"""GPU: frame-range sharding of ONE long file through the C ABI.  The ranks' ranges -- each decoded on the device from the
whole file's scan plus one warm-up frame (m3s_decode_run_range) -- concatenated in rank order must equal the whole-file decode
sample for sample and bit for bit, for product-encoded files and for the writer streams (bit reservoir, VBR, CRC, mono,
short / mixed blocks, trailing junk)."""
import numpy as np
import pytest

from conftest import golden_path, synth_wav

pytestmark = pytest.mark.gpu


def _whole(handle, blob):
    data = np.frombuffer(blob, np.uint8)
    sc = handle.decode_scan(data, [0, len(blob)])
    _, bits = handle.decode_reveal()
    pcm, _ = handle.decode_run()
    ch = max(int(sc["channels"][0]), 1)
    return sc, pcm[: int(sc["pcm_rows"][0]) * ch].reshape(-1, ch), bits[0]


def _sharded(handle, blob, world, exact=False):
    from mp3stego_b200 import shard
    parts = [shard.decode_frame_range(handle, blob, r, world, exact=exact) for r in range(world)]
    assert sum(p["count"] for p in parts) == parts[0]["n_frames"]
    return np.concatenate([p["pcm"] for p in parts]), "".join(p["bits"] for p in parts), parts


@pytest.mark.parametrize("bitrate", [128, 320])
def test_product_encoded_file_in_ranges(handle, oracle, bitrate):
    wav = synth_wav(900 + bitrate, 150)
    msg = "range shards keep every hidden bit"
    bits = oracle.str_to_bits(f"{len(msg)}#{msg}")
    res = handle.encode(wav.reshape(-1).astype(np.int16), [wav.shape[0]], 44100, bitrate, payloads=[bits])
    blob = bytes(res["mp3"][: int(res["out_len"][0])])
    sc, pcm, rbits = _whole(handle, blob)
    assert int(sc["n_frames"][0]) == 150 and not int(sc["status"][0])
    ref = oracle.decode(blob)
    for world in (2, 3, 8):
        p, b, parts = _sharded(handle, blob, world)
        assert b == rbits == ref["bits"]
        assert np.array_equal(p, pcm)
        assert np.abs(p.astype(np.int32) - ref["pcm16"].astype(np.int32)).max() <= 1
    assert oracle.reveal_parse(rbits) == msg


@pytest.mark.parametrize("name", ["stream_reservoir", "stream_vbr_32k_pad", "stream_mono_crc_48k", "stream_ms_stereo",
                                  "stream_long_alltables", "stream_short_mixed", "stream_is_only_bit",
                                  "fuzz_00", "fuzz_02", "fuzz_05", "fuzz_10", "fuzz_13", "edge_cut05", "edge_crc_on", "edge_late_switch"])
def test_writer_streams_in_ranges(handle, oracle, name):
    """main_data_begin > 0 (the halo must carry the reservoir), variable frame sizes, 21-byte side info, and the streams whose
    granules inherit scalefactors from earlier frames (whole-prefix halo).  Pinned on the ORACLE's whole-file decode: reveal bits
    equal, the float64 instantiation sample-exact, FP32 within 1 LSB (modulo the int16 wrap)."""
    blob = open(golden_path(name + ".mp3"), "rb").read()
    ref = oracle.decode(blob)
    sc, pcm, rbits = _whole(handle, blob)
    for world in (2, 3, 5):
        p, b, parts = _sharded(handle, blob, world)
        assert b == rbits == ref["bits"], name
        assert np.array_equal(p, pcm), name
        d = np.abs(p.astype(np.int32) - ref["pcm16"].reshape(p.shape).astype(np.int32))
        assert np.minimum(d, 65536 - d).max() <= 1, name
        pe, be, _ = _sharded(handle, blob, world, exact=True)
        assert be == ref["bits"] and np.array_equal(pe, ref["pcm16"].reshape(pe.shape)), name


def test_state_carry_flag(handle):
    """Mixed blocks / scfsi over a short granule 0 make a file's granules depend on arbitrarily old frames (A.D4)."""
    from mp3stego_b200 import _lib, shard
    st = {}
    for name in ("stream_short_mixed", "stream_long_alltables", "stream_reservoir"):
        blob = open(golden_path(name + ".mp3"), "rb").read()
        sc = handle.decode_scan(np.frombuffer(blob, np.uint8), [0, len(blob)])
        st[name] = int(sc["status"][0])
    assert st["stream_short_mixed"] & _lib.M3S_FILE_STATE_CARRY
    assert not st["stream_long_alltables"] & _lib.M3S_FILE_STATE_CARRY
    assert not st["stream_reservoir"] & _lib.M3S_FILE_STATE_CARRY
    assert shard.plan_frame_shard(100, _lib.M3S_FILE_STATE_CARRY, 3, 4) == dict(first=75, count=25, warm=1, compact_from=0)
    assert shard.plan_frame_shard(100, 0, 3, 4) == dict(first=75, count=25, warm=1, compact_from=65)


def test_trailing_junk_goes_to_the_last_range(handle):
    """An ID3v1 trailer stops the reference's parser and repeats the last frame's PCM once (A.D10): the last rank's rows."""
    wav = synth_wav(77, 40)
    res = handle.encode(wav.reshape(-1).astype(np.int16), [wav.shape[0]], 44100, 192)
    blob = bytes(res["mp3"][: int(res["out_len"][0])]) + b"TAG" + bytes(125)
    sc, pcm, rbits = _whole(handle, blob)
    assert int(sc["status"][0]) & 4 and pcm.shape[0] == 1152 * (int(sc["n_frames"][0]) + 1)
    p, b, parts = _sharded(handle, blob, 4)
    assert b == rbits and np.array_equal(p, pcm)
    assert parts[-1]["pcm"].shape[0] == 1152 * (parts[-1]["count"] + 1)
