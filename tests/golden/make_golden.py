#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED Python reference (development container only).

  python tests/golden/make_golden.py            # writes tests/golden/*.npz, *.json, *.mp3

The reference (/root/reference, read-only) is imported with a 7-line `bitarray` stand-in (the real
wheel is absent; it is used only at steganography.py:20-23).  Intermediate values are tapped by
wrapping the reference's own name-mangled methods; no reference source is copied.

Fixtures produced (all small):
  ref_test_mp3.npz     tests/test.mp3 decoded by the reference: int16 PCM (all 36 frames), float64 PCM
                       (first 6 + last 2 frames), integer spectra, per-granule side info, table ids, reveal bits
  ref_synth_*.npz      tone+noise WAV (seeded) -> reference encoder at 128/320 kbps, plain and hiding:
                       mp3 bytes, mdct / ix taps, side info, hide_str_offset; and the reference decode of
                       those bytes (int16 PCM, spectra, bits)
  ref_facade.json      sha256 of every facade artefact of SURVEY.md section 8(c) + revealed strings
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = "/tmp/refshim"
os.makedirs(SHIM, exist_ok=True)
with open(os.path.join(SHIM, "bitarray.py"), "w") as f:
    f.write("class bitarray(list):\n    def frombytes(self, b):\n        for byte in b:\n"
            "            for n in range(7, -1, -1):\n                self.append((byte >> n) & 1)\n")
sys.path.insert(0, SHIM)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from scipy.io import wavfile  # noqa: E402
import tqdm  # noqa: E402

# silence the progress bars
import mp3stego.decoder.MP3_Parser as _mp  # noqa: E402
import mp3stego.encoder.MP3_Encoder as _me  # noqa: E402
_mp.tqdm = lambda *a, **k: tqdm.tqdm(*a, **{**k, "disable": True})
_me.tqdm = lambda *a, **k: tqdm.tqdm(*a, **{**k, "disable": True})

from mp3stego import Steganography  # noqa: E402
from mp3stego.decoder.Frame import Frame  # noqa: E402
from mp3stego.decoder.ID3_Parser import ID3  # noqa: E402
from mp3stego.decoder.MP3_Parser import MP3Parser  # noqa: E402
from mp3stego.encoder.MP3_Encoder import MP3Encoder  # noqa: E402
from mp3stego.encoder.WAV_Reader import WavReader  # noqa: E402

SIDE_FIELDS = ["part2_3_length", "big_value", "global_gain", "scale_fac_compress", "window_switching",
               "block_type", "mixed_block_flag", "table_select", "sub_block_gain", "region0_count",
               "region1_count", "pre_flag", "scale_fac_scale", "count1table_select"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def ref_decode_taps(mp3_bytes):
    """Run MP3Parser over bytes; returns dict of taps."""
    data = list(mp3_bytes)
    id3 = ID3(data)
    offset = id3.offset if id3.is_valid else 0
    spectra, side, xr = [], [], []
    orig_set_main = Frame._Frame__set_main_data
    orig_imdct_mod = sys.modules["mp3stego.decoder.Frame"].imdct

    def tap_set_main(self, file_data, curr_offset):
        orig_set_main(self, file_data, curr_offset)
        spectra.append(np.array(self._Frame__samples, dtype=np.float64).astype(np.int32))  # [gr][ch][576]
        si = self.side_info
        row = np.zeros((2, 2, 18), dtype=np.int32)
        for gr in range(2):
            for ch in range(2):
                row[gr, ch, 0] = si.part2_3_length[gr][ch]
                row[gr, ch, 1] = si.big_value[gr][ch]
                row[gr, ch, 2] = si.global_gain[gr][ch]
                row[gr, ch, 3] = si.scale_fac_compress[gr][ch]
                row[gr, ch, 4] = si.window_switching[gr][ch]
                row[gr, ch, 5] = si.block_type[gr][ch]
                row[gr, ch, 6] = si.mixed_block_flag[gr][ch]
                row[gr, ch, 7:10] = si.table_select[gr][ch]
                row[gr, ch, 10:13] = si.sub_block_gain[gr][ch]
                row[gr, ch, 13] = si.region0_count[gr][ch]
                row[gr, ch, 14] = si.region1_count[gr][ch]
                row[gr, ch, 15] = si.pre_flag[gr][ch]
                row[gr, ch, 16] = si.scale_fac_scale[gr][ch]
                row[gr, ch, 17] = si.count1table_select[gr][ch]
        side.append(row)

    def tap_imdct(gr, ch, block_type, samples, sine_block, prev_samples):
        xr.append(np.array(samples[gr][ch], dtype=np.float64))
        return orig_imdct_mod(gr, ch, block_type, samples, sine_block, prev_samples)

    Frame._Frame__set_main_data = tap_set_main
    sys.modules["mp3stego.decoder.Frame"].imdct = tap_imdct
    try:
        p = MP3Parser(data, offset, "/dev/null")
        n = p.parse_file()
    finally:
        Frame._Frame__set_main_data = orig_set_main
        sys.modules["mp3stego.decoder.Frame"].imdct = orig_imdct_mod
    pcm = np.array(p._MP3Parser__pcm_data, dtype=np.float64)
    tables = np.array(p._MP3Parser__curr_frame.all_huffman_tables, dtype=np.uint8)
    ch = pcm.shape[1] if pcm.ndim == 2 else 2
    return dict(n_frames=n, pcm=pcm, pcm16=(pcm * 32767).astype(np.int16),
                spectra=np.array(spectra, dtype=np.int32), side=np.array(side, dtype=np.int32),
                xr=np.array(xr, dtype=np.float64).reshape(n, 2, ch, 576) if len(xr) == n * 2 * ch else np.zeros(0),
                tables=tables, bits=p.output_bits, bitrate=p.get_bitrate(),
                sampling_rate=p._MP3Parser__curr_frame.sampling_rate)


def ref_encode_taps(wav_path, bitrate, hide_bits):
    wr = WavReader(wav_path, bitrate)
    enc = MP3Encoder(wr, hide_str=hide_bits)
    mdct, ix, info, scfsi = [], [], [], []
    orig_fmt = MP3Encoder._MP3Encoder__format_bitstream

    def tap_fmt(self):
        orig_fmt(self)
        mdct.append(np.array(self._MP3Encoder__mdct_freq, dtype=np.int32).reshape(2, 2, 576))
        ix.append(np.array(self._MP3Encoder__l3_enc, dtype=np.int32))
        row = np.zeros((2, 2, 16), dtype=np.int32)
        for gr in range(2):
            for ch in range(2):
                gi = self._MP3Encoder__side_info.gr[gr].ch[ch].tt
                row[gr, ch] = [int(gi.part2_3_length), int(gi.big_values), int(gi.count1), int(gi.global_gain),
                               int(gi.table_select[0]), int(gi.table_select[1]), int(gi.table_select[2]),
                               int(gi.region0_count), int(gi.region1_count), int(gi.count1table_select),
                               int(gi.address1), int(gi.address2), int(gi.address3), int(gi.quantizerStepSize),
                               int(self._MP3Encoder__mpeg.padding), int(self._MP3Encoder__hide_str_offset)]
        info.append(row)
        scfsi.append(np.array(self._MP3Encoder__side_info.scfsi, dtype=np.int32))

    MP3Encoder._MP3Encoder__format_bitstream = tap_fmt
    try:
        enc.encode()
    finally:
        MP3Encoder._MP3Encoder__format_bitstream = orig_fmt
    out = bytes(enc._MP3Encoder__out_buffer)
    return dict(mp3=np.frombuffer(out, dtype=np.uint8), mdct=np.array(mdct), ix=np.array(ix),
                info=np.array(info), scfsi=np.array(scfsi), hide_str_offset=enc.hide_str_offset)


def synth_wav(seed, n_frames, sr=44100):
    """L = 0.4 sin(2 pi f_L t) + 0.05 N(0,1), R likewise; f ~ U[100, 5000]; *32767 -> int16."""
    rng = np.random.default_rng(seed)
    n = n_frames * 1152
    t = np.arange(n) / sr
    f = rng.uniform(100, 5000, size=2)
    x = np.stack([0.4 * np.sin(2 * np.pi * f[c] * t) + 0.05 * rng.standard_normal(n) for c in range(2)], axis=1)
    return (x * 32767).astype(np.int16)


def str_to_bits(s):
    return "".join(format(b, "08b") for b in s.encode("utf-8"))


def main():
    tmp = tempfile.mkdtemp(prefix="golden_")
    os.chdir(tmp)
    os.makedirs("tests", exist_ok=True)
    shutil.copy("/root/reference/tests/test.mp3", "tests/test.mp3")
    shutil.copy("/root/reference/tests/test.mp3", os.path.join(HERE, "test.mp3"))
    facade = {}
    s = Steganography(quiet=True)

    # ---- facade artefacts
    facade["test_mp3_sha256"] = sha("tests/test.mp3")
    facade["decode_returns"] = int(s.decode_mp3_to_wav("tests/test.mp3", "tests/out.wav"))
    facade["out_wav_sha256"] = sha("tests/out.wav")
    facade["out_wav_bytes"] = os.path.getsize("tests/out.wav")
    s.reveal_massage("tests/test.mp3", "tests/reveal0.txt")
    facade["reveal_test_mp3"] = open("tests/reveal0.txt", "rb").read().decode("utf-8")
    s.encode_wav_to_mp3("tests/out.wav", "tests/enc320.mp3", 320)
    facade["enc320_sha256"] = sha("tests/enc320.mp3")
    facade["enc320_bytes"] = os.path.getsize("tests/enc320.mp3")
    s.encode_wav_to_mp3("tests/out.wav", "tests/enc128.mp3", 128)
    facade["enc128_sha256"] = sha("tests/enc128.mp3")
    facade["enc128_bytes"] = os.path.getsize("tests/enc128.mp3")
    facade["hide_ddd_returns"] = bool(s.hide_message("tests/test.mp3", "tests/hid.mp3", "ddd"))
    facade["hid_sha256"] = sha("tests/hid.mp3")
    s.reveal_massage("tests/hid.mp3", "tests/reveal1.txt")
    facade["reveal_hid"] = open("tests/reveal1.txt", "rb").read().decode("utf-8")
    s.clear_file("tests/hid.mp3", "tests/cleared.mp3")
    facade["cleared_sha256"] = sha("tests/cleared.mp3")
    s.reveal_massage("tests/cleared.mp3", "tests/reveal2.txt")
    facade["reveal_cleared"] = open("tests/reveal2.txt", "rb").read().decode("utf-8")
    facade["hide_long_returns"] = bool(s.hide_message("tests/test.mp3", "tests/hid_long.mp3", "ddd" * 100))
    facade["hid_long_sha256"] = sha("tests/hid_long.mp3")
    json.dump(facade, open(os.path.join(HERE, "ref_facade.json"), "w"), indent=1, sort_keys=True)
    # out.wav int16 PCM is the encoder input for the facade composites: keep the WAV itself (166 KB)
    shutil.copy("tests/out.wav", os.path.join(HERE, "ref_test_out.wav"))
    for name in ("enc320", "enc128", "hid", "cleared", "hid_long"):
        shutil.copy("tests/%s.mp3" % name, os.path.join(HERE, "ref_test_%s.mp3" % name))

    # ---- decoder taps on tests/test.mp3
    d = ref_decode_taps(open("tests/test.mp3", "rb").read())
    keep = list(range(6)) + [34, 35]
    pcm64 = np.concatenate([d["pcm"][k * 1152:(k + 1) * 1152] for k in keep])
    np.savez_compressed(os.path.join(HERE, "ref_test_mp3.npz"), n_frames=d["n_frames"], pcm16=d["pcm16"],
                        pcm64_frames=np.array(keep), pcm64=pcm64, spectra=d["spectra"].astype(np.int16),
                        side=d["side"].astype(np.int16), tables=d["tables"], bits=np.array(d["bits"]),
                        bitrate=d["bitrate"], sampling_rate=d["sampling_rate"],
                        xr_frames=np.array([0, 1, 35]), xr=d["xr"][[0, 1, 35]])

    # ---- encoder taps on synthetic tone+noise
    msg = "mp3stego B200 parity: the quick brown fox jumps over the lazy dog 0123456789"
    cases = [("s11_128_plain", 11, 10, 128, ""), ("s11_128_hide", 11, 10, 128, str_to_bits("%d#%s" % (len(msg), msg))),
             ("s12_320_plain", 12, 8, 320, ""), ("s12_320_hide", 12, 8, 320, str_to_bits("5#hello")),
             ("s13_64_hide", 13, 6, 64, str_to_bits("%d#%s" % (len(msg), msg)))]
    for name, seed, nfr, br, bits in cases:
        pcm = synth_wav(seed, nfr)
        wavfile.write("tests/%s.wav" % name, 44100, pcm)
        e = ref_encode_taps("tests/%s.wav" % name, br, bits)
        dd = ref_decode_taps(bytes(e["mp3"]))
        np.savez_compressed(os.path.join(HERE, "ref_synth_%s.npz" % name), seed=seed, n_frames=nfr, bitrate=br,
                            hide_bits=np.array(bits), pcm_in=pcm, mp3=e["mp3"], mdct=e["mdct"][:3],
                            ix=e["ix"].astype(np.int16), info=e["info"], scfsi=e["scfsi"].astype(np.int8),
                            hide_str_offset=e["hide_str_offset"],
                            dec_n_frames=dd["n_frames"], dec_pcm16=dd["pcm16"], dec_spectra=dd["spectra"].astype(np.int16),
                            dec_bits=np.array(dd["bits"]), dec_tables=dd["tables"])
        print(name, "frames", nfr, "bytes", len(e["mp3"]), "hide_off", e["hide_str_offset"], "dec frames", dd["n_frames"])

    # a quiet / fading clip that exercises big_values == 0 probes (A.E6) and digital silence
    rng = np.random.default_rng(99)
    n = 10 * 1152
    env = np.concatenate([np.zeros(1152 * 2), np.linspace(0, 1, 1152 * 3) ** 4 * 0.01, np.full(1152 * 2, 2e-4),
                          np.zeros(1152), np.full(1152 * 2, 0.3)])
    x = np.stack([env * np.sin(2 * np.pi * 440 * np.arange(n) / 44100) + env * 0.1 * rng.standard_normal(n),
                  env * np.sin(2 * np.pi * 1000 * np.arange(n) / 44100)], axis=1)
    pcm = (x * 32767).astype(np.int16)
    wavfile.write("tests/quiet.wav", 44100, pcm)
    for name, br, bits in (("quiet_128_plain", 128, ""), ("quiet_128_hide", 128, str_to_bits("9#quietclip"))):
        e = ref_encode_taps("tests/quiet.wav", br, bits)
        dd = ref_decode_taps(bytes(e["mp3"]))
        np.savez_compressed(os.path.join(HERE, "ref_synth_%s.npz" % name), seed=99, n_frames=10, bitrate=br,
                            hide_bits=np.array(bits), pcm_in=pcm, mp3=e["mp3"], mdct=e["mdct"][:3],
                            ix=e["ix"].astype(np.int16), info=e["info"], scfsi=e["scfsi"].astype(np.int8),
                            hide_str_offset=e["hide_str_offset"],
                            dec_n_frames=dd["n_frames"], dec_pcm16=dd["pcm16"], dec_spectra=dd["spectra"].astype(np.int16),
                            dec_bits=np.array(dd["bits"]), dec_tables=dd["tables"])
        print(name, "bytes", len(e["mp3"]), "hide_off", e["hide_str_offset"],
              "bv0 granules", int((e["info"][:, :, :, 1] == 0).sum()))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
