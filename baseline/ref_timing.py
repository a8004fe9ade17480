#!/usr/bin/env python3
"""Time the UNMODIFIED Python/numba reference (mp3stego-lib 1.1.8, installed into baseline/_ref by __graft_entry__.build())
on this host, through its own public classes, on one core (the reference is single-threaded):

  decode+reveal   Steganography.reveal_massage(mp3, txt)                      (steganography.py:110-125)
  encode+hide     Encoder(wav, mp3, bitrate=128, hide_str=bits).encode()      (encoder.py:22-51; what hide_message calls)

Each operation is warmed once on a 4-frame clip (numba JIT excluded) and then timed on tone+noise clips of the length given
(reference throughput falls with file length -- O(N^2) list handling, BASELINE.md 2 -- so the length is part of the result).
Prints ONE JSON object.  bench.py runs this in a subprocess; nothing of the product is imported here."""
import argparse
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-dec", type=int, default=100)
    ap.add_argument("--frames-enc", type=int, default=50)
    ap.add_argument("--test-mp3", default="")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "mp3stego")):
        print(json.dumps({"unavailable": "baseline/_ref/mp3stego is missing (run __graft_entry__.build() where /root/reference exists)"}))
        return
    sys.path.insert(0, REF)
    try:
        import numba  # noqa: F401
        import numpy as np
        from scipy.io import wavfile
        import tqdm
        import mp3stego.decoder.MP3_Parser as _mp
        import mp3stego.encoder.MP3_Encoder as _me
        _mp.tqdm = lambda *x, **k: tqdm.tqdm(*x, **{**k, "disable": True})
        _me.tqdm = lambda *x, **k: tqdm.tqdm(*x, **{**k, "disable": True})
        from mp3stego import Steganography
        from mp3stego.encoder.encoder import Encoder
        from mp3stego.steganography import str_to_binary_str
    except Exception as e:   # numba or scipy absent on this box
        print(json.dumps({"unavailable": f"reference import failed: {type(e).__name__}: {e}"}))
        return

    def synth(n_frames, seed):
        rng = np.random.default_rng(seed)
        n = n_frames * 1152
        t = np.arange(n) / 44100.0
        f = rng.uniform(100, 5000, size=2)
        x = np.stack([0.4 * np.sin(2 * np.pi * f[c] * t) + 0.05 * rng.standard_normal(n) for c in range(2)], axis=1)
        return (x * 32767).astype(np.int16)

    out = {"python": sys.version.split()[0], "numba": numba.__version__, "cores": 1}
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        s = Steganography(quiet=True)

        def enc(n_frames, seed, bitrate, msg):
            wav, mp3 = os.path.join(d, f"c{seed}.wav"), os.path.join(d, f"c{seed}_{bitrate}.mp3")
            wavfile.write(wav, 44100, synth(n_frames, seed))
            bits = str_to_binary_str(str(len(msg)) + "#" + msg) if msg else ""
            t0 = time.perf_counter()
            Encoder(wav, mp3, bitrate=bitrate, hide_str=bits).encode(quiet=True)
            return time.perf_counter() - t0, mp3

        enc(4, 1, 128, "warm")                                   # JIT of the encoder
        _, m = enc(4, 2, 320, "")
        s.reveal_massage(m, os.path.join(d, "w.txt"))            # JIT of the decoder
        msg = "".join(chr(32 + (7 * i) % 95) for i in range(2 * a.frames_enc))   # ~16 bits per frame: beyond capacity
        t_enc, _ = enc(a.frames_enc, 11, 128, msg)
        out["encode_hide"] = {"value": a.frames_enc / t_enc, "unit": "frames/s", "frames": a.frames_enc, "seconds": t_enc,
                              "what": "Encoder(wav, mp3, bitrate=128, hide_str=bits).encode() on a tone+noise clip"}
        _, m320 = enc(a.frames_dec, 12, 320, "")
        t0 = time.perf_counter()
        s.reveal_massage(m320, os.path.join(d, "r.txt"))
        t_dec = time.perf_counter() - t0
        out["decode_reveal"] = {"value": a.frames_dec / t_dec, "unit": "frames/s", "frames": a.frames_dec, "seconds": t_dec,
                                "what": "Steganography.reveal_massage on a 320 kbps tone+noise clip (decode to WAV + reveal)"}
        if a.test_mp3 and os.path.exists(a.test_mp3):
            t0 = time.perf_counter()
            s.reveal_massage(a.test_mp3, os.path.join(d, "t.txt"))
            t = time.perf_counter() - t0
            out["configs0_test_mp3"] = {"value": 36 / t, "unit": "frames/s", "frames": 36, "seconds": t,
                                        "what": "Steganography.reveal_massage(tests/test.mp3): BASELINE configs[0]"}
        os.chdir(HERE)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
