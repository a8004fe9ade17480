#!/usr/bin/env python3
"""Diagnostic (not a test): decode every golden MP3 on the GPU and print where it differs from the oracle."""
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mp3-steganography-lib_b200"))
from mp3stego_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
cases = {"test.mp3": open(os.path.join(G, "test.mp3"), "rb").read()}
for p in sorted(glob.glob(os.path.join(G, "ref_synth_*.npz"))):
    cases[os.path.basename(p)] = np.load(p)["mp3"].tobytes()
for p in sorted(glob.glob(os.path.join(G, "ref_test_*.mp3"))):
    cases[os.path.basename(p)] = open(p, "rb").read()
for p in sorted(glob.glob(os.path.join(G, "stream_*.mp3"))):
    cases[os.path.basename(p)] = open(p, "rb").read()

h = _lib.Handle(0)
names = list(cases)
blob = b"".join(cases[n] for n in names)
offs = np.concatenate([[0], np.cumsum([len(cases[n]) for n in names])])
t0 = time.time()
sc = h.decode_scan(np.frombuffer(blob, np.uint8), offs)
ids, bits = h.decode_reveal()
pcm, sp = h.decode_run(spectra=True)
pcmf, _ = h.decode_run(as_float=True)
print("gpu batch decode of %d files took %.3fs, launches=%d" % (len(names), time.time() - t0, h.launches))
fb = np.concatenate([[0], np.cumsum(sc["n_frames"])])
eb = np.concatenate([[0], np.cumsum(sc["pcm_rows"] * np.maximum(sc["channels"], 1))])
bad = 0
for i, n in enumerate(names):
    ref = O.decode(cases[n])
    nf = int(sc["n_frames"][i])
    ok_nf = nf == ref["n_frames"]
    s = sp[fb[i]:fb[i + 1]].astype(np.int32)
    ok_sp = ok_nf and np.array_equal(s, ref["spectra"])
    ok_ids = ok_nf and np.array_equal(ids[fb[i]:fb[i + 1]], ref["tables"])
    ok_bits = bits[i] == ref["bits"]
    ch = max(int(sc["channels"][i]), 1)
    p16 = pcm[eb[i]:eb[i + 1]].reshape(-1, ch).astype(np.int32)
    pf = pcmf[eb[i]:eb[i + 1]].reshape(-1, ch)
    if p16.shape == ref["pcm16"].shape:
        d = np.abs(p16 - ref["pcm16"].astype(np.int32))
        df = np.abs(pf - ref["pcm"])
        msg = "pcm max %d LSB (%.3f%% off), float max %.2e" % (d.max(), (d > 0).mean() * 100, df.max())
        ok_pcm = d.max() <= 1 and df.max() <= 1e-5
    else:
        msg = "pcm shape %s vs %s" % (p16.shape, ref["pcm16"].shape)
        ok_pcm = False
    print("%-34s frames %d/%d spectra %s ids %s bits %s  %s  status %d" % (n, nf, ref["n_frames"], ok_sp, ok_ids, ok_bits, msg, sc["status"][i]))
    if not ok_sp and ok_nf:
        w = np.argwhere(s != ref["spectra"])
        print("   first spectra mismatches (frame, gr, ch, i):", w[:5].tolist(), "count", len(w))
        f0, g0, c0, i0 = w[0]
        print("   got", s[f0, g0, c0, max(0, i0 - 4):i0 + 8].tolist(), "ref", ref["spectra"][f0, g0, c0, max(0, i0 - 4):i0 + 8].tolist())
        print("   side", ref["side"][f0, g0, c0].tolist())
    if not ok_pcm and p16.shape == ref["pcm16"].shape:
        w = np.argwhere(d > 1)
        print("   first pcm mismatches (row, ch):", w[:5].tolist(), "count", len(w))
    bad += not (ok_sp and ok_ids and ok_bits and ok_pcm)
print("FAILED cases:", bad)
