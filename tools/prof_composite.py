"""Where the composites spend their host time: wall clock of the stages of batch.hide_batch / clear_batch (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mp3-steganography-lib_b200"))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from mp3stego_b200 import _lib, batch
import bench
h = _lib.Handle(0)
n, f = 256, 689
pcm = bench.synth_pcm_device(torch, n, f, 77, torch.device("cuda", 0)).reshape(-1)
r = h.encode(pcm, [f * 1152] * n, 44100, 320, compact=True)
mp3 = r["mp3"].cpu().numpy()
blobs = [bytes(mp3[int(r["mp3_off"][i]): int(r["mp3_off"][i]) + int(r["out_len"][i])]) for i in range(n)]
msgs = ["m%d" % i * 10 for i in range(n)]
T = {}
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return out
    setattr(obj, name, w)
for nm in ("decode", "encode", "decode_scan"):
    wrap(h, nm)
wrap(batch, "_concat")
for what in ("hide", "clear"):
    for rep in range(2):
        T.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if what == "hide": batch.hide_batch(h, blobs, msgs)
        else: batch.clear_batch(h, blobs)
        torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(what, "total %.1f ms" % (1e3 * tot), {k: round(1e3 * v, 1) for k, v in T.items()})
