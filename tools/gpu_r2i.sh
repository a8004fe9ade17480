#!/bin/bash
# round 2, pass i: the float64 instantiation of the fast hybrid kernel -- GPU tests (every exactness gate), composite throughput old vs new
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/tests_gpu.log
for v in fast direct; do
  if [ $v = direct ]; then export M3S_HYBRID_DIRECT=1; else unset M3S_HYBRID_DIRECT; fi
  timeout 600 python - <<'PY'
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "mp3-steganography-lib_b200")
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
batch.hide_batch(h, blobs, msgs)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): out, too = batch.hide_batch(h, blobs, msgs)
torch.cuda.synchronize(); th = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(3): cl = batch.clear_batch(h, blobs)
torch.cuda.synchronize(); tc = (time.perf_counter() - t0) / 3
# the exact decode alone, device resident
data = torch.from_numpy(np.frombuffer(b"".join(blobs), np.uint8).copy()).cuda()
off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])])
out_pcm = torch.empty(n * f * 2304 + 64, dtype=torch.int16, device="cuda")
h.decode(data, off, pcm=out_pcm, frames_bound=n * f + 8, exact=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): h.decode(data, off, pcm=out_pcm, frames_bound=n * f + 8, exact=True)
torch.cuda.synchronize(); td = (time.perf_counter() - t0) / 3
import hashlib
print(os.environ.get("M3S_HYBRID_DIRECT", "fast"), "hide %.0f frames/s clear %.0f frames/s exact decode %.3g frames/s" % (n * f / th, n * f / tc, n * f / td),
      "sha", hashlib.sha256(out_pcm.cpu().numpy().tobytes()).hexdigest()[:16], hashlib.sha256(b"".join(out)).hexdigest()[:16])
PY
done
