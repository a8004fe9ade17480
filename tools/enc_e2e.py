"""Host-buffer encode+hide under different chunk budgets / pipeline settings (diagnostic; no JSON contract).
usage: python tools/enc_e2e.py [clips] [frames] [budget ...]   -- pinned PCM in, pinned MP3 out, CUDA-event time of the whole call."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3445
    budgets = [int(a) for a in sys.argv[3:]] or [0]
    ge.build()
    from mp3stego_b200 import _lib
    dev = torch.device("cuda", 0)
    n_samp = frames * 1152
    pcm = torch.empty(clips * n_samp * 2, dtype=torch.int16, pin_memory=True)
    for lo in range(0, clips, 250):
        hi = min(clips, lo + 250)
        pcm[lo * n_samp * 2: hi * n_samp * 2] = bench.synth_pcm_device(torch, hi - lo, frames, 1000 + lo, dev).reshape(-1).cpu()
    pay, pay_off = bench.random_payload_bits(clips, bench.PAYLOAD_BITS_PER_FRAME * frames, 5)
    cap = int(_lib.load().m3s_encode_bound(n_samp, 44100, 128)) * clips + 64
    out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    ns = [n_samp] * clips
    floor_ms = pcm.numel() * 2 / 55.6e6
    for b in budgets:
        if b:
            os.environ["M3S_ENC_CHUNK_FRAMES"] = str(b)
        else:
            os.environ.pop("M3S_ENC_CHUNK_FRAMES", None)
        h = _lib.Handle(0)
        h.encode(pcm, ns, 44100, 128, payload_packed=(pay, pay_off), mp3_out=out)
        h.timing_enable(os.environ.get("ENC_E2E_TIMING", "0") == "1")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(2):
            h.encode(pcm, ns, 44100, 128, payload_packed=(pay, pay_off), mp3_out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        kt = h.timing()
        ks = ", ".join(f"{k} {v[0] / 2:.1f}" for k, v in sorted(kt.items()) if k.startswith("k_enc") and v[1])
        print(f"budget {b:8d}: {ms:8.1f} ms/pass = {clips * frames / ms / 1e3:6.2f} M frames/s  (H2D floor {floor_ms:.0f} ms)  [{ks}]", flush=True)
        del h


if __name__ == "__main__":
    main()
