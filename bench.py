#!/usr/bin/env python3
"""bench.py -- throughput of the mp3stego hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (CUDA path through the C ABI)
  python bench.py --impl reference [...]                          CPU arm: the oracle port on all host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...           one rank per GPU, files sharded by rank (weak scaling)

A step is ONE pass of decode+reveal over the whole per-GPU corpus (BASELINE.json configs[1]: 1,000 synthetic
320 kbps 44.1 kHz stereo 3-minute MP3s = 6.89 M frames), processed in HBM-sized waves of files.
  value        frames/s with the MP3 bytes already resident in HBM (device pointers through the C ABI)
  e2e          the same through the C ABI with HOST buffers: H2D of the MP3 bytes and D2H of PCM + reveal bits timed
  roofline     the dominant kernel's algorithmic bytes (5,652.9 B/frame) over its mean device time
  cpu_baseline the oracle port on one host core over a bounded sample of the same corpus
  encode_hide  (second half of the metric) BASELINE.json configs[2]: WAV -> 128 kbps MP3 hiding a full-capacity payload
The oracle is used ONLY by the cpu_baseline / --impl reference legs.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mp3-steganography-lib_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

DEC_BYTES_PER_FRAME = 1044.9 + 4608.0   # compressed frame @320k + int16 stereo PCM
ENC_BYTES_PER_FRAME = 4608.0 + 417.96   # int16 stereo PCM + compressed frame @128k
FRAMES_PER_FILE = 6890                  # 3 minutes at 44.1 kHz (7,937,280 samples)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# corpus
# ------------------------------------------------------------------------------------------------
def _frame_sizes(mp3: bytes):
    """Byte offsets / sizes of the MPEG-1 Layer III frames of a clean CBR clip (host-side header walk)."""
    rates = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
    srs = [44100, 48000, 32000]
    off, out = 0, []
    while off + 4 < len(mp3) and mp3[off] == 0xFF and mp3[off + 1] >= 0xE0:
        b2 = mp3[off + 2]
        fs = 144000 * rates[b2 >> 4] // srs[(b2 >> 2) & 3] + ((b2 >> 1) & 1)
        out.append((off, fs))
        off += fs
    return out


def _pad_clip(mp3: bytes) -> bytes:
    """The reference's bit writer drops the last 0-3 bytes (4-byte flush, MP3_Encoder.py:1370-1392): pad the last
    frame back to its nominal size so that clips can be laid end to end as one valid stream."""
    fr = _frame_sizes(mp3)
    end = fr[-1][0] + fr[-1][1]
    return mp3 + b"\x00" * max(0, end - len(mp3))


def fixture_clips_320():
    """Reference-encoded 320 kbps clips committed under tests/golden (every frame has main_data_begin = 0)."""
    g = os.path.join(ROOT, "tests", "golden")
    clips = [open(os.path.join(g, n), "rb").read() for n in
             ("test.mp3", "ref_test_enc320.mp3", "ref_test_hid.mp3", "ref_test_cleared.mp3", "ref_test_hid_long.mp3")]
    for n in ("ref_synth_s12_320_plain.npz", "ref_synth_s12_320_hide.npz"):
        clips.append(np.load(os.path.join(g, n))["mp3"].tobytes())
    return [np.frombuffer(_pad_clip(c), np.uint8) for c in clips], [len(_frame_sizes(c)) for c in clips]


def synth_pcm_device(torch, n_files, n_frames, seed0, device):
    """SURVEY.md 8(d) tone+noise WAVs generated on the device (setup only; torch is plumbing here):
    L/R = 0.4 sin(2 pi f t) + 0.05 N(0,1), f ~ U[100, 5000] Hz, *32767 -> int16, interleaved stereo."""
    n = n_frames * 1152
    g = torch.Generator(device=device)
    g.manual_seed(seed0)
    f = torch.rand((n_files, 1, 2), generator=g, device=device) * 4900.0 + 100.0
    t = (torch.arange(n, device=device, dtype=torch.float64) / 44100.0).reshape(1, n, 1)
    x = 0.4 * torch.sin((2 * np.pi) * f.double() * t).float()
    x += 0.05 * torch.randn((n_files, n, 2), generator=g, device=device)
    return (x * 32767.0).to(torch.int16).reshape(n_files, n * 2)


def build_corpus_host(n_files, frames_per_file, seed):
    """Concatenate fixture clips (random order per file) into n_files streams of ~frames_per_file frames."""
    clips, nfr = fixture_clips_320()
    rng = np.random.default_rng(seed)
    files, frames = [], []
    for _ in range(n_files):
        parts, tot = [], 0
        while tot < frames_per_file:
            i = int(rng.integers(0, len(clips)))
            if tot + nfr[i] > frames_per_file:
                # finish with leading frames of a clip
                fs = _frame_sizes(clips[i].tobytes())
                k = frames_per_file - tot
                parts.append(clips[i][: fs[k - 1][0] + fs[k - 1][1]])
                tot += k
            else:
                parts.append(clips[i])
                tot += nfr[i]
        files.append(np.concatenate(parts))
        frames.append(tot)
    return files, frames


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for k, nme in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        if sm:
            hi = [v for v in sm if v >= 0.5 * max(sm)]   # samples under load
            out["sm_mhz"] = float(np.median(hi))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# reference / cpu arm (the oracle port; the ONLY code in this file that touches oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_decode_worker(blob):
    from oracle import oracle as O
    r = O.decode(blob, 0, taps=False)
    return int(r["n_frames"]), len(r["bits"])


def cpu_decode_sample(blobs, procs):
    """Decode+reveal `blobs` with the oracle on `procs` processes; returns (frames, seconds)."""
    from oracle import oracle as O
    O.build()
    t0 = time.perf_counter()
    if procs <= 1:
        res = [_cpu_decode_worker(b) for b in blobs]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_decode_worker, blobs, chunksize=1)
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames_per_blob = 1378  # one fifth of a 3-minute file per worker per step keeps a step to a few seconds
    files, frames = build_corpus_host(cores, frames_per_blob, seed=12345)
    blobs = [f.tobytes() for f in files]
    for _ in range(args.warmup):
        cpu_decode_sample(blobs[: max(1, cores // 4)], cores)
    tot_f, tot_t = 0, 0.0
    for _ in range(args.steps):
        f, t = cpu_decode_sample(blobs, cores)
        tot_f += f
        tot_t += t
    v = tot_f / tot_t
    sample = f"{cores} clips x {frames_per_blob} frames of the 320 kbps corpus per step, one process per host core"
    line = {"impl": "reference", "metric": "decode+reveal throughput (MP3 frames/s)", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": corpus_config(args),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "audio_seconds_per_s": v * 1152 / 44100.0}
    print(json.dumps(line), flush=True)


def corpus_config(args):
    return {"workload": f"configs[1]: batch decode+reveal of {args.files} synthetic 320 kbps 44.1 kHz stereo "
                        f"{args.frames * 1152 / 44100.0:.0f}-s MP3s per GPU ({args.files * args.frames} frames)",
            "files_per_gpu": args.files, "frames_per_file": args.frames, "wave_files": args.wave,
            "l2": "inputs larger than L2 (each wave reads >= 0.9 GB of MP3 and writes >= 4 GB of PCM)",
            "corpus": "reference-encoded 320 kbps clips (tests/golden) laid end to end in seeded random order"}


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def run_product_arm(args):
    import torch
    import __graft_entry__ as ge
    ge.build()
    from mp3stego_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- corpus: per-rank shard of files (weak scaling: every rank holds `files` files)
    t0 = time.perf_counter()
    files, frames = build_corpus_host(args.files, args.frames, seed=1000 + rank)
    sizes = np.array([len(f) for f in files], np.int64)
    n_waves = (args.files + args.wave - 1) // args.wave
    waves = []
    for w in range(n_waves):
        lo, hi = w * args.wave, min(args.files, (w + 1) * args.wave)
        off = np.concatenate([[0], np.cumsum(sizes[lo:hi])])
        host = torch.empty(int(off[-1]) + 64, dtype=torch.uint8, pin_memory=True)
        hv = host.numpy()
        for i in range(lo, hi):
            hv[off[i - lo]: off[i - lo + 1]] = files[i]
        waves.append(dict(off=off, host=host, dev=host.to(dev), frames=int(sum(frames[lo:hi])), n=hi - lo))
    del files
    total_frames = sum(w["frames"] for w in waves)
    max_wave_frames = max(w["frames"] for w in waves)
    log(f"[rank {rank}] corpus: {args.files} files, {total_frames} frames, {sizes.sum() / 1e9:.2f} GB in {n_waves} waves "
        f"({time.perf_counter() - t0:.1f}s)")

    h = _lib.Handle(local)
    stream = torch.cuda.Stream(device=dev)   # the library launches on this stream, and so do the timing events
    h.set_stream(stream.cuda_stream)
    pcm_dev = torch.empty(max_wave_frames * 1152 * 2 + 64, dtype=torch.int16, device=dev)
    pcm_host = torch.empty(max_wave_frames * 1152 * 2 + 64, dtype=torch.int16, pin_memory=True)
    ids_dev = torch.empty(max_wave_frames * 12, dtype=torch.uint8, device=dev)
    bits_dev = torch.empty(max_wave_frames * 12, dtype=torch.uint8, device=dev)
    ids_host = np.zeros(max_wave_frames * 12, np.uint8)
    bits_host = np.zeros(max_wave_frames * 12, np.uint8)
    reveal_total = [0]

    def step_device():
        n = 0
        for w in waves:
            sc = h.decode_scan(w["dev"], w["off"])
            ln = h.decode_reveal_into(ids_dev, bits_dev)
            h.decode_run(pcm=pcm_dev)
            n += int(sc["n_frames"].sum())
            reveal_total[0] = int(ln.sum())
        return n

    def step_host():
        n = 0
        for w in waves:
            sc = h.decode_scan(w["host"], w["off"])
            h.decode_reveal_into(ids_host, bits_host)
            h.decode_run(pcm=pcm_host)
            n += int(sc["n_frames"].sum())
        return n

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        h.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        n = 0
        for _ in range(steps):
            n += fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms / 1e3, wall], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return n, float(t[0].item()), float(t[1].item())

    # ---- warm-up, then the timed device-resident region
    for _ in range(args.warmup):
        step_device()
    h.timing_enable(True)
    l0 = h.launches
    clocks = ClockSampler(local)
    n_dev, t_dev, wall_dev = timed(step_device, args.steps)
    clk = clocks.stop()
    launches = h.launches - l0
    ktimes = h.timing()
    h.timing_enable(False)
    assert n_dev == total_frames * args.steps, (n_dev, total_frames)

    # ---- end-to-end through the C ABI with host buffers
    step_host()
    n_e2e, t_e2e, _ = timed(step_host, args.steps)
    h2d = int(sizes.sum())
    d2h = int(total_frames * 1152 * 2 * 2 + 2 * 12 * total_frames)

    value = world * n_dev / t_dev
    e2e_value = world * n_e2e / t_e2e

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    dom = max(ktimes.items(), key=lambda kv: kv[1][0]) if ktimes else (None, (0.0, 0))
    roof = None
    if dom[0]:
        ms_total, n_l = dom[1]
        frames_per_launch = n_dev / n_l
        ach = DEC_BYTES_PER_FRAME * frames_per_launch / (ms_total / n_l * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom[0], {}).get("bytes_per_frame")
            if traffic is not None:
                traffic = traffic * frames_per_launch
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": ms_total / n_l,
                "frames_per_launch": frames_per_launch, "algorithmic_bytes_per_frame": DEC_BYTES_PER_FRAME,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(ktimes.items())},
                "whole_path_frac": value / world * DEC_BYTES_PER_FRAME / 1e9 / peak}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle port on one host core over a bounded sample of the same corpus
    sample_files, _ = build_corpus_host(1, min(args.frames, 6890 * 3), seed=777)
    sample_files = [sample_files[0].tobytes()] * (3 if args.frames >= 6890 else 1)
    cf, ct = cpu_decode_sample(sample_files, 1)
    cpu = {"value": cf / ct, "unit": "frames/s", "cores": 1, "kind": "port",
           "sample": f"{len(sample_files)} x {cf // len(sample_files)}-frame 320 kbps files of the same corpus, "
                     f"oracle/mp3stego_oracle.c decode+reveal, 1 thread of {os.cpu_count()} host cores, {ct:.1f} s"}

    line = {"metric": "decode+reveal throughput (MP3 frames/s)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": corpus_config(args),
            "audio_seconds_per_s": value * 1152 / 44100.0,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "wall_s_timed_region": wall_dev, "reveal_bits_last_wave": reveal_total[0]}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files", type=int, default=1000, help="files per GPU")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_FILE, help="frames per file")
    ap.add_argument("--wave", type=int, default=125, help="files per wave (bounds the device workspaces)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()
