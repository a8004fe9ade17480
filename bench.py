#!/usr/bin/env python3
"""bench.py -- throughput of the mp3stego hot path on B200 (contract: task brief; DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (CUDA path through the C ABI)
  python bench.py --impl reference [...]                          CPU arm: the oracle port on all host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...           one rank per GPU, files sharded by rank (weak scaling)

Corpus (per GPU, synthetic): `--files` tone+noise 44.1 kHz stereo WAVs of `--frames` frames, generated
on the device; the decode corpus is those WAVs encoded to 320 kbps by the product's own encoder (bit-exact with the
reference encoder -- tests/test_parity_encode.py), one MP3 per file.

A step is ONE pass over the whole per-GPU corpus, processed in waves of `--wave` files:
  decode+reveal (the JSON line's metric; BASELINE.json configs[1]: 1,000 x 3-minute 320 kbps files = 6.89 M frames)
      value        frames/s with the MP3 bytes resident in HBM (device pointers through the C ABI)
      e2e          the same through the C ABI with HOST buffers: H2D of MP3 bytes, D2H of PCM + table ids + reveal bits
      roofline     dominant kernel: algorithmic bytes (5,652.9 B/frame) over its mean device time
      cpu_baseline the oracle port on one host core over a bounded sample of the same corpus
  encode+hide  (key "encode_hide"; configs[2]: the same WAVs -> 128 kbps hiding a random payload beyond capacity)
      the same five figures for the encoder.
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
PAYLOAD_BITS_PER_FRAME = 14             # measured capacity is 11-12 bits/frame: the payload always exceeds it


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic corpus
# ------------------------------------------------------------------------------------------------
def synth_pcm_host(n_frames, seed):
    """SURVEY.md 8(d) tone+noise clip on the host (cpu / reference legs): int16 [n, 2]."""
    rng = np.random.default_rng(seed)
    n = n_frames * 1152
    t = np.arange(n) / 44100.0
    f = rng.uniform(100, 5000, size=2)
    x = np.stack([0.4 * np.sin(2 * np.pi * f[c] * t) + 0.05 * rng.standard_normal(n) for c in range(2)], axis=1)
    return (x * 32767).astype(np.int16)


def synth_pcm_device(torch, n_files, n_frames, seed, device):
    """The same recipe on the device (setup only; torch is plumbing): int16 [n_files, 2 * n] interleaved stereo."""
    n = n_frames * 1152
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f = torch.rand((n_files, 1, 2), generator=g, device=device, dtype=torch.float64) * 4900.0 + 100.0
    out = torch.empty((n_files, n, 2), dtype=torch.int16, device=device)
    step = max(1, (1 << 26) // max(n, 1))
    idx = torch.arange(n, device=device, dtype=torch.float64).reshape(1, n, 1)
    for lo in range(0, n_files, step):
        hi = min(n_files, lo + step)
        ph = torch.remainder(f[lo:hi] * (idx / 44100.0), 1.0).float()        # phase in turns, reduced in float64
        x = 0.4 * torch.sin(ph * (2 * np.pi))
        x += 0.05 * torch.randn((hi - lo, n, 2), generator=g, device=device)
        out[lo:hi] = (x * 32767.0).to(torch.int16)
    return out.reshape(n_files, n * 2)


def random_payload_bits(n_files, bits_per_file, seed):
    """Random 7-bit ASCII (rng.integers(32, 127)) as '0'/'1' chars, all files end to end + offsets."""
    rng = np.random.default_rng(seed)
    nbytes = (bits_per_file + 7) // 8
    chars = rng.integers(32, 127, size=(n_files, nbytes), dtype=np.uint8)
    bits = np.unpackbits(chars, axis=1)[:, :bits_per_file]
    packed = (bits + ord("0")).astype(np.uint8).reshape(-1)
    off = np.arange(n_files + 1, dtype=np.int64) * bits_per_file
    return packed, off


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
        sm, power, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                power.append(float(r[2]))
                for k, nme in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        if sm:
            hi = [v for v, p in zip(sm, power) if p >= 0.6 * max(power)] or sm   # samples under load
            out["sm_mhz"] = float(np.median(hi))
            out["samples"] = len(sm)
            out["power_w_max"] = max(power)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# reference / cpu arm (the oracle port; the ONLY code in this file that touches oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_decode_worker(blob):
    from oracle import oracle as O
    r = O.decode(blob, 0, taps=False)
    return int(r["n_frames"]), len(r["bits"])


def _cpu_encode_worker(args):
    from oracle import oracle as O
    pcm, bitrate, bits = args
    r = O.encode(pcm, 44100, bitrate, bits, taps=False)
    return int(r["n_frames"]), r["mp3"]


def _pool_map(fn, items, procs):
    if procs <= 1:
        return [fn(i) for i in items]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(fn, items, chunksize=1)


def cpu_sample_clips(n_clips, n_frames, seed):
    from oracle import oracle as O
    O.build()
    pcms = [synth_pcm_host(n_frames, seed + i) for i in range(n_clips)]
    return pcms


def cpu_baselines(n_frames, procs, n_clips):
    """Encode+hide @128k and decode+reveal @320k of n_clips tone+noise clips with the oracle on `procs` processes.
    Returns {"decode": (frames, s), "encode": (frames, s)}."""
    pcms = cpu_sample_clips(n_clips, n_frames, 4242)
    packed, off = random_payload_bits(n_clips, PAYLOAD_BITS_PER_FRAME * n_frames, 7)
    bits = [packed[off[i]:off[i + 1]].tobytes().decode("ascii") for i in range(n_clips)]
    mp3s = [m for _, m in _pool_map(_cpu_encode_worker, [(p, 320, "") for p in pcms], procs)]   # decode inputs (untimed)
    t0 = time.perf_counter()
    res = _pool_map(_cpu_encode_worker, [(p, 128, b) for p, b in zip(pcms, bits)], procs)
    t_enc = time.perf_counter() - t0
    f_enc = sum(r[0] for r in res)
    t0 = time.perf_counter()
    res = _pool_map(_cpu_decode_worker, mp3s, procs)
    t_dec = time.perf_counter() - t0
    f_dec = sum(r[0] for r in res)
    return {"decode": (f_dec, t_dec), "encode": (f_enc, t_enc)}


def corpus_config(args):
    return {"workload": f"configs[1]: batch decode+reveal of {args.files} synthetic 320 kbps 44.1 kHz stereo "
                        f"{args.frames * 1152 / 44100.0:.0f}-s MP3s per GPU ({args.files * args.frames} frames); "
                        f"encode_hide = configs[2]: the same WAVs -> 128 kbps hiding random ASCII beyond capacity",
            "files_per_gpu": args.files, "frames_per_file": args.frames, "wave_files": args.wave, "e2e_wave_files": args.e2e_wave, "e2e_workers": args.e2e_workers,
            "l2": "inputs larger than L2 (every wave streams >= 0.36 GB of MP3 and >= 1.6 GB of PCM; L2 is 126 MB)",
            "corpus": "tone+noise WAVs (SURVEY 8d) generated on device; MP3s produced from them by the product encoder "
                      "(byte-identical to the reference encoder's output)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_frames = 689      # one tenth of a 3-minute file per worker per step keeps a step to a few seconds
    for _ in range(args.warmup):
        cpu_baselines(60, cores, cores)
    dec_f = dec_t = enc_f = enc_t = 0.0
    for _ in range(args.steps):
        r = cpu_baselines(n_frames, cores, cores)
        dec_f += r["decode"][0]; dec_t += r["decode"][1]
        enc_f += r["encode"][0]; enc_t += r["encode"][1]
    v = dec_f / dec_t
    sample = f"{cores} clips x {n_frames} frames of the tone+noise corpus per step, one oracle process per host core"
    line = {"impl": "reference", "metric": "decode+reveal throughput (MP3 frames/s)", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dec_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": corpus_config(args),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "audio_seconds_per_s": v * 1152 / 44100.0,
            "encode_hide": {"value": enc_f / enc_t, "unit": "frames/s", "ms_per_step": 1e3 * enc_t / args.steps,
                            "e2e": {"value": enc_f / enc_t, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                            "cpu_baseline": {"value": enc_f / enc_t, "unit": "frames/s", "cores": cores, "kind": "port",
                                             "sample": sample}}}
    emit(line)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def roofline_of(ktimes, names, frames_total, bytes_per_frame, peak, peak_src, steps, whole_frac):
    ks = {k: v for k, v in ktimes.items() if k in names}
    if not ks:
        return None
    dom, (ms_total, n_l) = max(ks.items(), key=lambda kv: kv[1][0])
    fpl = frames_total / n_l
    ach = bytes_per_frame * fpl / (ms_total / n_l * 1e-3) / 1e9
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom, {}).get("bytes_per_frame")
        traffic = None if t is None else t * fpl
    except Exception:
        pass
    return {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": ms_total / n_l, "frames_per_launch": fpl,
            "algorithmic_bytes_per_frame": bytes_per_frame,
            "kernel_ms_per_step": {k: v[0] / steps for k, v in sorted(ks.items())}, "whole_path_frac": whole_frac}


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

    h = _lib.Handle(local)
    stream = torch.cuda.Stream(device=dev, priority=-1)   # the library launches on this stream, and so do the timing events
    h.set_stream(stream.cuda_stream)

    # ---- corpus: per-rank shard of files (weak scaling: every rank holds `files` files)
    t0 = time.perf_counter()
    nw = (args.files + args.wave - 1) // args.wave
    n_samp = args.frames * 1152
    pay_bits = PAYLOAD_BITS_PER_FRAME * args.frames
    pcm_all = torch.empty(args.files * n_samp * 2, dtype=torch.int16, device=dev)
    for w in range(nw):
        lo, hi = w * args.wave, min(args.files, (w + 1) * args.wave)
        pcm_all[lo * n_samp * 2: hi * n_samp * 2] = synth_pcm_device(torch, hi - lo, args.frames, 100000 * rank + 1000 + w, dev).reshape(-1)
    torch.cuda.synchronize()
    ns_all = [n_samp] * args.files
    res = h.encode(pcm_all, ns_all, 44100, 320, compact=True)                # decode corpus (setup, untimed)
    off_all = np.concatenate([res["mp3_off"], [res["mp3_off"][-1] + res["out_len"][-1]]]).astype(np.int64)
    mp3_all = res["mp3"]
    mp3_host_all = mp3_all.cpu().pin_memory()
    del res
    def make_waves(wave_files):
        ws = []
        for lo in range(0, args.files, wave_files):
            hi = min(args.files, lo + wave_files)
            b0, b1 = int(off_all[lo]), int(off_all[hi])
            ws.append(dict(n=hi - lo, frames=(hi - lo) * args.frames, off=off_all[lo:hi + 1] - b0,
                           mp3_dev=mp3_all[b0:b1], mp3_host=mp3_host_all[b0:b1]))
        return ws

    waves = make_waves(args.wave)             # device-resident leg: large waves (fewer latency-bound scan launches)
    e2e_waves = make_waves(args.e2e_wave)     # host leg: smaller waves keep both PCIe directions busy with short fill / drain
    pay_all, pay_off_all = random_payload_bits(args.files, pay_bits, 31 * rank + 5)
    pcm_host_all = None
    total_frames = args.files * args.frames
    max_wave_frames = max(w["frames"] for w in waves)
    max_e2e_frames = max(w["frames"] for w in e2e_waves)
    mp3_bytes = int(off_all[-1])
    log(f"[rank {rank}] corpus: {args.files} files x {args.frames} frames = {total_frames} frames, "
        f"{mp3_bytes / 1e9:.2f} GB MP3 @320k, {total_frames * 4608 / 1e9:.2f} GB PCM, {nw} decode waves ({time.perf_counter() - t0:.1f}s)")

    pcm_out_dev = torch.empty(max_wave_frames * 1152 * 2 + 64, dtype=torch.int16, device=dev)
    pcm_out_host = torch.empty(max_e2e_frames * 1152 * 2 + 64, dtype=torch.int16, pin_memory=True)
    ids_dev = torch.empty(max_wave_frames * 12, dtype=torch.uint8, device=dev)
    bits_dev = torch.empty(max_wave_frames * 12, dtype=torch.uint8, device=dev)
    ids_host = torch.empty(max_e2e_frames * 12, dtype=torch.uint8, pin_memory=True)
    bits_host = torch.empty(max_e2e_frames * 12, dtype=torch.uint8, pin_memory=True)
    enc_cap = int(_lib.load().m3s_encode_bound(n_samp, 44100, 128)) * args.files + 64
    enc_out_dev = torch.empty(enc_cap, dtype=torch.uint8, device=dev)
    enc_out_host = None
    check = {}

    def dec_device():
        n = 0
        for w in waves:
            sc = h.decode_scan(w["mp3_dev"], w["off"])
            ln = h.decode_reveal_into(ids_dev, bits_dev)
            h.decode_run(pcm=pcm_out_dev)
            n += int(sc["n_frames"].sum())
            check["reveal_bits"] = int(ln.sum())
        return n

    # e2e: host buffers through the C ABI.  A host application keeps PCIe busy in both directions by running one worker
    # thread per handle (handles are independent: own stream, own workspaces, own pinned PCM buffer); the waves are dealt
    # round-robin, so wave k's PCM goes home while wave k+1 is in the kernels and wave k+2's MP3 bytes come up.
    import threading
    state = dict(waves=e2e_waves, n=max(1, min(args.e2e_workers, len(e2e_waves))), trace=None)
    workers = []

    def ensure_workers(n, frames):
        for wk in workers:
            if wk["pcm"].numel() < frames * 1152 * 2 + 64:
                wk["pcm"] = torch.empty(frames * 1152 * 2 + 64, dtype=torch.int16, pin_memory=True)
                wk["ids"] = torch.empty(frames * 12, dtype=torch.uint8, pin_memory=True)
                wk["bits"] = torch.empty(frames * 12, dtype=torch.uint8, pin_memory=True)
        while len(workers) < n:
            workers.append(dict(h=h if not workers else _lib.Handle(local),
                                pcm=torch.empty(frames * 1152 * 2 + 64, dtype=torch.int16, pin_memory=True),
                                ids=torch.empty(frames * 12, dtype=torch.uint8, pin_memory=True),
                                bits=torch.empty(frames * 12, dtype=torch.uint8, pin_memory=True)))

    ensure_workers(state["n"], max_e2e_frames)

    def dec_host():
        n_workers, wv = state["n"], state["waves"]
        counts = [0] * n_workers
        errs = []

        def run(k):
            try:
                torch.cuda.set_device(local)
                wk = workers[k]
                for w in wv[k::n_workers]:
                    t0 = time.perf_counter()
                    sc = wk["h"].decode_scan(w["mp3_host"], w["off"])
                    t1 = time.perf_counter()
                    wk["h"].decode_reveal_into(wk["ids"], wk["bits"])
                    t2 = time.perf_counter()
                    wk["h"].decode_run(pcm=wk["pcm"])
                    t3 = time.perf_counter()
                    if state["trace"] is not None:
                        state["trace"].append((k, t0, t1, t2, t3))
                    counts[k] += int(sc["n_frames"].sum())
            except Exception as e:   # surfaced below: a failed worker must fail the bench
                errs.append(e)

        ts = [threading.Thread(target=run, args=(k,)) for k in range(1, n_workers)]
        for t in ts:
            t.start()
        run(0)
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
        return sum(counts)

    if args.e2e_sweep:    # diagnostic: e2e decode throughput over worker counts and wave sizes, with a per-call trace
        for wf in ((50,) if os.environ.get("M3S_TRACE") else (25, 50, 125)):
            state["waves"] = make_waves(wf)
            for nwk in ((2,) if os.environ.get("M3S_TRACE") else (1, 2, 3, 4, 6)):
                state["n"] = nwk
                ensure_workers(nwk, max(w["frames"] for w in state["waves"]))
                dec_host()
                state["trace"] = []
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = dec_host()
                dt = time.perf_counter() - t0
                tr = state["trace"]
                state["trace"] = None
                sc = np.mean([b - a for _, a, b, _, _ in tr]) * 1e3
                rv = np.mean([c - b for _, _, b, c, _ in tr]) * 1e3
                rn = np.mean([d - c for _, _, _, c, d in tr]) * 1e3
                log(f"[sweep] wave {wf:4d} files, {nwk} workers: {dt * 1e3:7.1f} ms/step = {n / dt / 1e6:6.2f} M frames/s; "
                    f"mean call ms: scan {sc:6.1f} reveal {rv:6.1f} run {rn:6.1f}")
        return

    # encode+hide: ONE call over all clips (the per-clip offset scan of the rate loop wants every clip's chain in flight together)
    # (the library walks it in frame windows to bound its intermediates)
    def enc_device():
        r = h.encode(pcm_all, ns_all, 44100, 128, payload_packed=(pay_all, pay_off_all), mp3_out=enc_out_dev)
        check["hide_off"] = int(r["hide_str_offset"].sum())
        check["enc_bytes"] = int(r["out_len"].sum())
        return total_frames

    # host leg of encode+hide: the same 1,000 clips (same call shape as the device-resident leg), each cut to `enc_e2e_frames` frames
    # when the box's RAM cannot pin the whole PCM corpus of every rank (8 x 31.7 GB on a 251 GB host)
    enc_e2e = dict(frames=args.frames)

    def enc_host():
        h.encode(pcm_host_all, [enc_e2e["frames"] * 1152] * args.files, 44100, 128, payload_packed=(pay_all, pay_off_all),
                 mp3_out=enc_out_host)
        return args.files * enc_e2e["frames"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        n = 0
        for _ in range(steps):
            n += fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1) / 1e3, wall], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return n, float(t[0].item()), float(t[1].item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    DEC_K = ("k_walk", "k_fscan", "k_sideinfo", "k_strip", "k_huff", "k_hybrid")
    ENC_K = ("k_enc_analysis", "k_enc_rate", "k_enc_resolve", "k_enc_emit", "k_enc_pack")

    def measure(dev_fn, host_fn, knames, bpf):
        for _ in range(args.warmup):
            dev_fn()
        h.timing_enable(True)
        l0 = h.launches
        clocks = ClockSampler(local)
        n_dev, t_dev, wall = timed(dev_fn, args.steps)
        clk = clocks.stop()
        launches = h.launches - l0
        kt = h.timing()
        h.timing_enable(False)
        assert n_dev == total_frames * args.steps, (n_dev, total_frames)
        host_fn()
        n_e2e, t_e2e, _ = timed(host_fn, args.steps)
        value = world * n_dev / t_dev
        roof = roofline_of(kt, knames, n_dev, bpf, peak, peak_src, args.steps, value / world * bpf / 1e9 / peak)
        return dict(value=value, ms_per_step=1e3 * t_dev / args.steps, e2e_value=world * n_e2e / t_e2e,
                    e2e_ms=1e3 * t_e2e / args.steps, launches=int(launches), clocks=clk, roofline=roof, wall=wall)

    D = measure(dec_device, dec_host, DEC_K, DEC_BYTES_PER_FRAME)
    log(f"[rank {rank}] decode+reveal: {D['value']:.4g} frames/s device-resident, {D['e2e_value']:.4g} e2e")
    E = None
    if not args.no_encode:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        avail = 64 << 30
        try:
            for ln in open("/proc/meminfo"):
                if ln.startswith("MemAvailable:"):
                    avail = int(ln.split()[1]) * 1024
        except OSError:
            pass
        budget = int(0.55 * avail / max(local_world, 1))           # bytes of PCM this rank may pin
        enc_e2e["frames"] = max(1, min(args.frames, budget // (args.files * 4608)))
        fe = enc_e2e["frames"]
        pcm_host_all = torch.empty(args.files * fe * 1152 * 2, dtype=torch.int16, pin_memory=True)
        pcm_host_all.view(args.files, fe * 1152 * 2).copy_(pcm_all.view(args.files, n_samp * 2)[:, : fe * 1152 * 2])
        enc_out_host = torch.empty(enc_cap, dtype=torch.uint8, pin_memory=True)
        log(f"[rank {rank}] encode e2e leg: {args.files} clips x {fe} frames from pinned host memory ({pcm_host_all.numel() * 2 / 1e9:.1f} GB)")
        torch.cuda.synchronize()
        E = measure(enc_device, enc_host, ENC_K, ENC_BYTES_PER_FRAME)
        log(f"[rank {rank}] encode+hide:   {E['value']:.4g} frames/s device-resident, {E['e2e_value']:.4g} e2e")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle port on one host core over a bounded sample of the same kind of corpus
    cb = cpu_baselines(min(args.frames, 6890), 1, 3 if args.frames >= 2000 else 1)
    ncore = os.cpu_count()

    def cpu_obj(key, what):
        f, t = cb[key]
        return {"value": f / t, "unit": "frames/s", "cores": 1, "kind": "port",
                "sample": f"{what} of {f} frames (tone+noise clips of {min(args.frames, 6890)} frames), oracle/mp3stego_oracle.c, "
                          f"1 thread of {ncore} host cores, {t:.1f} s"}

    line = {"metric": "decode+reveal throughput (MP3 frames/s)", "value": D["value"], "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": D["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": corpus_config(args),
            "audio_seconds_per_s": D["value"] * 1152 / 44100.0,
            "e2e": {"value": D["e2e_value"], "unit": "frames/s", "h2d_bytes_per_step": mp3_bytes,
                    "d2h_bytes_per_step": int(total_frames * (1152 * 2 * 2 + 24)), "ms_per_step": D["e2e_ms"]},
            "gpu_launches": D["launches"] + (E["launches"] if E else 0), "clocks": D["clocks"], "roofline": D["roofline"],
            "cpu_baseline": cpu_obj("decode", "decode+reveal @320k"),
            "check": check}
    if E:
        line["encode_hide"] = {
            "metric": "encode+hide throughput (MP3 frames/s)", "value": E["value"], "unit": "frames/s", "dtype": "int32",
            "ms_per_step": E["ms_per_step"], "audio_seconds_per_s": E["value"] * 1152 / 44100.0,
            "e2e": {"value": E["e2e_value"], "unit": "frames/s",
                    "h2d_bytes_per_step": int(args.files * enc_e2e["frames"] * 4608 + len(pay_all)),
                    "d2h_bytes_per_step": int(check.get("enc_bytes", 0) * enc_e2e["frames"] / args.frames), "ms_per_step": E["e2e_ms"],
                    "clips": args.files, "frames_per_clip": enc_e2e["frames"]},
            "gpu_launches": E["launches"], "clocks": E["clocks"], "roofline": E["roofline"],
            "cpu_baseline": cpu_obj("encode", "encode+hide @128k")}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, the compiler) write to fd 1; the contract is ONE JSON line there.  Everything but
    emit() goes to stderr from here on."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files", type=int, default=1000, help="files per GPU")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_FILE, help="frames per file")
    ap.add_argument("--wave", type=int, default=500, help="files per wave of the device-resident decode leg (bounds the workspaces)")
    ap.add_argument("--e2e-wave", type=int, default=50, help="files per wave of the host-buffer (e2e) decode leg")
    ap.add_argument("--no-encode", action="store_true", help="skip the encode+hide half")
    ap.add_argument("--e2e-sweep", action="store_true", help="diagnostic sweep of the decode e2e leg (no JSON line)")
    ap.add_argument("--e2e-workers", type=int, default=3, help="host worker threads (one handle each) of the decode e2e leg")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()
