"""CPU: the algorithm behind the parallel rate loop (k_enc_probe / k_enc_resolve), modelled in C on top of the
oracle and held to the oracle's own sequential iteration loop granule by granule: table choices, part2_3_length, big_values,
count1, step size, address1..3 after the granule and hide_str_offset."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import synth_wav

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    d = tmp_path_factory.mktemp("rate_model")
    exe = str(d / "rate_variants_model")
    subprocess.check_call(["gcc", "-O2", "-w", "-o", exe, os.path.join(ROOT, "tests", "model", "rate_variants_model.c"), "-lm"])
    return d, exe


def _clicks(seed, n, amp):
    rng = np.random.default_rng(seed)
    y = np.zeros((n, 2), np.int16)
    idx = rng.integers(0, n, size=n // 200)
    y[idx, 0] = amp
    y[idx[::2], 1] = -amp
    return y


def _run(model, name, pcm, bitrate, bits, payload_len=None):
    d, exe = model
    raw, pay = str(d / f"{name}.raw"), str(d / f"{name}.txt")
    np.ascontiguousarray(pcm, dtype=np.int16).tofile(raw)
    open(pay, "w").write(bits)
    args = [exe, raw, str(pcm.shape[0] // 1152), str(bitrate), pay] + ([] if payload_len is None else [str(payload_len)])
    out = subprocess.run(args, capture_output=True, text=True, check=True).stdout
    m = re.search(r"granules (\d+) silent (\d+) slow (\d+) bad (\d+)", out)
    assert m, out
    return dict(zip(("granules", "silent", "slow", "bad"), map(int, m.groups())))


def test_variants_and_resolve_reproduce_the_chain(model):
    rng = np.random.default_rng(5)
    n = 12 * 1152
    t = np.arange(n) / 44100.0
    clips = dict(
        tone=synth_wav(3, 40),
        silence=np.zeros((n, 2), np.int16),
        fade=(np.linspace(0, 1, n)[:, None] ** 6 * 300 * np.sin(2 * np.pi * 700 * t)[:, None] * np.ones((1, 2))).astype(np.int16),
        loud=rng.integers(-32768, 32767, size=(n, 2)).astype(np.int16),
        square=(np.sign(np.sin(2 * np.pi * 90 * t)) * 32000).astype(np.int16)[:, None] * np.ones((1, 2), np.int16),
        tiny=rng.integers(-2, 3, size=(n, 2)).astype(np.int16),
        hush=np.random.default_rng(6).integers(-6, 7, size=(n, 2)).astype(np.int16),
        clicks=_clicks(48, n, 48),
        clicks1=_clicks(1, n, 1),
    )
    bits = "1100101" * 400
    slow_seen = 0
    for name, pcm in clips.items():
        for br in (32, 128, 320):
            r = _run(model, f"{name}{br}", pcm, br, bits)
            assert r["bad"] == 0, (name, br, r)
            slow_seen += r["slow"]
    assert slow_seen > 0          # the quiet clips do exercise the stale-address path (A.E6)
    # payloads that end inside the clip (variants with 2, 1 and 0 bits left), and plain encodes
    for name, plen in (("tone", 37), ("tone", 8), ("tone", 0), ("clicks", 50), ("hush", 3)):
        r = _run(model, f"{name}_p{plen}", clips[name], 128, bits, payload_len=plen)
        assert r["bad"] == 0, (name, plen, r)


def test_probe_bounds_never_decide_wrongly(tmp_path):
    """Model of k_enc_probe's "light" probes (csrc/m3s_encode.cu: probe_row): certified lower / upper bounds on a probe's bit count
    decide about 3 of the 7 binary-search probes of a granule at 128 kbps without running them -- and never differently from the
    true count, with no payload and with random payloads at random offsets (every swap variant meets every probe)."""
    exe = str(tmp_path / "probe_bounds_model")
    subprocess.check_call(["gcc", "-O2", "-w", "-o", exe, os.path.join(ROOT, "tests", "model", "probe_bounds_model.c"), "-lm"])
    rng = np.random.default_rng(9)
    n = 12 * 1152
    clips = dict(tone=synth_wav(4, 60), loud=rng.integers(-32768, 32767, size=(n, 2)).astype(np.int16),
                 hush=rng.integers(-6, 7, size=(n, 2)).astype(np.int16))
    decided = {}
    for name, pcm in clips.items():
        raw = str(tmp_path / f"{name}.raw")
        np.ascontiguousarray(pcm, dtype=np.int16).tofile(raw)
        for br in (64, 128, 320):
            for seed in ((), ("1",), ("2",), ("3",)):      # no payload, then three random payloads
                out = subprocess.run([exe, raw, str(pcm.shape[0] // 1152), str(br), *seed], capture_output=True, text=True, check=True).stdout
                m = re.search(r"probes\(bin search\) (\d+)  decided by LB (\d+)  by UB (\d+)  violations (\d+)", out)
                assert m, out
                probes, lb, ub, viol = map(int, m.groups())
                assert viol == 0, (name, br, seed, out)
                if not seed:
                    decided[(name, br)] = (lb + ub) / max(probes, 1)
    assert decided[("tone", 128)] > 0.35      # ~3 of 7 probes on the benchmark's kind of clip
