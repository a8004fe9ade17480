#!/usr/bin/env python3
"""Test-bitstream writer + golden generator for BASELINE configs[3] (development container only).

The reference ENCODER cannot emit short / mixed blocks, MS stereo, scalefactors, scfsi, the bit reservoir, CRC, mono or
most Huffman tables (SURVEY A.E3/A.E7), so this tool writes MPEG-1 Layer III streams that use them, from random integer
spectra, and then decodes them with the UNMODIFIED reference DECODER to produce the golden outputs:

  tests/golden/stream_<name>.mp3        the stream
  tests/golden/ref_stream_<name>.npz    reference decode: n_frames, int16 PCM, integer spectra, table ids, reveal bits

    python tests/golden/make_streams.py

Code books come from the reference's decoder tables (the data source), the bit layout from ISO 11172-3 as the
reference parses it (FrameSideInformation.py:39-137, Frame.py:365-559).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets up the bitarray shim + reference import path, silences tqdm)

import numpy as np  # noqa: E402
from mp3stego.decoder import tables as dt  # noqa: E402

BITRATES = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
SR_CODE = {44100: 0, 48000: 1, 32000: 2}
SFB_LONG = {44100: dt.band_index_table.long_44, 48000: dt.band_index_table.long_48, 32000: dt.band_index_table.long_32}


class Bits:
    def __init__(self):
        self.b = []

    def put(self, v, n):
        for i in range(n - 1, -1, -1):
            self.b.append((int(v) >> i) & 1)

    def __len__(self):
        return len(self.b)

    def tobytes(self):
        bits = self.b + [0] * ((-len(self.b)) % 8)
        return bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))


def huff_pair(w, t, x, y):
    """One big-values pair with table t (1..31, not 4/14)."""
    dim = dt.big_value_max[t]
    lin = dt.big_value_linbit[t]
    ax, ay = abs(x), abs(y)
    cx, cy = min(ax, dim - 1), min(ay, dim - 1)
    flat = dt.big_value_table[t]
    i = 2 * dim * cx + 2 * cy
    code, ln = flat[i], flat[i + 1]
    w.put(code >> (32 - ln), ln)
    for a, c, v in ((ax, cx, x), (ay, cy, y)):
        if lin and c == dim - 1:
            w.put(a - c, lin)
        if c > 0:
            w.put(1 if v < 0 else 0, 1)


def huff_quad(w, sel, q):
    a = [abs(v) for v in q]
    if sel == 1:
        w.put(((1 - a[0]) << 3) | ((1 - a[1]) << 2) | ((1 - a[2]) << 1) | (1 - a[3]), 4)
    else:
        e = 8 * a[0] + 4 * a[1] + 2 * a[2] + a[3]
        ln = dt.quad_table_1.h_len[e]
        w.put(dt.quad_table_1.h_cod[e] >> (32 - ln), ln)
    for v in q:
        if v:
            w.put(1 if v < 0 else 0, 1)


def table_limit(t):
    if t in (0, 4, 14):
        return 0
    return dt.big_value_max[t] - 1 + ((1 << dt.big_value_linbit[t]) - 1 if dt.big_value_linbit[t] else 0)


def make_granule(rng, sr, gr, ch, opts, scfsi):
    """Random side info + spectra for one granule-channel; returns (side dict, main-data Bits)."""
    long_win = SFB_LONG[sr]
    si = dict(scalefac_compress=int(rng.integers(0, 16)) if opts.get("scalefac", True) else 0,
              global_gain=int(rng.integers(opts.get("gain_lo", 135), opts.get("gain_hi", 168))),
              preflag=int(rng.integers(0, 2)) if opts.get("scalefac", True) else 0,
              scalefac_scale=int(rng.integers(0, 2)) if opts.get("scalefac", True) else 0,
              count1table_select=int(rng.integers(0, 2)))
    ws = 1 if opts.get("switching") and rng.random() < 0.7 else 0
    si["window_switching"] = ws
    tabs_pool = opts.get("tables", [t for t in range(1, 32) if t not in (4, 14)])
    pick = lambda: int(rng.choice(tabs_pool)) if rng.random() > 0.08 else int(rng.choice([0, 4, 14]))  # noqa: E731
    if ws:
        bt = int(rng.choice(opts.get("block_types", [1, 2, 3])))
        si["block_type"] = bt
        si["mixed_block_flag"] = int(rng.integers(0, 2)) if bt == 2 and opts.get("mixed", True) else 0
        si["table_select"] = [pick(), pick()]
        si["subblock_gain"] = [int(rng.integers(0, 4)) for _ in range(3)]
        r0 = 8 if bt == 2 else 7
        r1 = 20 - r0
        if bt == 2:
            bounds = (36, 576)
        else:
            bounds = (long_win[r0 + 1], long_win[min(r0 + r1 + 2, 22)])
    else:
        si["block_type"] = 0
        si["mixed_block_flag"] = 0
        si["table_select"] = [pick(), pick(), pick()]
        r0 = int(rng.integers(0, 16))
        r1 = int(rng.integers(0, 8))
        while r0 + r1 + 2 > 22:
            r0 = int(rng.integers(0, 16))
            r1 = int(rng.integers(0, 8))
        si["region0_count"], si["region1_count"] = r0, r1
        bounds = (long_win[r0 + 1], long_win[r0 + r1 + 2])
    w = Bits()
    # ---- part2: scalefactors (Frame.py:365-441)
    sl1, sl2 = dt.slen[si["scalefac_compress"]]
    rnd = lambda n: int(rng.integers(0, 1 << n)) if n else 0  # noqa: E731
    if ws and si["block_type"] == 2:
        if si["mixed_block_flag"]:
            for _ in range(8):
                w.put(rnd(sl1), sl1)
            for _ in range(3, 6):
                for _w in range(3):
                    w.put(rnd(sl1), sl1)
        else:
            for _ in range(6):
                for _w in range(3):
                    w.put(rnd(sl1), sl1)
        for _ in range(6, 12):
            for _w in range(3):
                w.put(rnd(sl2), sl2)
    elif gr == 0:
        for _ in range(11):
            w.put(rnd(sl1), sl1)
        for _ in range(10):
            w.put(rnd(sl2), sl2)
    else:
        for i, (lo, hi) in enumerate(((0, 6), (6, 11), (11, 16), (16, 21))):
            if not scfsi[ch][i]:
                for _ in range(lo, hi):
                    w.put(rnd(sl1 if i < 2 else sl2), sl1 if i < 2 else sl2)
    # ---- part3: big values + count1 (Frame.py:443-559)
    max_bv = opts.get("max_bv", 200)
    bv = int(rng.integers(0, max_bv + 1))
    si["big_values"] = bv
    amp = opts.get("amp", 12)
    for p in range(bv):
        s = 2 * p
        t = si["table_select"][0] if s < bounds[0] else (si["table_select"][1] if s < bounds[1] else
                                                          (si["table_select"][2] if len(si["table_select"]) > 2 else opts["_stale_t2"]))
        lim = table_limit(t)
        if lim == 0:
            continue   # table 0 / 4 / 14: zeros, no bits
        hi = min(lim, amp if s < 60 else max(2, amp // 3))
        if dt.big_value_linbit[t] and rng.random() < 0.05:
            hi = min(lim, 15 + (1 << min(dt.big_value_linbit[t], 6)))
        x = int(rng.integers(-hi, hi + 1))
        y = int(rng.integers(-hi, hi + 1))
        huff_pair(w, t, x, y)
    nq_max = max(0, (572 - 2 * bv) // 4)            # the reference stops at sample + 4 < 576 (A.D5)
    nq = int(rng.integers(0, min(nq_max, opts.get("max_quads", 60)) + 1))
    for _ in range(nq):
        q = [int(v) for v in rng.integers(-1, 2, size=4)]
        huff_quad(w, si["count1table_select"], q)
    if opts.get("stuff_ones") and si["count1table_select"] == 0 and rng.random() < 0.5:
        for _ in range(int(rng.integers(1, 9))):
            w.put(1, 1)                              # table A '1' = zero quad: stuffing as the reference encoder writes it
    si["part2_3_length"] = len(w)
    assert len(w) < 4096
    return si, w


def side_info_bits(frame, nch):
    w = Bits()
    w.put(frame["main_data_begin"], 9)
    w.put(0, 3 if nch == 2 else 5)
    for ch in range(nch):
        for b in range(4):
            w.put(frame["scfsi"][ch][b], 1)
    for gr in range(2):
        for ch in range(nch):
            s = frame["gr"][gr][ch]
            w.put(s["part2_3_length"], 12)
            w.put(s["big_values"], 9)
            w.put(s["global_gain"], 8)
            w.put(s["scalefac_compress"], 4)
            w.put(s["window_switching"], 1)
            if s["window_switching"]:
                w.put(s["block_type"], 2)
                w.put(s["mixed_block_flag"], 1)
                for t in s["table_select"]:
                    w.put(t, 5)
                for g in s["subblock_gain"]:
                    w.put(g, 3)
            else:
                for t in s["table_select"]:
                    w.put(t, 5)
                w.put(s["region0_count"], 4)
                w.put(s["region1_count"], 3)
            w.put(s["preflag"], 1)
            w.put(s["scalefac_scale"], 1)
            w.put(s["count1table_select"], 1)
    assert len(w) == (256 if nch == 2 else 136), len(w)
    return w.tobytes()


def make_stream(seed, n_frames, sr=44100, bitrate=128, mode=0, mode_ext=0, crc=False, opts=None, vbr=None, reservoir=False):
    """mode: 0 stereo, 1 joint stereo, 3 mono.  Returns the stream bytes."""
    opts = dict(opts or {})
    rng = np.random.default_rng(seed)
    nch = 1 if mode == 3 else 2
    hdr_len = 4 + (2 if crc else 0) + (32 if nch == 2 else 17)
    frames = []
    stale_t2 = [[0, 0], [0, 0]]   # table_select[gr][ch][2] as the reference's persistent side-info object holds it (A.D3)
    R = 0                         # bytes of reservoir in front of the next frame (its main_data_begin)
    payload = b""
    for f in range(n_frames):
        br = int(rng.choice(vbr)) if vbr else bitrate
        pad = int(rng.integers(0, 2)) if opts.get("padding") else 0
        size = 144000 * br // sr + pad
        cap = size - hdr_len
        scfsi = [[int(rng.integers(0, 2)) if opts.get("scfsi", True) else 0 for _ in range(4)] for _ in range(nch)]
        o = dict(opts)
        while True:   # shrink the granules until the frame's main data fits the reservoir + its own payload area
            grs = [[None] * nch for _ in range(2)]
            md = Bits()
            st = [row[:] for row in stale_t2]
            for gr in range(2):
                for ch in range(nch):
                    o["_stale_t2"] = st[gr][ch]
                    si, w = make_granule(rng, sr, gr, ch, o, scfsi)
                    if not si["window_switching"]:
                        st[gr][ch] = si["table_select"][2]
                    grs[gr][ch] = si
                    md.b += w.b
            mdb = md.tobytes()
            if len(mdb) <= R + cap:
                break
            o["max_bv"] = max(4, int(o.get("max_bv", 200) * 0.8))
            o["max_quads"] = max(2, int(o.get("max_quads", 60) * 0.8))
        stale_t2 = st
        M = len(mdb)
        if reservoir:
            extra = max(0, R + cap - 511 - M)
            if rng.random() < 0.25:   # sometimes drain the reservoir completely
                extra = R + cap - M
        else:
            extra = R + cap - M
        frames.append(dict(bitrate=br, pad=pad, size=size, scfsi=scfsi, gr=grs, cap=cap, main_data_begin=R))
        payload += mdb + b"\x00" * extra   # ancillary / padding bytes after the last granule: ignored by the decoder
        R = R + cap - M - extra
        assert 0 <= R <= 511
    out = b""
    pos = 0
    for fr in frames:
        h = Bits()
        h.put(0x7FF, 11)
        h.put(3, 2)
        h.put(1, 2)
        h.put(0 if crc else 1, 1)
        h.put(BITRATES.index(fr["bitrate"]), 4)
        h.put(SR_CODE[sr], 2)
        h.put(fr["pad"], 1)
        h.put(0, 1)
        h.put(mode, 2)
        h.put(mode_ext, 2)
        h.put(0, 1)
        h.put(1, 1)
        h.put(0, 2)
        out += h.tobytes() + (b"\xAB\xCD" if crc else b"") + side_info_bits(fr, nch)
        chunk = payload[pos:pos + fr["cap"]]
        out += chunk + b"\x00" * (fr["cap"] - len(chunk))
        pos += fr["cap"]
    return out


STREAMS = {
    # name: kwargs
    "long_alltables": dict(seed=1, n_frames=8, bitrate=192, opts=dict(max_bv=140, stuff_ones=True)),
    "reservoir": dict(seed=2, n_frames=12, bitrate=128, reservoir=True, opts=dict(max_bv=110, max_quads=40)),
    "short_mixed": dict(seed=3, n_frames=10, bitrate=192, opts=dict(switching=True, max_bv=130)),
    "ms_stereo": dict(seed=4, n_frames=8, bitrate=160, mode=1, mode_ext=3, reservoir=True, opts=dict(switching=True, max_bv=110)),
    "is_only_bit": dict(seed=5, n_frames=4, bitrate=160, mode=1, mode_ext=1, opts=dict(max_bv=110)),
    "mono_crc_48k": dict(seed=6, n_frames=8, sr=48000, bitrate=96, mode=3, crc=True, reservoir=True, opts=dict(switching=True, max_bv=120)),
    "vbr_32k_pad": dict(seed=7, n_frames=10, sr=32000, vbr=[96, 128, 160, 320], reservoir=True, opts=dict(padding=True, max_bv=90, switching=True)),
    "loud_wrap": dict(seed=8, n_frames=4, bitrate=320, opts=dict(max_bv=200, amp=15, gain_lo=175, gain_hi=190, scalefac=False)),
}


def fuzz_config(seed):
    """A random combination of everything the writer can vary (BASELINE configs[3] as a corpus rather than eight hand-picked
    streams): sample rate, CBR / VBR, padding, mono / stereo / joint stereo with either mode-extension bit, CRC, bit reservoir,
    window switching with short / mixed / start / stop blocks, scalefactors + scfsi, loud spectra."""
    rng = np.random.default_rng(1000 + seed)
    sr = int(rng.choice([44100, 48000, 32000]))
    mode = 3 if seed % 5 == 2 else int(rng.choice([0, 0, 1, 1]))   # every fifth stream is mono (17-byte side info)
    kw = dict(seed=seed, n_frames=int(rng.integers(16, 29)), sr=sr, mode=mode, mode_ext=int(rng.integers(0, 4)) if mode == 1 else 0,
              crc=bool(rng.integers(0, 2)), reservoir=bool(rng.integers(0, 2)))
    rates = [64, 96, 128, 160, 192, 256, 320] if mode != 3 else [48, 64, 96, 128, 160]
    if rng.random() < 0.3:
        kw["vbr"] = [int(v) for v in rng.choice(rates, size=3, replace=False)]
    else:
        kw["bitrate"] = int(rng.choice(rates))
    lo = min(kw.get("vbr", [kw.get("bitrate", 128)]))
    opts = dict(switching=bool(rng.random() < 0.7), padding=bool(rng.integers(0, 2)), scfsi=bool(rng.random() < 0.8),
                stuff_ones=bool(rng.integers(0, 2)), max_bv=int(min(200, 40 + lo * (2 if mode == 3 else 1) // 2)),
                max_quads=int(rng.integers(10, 61)))
    if rng.random() < 0.2:
        opts.update(amp=15, gain_lo=170, gain_hi=186)
    kw["opts"] = opts
    return kw


N_FUZZ = 16


def main_fuzz():
    """tests/golden/fuzz_NN.mp3 + ref_fuzz.json: digests of what the UNMODIFIED reference decoder makes of each stream."""
    import hashlib
    import json
    digests = {}
    for k in range(N_FUZZ):
        kw = fuzz_config(k)
        data = make_stream(**kw)
        d = MG.ref_decode_taps(data)
        name = "fuzz_%02d" % k
        open(os.path.join(HERE, name + ".mp3"), "wb").write(data)
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
        digests[name] = dict(n_frames=int(d["n_frames"]), bitrate=int(d["bitrate"]), sampling_rate=int(d["sampling_rate"]),
                             channels=int(d["pcm16"].shape[1]) if d["pcm16"].ndim == 2 else 1,
                             pcm16_sha256=sha(d["pcm16"].astype(np.int16)), spectra_sha256=sha(d["spectra"].astype(np.int16)),
                             tables_sha256=sha(np.asarray(d["tables"], np.uint8)), bits=d["bits"],
                             config={k2: (v if not isinstance(v, dict) else v) for k2, v in kw.items()})
        print("%-8s bytes %6d frames %3d sr %5d mode %d |pcm|max %.3f bits %d" % (
            name, len(data), d["n_frames"], kw["sr"], kw["mode"], np.abs(d["pcm"]).max() if d["pcm"].size else 0, len(d["bits"])))
    json.dump(digests, open(os.path.join(HERE, "ref_fuzz.json"), "w"), indent=1, sort_keys=True)


def frame_table(data, start=0):
    """(position, size, main_data_begin) of every frame found by the reference's own walk rule (offset += frame_size)."""
    out, pos = [], start
    while len(data) > pos + 4 and data[pos] == 0xFF and data[pos + 1] >= 0xE0:
        b1, b2, b3 = data[pos + 1], data[pos + 2], data[pos + 3]
        br = BITRATES[b2 >> 4] if b2 >> 4 else 320
        sr = {0: 44100, 1: 48000, 2: 32000}[(b2 >> 2) & 3]
        size = 144000 * br // sr + ((b2 >> 1) & 1)
        si = pos + 4 + (0 if b1 & 1 else 2)
        mdb = (data[si] << 1) | (data[si + 1] >> 7)
        out.append((pos, size, mdb))
        pos += size
    return out


def cut_at(data, k):
    """The stream from its k-th frame on (what a file cut out of a longer stream looks like)."""
    return data[frame_table(data)[k][0]:]


def id3v2(n):
    """A syntactically valid ID3v2.3 tag of 10 + n bytes (decoder.py:29-33 honours the synchsafe size)."""
    body = bytes((37 * i + 11) & 0xFF for i in range(n))
    return b"ID3\x03\x00\x00" + bytes([(n >> 21) & 0x7F, (n >> 14) & 0x7F, (n >> 7) & 0x7F, n & 0x7F]) + body


def edge_cases():
    """Streams whose first frames point into a bit reservoir that is not there (files cut out of a stream, with and without an
    ID3v2 tag in front), and streams whose header + side-info length changes under a live reservoir (mono <-> stereo, CRC on /
    off): the cases in which the reference's main-data assembly (Frame.py:318-363, A.D9) is NOT 'the bytes of the earlier payloads'."""
    res = make_stream(**STREAMS["reservoir"])
    ft = frame_table(res)
    cases = {}
    ks = [k for k in range(1, len(ft)) if ft[k][2] > 0]
    for k in ks[:6]:
        cases["cut%02d" % k] = cut_at(res, k)
    k = ks[1]
    cases["cut%02d_id3_small" % k] = id3v2(40) + cut_at(res, k)      # the tag is shorter than the reach: Python negative-index slices
    cases["cut%02d_id3_large" % k] = id3v2(900) + cut_at(res, k)     # reads tag bytes as main data
    vbr = make_stream(**STREAMS["vbr_32k_pad"])
    kv = [k for k, f in enumerate(frame_table(vbr)) if f[2] > 0 and k > 0]
    for k in kv[:2]:
        cases["vbrcut%02d" % k] = cut_at(vbr, k)
    mono = make_stream(seed=21, n_frames=6, sr=44100, bitrate=96, mode=3, reservoir=True, opts=dict(switching=True, max_bv=100))
    mono_crc = make_stream(seed=22, n_frames=6, sr=44100, bitrate=96, mode=3, crc=True, reservoir=True, opts=dict(switching=True, max_bv=100))
    st = make_stream(seed=23, n_frames=9, sr=44100, bitrate=160, mode=1, mode_ext=2, reservoir=True,
                     opts=dict(switching=True, max_bv=110, block_types=[2], mixed=True))
    st_crc = make_stream(seed=24, n_frames=9, sr=44100, bitrate=160, crc=True, reservoir=True, opts=dict(switching=True, max_bv=110))
    def first_live(d):
        return [k for k, f in enumerate(frame_table(d)) if f[2] > 0 and k > 0][0]
    # (a channel-count change inside one file is outside the reference's domain: MP3Parser.parse_file raises ValueError when it
    #  stacks PCM rows of different widths, MP3_Parser.py:83 -- so only the CRC flag can change C under a live reservoir)
    cases["crc_on"] = make_stream(**STREAMS["reservoir"]) + cut_at(st_crc, first_live(st_crc))   # C 36 -> 38
    cases["crc_off_mono"] = mono_crc + cut_at(mono, first_live(mono))         # C 23 -> 21
    long_crc = make_stream(seed=25, n_frames=14, sr=44100, bitrate=128, crc=True, reservoir=True, opts=dict(max_bv=100))
    cases["late_switch"] = long_crc + cut_at(st, first_live(st))             # C 38 -> 36 after frame 9: dynamic assembly slots
    return cases


def main_edge():
    import hashlib
    import json
    digests = {}
    for name, data in edge_cases().items():
        d = MG.ref_decode_taps(data)
        fn = "edge_" + name
        open(os.path.join(HERE, fn + ".mp3"), "wb").write(data)
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
        ch = int(d["pcm16"].shape[1]) if d["pcm16"].ndim == 2 else 1
        digests[fn] = dict(n_frames=int(d["n_frames"]), bitrate=int(d["bitrate"]), sampling_rate=int(d["sampling_rate"]), channels=ch,
                           pcm16_sha256=sha(d["pcm16"].astype(np.int16)), spectra_sha256=sha(d["spectra"].astype(np.int16)),
                           tables_shape=list(d["tables"].shape), tables_sha256=sha(np.asarray(d["tables"], np.uint8)), bits=d["bits"])
        print("%-28s bytes %6d frames %3d ch %d |pcm|max %.3g bits %d" % (fn, len(data), d["n_frames"], ch,
                                                                         np.abs(d["pcm"]).max() if d["pcm"].size else 0, len(d["bits"])))
    json.dump(digests, open(os.path.join(HERE, "ref_edge.json"), "w"), indent=1, sort_keys=True)


def main():
    if "--fuzz" in sys.argv:
        return main_fuzz()
    if "--edge" in sys.argv:
        return main_edge()
    for name, kw in STREAMS.items():
        data = make_stream(**kw)
        d = MG.ref_decode_taps(data)
        path = os.path.join(HERE, "stream_%s.mp3" % name)
        open(path, "wb").write(data)
        np.savez_compressed(os.path.join(HERE, "ref_stream_%s.npz" % name), n_frames=d["n_frames"], pcm16=d["pcm16"],
                            spectra=d["spectra"].astype(np.int16), tables=d["tables"], bits=np.array(d["bits"]),
                            bitrate=d["bitrate"], sampling_rate=d["sampling_rate"],
                            pcm_absmax=float(np.abs(d["pcm"]).max()) if d["pcm"].size else 0.0)
        print("%-16s bytes %6d frames %3d |pcm|max %.3f bits %d" % (name, len(data), d["n_frames"],
                                                                    np.abs(d["pcm"]).max() if d["pcm"].size else 0, len(d["bits"])))


if __name__ == "__main__":
    main()
