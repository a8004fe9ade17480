"""Batch entry points over the C ABI: what the facade's five methods do (steganography.py:80-182), for many files per
call and without leaving the GPU between the decode and the encode half of the composites.

  reveal_batch   reveal_massage for N files: frame walk + side-info scan + reveal bits only (D0 + D4, ~36 B/frame of traffic);
                 no Huffman decode, no synthesis, no temp WAV
  hide_batch     hide_message for N files: decode (float64 instantiation, so the int16 PCM equals the reference's WAV sample
                 for sample) -> encode + hide at each file's own bitrate; the PCM stays in HBM
  clear_batch    clear_file: the same with no payload
The semantics per file are the facade's: message framing '<len>#<message>' as utf-8 bits (steganography.py:10-24,88-90),
bitrate of the last frame (MP3_Parser.py:93-97), too_long = hide_str_offset < len(bits) - 1 (encoder.py:49-51),
ID3v2 skip (decoder.py:29-33).  torch is used for the device buffers only."""
from typing import List, Sequence, Tuple

import numpy as np

from mp3stego_b200 import _lib
from mp3stego_b200.decoder import id3_offset, parse_reveal
from mp3stego_b200.steganography import str_to_binary_str


def _concat(blobs: Sequence[bytes]):
    data = np.frombuffer(b"".join(blobs), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(b) for b in blobs])]).astype(np.int64)
    audio = [id3_offset(np.frombuffer(b[:10], np.uint8)) if len(b) >= 10 else 0 for b in blobs]
    return data, off, audio


def reveal_batch(handle: "_lib.Handle", blobs: Sequence[bytes]) -> List[str]:
    """Hidden strings of N MP3 files ('' where there is none), from the side info alone."""
    if not blobs:
        return []
    data, off, audio = _concat(blobs)
    sc = handle.decode_scan(data, off, audio)
    for i in range(len(blobs)):   # the reference facade exits / raises on these; a batch must not turn them into ''
        _lib.raise_for_status(int(sc["status"][i]), f"file {i}")
        if sc["status"][i] & _lib.M3S_FILE_NO_SYNC:
            raise ValueError(f"file {i}: no MPEG sync word at the audio start (MP3Parser is not valid, MP3_Parser.py:36-44)")
    _, bits = handle.decode_reveal()
    return [parse_reveal(b) for b in bits]


def _transcode(handle, blobs, payloads) -> Tuple[List[bytes], List[int], List[int]]:
    import torch
    data, off, audio = _concat(blobs)
    dev = torch.device("cuda", handle.device)
    L = _lib.load()
    fb = np.zeros(1, np.int64)
    au, au_p = _lib._i64(audio)
    offc = np.ascontiguousarray(off, np.int64)
    elems = int(L.m3s_decode_bound(_lib._ptr(data), offc.ctypes.data_as(_lib._c_i64p), au_p, len(blobs), fb.ctypes.data_as(_lib._c_i64p)))
    d_data = torch.from_numpy(data.copy()).to(dev)
    pcm = torch.empty(max(elems, 2) + 2, dtype=torch.int16, device=dev)
    try:     # one self-pipelining call: scan of wave k+1 under the float64 synthesis of wave k, nothing leaves the device
        sc = handle.decode(d_data, off, audio, pcm=pcm, frames_bound=int(fb[0]) + 1, exact=True)
    except _lib.M3SError as e:
        if "hold" not in str(e):
            raise
        s0 = handle.decode_scan(d_data, off, audio)   # VBR input: size from an exact scan
        pcm = torch.empty(int((s0["pcm_rows"] * np.maximum(s0["channels"], 1)).sum()) + 2, dtype=torch.int16, device=dev)
        sc = handle.decode(d_data, off, audio, pcm=pcm, frames_bound=int(s0["n_frames"].sum()) + 1, exact=True)
    for i in range(len(blobs)):
        _lib.raise_for_status(int(sc["status"][i]), f"file {i}")
        if sc["status"][i] & _lib.M3S_FILE_NO_SYNC or sc["n_frames"][i] == 0:
            raise ValueError(f"file {i}: not an MPEG-1 Layer III stream the reference can decode")
        if sc["channels"][i] != 2:
            raise IndexError(f"file {i}: the reference encoder only handles stereo input (WAV_Reader / MP3_Encoder.py:611-614)")
    rows = sc["pcm_rows"].astype(np.int64)
    pcm_off = sc["pcm_off"]
    out: List[bytes] = [b""] * len(blobs)
    hoff = [0] * len(blobs)
    # one encode call per (sample rate, bitrate) group: the encoder takes one rate pair per batch
    keys = sorted({(int(sc["sample_rate"][i]), int(sc["bitrate"][i]) // 1000) for i in range(len(blobs))})
    for sr, kbps in keys:
        idx = [i for i in range(len(blobs)) if (int(sc["sample_rate"][i]), int(sc["bitrate"][i]) // 1000) == (sr, kbps)]
        res = handle.encode(pcm, [int(rows[i]) for i in idx], sr, kbps, payloads=[payloads[i] for i in idx],
                            pcm_off=[int(pcm_off[i]) for i in idx], compact=True)
        mp3 = res["mp3"].cpu().numpy()
        for n, i in enumerate(idx):
            o = int(res["mp3_off"][n])
            out[i] = bytes(mp3[o:o + int(res["out_len"][n])])
            hoff[i] = int(res["hide_str_offset"][n])
    return out, hoff, [int(b) // 1000 for b in sc["bitrate"]]


def hide_batch(handle: "_lib.Handle", blobs: Sequence[bytes], messages: Sequence[str]) -> Tuple[List[bytes], List[bool]]:
    """hide_message for N (mp3 bytes, message) pairs -> (new mp3 bytes, too_long flags)."""
    if len(blobs) != len(messages):
        raise ValueError("one message per file")
    if not blobs:
        return [], []
    bits = [str_to_binary_str(str(len(m)) + "#" + m) for m in messages]
    out, hoff, _ = _transcode(handle, blobs, bits)
    return out, [hoff[i] < len(bits[i]) - 1 for i in range(len(blobs))]


def clear_batch(handle: "_lib.Handle", blobs: Sequence[bytes]) -> List[bytes]:
    """clear_file for N files: decode and re-encode at the file's own bitrate with no payload."""
    if not blobs:
        return []
    return _transcode(handle, blobs, [""] * len(blobs))[0]
