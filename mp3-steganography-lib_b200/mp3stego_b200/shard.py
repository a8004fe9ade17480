"""Multi-GPU sharding of the codec hot path: one process per GPU, files are independent units, so a
corpus is partitioned BY FILE with no data-path collective; the only collectives are the off-path gather of small
per-file results and the max-over-ranks of the step time.  Works with any torch.distributed backend (nccl on the GPU
box, gloo in the CPU tests); every function is also usable without torch.distributed (world = 1).

The reference has no counterpart (it is single-process); the units follow its own: a file is what one
MP3Parser (MP3_Parser.py:21-85) or MP3Encoder (MP3_Encoder.py:596-650) instance consumes.
"""
from typing import Dict, List, Sequence

import numpy as np


def partition_files(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first partition of file indices over `world` ranks by byte size (every rank computes the same
    answer from the same sizes: no communication).  Returns, per rank, its file indices in ascending order."""
    if world <= 0:
        raise ValueError("world must be positive")
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world, np.int64)
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))          # ties -> lowest rank: deterministic
        shards[r].append(int(i))
        load[r] += int(sizes[i])
    return [sorted(s) for s in shards]


def frame_ranges(n_frames: int, parts: int) -> List[Dict[str, int]]:
    """Split one long file into `parts` frame ranges for range-sharded DECODE.  Every range but the first re-decodes one
    warm-up frame (two granules: overlap-add tail + the 15 earlier V vectors of the 1024-sample synthesis fifo,
    Frame.py:81-92,150-153) whose PCM is discarded, which reproduces the full decode exactly (measured
    with the reference itself).  The frame walk and side-info scan still run over the whole file (no resync in the
    reference), which is what m3s_decode_scan does.  decode_frame_range() below is the driver that uses it."""
    parts = max(1, min(parts, max(n_frames, 1)))
    edges = [n_frames * k // parts for k in range(parts + 1)]
    return [dict(first=edges[k], count=edges[k + 1] - edges[k], warm=1 if edges[k] > 0 else 0)
            for k in range(parts) if edges[k + 1] > edges[k]]


def plan_frame_shard(n_frames: int, status: int, rank: int, world: int) -> Dict[str, int]:
    """The frame range rank `rank` of `world` decodes of ONE long file: dict(first, count, warm, compact_from).  `warm` = 1 when a
    warm-up frame is decoded in front of the range (PCM discarded); `compact_from` = the first frame whose main data has to be
    compacted for the range: 9 frames in front of the warm-up frame cover its bit reservoir (511 bytes over the smallest legal
    payload, Frame.py:306-309), and a file whose granules inherit scalefactors from earlier frames (M3S_FILE_STATE_CARRY) needs
    its whole prefix.  Informational: m3s_decode_run_range derives the same numbers itself.  Every rank computes every rank's
    plan from the same scan result: no communication."""
    from . import _lib
    rs = frame_ranges(n_frames, world)
    if rank >= len(rs):
        return dict(first=n_frames, count=0, warm=0, compact_from=n_frames)
    r = rs[rank]
    lo = max(r["first"] - r["warm"], 0)
    return dict(first=r["first"], count=r["count"], warm=r["warm"],
                compact_from=0 if status & _lib.M3S_FILE_STATE_CARRY else max(lo - 9, 0))


_H0 = frozenset((3, 6, 8, 11, 12, 15, 17, 19, 21, 23, 24, 26, 28, 30))   # tables that reveal a '0' (decoder/util.py:3)
_REVEAL_CHAR = np.array([ord("0") if t in _H0 else ord("1") for t in range(256)], np.uint8)


def decode_frame_range(handle, blob, rank: int, world: int, audio_start: int = 0, exact: bool = False, pcm=None):
    """Range-sharded decode+reveal of one long file: rank `rank` of `world` returns the PCM rows and the reveal
    bits of ITS frame range only; the ranges of all ranks, concatenated in rank order, equal the whole-file decode bit for bit.

    The frame walk and the side-info scan (D0 + D4: no Huffman decode, ~36 bytes read per frame) run over the WHOLE file on
    every rank, because the reference never resynchronises: frame positions, bit-reservoir cursors, the reveal bits and the
    carried table ids of window-switched granules (A.D3) all come from that scan.  The expensive part (D1-D3) runs on the
    range plus one warm-up frame, cut on the device from the bytes already uploaded (m3s_decode_run_range).
    `blob`: bytes, a uint8 numpy array or a uint8 torch tensor (host or cuda); `pcm`: optional output buffer (same side)."""
    data = np.frombuffer(blob, np.uint8) if isinstance(blob, (bytes, bytearray)) else blob
    nbytes = int(data.numel() if hasattr(data, "numel") else data.size)
    sc = handle.decode_scan(data, [0, nbytes], [audio_start])
    n_frames, status, ch = int(sc["n_frames"][0]), int(sc["status"][0]), max(int(sc["channels"][0]), 1)
    meta = dict(n_frames=n_frames, sample_rate=int(sc["sample_rate"][0]), channels=ch, bitrate=int(sc["bitrate"][0]), status=status)
    plan = plan_frame_shard(n_frames, status, rank, world)
    first, count = plan["first"], plan["count"]
    if count == 0:
        return dict(meta, first=first, count=0, pcm=np.zeros((0, ch), np.int16), bits="")
    ids, _ = handle.decode_reveal()
    own = ids[first:first + count].reshape(-1)
    bits = _REVEAL_CHAR[own[own != 0]].tobytes().decode("ascii")     # zeros skipped, H0 -> '0' (decoder/util.py:67-81)
    out, rows = handle.decode_run_range(0, first, count, pcm=pcm, exact=exact)
    return dict(meta, first=first, count=count, pcm=out[: rows * ch].reshape(rows, ch), bits=bits)


def rank_world():
    """(rank, world) of the current process; (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def gather_results(local: Dict[int, object]) -> Dict[int, object]:
    """Merge {file_index: small result} dicts of all ranks (off the data path: reveal strings, frame counts, checksums;
    bulk PCM / MP3 bytes stay with the rank that produced them).  Raises if two ranks claim the same file."""
    rank, world = rank_world()
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local)
    merged: Dict[int, object] = {}
    for r, p in enumerate(parts):
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"file {k} was processed by more than one rank (second: {r})")
            merged[k] = v
    return merged


def max_over_ranks(seconds: float, device=None) -> float:
    """Step time of the job = the slowest rank's (the bench contract: device time, max over ranks)."""
    rank, world = rank_world()
    if world == 1:
        return float(seconds)
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def decode_reveal_sharded(handle, blobs: Sequence[bytes], audio_start: Sequence[int] = None, wave_bytes: int = 1 << 30):
    """decode+reveal of this rank's shard of `blobs` (the whole corpus, identical on every rank) through the C ABI.
    Returns {file_index: dict(n_frames, sample_rate, channels, bitrate, bits, pcm)} for the local files only; `pcm` is an
    int16 numpy array [rows, channels].  Files go to the GPU in waves of about `wave_bytes` to bound the workspaces."""
    rank, world = rank_world()
    mine = partition_files([len(b) for b in blobs], world)[rank]
    out = {}
    i = 0
    while i < len(mine):
        j, tot = i, 0
        while j < len(mine) and (j == i or tot + len(blobs[mine[j]]) <= wave_bytes):
            tot += len(blobs[mine[j]])
            j += 1
        idx = mine[i:j]
        data = np.frombuffer(b"".join(blobs[k] for k in idx), np.uint8)
        off = np.concatenate([[0], np.cumsum([len(blobs[k]) for k in idx])]).astype(np.int64)
        sc = handle.decode_scan(data, off, None if audio_start is None else [audio_start[k] for k in idx])
        _, bits = handle.decode_reveal()
        pcm, _ = handle.decode_run()
        pos = 0
        for n, k in enumerate(idx):
            ch = max(int(sc["channels"][n]), 1)
            elems = int(sc["pcm_rows"][n]) * ch
            out[k] = dict(n_frames=int(sc["n_frames"][n]), sample_rate=int(sc["sample_rate"][n]), channels=ch,
                          bitrate=int(sc["bitrate"][n]), bits=bits[n], pcm=pcm[pos:pos + elems].reshape(-1, ch))
            pos += elems
        i = j
    return out
