"""CPU: the N>1 host logic (file partition, off-path gather, max-over-ranks step time) on a world_size-2 gloo group.
The codec work of each rank is done by the oracle here (no GPU in this suite); on the GPU box the same functions run
under nccl around the C-ABI calls (bench.py, shard.decode_reveal_sharded)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mp3-steganography-lib_b200")


def test_partition_is_exact_cover_and_balanced():
    sys.path.insert(0, PKG)
    from mp3stego_b200 import shard
    rng = np.random.default_rng(3)
    sizes = rng.integers(1_000, 9_000_000, size=1000)
    for world in (1, 2, 3, 8):
        parts = shard.partition_files(sizes, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(sizes)))
        loads = [int(sizes[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= int(sizes.max())        # greedy longest-first bound
    assert shard.partition_files([], 2) == [[], []]
    assert shard.partition_files([5, 5, 5], 2) == [[0, 2], [1]]   # ties go to the lowest rank: every rank agrees


def test_frame_ranges_cover_and_warm_up():
    sys.path.insert(0, PKG)
    from mp3stego_b200 import shard
    for n, parts in ((6890, 8), (5, 8), (1, 1), (0, 4), (1148, 3)):
        r = shard.frame_ranges(n, parts)
        assert sum(x["count"] for x in r) == n
        pos = 0
        for x in r:
            assert x["first"] == pos and x["warm"] == (1 if pos > 0 else 0)
            pos += x["count"]


def test_frame_shard_plans_agree_and_cover():
    """Every rank derives every rank's frame range of a long file from the same scan result: exact cover, the halo rule, and the
    whole-prefix halo of files whose granules inherit state from earlier frames."""
    sys.path.insert(0, PKG)
    from mp3stego_b200 import shard
    CARRY = 8   # M3S_FILE_STATE_CARRY (include/mp3stego_b200.h)
    for n, world in ((6890, 8), (25, 4), (3, 8), (0, 2), (1148, 3)):
        for status in (0, 4, CARRY):
            plans = [shard.plan_frame_shard(n, status, r, world) for r in range(world)]
            pos = 0
            for p in plans:
                if p["count"]:
                    assert p["first"] == pos
                    assert p["warm"] == (1 if p["first"] > 0 else 0)
                    assert p["compact_from"] == (0 if status & CARRY else max(p["first"] - p["warm"] - 9, 0))
                    pos += p["count"]
            assert pos == n


def _worker(rank, world, port, blobs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, PKG)
    import torch.distributed as dist
    from mp3stego_b200 import shard
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.partition_files([len(b) for b in blobs], world)[rank]
        local = {}
        for k in mine:
            r = O.decode(blobs[k], 0, taps=False)
            local[k] = (int(r["n_frames"]), r["bits"], int(np.asarray(r["pcm16"], np.int64).sum()))
        merged = shard.gather_results(local)
        t = shard.max_over_ranks(1.0 + rank)
        q.put((rank, mine, merged, t))
    finally:
        dist.destroy_process_group()


def test_sharded_decode_reveal_world2_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    O.build()
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.mp3")))
    assert len(files) >= 4
    blobs = [open(f, "rb").read() for f in files]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blobs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {r: mine for r, mine, _, _ in got}
    assert sorted(shards[0] + shards[1]) == list(range(len(blobs))) and shards[0] and shards[1]
    single = {}
    for k, b in enumerate(blobs):
        r = O.decode(b, 0, taps=False)
        single[k] = (int(r["n_frames"]), r["bits"], int(np.asarray(r["pcm16"], np.int64).sum()))
    for _, _, merged, t in got:
        assert merged == single               # every rank sees the whole job's results, identical to one process
        assert t == pytest.approx(2.0)        # max over ranks of (1.0, 2.0)


class _OracleHandle:
    """Stands in for _lib.Handle in the CPU suite: the same four calls decode_frame_range makes, answered by the oracle
    (whole-file sequential decode), so the host logic of range sharding is exercised without a GPU."""

    def __init__(self, O):
        self.O, self.r = O, None

    def decode_scan(self, data, file_off, audio_start=None):
        blob = bytes(np.asarray(data, np.uint8)[int(file_off[0]):int(file_off[1])])
        self.r = self.O.decode(blob, 0 if audio_start is None else int(audio_start[0]), taps=False)
        r = self.r
        return dict(n_frames=np.array([r["n_frames"]], np.int64), pcm_rows=np.array([r["pcm16"].shape[0]], np.int64),
                    sample_rate=np.array([r["sampling_rate"]], np.int32), channels=np.array([r["channels"]], np.int32),
                    bitrate=np.array([r["bit_rate"]], np.int32), status=np.array([0], np.int32))

    def decode_reveal(self):
        return self.r["tables"], [self.r["bits"]]

    def decode_frame_pos(self):
        return self.r["frame_off"]

    def decode_run(self, exact=False):
        return self.r["pcm16"].reshape(-1), None

    def decode_run_range(self, file_index, first, count, pcm=None, exact=False):
        ch = max(self.r["channels"], 1)
        rows = self.r["pcm16"].shape[0] - first * 1152 if first + count == self.r["n_frames"] else count * 1152
        return self.r["pcm16"][first * 1152: first * 1152 + rows].reshape(-1), rows


def _range_worker(rank, world, port, blob, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, PKG)
    import torch.distributed as dist
    from mp3stego_b200 import shard
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = shard.decode_frame_range(_OracleHandle(O), blob, rank, world)
        merged = shard.gather_results({rank: (part["first"], part["count"], part["bits"], part["pcm"].tobytes())})
        q.put((rank, merged))
    finally:
        dist.destroy_process_group()


def test_frame_range_sharded_decode_world2_gloo():
    """One long file split by frame range over two ranks (each scans the whole file and decodes its range + one warm-up frame): the ranges,
    gathered, are the whole-file decode sample for sample and bit for bit."""
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synth_wav
    from oracle import oracle as O
    O.build()
    wav = synth_wav(31, 48)
    bits = O.str_to_bits("11#frame range")
    blob = O.encode(wav, 44100, 128, bits, taps=False)["mp3"]
    whole = O.decode(blob, 0, taps=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_range_worker, args=(r, 2, port, blob, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, merged in got:
        assert sorted(merged) == [0, 1]
        assert merged[0][0] == 0 and merged[1][0] == merged[0][1] and merged[0][1] + merged[1][1] == whole["n_frames"]
        assert merged[0][2] + merged[1][2] == whole["bits"]
        assert merged[0][3] + merged[1][3] == whole["pcm16"].tobytes()
    assert O.reveal_parse(whole["bits"]) == "frame range"
