"""world_size-2 gloo test (CPU) of the range-sharding logic: each rank counts its slice of the windows on its
slice of the genome (with the CPU oracle standing in for the kernels), totals are all-reduced and rows gathered;
the result must equal the single-process run bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from digdriver_b200 import sharding
    from digdriver_b200.genome import Genome, tile_windows
    from oracle import dig_oracle
    lengths = np.array([50_003, 31_000, 12_345], dtype=np.int64)
    seqs = [dig_oracle.synth_genome(int(o), int(n), 5) for o, n in zip(np.cumsum(lengths) - lengths, lengths)]
    for s in seqs:
        s[100:140] = ord("N")
    g = Genome(["chr1", "chr2", "chr3"], seqs)
    wins = tile_windows([0, 1, 2], lengths, 1000)
    wins = np.concatenate([wins, [[2, 12_000, 12_345], [0, 50_000, 51_000]]])      # chromosome-end windows
    coll = sharding.Collectives()
    lo, hi = sharding.partition_windows(wins[:, 1], wins[:, 2], world)[rank]
    mine = wins[lo:hi]
    for (u, d) in ((1, 1), (2, 2)):
        sub, c, s, e = sharding.slice_genome(g, mine[:, 0], mine[:, 1], mine[:, 2], halo=max(u, d))
        off = np.concatenate([[0], np.cumsum(sub.lengths)[:-1]])
        counts, _ = dig_oracle.count_regions(np.concatenate(sub.seqs) if len(sub.seqs) else np.zeros(0, np.uint8),
                                             off, sub.lengths, c, s, e, u, d)
        totals = coll.all_reduce_sum(torch.from_numpy(counts.sum(axis=0)))
        rows = coll.gather_rows(torch.from_numpy(counts))
        # the known-sizes form (no size exchange, no host reads: the one used inside the CUDA graph of bench.py)
        sizes = [b - a for a, b in sharding.partition_windows(wins[:, 1], wins[:, 2], world)]
        rows2 = coll.gather_rows(torch.from_numpy(counts), sizes=sizes)
        if rank == 0:
            assert torch.equal(rows, rows2)
            full_off = np.concatenate([[0], np.cumsum(lengths)[:-1]])
            want, _ = dig_oracle.count_regions(np.concatenate(seqs), full_off, lengths, wins[:, 0], wins[:, 1],
                                               wins[:, 2], u, d)
            ret["ok_%d" % u] = bool(np.array_equal(rows.numpy(), want) and
                                    np.array_equal(totals.numpy(), want.sum(axis=0)))
    dist.destroy_process_group()


def test_range_sharded_scan_world2():
    from oracle import dig_oracle
    dig_oracle.build()
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret.get("ok_1") is True and ret.get("ok_2") is True


def test_partition_windows_balanced():
    sys.path.insert(0, ROOT)
    from digdriver_b200 import sharding
    s = np.arange(0, 100_000, 1000)
    parts = sharding.partition_windows(s, s + 1000, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 100
    assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
    assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    assert sharding.partition_windows([], [], 2) == [(0, 0), (0, 0)]
