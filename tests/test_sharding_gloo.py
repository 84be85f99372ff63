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


def _element_worker(rank, world, port, ret):
    """Config 4 sharded: windows by genomic range, elements by their first block, window counts all-gathered."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from digdriver_b200 import sharding
    from digdriver_b200.genome import Genome, tile_windows
    from oracle import dig_oracle
    W = 1000
    lengths = np.array([40_000, 25_000], dtype=np.int64)
    seqs = [dig_oracle.synth_genome(int(o), int(n), 9) for o, n in zip(np.cumsum(lengths) - lengths, lengths)]
    g = Genome(["chr1", "chr2"], seqs)
    wins = tile_windows([0, 1], lengths, W)
    rng = np.random.default_rng(77)                                        # same inputs on every rank
    E = 60
    e_chrom = rng.integers(1, 3, E)
    nb = rng.integers(1, 4, E)
    ptr = np.concatenate([[0], np.cumsum(nb)])
    bs = np.concatenate([np.sort(rng.integers(100, (lengths[c - 1] // W - 2) * W, k)) for c, k in zip(e_chrom, nb)])
    be = bs + rng.integers(20, 1500, len(bs))                              # some blocks span a window / shard boundary
    e_strand = rng.choice([-1, 1], E).astype(np.int8)
    L = rng.integers(0, 9, (E, 192)).astype(np.float64)
    n_win = len(wins)
    y_pred, std = rng.gamma(2.0, 10.0, n_win), rng.uniform(0.5, 5.0, n_win)
    y_true, flag = rng.poisson(20, n_win).astype(float), rng.random(n_win) < 0.1
    d_pr = rng.lognormal(np.log(1e-6), 1.0, 192)
    win_index = {(int(c) + 1, int(s)): i for i, (c, s, e) in enumerate(wins)}
    coll = sharding.Collectives()
    parts = sharding.partition_windows(wins[:, 1], wins[:, 2], world)
    lo, hi = parts[rank]
    mine = wins[lo:hi]
    sub, c, s, e = sharding.slice_genome(g, mine[:, 0], mine[:, 1], mine[:, 2], halo=1)
    off = np.concatenate([[0], np.cumsum(sub.lengths)[:-1]])
    local, _ = dig_oracle.count_regions(np.concatenate(sub.seqs), off, sub.lengths, c, s, e, 1, 1)
    table = sharding.all_gather_rows(coll, torch.from_numpy(local), [b - a for a, b in parts]).numpy()
    owner = sharding.partition_elements(e_chrom - 1, bs[ptr[:-1]], wins[:, 0], wins[:, 1], wins[:, 2], parts)
    sel = np.flatnonzero(owner == rank)
    sp = np.concatenate([[0], np.cumsum(nb[sel])])
    take = np.concatenate([np.arange(ptr[i], ptr[i + 1]) for i in sel]) if len(sel) else np.zeros(0, dtype=np.int64)
    res = dig_oracle.element_transfer(e_chrom[sel], e_strand[sel], sp, bs[take], be[take], L[sel], W, win_index, table,
                                      y_pred, std, y_true, flag, d_pr)
    rows = torch.from_numpy(np.stack([sel.astype(np.float64), res["MU"], res["SIGMA"], res["P_SUM"],
                                      res["R_SIZE"].astype(np.float64)], axis=1))
    got = coll.gather_rows(rows)
    if rank == 0:
        got = got.numpy()
        got = got[np.argsort(got[:, 0])]
        full_off = np.concatenate([[0], np.cumsum(lengths)[:-1]])
        full, _ = dig_oracle.count_regions(np.concatenate(seqs), full_off, lengths, wins[:, 0], wins[:, 1], wins[:, 2], 1, 1)
        want = dig_oracle.element_transfer(e_chrom, e_strand, ptr, bs, be, L, W, win_index, full, y_pred, std, y_true,
                                           flag, d_pr)
        ret["all_owned"] = bool(np.array_equal(got[:, 0], np.arange(E)) and np.all(owner >= 0))
        ret["both_ranks_work"] = bool(0 < (owner == 0).sum() < E)
        ret["table"] = bool(np.array_equal(table, full))
        ret["equal"] = bool(np.array_equal(got[:, 1], want["MU"]) and np.array_equal(got[:, 2], want["SIGMA"]) and
                            np.array_equal(got[:, 3], want["P_SUM"]) and np.array_equal(got[:, 4], want["R_SIZE"]))
    dist.destroy_process_group()


def test_range_sharded_element_model_world2():
    from oracle import dig_oracle
    dig_oracle.build()
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_element_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {"all_owned": True, "both_ranks_work": True, "table": True, "equal": True}


def test_partition_elements_edges():
    sys.path.insert(0, ROOT)
    from digdriver_b200 import sharding
    wc = np.array([0, 0, 0, 1, 1]); ws = np.array([0, 1000, 2000, 0, 1000]); we = ws + 1000
    parts = [(0, 2), (2, 5)]
    owner = sharding.partition_elements([0, 0, 0, 1, 1, 1, 2], [0, 1999, 2000, 500, 1999, 2000, 10], wc, ws, we, parts)
    assert owner.tolist() == [0, 0, 1, 1, 1, -1, -1]              # chr2:2000 and chr3 lie in no window
    assert sharding.partition_elements([], [], wc, ws, we, parts).tolist() == []


def _gathered_worker(rank, world, port, ret):
    """ONE genome over two ranks the way bench.py's strong mode does it: each rank counts its slice into its block of
    the exchange buffer (rows + partial totals), ONE all_gather_into_tensor, and the element stage resolves windows
    through the remapped window map -- straddling elements included."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from digdriver_b200 import sharding
    from digdriver_b200.genome import Genome, tile_windows
    from oracle import dig_oracle
    W = 1000
    lengths = np.array([40_500, 23_000], dtype=np.int64)
    seqs = [dig_oracle.synth_genome(int(o), int(n), 9) for o, n in zip(np.cumsum(lengths) - lengths, lengths)]
    g = Genome(["chr1", "chr2"], seqs)
    wins = tile_windows([0, 1], lengths, W)
    parts = sharding.partition_windows(wins[:, 1], wins[:, 2], world)
    gt = sharding.GatheredTable(parts)
    lo, hi = parts[rank]
    mine = wins[lo:hi]
    sub, c, s, e = sharding.slice_genome(g, mine[:, 0], mine[:, 1], mine[:, 2], halo=2)
    off = np.concatenate([[0], np.cumsum(sub.lengths)[:-1]])
    cat = np.concatenate(sub.seqs)
    c3, _ = dig_oracle.count_regions(cat, off, sub.lengths, c, s, e, 1, 1)
    c5, _ = dig_oracle.count_regions(cat, off, sub.lengths, c, s, e, 2, 2)
    local = torch.zeros((gt.block_rows, 64), dtype=torch.int32)
    rows, t5, t3, sub = gt.local_views(local)
    rows[: hi - lo] = torch.from_numpy(c3.astype(np.int32))
    t5.copy_(torch.from_numpy(c5.sum(axis=0)))
    t3.copy_(torch.from_numpy(c3.sum(axis=0)))
    sub.copy_(torch.arange(192, dtype=torch.int64) * (rank + 1) + (1 << 33))          # partial substitution counts
    gathered = torch.zeros((world, gt.block_rows, 64), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered.view(-1, 64), local)
    tot = gt.summed_totals(gathered).numpy()
    full_off = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    f3, _ = dig_oracle.count_regions(np.concatenate(seqs), full_off, lengths, wins[:, 0], wins[:, 1], wins[:, 2], 1, 1)
    f5, _ = dig_oracle.count_regions(np.concatenate(seqs), full_off, lengths, wins[:, 0], wins[:, 1], wins[:, 2], 2, 2)
    table = gathered.view(-1, 64).numpy()
    row_of = gt.row_of_window(len(wins))
    ok_rows = bool(np.array_equal(table[row_of], f3))
    ok_tot = bool(np.array_equal(tot[:1024], f5.sum(axis=0)) and np.array_equal(tot[1024:1088], f3.sum(axis=0)) and
                  np.array_equal(tot[1088:], np.arange(192) * 3 + (1 << 34)) and len(tot) == 1280)
    # the window map resolves (chromosome, window number) to gathered rows; an element straddling the cut sees both
    moff, wmap = gt.window_map(wins[:, 0], wins[:, 1], W, 2)
    cut = parts[0][1]
    c_cut, s_cut = int(wins[cut, 0]), int(wins[cut, 1])
    ok_map = s_cut >= W and int(wins[cut - 1, 0]) == c_cut
    if ok_map:
        r_left = wmap[moff[c_cut] + s_cut // W - 1]
        r_right = wmap[moff[c_cut] + s_cut // W]
        ok_map = bool(np.array_equal(table[r_left], f3[cut - 1]) and np.array_equal(table[r_right], f3[cut]) and
                      r_left // gt.block_rows == 0 and r_right // gt.block_rows == 1)
    ret["r%d" % rank] = (ok_rows, ok_tot, bool(ok_map))
    dist.destroy_process_group()


def test_gathered_table_world2():
    from oracle import dig_oracle
    dig_oracle.build()
    world = 2
    port = 33500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gathered_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {"r0": (True, True, True), "r1": (True, True, True)}
