"""Config 4 (noncoding elements, genome range-sharded over the GPUs of one box) with the real kernels over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/sharded_elements_nccl.py [--bases 400000000] [--elements 100000]

Every rank scans its genomic range of windows (K2), the [Nw, 64] window-count table is all-gathered, elements are
assigned by their first block (sharding.partition_elements), each rank runs K4 block counts + K6 + the NB test for its
own elements and rank 0 gathers the rows.  Rank 0 then repeats the whole job alone and the two results must be
bit-identical.  Prints one JSON line with device-timed (max over ranks) figures."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class ElementShard:
    """Device-resident CSR of the elements one rank owns (built once, outside the timed job)."""

    def __init__(self, elts, sel, dev, W):
        from digdriver_b200 import kernels
        e_chrom, e_strand, ptr, bs, be, obs = elts
        nb = np.diff(ptr)[sel]
        sp = np.concatenate([[0], np.cumsum(nb)])
        take = (np.repeat(ptr[:-1][sel] - sp[:-1], nb) + np.arange(sp[-1])) if len(sel) else np.zeros(0, dtype=np.int64)
        owner = np.repeat(np.arange(len(sel)), nb)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)     # noqa: E731
        self.chrom, self.strand = t(e_chrom[sel], torch.int32), t(e_strand[sel], torch.int8)
        self.ptr, self.bs, self.be = t(sp, torch.int64), t(bs[take], torch.int64), t(be[take], torch.int64)
        self.blk_chrom, self.blk_strand = t(e_chrom[sel][owner], torch.int32), t(e_strand[sel][owner], torch.int8)
        self.obs, self.ids = t(obs[sel], torch.float64), t(sel, torch.float64)
        self.max_span = kernels.element_max_span(sp, bs[take], be[take], W)


def element_rows(dg, table, sh, rp, d_pr, W, wmap, sink=None):
    """K4 + K6 + K7 for one shard's elements: float64 [n, 6] = id, MU, SIGMA, P_SUM, R_SIZE, PVAL."""
    from digdriver_b200 import kernels
    dev = dg.device
    blk, _ = kernels.count_contexts(dg, sh.blk_chrom, sh.bs, sh.be, 1, 1, strand=sh.blk_strand)
    pre = kernels.element_transfer(sh.chrom, sh.strand, sh.ptr, sh.bs, sh.be, W, wmap[0], wmap[1], table, rp["y_pred"],
                                   rp["std"], rp["y_true"], rp["flag"], d_pr, blk_counts=blk, device=dev,
                                   max_span=sh.max_span, status_sink=sink)
    mu, sigma, p = pre["MU"][0], pre["SIGMA"][0], pre["P"][0][:, 0]
    alpha, theta = mu ** 2 / sigma ** 2, sigma ** 2 / mu
    _, pval = kernels.nb_burden_test(sh.obs, alpha, theta, p, dev, want_exp=False)
    return torch.stack([sh.ids, mu, sigma, p, pre["R_SIZE"].to(torch.float64), pval], dim=1)


_ALL = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=int, default=400_000_000)
    ap.add_argument("--elements", type=int, default=100_000)
    ap.add_argument("--window", type=int, default=10_000)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from digdriver_b200 import genome as G, kernels, sharding
    W = args.window
    lengths = G.hg19_like_lengths(args.bases)
    dg = G.DeviceGenome.synthetic(["chr%d" % (i + 1) for i in range(22)], lengths, seed=4, device=dev)
    wins = G.tile_windows(np.arange(22), lengths, W)
    n_win = len(wins)
    rng = np.random.default_rng(44)                                   # identical inputs on every rank
    E = args.elements
    usable = (lengths - 1) // W * W - 2500
    e_chrom = np.sort(rng.choice(22, E, p=usable / usable.sum()))
    nb = rng.integers(1, 4, E)
    ptr = np.concatenate([[0], np.cumsum(nb)])
    first = (rng.random(E) * (usable[e_chrom] - 10)).astype(np.int64) + 5
    owner_blk = np.repeat(np.arange(E), nb)
    step = rng.integers(250, 2500, len(owner_blk))
    rel = np.cumsum(step) - step
    rel -= rel[ptr[:-1]][owner_blk]
    bs = first[owner_blk] + rel
    be = bs + rng.integers(200, 2000, len(bs))                        # 200-2000 bp blocks, some across window borders
    be = np.minimum(be, (lengths[e_chrom][owner_blk] - 1) // W * W - 1)
    bs = np.minimum(bs, be - 1)
    e_strand = rng.choice([-1, 1], E).astype(np.int8)
    obs = rng.poisson(1.5, E).astype(np.float64)
    rp = {"y_pred": rng.gamma(2.0, 10.0, n_win), "std": rng.uniform(0.5, 5.0, n_win),
          "y_true": rng.poisson(20, n_win).astype(np.float64), "flag": rng.random(n_win) < 0.1}
    d_pr = rng.lognormal(np.log(1e-6), 1.0, 192)
    wmap = kernels.build_window_map(wins[:, 0], wins[:, 1], W, 22)
    elts = (e_chrom, e_strand, ptr, bs, be, obs)
    coll = sharding.Collectives()
    parts = sharding.partition_windows(wins[:, 1], wins[:, 2], world)
    lo, hi = parts[rank]
    owner = sharding.partition_elements(e_chrom, bs[ptr[:-1]], wins[:, 0], wins[:, 1], wins[:, 2], parts)
    assert np.all(owner >= 0)
    sel = np.flatnonzero(owner == rank)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)         # noqa: E731
    rp = {k: t(v.astype(np.uint8) if v.dtype == bool else v, torch.uint8 if v.dtype == bool else torch.float64)
          for k, v in rp.items()}
    d_pr = t(d_pr, torch.float64)
    wmap = (t(wmap[0], torch.int64), t(wmap[1], torch.int32))
    w_dev = (t(wins[:, 0], torch.int32), t(wins[:, 1], torch.int64), t(wins[:, 2], torch.int64))
    mine = ElementShard(elts, sel, dev, W)
    sizes = [b - a for a, b in parts]

    sink = []                                             # device status words, read once after the timed job
    n_own = [int((owner == r).sum()) for r in range(world)]

    def sharded():
        local_counts, _ = kernels.count_contexts(dg, w_dev[0][lo:hi], w_dev[1][lo:hi], w_dev[2][lo:hi], 1, 1)
        table = sharding.all_gather_rows(coll, local_counts, sizes)
        rows = element_rows(dg, table, mine, rp, d_pr, W, wmap, sink)
        return table, coll.gather_rows(rows, sizes=n_own)

    for _ in range(2):
        sharded()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    table, got = sharded()
    e1.record()
    torch.cuda.synchronize()
    kernels.check_deferred(sink)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    coll.all_reduce_max(ms)
    ok = None
    if rank == 0:
        full, _ = kernels.count_contexts(dg, w_dev[0], w_dev[1], w_dev[2], 1, 1)
        _ALL.append(ElementShard(elts, np.arange(E), dev, W))
        want = element_rows(dg, full, _ALL[0], rp, d_pr, W, wmap)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        kernels.count_contexts(dg, w_dev[0], w_dev[1], w_dev[2], 1, 1)
        element_rows(dg, full, _ALL[0], rp, d_pr, W, wmap)
        s1.record()
        torch.cuda.synchronize()
        ms_single = s0.elapsed_time(s1)
        got = got[torch.argsort(got[:, 0])]
        same = lambda a, b: bool(torch.equal(torch.isnan(a), torch.isnan(b)) and   # noqa: E731
                                 torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)))
        table_ok, rows_ok = bool(torch.equal(table, full)), same(got, want)
        ok = table_ok and rows_ok
        n_nan = int(torch.isnan(want).any(dim=1).sum().item())       # elements whose windows hold no countable base
        print(json.dumps({"workload": "config 4: %d elements (1-3 blocks of 200-2000 bp) on a %.0f Mb genome, %d kb windows, "
                                      "range-sharded x%d" % (E, args.bases / 1e6, W // 1000, world),
                          "n_gpus": world, "bit_identical_to_single_gpu": ok, "window_table_identical": table_ok,
                          "element_rows_identical": rows_ok, "elements_with_nan_rows": n_nan, "ms_sharded_job": float(ms.item()), "ms_same_job_one_gpu": ms_single,
                          "elements_per_s": E / (float(ms.item()) / 1e3), "elements_per_rank": [int((owner == r).sum()) for r in range(world)]}))
    if world > 1:
        dist.barrier()
    os._exit(0 if ok in (None, True) else 1)


if __name__ == "__main__":
    main()
