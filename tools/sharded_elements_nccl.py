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


def element_rows(dg, wins_t, table, elts, sel, rp, d_pr, W, wmap):
    """K4 + K6 + K7 for the elements `sel`: float64 [len(sel), 6] = id, MU, SIGMA, P_SUM, R_SIZE, PVAL."""
    from digdriver_b200 import kernels
    e_chrom, e_strand, ptr, bs, be, obs = elts
    dev = dg.device
    nb = np.diff(ptr)[sel]
    sp = np.concatenate([[0], np.cumsum(nb)])
    take = np.concatenate([np.arange(ptr[i], ptr[i + 1]) for i in sel]) if len(sel) else np.zeros(0, dtype=np.int64)
    owner = np.repeat(np.arange(len(sel)), nb)
    blk, _ = kernels.count_contexts(dg, e_chrom[sel][owner], bs[take], be[take], 1, 1, strand=e_strand[sel][owner])
    pre = kernels.element_transfer(e_chrom[sel].astype(np.int32), e_strand[sel], sp, bs[take], be[take], W, wmap[0], wmap[1],
                                   table, rp["y_pred"], rp["std"], rp["y_true"], rp["flag"], d_pr, blk_counts=blk,
                                   device=dev)
    mu, sigma, p = pre["MU"][0], pre["SIGMA"][0], pre["P"][0][:, 0]
    alpha, theta = mu ** 2 / sigma ** 2, sigma ** 2 / mu
    _, pval = kernels.nb_burden_test(torch.from_numpy(obs[sel]).to(dev), alpha, theta, p, dev, want_exp=False)
    ids = torch.from_numpy(sel.astype(np.float64)).to(dev)
    return torch.stack([ids, mu, sigma, p, pre["R_SIZE"].to(torch.float64), pval], dim=1)


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

    def sharded():
        local_counts, _ = kernels.count_contexts(dg, wins[lo:hi, 0], wins[lo:hi, 1], wins[lo:hi, 2], 1, 1)
        table = sharding.all_gather_rows(coll, local_counts, [b - a for a, b in parts])
        rows = element_rows(dg, wins, table, elts, sel, rp, d_pr, W, wmap)
        return table, coll.gather_rows(rows)

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
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    coll.all_reduce_max(ms)
    ok = None
    if rank == 0:
        full, _ = kernels.count_contexts(dg, wins[:, 0], wins[:, 1], wins[:, 2], 1, 1)
        want = element_rows(dg, wins, full, elts, np.arange(E), rp, d_pr, W, wmap)
        got = got[torch.argsort(got[:, 0])]
        ok = bool(torch.equal(table, full) and torch.equal(got, want))
        print(json.dumps({"workload": "config 4: %d elements (1-3 blocks of 200-2000 bp) on a %.0f Mb genome, %d kb windows, "
                                      "range-sharded x%d" % (E, args.bases / 1e6, W // 1000, world),
                          "n_gpus": world, "bit_identical_to_single_gpu": ok, "ms_sharded_job": float(ms.item()),
                          "elements_per_s": E / (float(ms.item()) / 1e3), "elements_per_rank": [int((owner == r).sum()) for r in range(world)]}))
    if world > 1:
        dist.barrier()
    os._exit(0 if ok in (None, True) else 1)


if __name__ == "__main__":
    main()
