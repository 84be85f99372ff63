"""Host-buffer pipeline (digdriver_b200/host_pipeline.py): FASTA / pinned host genome -> device -> host count tables.
Parity: the pipelined, per-chromosome, uint16-shipping path must equal the one-shot device path and the CPU oracle bit for
bit, and a packed-cache hit must equal the cold (parse + pack) path."""
import os

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _fasta(tmp_path, seqs, width=61):
    p = tmp_path / "g.fa"
    with open(p, "w") as f:
        for name, s in seqs.items():
            f.write(">%s some description\n" % name)
            t = s.tobytes().decode()
            for i in range(0, len(t), width):
                f.write(t[i:i + width] + "\n")
    return str(p)


def _genome(oracle, lengths, seed=11):
    seqs, off = {}, 0
    for i, n in enumerate(lengths):
        s = oracle.synth_genome(off, n, seed).copy()
        seqs["chr%d" % (i + 1)] = s
        off += (n + 127) // 128 * 128
    return seqs


def test_host_scan_matches_device_path_and_oracle(oracle):
    from digdriver_b200 import host_pipeline as hp, kernels
    from digdriver_b200.genome import DeviceGenome, Genome, tile_windows
    lengths = np.array([410_000, 133_333, 250_001], dtype=np.int64)
    seqs = _genome(oracle, lengths)
    g = Genome(list(seqs), list(seqs.values()))
    W = 10_000
    wins = tile_windows(np.arange(3), lengths, W)
    dg = DeviceGenome.from_genome(g, "cuda:0")
    want5, want3, wt5, wt3 = kernels.count_contexts_fused53(dg, wins[:, 0], wins[:, 1], wins[:, 2], want_totals=True)
    hg = hp.HostGenome.from_genome(g)
    hs = hp.HostScan(hg, wins, "cuda:0").run()
    assert hs.narrow and hs.host_counts.dtype == torch.uint16
    assert np.array_equal(hs.host_counts.numpy().astype(np.int64), want5.cpu().numpy())
    assert np.array_equal(hs.host_counts3.numpy().astype(np.int64), want3.cpu().numpy())
    assert np.array_equal(hs.host_totals.numpy(), torch.cat([wt5, wt3]).cpu().numpy())
    # packed source (what the cache holds): same rows, a third of the upload
    hg2 = hp.HostGenome.from_device(hs.genome)
    hs2 = hp.HostScan(hg2, wins, "cuda:0").run()
    assert hs2.h2d_bytes * 2 < hs.h2d_bytes
    assert np.array_equal(hs2.host_counts.numpy(), hs.host_counts.numpy())
    assert np.array_equal(hs2.host_counts3.numpy(), hs.host_counts3.numpy())
    assert torch.equal(hs2.genome.packed2, dg.packed2) and torch.equal(hs2.genome.nmask, dg.nmask)
    # int32 shipping and a single (n_up, n_down) table
    hs3 = hp.HostScan(hg2, wins, "cuda:0", tables=(1, 1), narrow=False).run()
    assert hs3.host_counts.dtype == torch.int32 and np.array_equal(hs3.host_counts.numpy(), want3.cpu().numpy())
    # oracle on a sample of rows
    seq = np.full(hg.n_bases, ord("N"), dtype=np.uint8)
    for o, s in zip(hg.chrom_off, seqs.values()):
        seq[int(o):int(o) + len(s)] = s
    rows = np.arange(0, len(wins), 7)
    oc, _ = oracle.count_regions(seq, hg.chrom_off, lengths, wins[rows, 0], wins[rows, 1], wins[rows, 2], 2, 2)
    assert np.array_equal(hs.host_counts.numpy()[rows].astype(np.int64), oc)
    # windows grouped by chromosome in another order, and a chromosome without windows
    perm = np.concatenate([np.flatnonzero(wins[:, 0] == 2), np.flatnonzero(wins[:, 0] == 0)])
    hs4 = hp.HostScan(hg2, wins[perm], "cuda:0").run()
    assert np.array_equal(hs4.host_counts.numpy(), hs.host_counts.numpy()[perm])
    assert torch.equal(hs4.genome.packed2, dg.packed2)
    with pytest.raises(ValueError):
        hp.HostScan(hg2, wins[[0, len(wins) - 1, 1]], "cuda:0")


def test_long_regions_fall_back_to_int32(oracle):
    from digdriver_b200 import host_pipeline as hp
    from digdriver_b200.genome import Genome
    lengths = np.array([1_300_000], dtype=np.int64)
    seqs = _genome(oracle, lengths, seed=5)
    hg = hp.HostGenome.from_genome(Genome(list(seqs), list(seqs.values())))
    wins = np.array([[0, 0, 1_250_000], [0, 1000, 2000]], dtype=np.int64)
    hs = hp.HostScan(hg, wins, "cuda:0", tables=(1, 0)).run()
    assert not hs.narrow and hs.host_counts.dtype == torch.int32          # "auto": a 1.25 Mb region may exceed 65535
    assert int(hs.host_counts.numpy()[0].max()) > 65535                    # and with 16 bins it does
    forced = hp.HostScan(hg, wins, "cuda:0", tables=(1, 0), narrow=True).run()
    assert not forced.narrow and forced.host_counts.dtype == torch.int32   # the kernel objected, int32 rows shipped
    assert np.array_equal(forced.host_counts.numpy(), hs.host_counts.numpy())


def test_packed_cache_hit_equals_cold_path(tmp_path, oracle):
    from digdriver_b200 import host_pipeline as hp, storage
    from digdriver_b200.genome import tile_windows
    from digdriver_b200.sequence_model import sequence_tools as st
    lengths = np.array([120_500, 64_000], dtype=np.int64)
    seqs = _genome(oracle, lengths, seed=3)
    fa = _fasta(tmp_path, seqs)
    wins = tile_windows(np.arange(2), lengths, 4096)
    cold5, coldt, g_cold, ex = hp.count_contexts_from_fasta(fa, wins[:, 0], wins[:, 1], wins[:, 2], 2, 2, also_tri=True)
    assert not ex["cache_hit"] and os.path.exists(fa + ".dig2bit/meta.json")
    hot5, hott, g_hot, ex2 = hp.count_contexts_from_fasta(fa, wins[:, 0], wins[:, 1], wins[:, 2], 2, 2, also_tri=True)
    assert ex2["cache_hit"] and ex2["h2d_bytes"] * 2 < ex["h2d_bytes"]
    assert np.array_equal(hot5, cold5) and np.array_equal(hott, coldt)
    assert np.array_equal(ex2["counts3"], ex["counts3"]) and np.array_equal(ex2["totals3"], ex["totals3"])
    assert torch.equal(g_hot.packed2, g_cold.packed2) and torch.equal(g_hot.nmask, g_cold.nmask)
    assert g_hot.names == ["chr1", "chr2"] and list(g_hot.chrom_len) == list(lengths)
    # a rewritten FASTA invalidates the cache
    seqs2 = dict(seqs)
    seqs2["chr2"] = seqs["chr2"][::-1].copy()
    fa2 = _fasta(tmp_path, seqs2)
    os.utime(fa2, ns=(1, 1))
    again5, _, _, ex3 = hp.count_contexts_from_fasta(fa2, wins[:, 0], wins[:, 1], wins[:, 2], 2, 2)
    assert not ex3["cache_hit"]
    n1 = int((wins[:, 0] == 0).sum())
    assert np.array_equal(again5[:n1], cold5[:n1]) and not np.array_equal(again5[n1:], cold5[n1:])
    # the reference-facing entry point goes through the same cache
    st._GENOME_CACHE.clear()
    g = st.get_device_genome(fa2)
    assert torch.equal(g.packed2[: g_cold.packed2.numel() // 2], g_cold.packed2[: g_cold.packed2.numel() // 2])
    # CLI: fused run writes both stores; rows equal two separate runs
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("cli_DigPreprocess", os.path.join(root, "scripts", "DigPreprocess.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    bed = tmp_path / "w.bed"
    with open(bed, "w") as f:
        for c, s, e in wins:
            f.write("%d\t%d\t%d\n" % (c + 1, s, e))
    out5, out3, sep3 = str(tmp_path / "c5"), str(tmp_path / "c3"), str(tmp_path / "s3")
    a = m.parse_args("countGenomeContext %s %s --bed %s --up 2 --down 2 --fout-tri %s" % (fa, out5, bed, out3))
    a.func(a)
    a = m.parse_args("countGenomeContext %s %s --bed %s --up 1 --down 1 --no-cache" % (fa, sep3, bed))
    a.func(a)
    t5 = storage.Store(out5, "r").read_table("all_window_genome_counts")
    t3 = storage.Store(out3, "r").read_table("all_window_genome_counts")
    s3 = storage.Store(sep3, "r").read_table("all_window_genome_counts")
    # fa == fa2 on disk (same path): rows are those of the rewritten genome
    assert np.array_equal(t5.values, again5.astype(np.int64))
    assert np.array_equal(t3.values, s3.values) and list(t3.columns) == list(s3.columns)
    assert np.array_equal(storage.Store(out3, "r").read_table("genome_counts").values, s3.values.sum(axis=0))
    assert storage.Store(out3, "r").get_attrs()["n_up"] == 1
