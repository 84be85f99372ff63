#!/usr/bin/env python
"""Golden vectors for the per-position hotspot test, produced by the UNMODIFIED reference
functions nb_model.apply_nb_to_region / nb_pvalue_exact and sequence_tools.base_probabilities_by_region.

Build container only (needs /root/reference); writes tests/golden/position.npz.  The reference reads mutations
through pysam.TabixFile; a duck-typed in-memory object with the same fetch(chrom, start, end) -> tab-separated rows
contract (rows whose [START, END) overlaps the query) stands in for it.

    python tests/golden/make_golden_position.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


class FakeTabix:
    def __init__(self, rows):
        self.rows = rows            # (chrom str, start, end, ref, alt, sample)

    def fetch(self, chrom, start, end):
        return ["\t".join(map(str, r)) for r in self.rows if r[0] == chrom and r[1] < end and r[2] > start]


def main():
    ref = rh.load_reference()
    st, nb = ref.sequence_tools, ref.nb_model
    rng = np.random.default_rng(161616)
    n = 30011
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    seq = np.where(rng.random(n) < 0.5, seq | 0x20, seq).astype(np.uint8)
    for a, ln in ((4000, 37), (9000, 1), (9003, 2), (20000, 800)):
        seq[a:a + ln] = ord("N")
    seq[-2:] = ord("N")
    rh.register_fasta("pos.fa", {"chr1": seq.tobytes().decode()})
    fasta = sys.modules["pysam"].FastaFile("pos.fa")
    # mutations: mostly singletons, some recurrent positions (hotspots), some in N runs, one indel row
    pos = rng.integers(0, n - 1, 1500)
    pos = np.concatenate([pos, np.repeat(rng.integers(0, n - 1, 12), rng.integers(2, 9, 12)), [4010, 4010, 9000]])
    rows = [("1", int(p), int(p) + 1, "A", "C", "S%d" % (i % 17)) for i, p in enumerate(pos)]
    rows.append(("1", 12000, 12004, "ACGT", "A", "S3"))
    tabix = FakeTabix(rows)
    out = {"seq": seq, "mut_start": np.array([r[1] for r in rows], dtype=np.int64)}
    cases = []
    regions = [(0, 1000), (1000, 2000), (3900, 4100), (8990, 9010), (19990, 20810), (20100, 20200), (29000, 30011),
               (29500, 31000), (15000, 15001), (2, 9)]
    ci = 0
    for (u, d) in ((2, 2), (1, 1)):
        kmers = list(st.mk_context_sequences(u, d).keys())
        sp = rng.lognormal(np.log(1e-6), 1.0, len(kmers))
        S_probs = dict(zip(kmers, sp))
        out["s_prob_%d_%d" % (u, d)] = sp
        for binsize in (1, 50, 7):
            for (s, e) in regions:
                if 0 < s < u:
                    continue
                mu = float(rng.gamma(2.0, 5.0)) + 0.05
                sigma = mu * float(rng.uniform(0.05, 0.9))
                with np.errstate(all="ignore"):
                    pvals, poss, obss, exps, pts = nb.apply_nb_to_region(1, s, e, mu, sigma, S_probs, tabix, fasta,
                                                                         n_up=u, n_down=d, binsize=binsize)
                cases.append((u, d, binsize, s, e, mu, sigma, len(pvals)))
                out["pval_%d" % ci] = np.asarray(pvals, dtype=np.float64)
                out["pos_%d" % ci] = np.asarray(poss, dtype=np.float64)
                out["obs_%d" % ci] = np.asarray(obss, dtype=np.int64)
                out["exp_%d" % ci] = np.asarray(exps, dtype=np.float64)
                out["pt_%d" % ci] = np.asarray(pts, dtype=np.float64)
                ci += 1
    out["cases"] = np.array(cases, dtype=np.float64)
    # nb_pvalue_exact on a grid that covers both tails, the pmf fallback and the edge cases
    ks, als, ps = [], [], []
    for k in (0, 1, 2, 5, 17, 60, 300, 2000):
        for a in (1e-3, 0.3, 1.0, 7.5, 120.0, 1e5):
            for p in (1e-9, 1e-3, 0.2, 0.5, 0.93, 0.999999, 1.0):
                ks.append(k); als.append(a); ps.append(p)
    with np.errstate(all="ignore"):
        ex = np.array([nb.nb_pvalue_exact(k, a, p) for k, a, p in zip(ks, als, ps)], dtype=np.float64)
    out["ex_k"], out["ex_alpha"], out["ex_p"], out["ex_pval"] = np.array(ks, float), np.array(als), np.array(ps), ex
    np.savez_compressed(os.path.join(HERE, "position.npz"), **out)
    print("wrote position.npz: %d region cases, %d exact-p cases" % (ci, len(ks)))


if __name__ == "__main__":
    main()
