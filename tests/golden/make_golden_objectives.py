#!/usr/bin/env python
"""Golden vectors for the per-window observed counts.  scripts/DataExtractor.py cannot be
imported here (bbi / pysam / h5py / pybedtools at module level and bedtools binaries), so add_objectives' pandas
steps are re-typed in oracle.window_objectives (GLUE) while the three filter functions it calls are the UNMODIFIED
reference functions from DIGDriver/data_tools/mutation_tools.py (filter_samples_by_stdev prints the stdev).

    python tests/golden/make_golden_objectives.py        (build container only)
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh, dig_oracle  # noqa: E402


def main():
    ref = rh.load_reference()
    mt = ref.mutation_tools
    filters = (mt.cap_muts_per_element_per_sample, mt.filter_samples_by_stdev, mt.filter_hypermut_samples)
    rng = np.random.default_rng(4242)
    sizes = {1: 100_500, 2: 61_000}
    W = 10_000
    idx = []
    for c, n in sizes.items():
        i = 0
        while i + W < n:
            idx.append((c, i, i + W))
            i += W
    idx = np.array(idx, dtype=np.int64)
    n_mut = 3000
    chrom = rng.choice([1, 2, 3], n_mut, p=[0.6, 0.35, 0.05])
    start = np.array([rng.integers(0, sizes.get(int(c), 50_000)) for c in chrom])
    is_indel = rng.random(n_mut) < 0.08
    end = start + np.where(is_indel, rng.integers(1, 30, n_mut), 1)
    w = 1.0 / np.arange(1, 41) ** 1.3                    # Zipf-ish sample sizes: a few hypermutators
    sample = rng.choice(40, n_mut, p=w / w.sum())
    df = pd.DataFrame({"CHROM": chrom.astype(str), "START": start, "END": end,
                       "REF": rng.choice(list("ACGT"), n_mut), "ALT": rng.choice(list("ACGT"), n_mut),
                       "SAMPLE": ["S%02d" % s for s in sample], "GENE": ".",
                       "ANNOT": np.where(is_indel, "INDEL", "Noncoding")})
    df = pd.concat([df, df.iloc[:150]], ignore_index=True)                    # duplicated rows
    df.loc[len(df)] = ["1", 9_995, 10_010, "A", "-", "S03", ".", "INDEL"]     # indel straddling two windows
    out = {"idx": idx}
    for k in df.columns:
        v = np.asarray(df[k].values)
        out["mut_" + k] = v if v.dtype.kind in "iuf" else np.array([str(e) for e in v], dtype=str)
    cases = [(None, None, None), (2, None, None), (None, 2.5, None), (None, None, 12), (3, 3.0, 13), (None, 5.0, 8)]
    for i, (cap, std, mx) in enumerate(cases):
        out["y_%d" % i] = dig_oracle.window_objectives(df.copy(), idx, cap, std, mx, filters=filters)
    out["cases"] = np.array([[np.nan if v is None else v for v in c] for c in cases], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "objectives.npz"), **out)
    print("wrote objectives.npz", [int(out["y_%d" % i].sum()) for i in range(len(cases))])


if __name__ == "__main__":
    main()
