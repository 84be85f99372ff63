"""GPU parity tests for the per-window observed counts (Y_TRUE): K5 with windows as elements,
against goldens built with the reference's own sample filters and against the pandas oracle."""
import os

import numpy as np
import pandas as pd
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

COLS = ("CHROM", "START", "END", "REF", "ALT", "SAMPLE", "GENE", "ANNOT")


def _frame(z):
    return pd.DataFrame({k: z["mut_" + k] for k in COLS})


def test_window_counts_golden():
    from digdriver_b200.data_tools import objectives
    z = golden("objectives")
    df = _frame(z)
    for i, c in enumerate(z["cases"]):
        cap, std, mx = [None if np.isnan(v) else v for v in c]
        got = objectives.window_mutation_counts(df, z["idx"], cap, std, mx)
        assert np.array_equal(got, z["y_%d" % i]), (i, got, z["y_%d" % i])


def test_add_objectives_cli_and_tiling(tmp_path, oracle):
    """DataExtractor.py addObjectives through the CLI shim on a mutation FILE and a directory store, 1 kb windows,
    20 k mutations; result equals the oracle's add_objectives restatement."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import importlib
    de = importlib.import_module("DataExtractor")
    from digdriver_b200.data_tools import objectives
    from digdriver_b200.storage import Store
    rng = np.random.default_rng(9)
    sizes = {1: 250_300, 2: 99_999, 3: 1000}
    idx = objectives.tile_windows(sizes, 1000)
    assert len(idx) == 250 + 99 and idx[-1].tolist() == [2, 98_000, 99_000]       # while i + W < size
    n = 20_000
    chrom = rng.choice([1, 2], n, p=[0.7, 0.3])
    start = (rng.random(n) * np.where(chrom == 1, 250_300, 99_999)).astype(np.int64)
    indel = rng.random(n) < 0.05
    df = pd.DataFrame({"CHROM": chrom.astype(str), "START": start, "END": start + np.where(indel, 5, 1),
                       "REF": rng.choice(list("ACGT"), n), "ALT": rng.choice(list("ACGT"), n),
                       "SAMPLE": ["P%d" % s for s in rng.zipf(1.6, n) % 60], "GENE": ".",
                       "ANNOT": np.where(indel, "INDEL", "Noncoding")})
    df = pd.concat([df, df.iloc[:500]], ignore_index=True)
    f_mut = tmp_path / "COHORT.annot.txt"
    df.to_csv(f_mut, sep="\t", header=False, index=False)
    store = tmp_path / "data.h5"
    Store(str(store), "w").write_array("idx", idx.astype(np.int32))
    args = de.parse_args(["addObjectives", str(store), str(f_mut), "--sample-filter-stdev", "2.0",
                          "--max-muts-per-sample", "200", "--suffix", "_x"])
    args.func(args)
    got = Store(str(store), "r").read_array("COHORT_x")
    assert got.dtype == np.float64
    want = oracle.window_objectives(df.copy(), idx, None, 2.0, 200)
    assert np.array_equal(got.astype(np.int64), want)
    assert 0 < want.sum() < len(df)
    # unfiltered: every de-duplicated SNV row inside a tiled window counts once
    plain = objectives.window_mutation_counts(str(f_mut), idx)
    assert np.array_equal(plain, oracle.window_objectives(df.copy(), idx))


def test_muts_per_sample_per_element_table(oracle):
    """tabulate_muts_per_sample_per_element (mutation_tools.py:191-230): the (ELT, SAMPLE) rows read back from the K5
    hash table equal the pandas restatement, and the reference-named filters applied to it reproduce add_objectives."""
    from digdriver_b200.data_tools import mutation_tools as mt
    z = golden("objectives")
    df = _frame(z)
    idx = z["idx"]
    blocks = pd.DataFrame({"CHROM": idx[:, 0].astype(str), "START": idx[:, 1], "END": idx[:, 2],
                           "ELT": ["%d:%d-%d" % tuple(r) for r in idx]})
    got = mt.tabulate_muts_per_sample_per_element(df, blocks, drop_duplicates=True)
    want = oracle.muts_per_sample_per_element(df, blocks, drop_duplicates=True)
    want = want.sort_values(["ELT", "SAMPLE"], kind="stable").reset_index(drop=True)
    assert list(got.columns) == ['ELT', 'SAMPLE', 'OBS_SNV', 'OBS_INDEL', 'OBS_MUT']
    assert got.ELT.tolist() == want.ELT.tolist() and got.SAMPLE.tolist() == want.SAMPLE.tolist()
    for c in ("OBS_SNV", "OBS_INDEL", "OBS_MUT"):
        assert np.array_equal(got[c].values, want[c].values), c
    t = mt.filter_hypermut_samples(mt.filter_samples_by_stdev(mt.cap_muts_per_element_per_sample(got.copy(), 3), 3.0), 13)
    y = t.groupby("ELT").OBS_SNV.sum().reindex(blocks.ELT).fillna(0).astype(int).values
    assert np.array_equal(y, z["y_4"])                      # golden case (cap 3, stdev 3.0, max 13)
