"""Per-window observed mutation counts (the Y_TRUE track) and window tiling: drop-in for the mutation branch of
``add_objectives`` and the tiling rule of ``extract_high_mappability`` in the reference's scripts/DataExtractor.py
(:55-81, :525-572) -- the step immediately before the hot path.

The counting is K5 (``kernels.tabulate_elements``) with the windows as one-block elements; nothing here runs on
the CPU except reading the mutation file and the O(n_sample) filter thresholds.
"""
import numpy as np
import pandas as pd

from .. import kernels
from . import mutation_tools


def tile_windows(chrom_sizes, window, overlap=0):
    """``idx`` rows (CHROM, START, END) as the reference tiles them: from 0 in steps of ``window - overlap`` while
    ``i + window < chrom_size`` (DataExtractor.py:70-77).  chrom_sizes: {chrom label: length} in order."""
    rows = []
    for chrom, size in chrom_sizes.items():
        i = 0
        while i + window < size:
            rows.append((chrom, i, i + window))
            i += window - overlap
    return np.array(rows, dtype=np.int64).reshape(-1, 3)


def window_mutation_counts(f_mut, idx, max_muts_per_elt_per_sample=None, sample_filter_stdev=None,
                           max_muts_per_sample=None, device=None):
    """OBS_SNV per window exactly as add_objectives builds it (DataExtractor.py:540-559):

    * rows of ``tabulate_muts_per_sample_per_element(mut_file, windows, bed12=False, drop_duplicates=True)``;
    * ``cap_muts_per_element_per_sample`` caps the OBS_MUT column only (mutation_tools.py:318-326), so it does NOT
      change the OBS_SNV sums -- reproduced (the argument is accepted and has no effect on the result);
    * ``filter_samples_by_stdev`` / ``filter_hypermut_samples`` see ``SAMPLE.value_counts()`` of that table, i.e. the
      number of WINDOWS a sample hits (not its mutation count): samples above ``stdev * cutoff`` (pandas std, ddof=1)
      or above ``max_muts_per_sample`` are dropped;
    * the remaining OBS_SNV are summed per window; windows without mutations get 0.

    f_mut: path of a mutation file or a DataFrame with its columns in file order (>= 6 columns: CHROM START END REF ALT
    SAMPLE [GENE ANNOT ...]).  idx: int array [n, 3] of (CHROM, START, END).  Returns int64 [n]."""
    import torch
    dev = device or torch.device("cuda", torch.cuda.current_device())
    mut = mutation_tools._read_raw_mutations(f_mut) if not isinstance(f_mut, pd.DataFrame) else f_mut.copy()
    mut.columns = range(mut.shape[1])
    mut = mut.drop_duplicates([0, 1, 2, 3, 4, 5])
    idx = np.asarray(idx, dtype=np.int64).reshape(-1, 3)
    n_win = len(idx)
    if n_win == 0:
        return np.zeros(0, dtype=np.int64)
    samples, sample_id = np.unique(mut[5].astype(str).values, return_inverse=True)
    mc, bc = mutation_tools._chrom_codes(mut[0].values, idx[:, 0])
    is_indel = (mut[7].values == 'INDEL') if mut.shape[1] > 7 else np.zeros(len(mut), dtype=bool)
    args = ((bc << 32) | idx[:, 1], (bc << 32) | idx[:, 2], np.arange(n_win, dtype=np.int32),
            (mc << 32) | mut[1].values.astype(np.int64), (mc << 32) | mut[2].values.astype(np.int64), sample_id,
            is_indel, n_win, len(samples))
    limit = None
    if sample_filter_stdev or max_muts_per_sample:
        _, rows = kernels.tabulate_elements(*args, device=dev, sample_rows_mode=True)
        rows = rows.cpu().numpy().astype(np.float64)
        present = rows[rows > 0]
        limits = []
        if sample_filter_stdev:
            std = float(np.std(present, ddof=1)) if len(present) > 1 else float("nan")
            limits.append(std * sample_filter_stdev)
        if max_muts_per_sample:
            limits.append(float(max_muts_per_sample))
        limits = [v for v in limits if not np.isnan(v)]       # `cnt > nan` is False in pandas: nobody is dropped
        if limits:
            limit = min(limits)
    if limit is None:
        obs, _ = kernels.tabulate_elements(*args, device=dev, sample_rows_mode=True)
    else:
        # "rows > limit" for integer row counts == "rows > floor(limit)"
        obs, _ = kernels.tabulate_elements(*args, max_muts_per_sample=int(np.floor(limit)), device=dev,
                                           sample_rows_mode=True)
    return obs[:, 1].cpu().numpy().astype(np.int64)


def add_objectives(h5_file, mut_file, max_muts_per_sample=None, sample_filter_stdev=None,
                   max_muts_per_elt_per_sample=None, suffix='', cnv=False):
    """``DataExtractor.py addObjectives`` (:525-572): reads ``idx`` from the archive, counts, and stores the track
    under the cohort name derived from the mutation file name (float dataset, as the reference writes it)."""
    import os
    from ..storage import Store
    if cnv:
        raise NotImplementedError("CNV tracks (fetch_cnv_region_avg) are outside the hot path")
    st = Store(h5_file, "a")
    idx = st.read_array("idx")
    counts = window_mutation_counts(mut_file, idx, max_muts_per_elt_per_sample, sample_filter_stdev,
                                    max_muts_per_sample)
    cancer = os.path.basename(str(mut_file)).split('.annot')[0].split('.txt')[0].split('.bed')[0] + suffix
    st.write_array(cancer, counts.astype(np.float64), dtype=float)
    return cancer, counts
