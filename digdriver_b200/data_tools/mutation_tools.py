"""Drop-in for DIGDriver/data_tools/mutation_tools.py: mutation-file reader and observed-count
tabulation.  The bedtools subprocess + pandas group-bys of the reference are replaced by K5
(interval stabbing + hash-table aggregation on the GPU); file parsing stays pandas.
"""
import csv
import gzip
import os

import numpy as np
import pandas as pd

from .. import kernels

AUTOSOMES = [str(i) for i in range(1, 23)]


def read_mutation_file(path, drop_sex=True, drop_duplicates=False, unique_indels=True):
    """Reference :45-104: headerless TSV whose column count (5-11) selects the schema."""
    try:
        with open(path) as f:
            first_row = next(csv.reader(f, delimiter='\t', skipinitialspace=True))
    except UnicodeDecodeError:
        with gzip.open(path, 'rt') as f:
            first_row = next(csv.reader(f, delimiter='\t', skipinitialspace=True))
    schemas = {
        5: ['CHROM', 'POS', 'REF', 'ALT', 'SAMPLE'],
        6: ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE'],
        7: ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'ANNOT'],
        8: ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'GENE', 'ANNOT'],
        9: ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'ANNOT', 'MUT_TYPE', 'CONTEXT'],
        10: ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'GENE', 'ANNOT', 'MUT_TYPE', 'CONTEXT'],
        11: ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'GENE', 'ANNOT', 'MUT_TYPE', 'CONTEXT', 'STRAND'],
    }
    cols = schemas[len(first_row)]
    dtype = {c: (int if c in ('POS', 'START', 'END') else str) for c in cols}
    df = pd.read_csv(path, sep="\t", low_memory=False, names=cols, dtype=dtype)
    if drop_sex:
        if set(df.CHROM.unique()) - set(AUTOSOMES):
            print('Restricting to autosomes')
            df = df[df.CHROM.isin(AUTOSOMES)]
        df = df.assign(CHROM=df.CHROM.astype(int))
    if drop_duplicates:
        df = drop_duplicate_mutations(df)
    if unique_indels and 'ANNOT' in df.columns:
        df = get_unique_indels(df)
    return df


def drop_duplicate_mutations(df_mut):
    return df_mut.drop_duplicates(['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE'])


def get_unique_indels(df_mut):
    df_indel = df_mut[df_mut.ANNOT == 'INDEL']
    df_snv = df_mut[df_mut.ANNOT != 'INDEL']
    subset = [c for c in ['CHROM', 'START', 'END', 'REF', 'ALT', 'GENE'] if c in df_mut.columns]
    return pd.concat([df_snv, df_indel.drop_duplicates(subset=subset)])


def filter_hypermut_samples(df_mut, max_muts_per_sample, return_blacklist=False):
    """Remove samples that have more mutations than the threshold (reference :293-304)."""
    sample_cnt = df_mut.SAMPLE.value_counts()
    samples_blacklist = sample_cnt[sample_cnt > max_muts_per_sample].index.to_list()
    df_whitelist = df_mut[~df_mut.SAMPLE.isin(samples_blacklist)]
    if return_blacklist:
        return df_whitelist, samples_blacklist
    return df_whitelist


def bed12_boundaries(f_bed):
    """Reference :383-414: CHROM (int, autosomes), ELT, STRAND, BLOCK_STARTS, BLOCK_ENDS per bed12 row."""
    names = ['CHROM', 'START', 'END', "ELT", "SCORE", "STRAND", 'thickStart', 'thickEnd', 'rgb', 'blockCount',
             'blockSizes', 'blockStarts']
    df = pd.read_table(f_bed, names=names, low_memory=False)
    df['CHROM'] = df.CHROM.astype(str).map(lambda x: x[3:] if x.startswith('chr') else x)
    df = df[df.CHROM.isin(AUTOSOMES)].copy()
    df['CHROM'] = df.CHROM.astype(int)
    starts, ends = [], []
    for s0, bs, bz in zip(df.START.values, df.blockStarts.astype(str).values, df.blockSizes.astype(str).values):
        st = [int(x) + int(s0) for x in bs.strip(',').split(',')]
        sz = [int(x) for x in bz.strip(',').split(',')]
        starts.append(st)
        ends.append([a + b for a, b in zip(st, sz)])
    df['BLOCK_STARTS'] = starts
    df['BLOCK_ENDS'] = ends
    return df[['CHROM', 'ELT', 'STRAND', 'BLOCK_STARTS', 'BLOCK_ENDS']]


def _chrom_codes(*series):
    """Shared integer code per chromosome label across several columns ('chr' prefix kept as written,
    exactly like bedtools' string comparison)."""
    labels = pd.unique(np.concatenate([np.asarray(s).astype(str) for s in series]))
    code = {c: i for i, c in enumerate(labels)}
    return [np.array([code[str(v)] for v in s], dtype=np.int64) for s in series]


def _read_bed_blocks(f_elt_bed, bed12):
    bed = pd.read_table(f_elt_bed, header=None, low_memory=False)
    if not bed12:
        return pd.DataFrame({'CHROM': bed[0].astype(str), 'START': bed[1].astype(np.int64),
                             'END': bed[2].astype(np.int64), 'ELT': bed[3].astype(str)})
    rows = []
    for r in bed.itertuples(index=False):
        sizes = [int(x) for x in str(r[10]).strip(',').split(',')]
        starts = [int(x) for x in str(r[11]).strip(',').split(',')]
        for sz, st in zip(sizes, starts):
            rows.append((str(r[0]), int(r[1]) + st, int(r[1]) + st + sz, str(r[3])))
    return pd.DataFrame(rows, columns=['CHROM', 'START', 'END', 'ELT'])


def _read_raw_mutations(f_mut):
    """The mutation file as bedtools sees it: every row, CHROM as written."""
    try:
        df = pd.read_table(f_mut, header=None, low_memory=False, dtype={0: str})
    except UnicodeDecodeError:
        df = pd.read_table(f_mut, header=None, low_memory=False, dtype={0: str}, compression='gzip')
    return df


def _tabulate(f_mut, f_elt_bed, bed12, drop_duplicates, max_muts_per_sample, max_muts_per_elt_per_sample):
    """K5 behind tabulate_muts_per_sample_per_element (:191-230) + tabulate_mutations_in_element (:155-189)."""
    mut = _read_raw_mutations(f_mut)
    blocks = _read_bed_blocks(f_elt_bed, bed12)
    if drop_duplicates:
        # drop_duplicates([0..5, 13]) after the intersect == de-duplicate mutation rows first (:208)
        mut = mut.drop_duplicates([0, 1, 2, 3, 4, 5])
    elts, elt_id = np.unique(blocks.ELT.values, return_inverse=True)
    samples, sample_id = np.unique(mut[5].astype(str).values, return_inverse=True)
    mc, bc = _chrom_codes(mut[0].values, blocks.CHROM.values)
    is_indel = (mut[7].values == 'INDEL') if mut.shape[1] > 7 else np.zeros(len(mut), dtype=bool)
    obs, stot = kernels.tabulate_elements(
        (bc << 32) | blocks.START.values.astype(np.int64), (bc << 32) | blocks.END.values.astype(np.int64), elt_id,
        (mc << 32) | mut[1].values.astype(np.int64), (mc << 32) | mut[2].values.astype(np.int64), sample_id, is_indel,
        len(elts), len(samples), max_muts_per_sample=int(min(max_muts_per_sample, 2 ** 62)),
        max_per_elt_per_sample=int(min(max_muts_per_elt_per_sample, 2 ** 62)))
    obs = obs.cpu().numpy()
    stot = stot.cpu().numpy()
    blacklist = list(samples[stot > max_muts_per_sample])
    df = pd.DataFrame(obs.astype(np.float64), index=pd.Index(elts, name='ELT'),
                      columns=['OBS_SAMPLES', 'OBS_SNV', 'OBS_INDEL'])
    return df[df.OBS_SAMPLES > 0], blacklist


def tabulate_muts_per_sample_per_element(f_mut, f_elt_bed, bed12=False, drop_duplicates=False, unique_indels=True):
    """Reference :191-230: one row per (ELT, SAMPLE) with OBS_SNV, OBS_INDEL, OBS_MUT -- the K5 hash table read back."""
    mut = _read_raw_mutations(f_mut) if not isinstance(f_mut, pd.DataFrame) else f_mut.copy()
    mut.columns = range(mut.shape[1])
    blocks = _read_bed_blocks(f_elt_bed, bed12) if not isinstance(f_elt_bed, pd.DataFrame) else f_elt_bed
    empty = pd.DataFrame({'ELT': [], 'SAMPLE': [], 'OBS_SNV': [], 'OBS_INDEL': [], 'OBS_MUT': []})
    if len(mut) == 0 or len(blocks) == 0:
        return empty
    if drop_duplicates:
        mut = mut.drop_duplicates([0, 1, 2, 3, 4, 5])
    elts, elt_id = np.unique(blocks.ELT.values.astype(str), return_inverse=True)
    samples, sample_id = np.unique(mut[5].astype(str).values, return_inverse=True)
    mc, bc = _chrom_codes(mut[0].values, blocks.CHROM.values)
    is_indel = (mut[7].values == 'INDEL') if mut.shape[1] > 7 else np.zeros(len(mut), dtype=bool)
    _, _, (keys, snv, ind) = kernels.tabulate_elements(
        (bc << 32) | blocks.START.values.astype(np.int64), (bc << 32) | blocks.END.values.astype(np.int64), elt_id,
        (mc << 32) | mut[1].values.astype(np.int64), (mc << 32) | mut[2].values.astype(np.int64), sample_id, is_indel,
        len(elts), len(samples), return_table=True)
    keys = keys.cpu().numpy().astype(np.uint64)
    used = keys != 0
    if not used.any():
        return empty
    k = keys[used] - np.uint64(1)
    df = pd.DataFrame({'ELT': elts[(k >> np.uint64(32)).astype(np.int64)],
                       'SAMPLE': samples[(k & np.uint64(0xFFFFFFFF)).astype(np.int64)],
                       'OBS_SNV': snv.cpu().numpy()[used].astype(np.float64),
                       'OBS_INDEL': ind.cpu().numpy()[used].astype(np.float64)})
    df['OBS_MUT'] = df.OBS_SNV + df.OBS_INDEL
    return df.sort_values(['ELT', 'SAMPLE'], kind='stable').reset_index(drop=True)


def filter_samples_by_stdev(df_mut, stdev_cutoff):
    """Reference :306-316: drop samples whose row count exceeds stdev_cutoff * stdev of the per-sample row counts."""
    sample_cnt = df_mut.SAMPLE.value_counts()
    stdev = sample_cnt.std()
    print(stdev)
    samples_blacklist = sample_cnt[sample_cnt > stdev * stdev_cutoff].index.to_list()
    return df_mut[~df_mut.SAMPLE.isin(samples_blacklist)]


def cap_muts_per_element_per_sample(df_mut_elt_samp, max_muts_per_elt_per_sample):
    """Reference :318-326: caps the OBS_MUT column (only) of the per-sample-per-element table."""
    mask = df_mut_elt_samp.OBS_MUT > max_muts_per_elt_per_sample
    df_mut_elt_samp.loc[mask, 'OBS_MUT'] = max_muts_per_elt_per_sample
    return df_mut_elt_samp


def tabulate_mutations_in_element(f_mut, f_elt_bed, bed12=False, drop_duplicates=False, all_elements=False,
                                  max_muts_per_sample=1e9, max_muts_per_elt_per_sample=3e9, return_blacklist=False):
    """Reference :155-189: OBS_SAMPLES, OBS_SNV, OBS_INDEL per element (index ELT)."""
    df_summary, blacklist = _tabulate(f_mut, f_elt_bed, bed12, drop_duplicates, max_muts_per_sample,
                                      max_muts_per_elt_per_sample)
    if all_elements:
        df_bed = pd.read_csv(f_elt_bed, sep="\t", header=None).set_index(3)
        df_bed.index.rename('ELT', inplace=True)
        df_summary = df_bed.merge(df_summary, left_index=True, right_index=True, how='left')
        for c in ('OBS_SNV', 'OBS_INDEL', 'OBS_SAMPLES'):
            df_summary[c] = df_summary[c].fillna(0)
    out = df_summary[['OBS_SAMPLES', 'OBS_SNV', 'OBS_INDEL']]
    if return_blacklist:
        return out, blacklist
    return out


def mutations_per_gene(df_mut_cds, max_muts_per_gene_per_sample=3e9, return_sample_counts=False):
    """Reference :329-361: OBS_{MIS,NONS,SYN,SPL,INDEL} per gene (index GENE), per-(gene,sample,class) counts
    capped.  With return_sample_counts also the distinct-sample counts of transfer_gene_model
    (transfer_tools.py:243-265)."""
    genes, gid = np.unique(df_mut_cds.GENE.values.astype(str), return_inverse=True)
    samples, sid = np.unique(df_mut_cds.SAMPLE.values.astype(str), return_inverse=True)
    cls_map = {"Synonymous": 0, "Missense": 1, "Nonsense": 2, "Essential_Splice": 3, "INDEL": 4}
    cls = np.array([cls_map.get(a, 255) for a in df_mut_cds.ANNOT.values], dtype=np.uint8)
    obs, nsamp = kernels.tabulate_genes(gid, sid, cls, len(genes),
                                        max_per_gene_per_sample=int(min(max_muts_per_gene_per_sample, 2 ** 62)))
    obs, nsamp = obs.cpu().numpy(), nsamp.cpu().numpy()
    idx = pd.Index(genes, name='GENE')
    df_counts = pd.DataFrame({'OBS_SPL': obs[:, 3], 'OBS_INDEL': obs[:, 4], 'OBS_MIS': obs[:, 1],
                              'OBS_NONS': obs[:, 2], 'OBS_SYN': obs[:, 0]}, index=idx)
    if return_sample_counts:
        df_ns = pd.DataFrame(nsamp, index=idx, columns=['N_SAMP_SYN', 'N_SAMP_MIS', 'N_SAMP_NONS', 'N_SAMP_SPL',
                                                        'N_SAMP_TRUNC', 'N_SAMP_NONSYN', 'N_SAMP_INDEL'])
        return df_counts, df_ns
    return df_counts


def tabulate_nonc_mutations_at_sites(f_sites, f_mut, return_sites=False):
    """Reference :233-275: mutations matching a site exactly on (CHROM, START, END, REF, ALT, GENE, ANNOT,
    MUT_TYPE, CONTEXT); OBS_SNV = matched rows, OBS_SAMPLES = distinct samples, per site-set (ELT)."""
    df_sites = read_mutation_file(f_sites)
    df_sites = df_sites.rename({"SAMPLE": "ELT"}, axis=1)
    assert ('GENE' in df_sites.columns and 'ANNOT' in df_sites.columns and 'MUT_TYPE' in df_sites.columns)
    if 'STRAND' not in df_sites.columns:
        print("WARNING: strand column not detected in sites file. Defaulting all sites to + strand.")
        df_sites['STRAND'] = "+"
    df_mut = read_mutation_file(f_mut, drop_duplicates=False)
    assert ('GENE' in df_mut.columns and 'ANNOT' in df_mut.columns and 'MUT_TYPE' in df_mut.columns)
    if len(df_mut[df_mut.ANNOT == 'INDEL']):
        print('WARNING: INDELS found in mutation file. Dig sites model is only applicable to SNVs. '
              'INDELS will be dropped.')
    df_mut = df_mut[df_mut.ANNOT != 'INDEL']
    on = ['CHROM', 'START', 'END', 'REF', 'ALT', 'GENE', 'ANNOT', 'MUT_TYPE', 'CONTEXT']
    # exact match == zero-length-interval stabbing on a key; the (ELT, SAMPLE) aggregation is K5's table
    key_s = df_sites[on].astype(str).agg('|'.join, axis=1)
    key_m = df_mut[on].astype(str).agg('|'.join, axis=1)
    ukeys, inv = np.unique(np.concatenate([key_s.values, key_m.values]), return_inverse=True)
    ks, km = inv[:len(key_s)].astype(np.int64), inv[len(key_s):].astype(np.int64)
    elts, elt_id = np.unique(df_sites.ELT.values.astype(str), return_inverse=True)
    samples, sid = np.unique(df_mut.SAMPLE.values.astype(str), return_inverse=True)
    obs, _ = kernels.tabulate_elements(ks, ks + 1, elt_id, km, km + 1, sid, np.zeros(len(km), dtype=bool),
                                       len(elts), len(samples))
    obs = obs.cpu().numpy()
    # K5 counts a mutation once per site-set; the reference's merge counts it once per matching site ROW
    # (:259-261), so site rows repeated inside one set add their mutation count again (integer glue)
    pair, mult = np.unique(np.stack([ks, elt_id.astype(np.int64)]), axis=1, return_counts=True)
    if np.any(mult > 1):
        per_key = np.bincount(km, minlength=len(ukeys))
        extra = (mult - 1) * per_key[pair[0]]
        np.add.at(obs[:, 1], pair[1], extra)
    keep = obs[:, 0] > 0
    counts = pd.DataFrame({'ELT': elts[keep], 'OBS_SAMPLES': obs[keep, 0], 'OBS_SNV': obs[keep, 1]})
    if return_sites:
        df_sites = df_sites.drop(columns=['GENE', 'ANNOT', 'REF', 'ALT']).rename(columns={'ELT': 'GENE'}).set_index('GENE')
        return counts, df_sites
    return counts


def tabulate_sites_in_element(f_sites, f_mut):
    df_res = tabulate_nonc_mutations_at_sites(f_sites, f_mut).set_index('ELT')
    return df_res[['OBS_SAMPLES', 'OBS_SNV']]


# ---------------------------------------------------------------------------------------------
# bedtools-intersect front ends of the reference, on the overlap-join kernel (dig_overlap_count / dig_overlap_fill)
# ---------------------------------------------------------------------------------------------

_BED_FIELDS = ['chrom', 'start', 'end', 'name', 'score', 'strand', 'thickStart', 'thickEnd', 'itemRgb', 'blockCount',
               'blockSizes', 'blockStarts']


def _overlap_join(a_chrom, a_start, a_end, b_chrom, b_start, b_end):
    """(index into a, index into b) for every pair with a.start < b.end and b.start < a.end on the same chromosome
    label (string comparison, as bedtools does)."""
    ac, bc = _chrom_codes(np.asarray(a_chrom), np.asarray(b_chrom))
    return kernels.overlap_pairs((bc << 32) | np.asarray(b_start, dtype=np.int64), (bc << 32) | np.asarray(b_end, dtype=np.int64),
                                 (ac << 32) | np.asarray(a_start, dtype=np.int64), (ac << 32) | np.asarray(a_end, dtype=np.int64))


def restrict_mutations_by_bed(df_mut, df_bed, unique=True, remove_X=True, replace_cols=False):
    """Reference :8-31 (`bed_mut.intersect(bed_bed)`): one row per overlapping (mutation, interval) pair, START / END
    clipped to the overlap, columns named like pybedtools' to_dataframe(); exact duplicate rows dropped when unique."""
    if remove_X:
        df_mut = df_mut[df_mut.iloc[:, 0] != "X"]
        df_bed = df_bed[df_bed.iloc[:, 0] != "X"]
    im, ib = _overlap_join(df_mut.iloc[:, 0].astype(str).values, df_mut.iloc[:, 1].values, df_mut.iloc[:, 2].values,
                           df_bed.iloc[:, 0].astype(str).values, df_bed.iloc[:, 1].values, df_bed.iloc[:, 2].values)
    df_inter = df_mut.iloc[im].reset_index(drop=True)
    c1, c2 = df_inter.columns[1], df_inter.columns[2]
    df_inter[c1] = np.maximum(df_inter[c1].values, df_bed.iloc[ib, 1].values)
    df_inter[c2] = np.minimum(df_inter[c2].values, df_bed.iloc[ib, 2].values)
    if df_inter.shape[1] <= len(_BED_FIELDS):
        df_inter.columns = _BED_FIELDS[:df_inter.shape[1]]
    if unique:
        df_inter = df_inter.drop_duplicates()
    if replace_cols:
        df_inter.columns = df_mut.columns
    return df_inter


def _raw_rows_to_mutation_frame(raw, drop_duplicates, drop_sex):
    """read_mutation_file applied to already-parsed rows (what the reference does by re-reading bedtools' temp file)."""
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".tsv", delete=False) as f:
        raw.to_csv(f, sep="\t", header=False, index=False)
        name = f.name
    try:
        if len(raw) == 0:
            return pd.DataFrame()
        return read_mutation_file(name, drop_duplicates=drop_duplicates, drop_sex=drop_sex)
    finally:
        os.remove(name)


def restrict_mutations_by_bed_efficient(f_mut, f_bed, bed12=False, drop_duplicates=False, drop_sex=False,
                                        replace_cols=False):
    """Reference :33-43 (`intersect -wa`): the mutation rows overlapping the bed file (bed12: its blocks), one copy
    per overlapping interval, then read_mutation_file's clean-up."""
    mut = _read_raw_mutations(f_mut)
    blocks = _read_bed_blocks(f_bed, bed12) if bed12 else pd.read_table(f_bed, header=None, low_memory=False,
                                                                         dtype={0: str}).rename(
        columns={0: 'CHROM', 1: 'START', 2: 'END'})
    im, _ = _overlap_join(mut[0].values, mut[1].values, mut[2].values, blocks.CHROM.astype(str).values,
                          blocks.START.values, blocks.END.values)
    return _raw_rows_to_mutation_frame(mut.iloc[im], drop_duplicates, drop_sex)


def mutations_by_element(f_mut, f_elt_bed, bed12=False, drop_duplicates=False):
    """Reference :363-381 (`intersect -wa -wb`): every (mutation, element interval) pair with the element's name;
    expects the 10-column annotated mutation format."""
    mut = _read_raw_mutations(f_mut)
    blocks = _read_bed_blocks(f_elt_bed, bed12)
    im, ib = _overlap_join(mut[0].values, mut[1].values, mut[2].values, blocks.CHROM.values, blocks.START.values,
                           blocks.END.values)
    df_hits = mut.iloc[im, :10].reset_index(drop=True)
    df_hits[13] = blocks.ELT.values[ib]
    if drop_duplicates:
        df_hits = df_hits.drop_duplicates([0, 1, 2, 3, 4, 5, 13])
    df_hits.columns = ['CHROM', 'START', 'END', 'REF', 'ALT', 'SAMPLE', 'GENE', 'ANNOT', 'TYPE', 'CONTEXT', 'ELT']
    return df_hits


def tabulate_nonc_mutations_split(f_nonc_bed, f_mut):
    """Reference :120-153 (`bed6().intersect(mut, wao=True)` + pivot): per (CHROM, ELT, STRAND) the sorted distinct
    block starts / ends, the number of distinct samples hitting any block and the overlapping base pairs summed over
    (block, mutation) pairs.  Returns (None, df_whole) like the reference."""
    df_nonc = pd.read_table(f_nonc_bed, names=['CHROM', 'START', 'END', "ELT", "SCORE", "STRAND", 'thickStart',
                                               'thickEnd', 'rgb', 'blockCount', 'blockSizes', 'blockStarts'],
                            low_memory=False)
    df_nonc['CHROM'] = df_nonc.CHROM.astype(str)
    df_nonc = df_nonc[df_nonc.CHROM.isin(AUTOSOMES)].reset_index(drop=True)
    df_mut = read_mutation_file(f_mut, drop_duplicates=True)
    assert ('GENE' in df_mut.columns and 'ANNOT' in df_mut.columns and 'MUT_TYPE' in df_mut.columns)
    rows = []
    for r in df_nonc.itertuples(index=False):
        sizes = [int(x) for x in str(r.blockSizes).strip(',').split(',')]
        starts = [int(x) for x in str(r.blockStarts).strip(',').split(',')]
        for sz, st in zip(sizes, starts):
            rows.append((int(r.CHROM), int(r.START) + st, int(r.START) + st + sz, r.ELT, r.STRAND))
    blk = pd.DataFrame(rows, columns=['CHROM', 'START', 'END', 'ELT', 'STRAND'])
    im, ib = _overlap_join(df_mut.CHROM.astype(str).values, df_mut.START.values, df_mut.END.values,
                           blk.CHROM.astype(str).values, blk.START.values, blk.END.values)
    ov = np.minimum(df_mut.END.values[im], blk.END.values[ib]) - np.maximum(df_mut.START.values[im], blk.START.values[ib])
    grp = blk.groupby(['CHROM', 'ELT', 'STRAND'], sort=True)
    gid = grp.ngroup().values
    hits = pd.DataFrame({'G': gid[ib], 'SAMPLE': df_mut.SAMPLE.values[im], 'OV': ov})
    n_samp = hits.groupby('G').SAMPLE.nunique()
    n_bp = hits.groupby('G').OV.sum()
    df_whole = grp.agg(BLOCK_STARTS=('START', lambda x: sorted(set(x))), BLOCK_ENDS=('END', lambda x: sorted(set(x)))).reset_index()
    g_of_row = np.arange(len(df_whole))
    df_whole['OBS_SAMPLES'] = n_samp.reindex(g_of_row).fillna(0).astype(np.int64).values
    df_whole['OBS_MUT'] = n_bp.reindex(g_of_row).fillna(0).astype(np.int64).values
    return None, df_whole


def _genic_fill_empty_cols(df_gene_counts):
    """Ensures that every mutation class has a column (reference :282-291)."""
    for c in {'Essential_Splice', 'Missense', 'Nonsense', 'Stop_loss', 'Synonymous'} - set(df_gene_counts.columns):
        df_gene_counts[c] = 0
