"""Drop-in for DIGDriver/sequence_model/genic_driver_tools.py: element / gene "pretrain" -- the set of
windows an element overlaps, the sums of the region model over those windows and the context-weighted
mutability fraction.  The per-element Python loops (h5py row reads + pandas .loc lookups) are replaced
by kernel K6 (one warp per element); see csrc/transfer.cu.
"""
import math

import numpy as np
import pandas as pd
import torch

from .. import kernels, storage
from . import sequence_tools


def reverse_complement(seq):
    return sequence_tools.reverse_complement(seq)


def trip_to_str(trip):
    return 'chr{}:{}-{}'.format(trip[0], trip[1], trip[2])


def get_ideal_overlaps(chrom, intervals, window):
    """Windows overlapped by a set of intervals (reference :275-283).  The reference returns list(set(...)) in
    hash order; the same set is returned here in ascending order."""
    region = set()
    for i in np.asarray(intervals).T:
        low = math.floor(i[0].min() / window) * window
        high = math.ceil(i[1].max() / window) * window
        borders = np.arange(low, high + window, window)
        for j in range(len(borders) - 1):
            region.add((chrom, int(borders[j]), int(borders[j + 1])))
    return sorted(region)


def get_elt_ideal_overlaps(chrom, start, end, window):
    return [(int(c), s, e) for c, s, e in get_ideal_overlaps(chrom, np.array([[start], [end]]), window)]


class RegionModel:
    """region_params (reference: region_model_tools.py:93-126) held as device-ready arrays + the dense
    (chromosome, window number) -> row map used by K6."""

    def __init__(self, df):
        self.df = df
        self.window = int(df.iloc[0]['END'] - df.iloc[0]['START'])
        chrom = df.CHROM.values.astype(np.int64)
        self.n_chrom = int(chrom.max()) + 1 if len(chrom) else 1        # index by chromosome number itself
        self.win_map_off, self.win_map = kernels.build_window_map(chrom, df.START.values, self.window, self.n_chrom)
        self.y_pred = df.Y_PRED.values.astype(np.float64)
        self.std = df.STD.values.astype(np.float64)
        self.y_true = df.Y_TRUE.values.astype(np.float64)
        self.flag = df.FLAG.values.astype(bool) if 'FLAG' in df.columns else np.zeros(len(df), dtype=bool)


def _one_cohort(out):
    o = {k: v.cpu().numpy() for k, v in out.items()}
    return {"MU": o["MU"][0], "SIGMA": o["SIGMA"][0], "R_OBS": o["R_OBS"][0], "FLAG": o["FLAG"][0].astype(bool),
            "R_SIZE": o["R_SIZE"], "ELT_SIZE": o["ELT_SIZE"], "P": o["P"][0], "N_WIN": o["N_WIN"]}


def transfer_elements(region_model, win_counts64, d_pr, elt_chrom, elt_strand, blk_ptr, blk_start, blk_end,
                      blk_counts=None, L_elt=None):
    """K6 for a batch of elements against one region model.  elt_chrom holds chromosome NUMBERS (1..22).
    Returns a dict of host arrays (MU, SIGMA, R_OBS, FLAG, R_SIZE, ELT_SIZE, P [E, n_col], N_WIN)."""
    rm = region_model
    dev = torch.device("cuda", torch.cuda.current_device())
    out = kernels.element_transfer(np.asarray(elt_chrom, dtype=np.int32), np.asarray(elt_strand, dtype=np.int8),
                                   blk_ptr, blk_start, blk_end, rm.window, rm.win_map_off, rm.win_map, win_counts64,
                                   rm.y_pred, rm.std, rm.y_true, rm.flag, d_pr, blk_counts=blk_counts, L_elt=L_elt,
                                   device=dev)
    return _one_cohort(out)


def get_region_params_direct(df, overlaps, window):
    """mu, sigma, R_obs, FLAG over a list of (chrom, start, end) windows (reference :258-272), via K6."""
    rm = df if isinstance(df, RegionModel) else RegionModel(df)
    ov = sorted(overlaps)
    if len(ov) == 0:
        return 0, 0.0, 0, False
    chrom = int(ov[0][0])
    starts = np.array([o[1] for o in ov], dtype=np.int64)
    n = len(df) if not isinstance(df, RegionModel) else len(rm.df)
    res = transfer_elements(rm, np.zeros((n, 64), dtype=np.int32), np.ones(192), [chrom], [1],
                            [0, len(ov)], starts, starts + 1, L_elt=np.zeros((1, 192, 1)))
    return res["MU"][0], res["SIGMA"][0], res["R_OBS"][0], bool(res["FLAG"][0])


def get_region_params(df, chrom, intervals, window):
    """Reference :235-250."""
    return get_region_params_direct(df, get_ideal_overlaps(chrom, intervals, window), window)


def _strand_code(values):
    return np.array([-1 if (s == '-' or s == '-1' or s == -1) else 1 for s in values], dtype=np.int8)


def _indel_params(region_model_indels, win_counts64, d_pr, chrom, strand, ptr, bs, be, res):
    """MU / SIGMA / R_OBS of the same elements against region_params_indels (--indels-direct, reference :161-164 and
    :383-386); without it the SNV parameters are copied.  The window counts do not enter these three outputs, so a
    zero table of the indel model's own size is passed when the SNV one does not fit it."""
    if region_model_indels is None:
        return res["MU"], res["SIGMA"], res["R_OBS"]
    n = len(region_model_indels.df)
    wc = win_counts64 if len(win_counts64) == n else np.zeros((n, 64), dtype=np.int32)
    E = len(chrom)
    ind = transfer_elements(region_model_indels, wc, d_pr, chrom, strand, ptr, bs, be, L_elt=np.zeros((E, 192, 1)))
    return ind["MU"], ind["SIGMA"], ind["R_OBS"]


def nonc_model_arrays(df_elts, L_contexts, region_model, win_counts64, d_pr, f_fasta=None, region_model_indels=None):
    """Element pretrain on in-memory inputs: preprocess_nonc (sequence_tools.py:596-644) + nonc_model
    (reference :300-431) == the loop body of DIG_onthefly (onthefly_tools.py:109-165).

    df_elts: bed12_boundaries() frame.  L_contexts: precount_region_contexts_parallel() frame (index
    'chr{c}:{s}-{e}', 192 columns).  Returns the pretrain DataFrame of reference :404-417."""
    E = len(df_elts)
    ptr = np.zeros(E + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(b) for b in df_elts.BLOCK_STARTS])
    bs = np.array([x for b in df_elts.BLOCK_STARTS for x in b], dtype=np.int64)
    be = np.array([x for b in df_elts.BLOCK_ENDS for x in b], dtype=np.int64)
    owner = np.repeat(np.arange(E), np.diff(ptr))
    chrom = df_elts.CHROM.values.astype(np.int64)
    keys = ['chr{}:{}-{}'.format(c, s, e) for c, s, e in zip(chrom[owner], bs, be)]
    Lb = L_contexts.loc[keys].values                          # raises KeyError like the reference (:637)
    L = np.zeros((E, 192), dtype=np.float64)
    np.add.at(L, owner, Lb)
    strand = _strand_code(df_elts.STRAND.values)
    res = transfer_elements(region_model, win_counts64, d_pr, chrom, strand, ptr, bs, be, L_elt=L.reshape(E, 192, 1))
    mu_ind, sigma_ind, r_ind = _indel_params(region_model_indels, win_counts64, d_pr, chrom, strand, ptr, bs, be, res)
    elt_size = (L.sum(axis=1) / 3).astype(np.int64)           # int(np.sum(L) / 3)  (:380)
    with np.errstate(divide='ignore', invalid='ignore'):
        p_indel = elt_size / res["R_SIZE"].astype(np.float64)
    return pd.DataFrame({
        'ELT': df_elts.ELT.values, 'ELT_SIZE': elt_size, 'FLAG': res["FLAG"], 'R_SIZE': res["R_SIZE"],
        'R_OBS': res["R_OBS"], 'R_INDEL': r_ind, 'MU': res["MU"], 'SIGMA': res["SIGMA"],
        'MU_INDEL': mu_ind, 'SIGMA_INDEL': sigma_ind, 'P_SUM': res["P"][:, 0], 'P_INDEL': p_indel})


def _window_counts_for(region_model, f_fasta):
    """Trinucleotide counts of every window of the region model (K2)."""
    df = region_model.df
    g = sequence_tools.get_device_genome(f_fasta)
    cidx = g.chrom_indices(df.CHROM.values)
    counts, _ = kernels.count_contexts(g, cidx, df.START.values, df.END.values, 1, 1)
    return counts


def nonc_model_parallel(f_pretrained, f_nonc_data, nonc_L_key, N_procs=1, indels_direct=False):
    """Reference :434-461 on the directory/HDF5 store written by preprocess_element_model."""
    pre = storage.Store(f_pretrained, "r")
    rm = RegionModel(pre.read_table('region_params'))
    rm_ind = RegionModel(pre.read_table('region_params_indels')) if indels_direct else None     # reference :316-317
    d_pr = sequence_tools.d_pr_from_model192(pre.read_table('sequence_model_192'))
    data = storage.Store(f_nonc_data, "r")
    wkey = 'window_{}'.format(rm.window)
    win_idx = data.read_array('{}/full_window_si_index'.format(wkey))
    win_vals = data.read_array('{}/full_window_si_values'.format(wkey))
    # align the stored window counts to the region model's rows
    key = {tuple(r): i for i, r in enumerate(map(tuple, win_idx))}
    rows = np.array([key[(int(c), int(s), int(e))] for c, s, e in zip(rm.df.CHROM, rm.df.START, rm.df.END)])
    win_counts = win_vals[rows].astype(np.int32)
    if data.has('{}/{}/sites'.format(wkey, nonc_L_key)):               # preprocess_element_model --f-sites
        df_sites = data.read_table('{}/{}/sites'.format(wkey, nonc_L_key))
        if indels_direct:
            raise NotImplementedError("--indels-direct is not defined for site sets (the reference's sites model copies "
                                      "the SNV parameters, genic_driver_tools.py:660)")
        return sites_model_arrays(df_sites.reset_index(drop=True), rm, win_counts, d_pr)
    df_elts = data.read_table('{}/{}/elements'.format(wkey, nonc_L_key))
    df_elts['BLOCK_STARTS'] = [list(map(int, s.split(','))) for s in df_elts.BLOCK_STARTS]
    df_elts['BLOCK_ENDS'] = [list(map(int, s.split(','))) for s in df_elts.BLOCK_ENDS]
    L_contexts = data.read_table('{}/{}/L_contexts'.format(wkey, nonc_L_key))
    return nonc_model_arrays(df_elts, L_contexts, rm, win_counts, d_pr, region_model_indels=rm_ind)


def genic_model_arrays(genes, region_model, win_counts64, d_pr, region_model_indels=None):
    """genic_model (reference :31-203) on in-memory inputs.  ``genes``: pipeline.GeneTable with chromosome
    NUMBERS in chrom_idx.  Returns the genic pretrain DataFrame (reference :170-201)."""
    res = transfer_elements(region_model, win_counts64, d_pr, genes.chrom_idx, genes.strand, genes.blk_ptr,
                            genes.blk_start, genes.blk_end, L_elt=genes.L)
    P = res["P"]
    mu_ind, sigma_ind, r_ind = _indel_params(region_model_indels, win_counts64, d_pr, genes.chrom_idx, genes.strand,
                                             genes.blk_ptr, genes.blk_start, genes.blk_end, res)
    with np.errstate(divide='ignore', invalid='ignore'):
        p_indel = res["ELT_SIZE"] / res["R_SIZE"].astype(np.float64)
    df = pd.DataFrame({'CHROM': [str(c) for c in genes.chrom_idx], 'GENE': genes.names,
                       'GENE_LENGTH': res["ELT_SIZE"], 'R_SIZE': res["R_SIZE"], 'R_OBS': res["R_OBS"],
                       'R_INDEL': r_ind, 'MU': res["MU"], 'SIGMA': res["SIGMA"], 'MU_INDEL': mu_ind,
                       'SIGMA_INDEL': sigma_ind, 'FLAG': res["FLAG"], 'P_MIS': P[:, 1], 'P_NONS': P[:, 2],
                       'P_SILENT': P[:, 0], 'P_SPLICE': P[:, 3], 'P_TRUNC': P[:, 2] + P[:, 3], 'P_INDEL': p_indel})
    return df


def genic_model_parallel(f_pretrained_str, f_genic_str, N_procs=1, counts_key="window_10kb/counts",
                         indels_direct=False, f_fasta=None, genes_lst=None):
    """Reference :206-226.  f_genic is a store with the f_genic layout flattened to arrays: 'genes' table
    (GENE, CHROM, STRAND), 'cds_ptr', 'cds_start', 'cds_end' (inclusive intervals) and 'L_data' [E,192,4];
    the window trinucleotide counts come from f_pretrained's 'window_counts_64' (written by
    DigPretrain.py genicModel from the genome-count file) or are scanned from f_fasta."""
    from ..pipeline import GeneTable
    pre = storage.Store(f_pretrained_str, "r")
    rm = RegionModel(pre.read_table('region_params'))
    rm_ind = RegionModel(pre.read_table('region_params_indels')) if indels_direct else None     # reference :43-44
    d_pr = sequence_tools.d_pr_from_model192(pre.read_table('sequence_model_192'))
    gs = storage.Store(f_genic_str, "r")
    meta = gs.read_table('genes')
    keep = ~meta.CHROM.astype(str).isin(['X', 'Y']).values            # reference :89-90
    ptr = gs.read_array('cds_ptr')
    sel = np.flatnonzero(keep)
    if genes_lst is not None:                                         # genic_model's chunk, in the caller's order
        row_of = {str(g): i for i, g in enumerate(meta.GENE.values)}
        rows = np.array([row_of[str(g)] for g in genes_lst], dtype=np.int64)      # KeyError for an unknown gene
        sel = rows[keep[rows]]
    lens = np.diff(ptr)[sel]
    new_ptr = np.zeros(len(sel) + 1, dtype=np.int64)
    new_ptr[1:] = np.cumsum(lens)
    take = np.concatenate([np.arange(ptr[i], ptr[i + 1]) for i in sel]) if len(sel) else np.zeros(0, dtype=np.int64)
    genes = GeneTable(chrom_idx=meta.CHROM.values[sel].astype(np.int32), strand=_strand_code(meta.STRAND.values[sel]),
                      blk_ptr=new_ptr, blk_start=gs.read_array('cds_start')[take], blk_end=gs.read_array('cds_end')[take],
                      L=gs.read_array('L_data')[sel], names=list(meta.GENE.values[sel]))
    if pre.has('window_counts_64'):
        wc = pre.read_array('window_counts_64').astype(np.int32)
    else:
        wc = _window_counts_for(rm, f_fasta)
    return genic_model_arrays(genes, rm, wc, d_pr, region_model_indels=rm_ind)


def sites_model_arrays(df_sites, region_model, win_counts64, d_pr):
    """preprocess_sites (sequence_tools.py:647-711) + nonc_model (reference :300-431) for site sets.

    df_sites: a sites file read with mutation_tools.read_mutation_file -- its SAMPLE column holds the site-set
    name (ELT), MUT_TYPE / CONTEXT give each site's substitution, STRAND (optional) flips it.  Per set:
    L[j] = number of sites with substitution j (K5's dig_site_counts), windows = overlaps of the sites' own
    (START, END) intervals, strand = the set's first STRAND value.  Returns the pretrain DataFrame."""
    from .sequence_tools import mk_trans_idx
    df = df_sites.rename(columns={'SAMPLE': 'ELT'}) if 'ELT' not in df_sites.columns else df_sites
    if 'STRAND' not in df.columns:
        df = df.assign(STRAND='.')
    sub_pos = {n: i for i, n in enumerate(mk_trans_idx(1, 1))}
    elts, elt_id = np.unique(df.ELT.values.astype(str), return_inverse=True)
    order = np.argsort(elt_id, kind="stable")
    df = df.iloc[order]
    elt_id = elt_id[order]
    first = np.concatenate([[0], np.flatnonzero(np.diff(elt_id)) + 1])
    strand = _strand_code(df.STRAND.values[first])                       # list(group['STRAND'])[0]  (:679)
    minus = strand[elt_id] < 0
    ctx = df.CONTEXT.astype(str).values
    mut = df.MUT_TYPE.astype(str).values
    sub = np.full(len(df), -1, dtype=np.int32)
    for i, (m, c, neg) in enumerate(zip(mut, ctx, minus)):
        if 'nan' in c or len(c) != 3 or len(m) != 3:
            continue                                                      # 'nan' contexts are skipped (:700-701)
        name = c + '>' + c[0] + m[2] + c[2]
        if neg:                                                           # strand flips the substitution (:681-683)
            name = reverse_complement(c) + '>' + reverse_complement(c[0] + m[2] + c[2])
        sub[i] = sub_pos.get(name, -1)
    E = len(elts)
    L = kernels.site_counts(elt_id, sub, E, 192, device=torch.device("cuda", torch.cuda.current_device()))
    ptr = np.concatenate([first, [len(df)]]).astype(np.int64)
    chrom = df.CHROM.values[first].astype(np.int64)
    res = transfer_elements(region_model, win_counts64, d_pr, chrom, strand, ptr, df.START.values.astype(np.int64),
                            df.END.values.astype(np.int64), L_elt=L.to(torch.float64).reshape(E, 192, 1))
    Lh = L.cpu().numpy()
    elt_size = (Lh.sum(axis=1) / 3).astype(np.int64)
    with np.errstate(divide='ignore', invalid='ignore'):
        p_indel = elt_size / res["R_SIZE"].astype(np.float64)
    return pd.DataFrame({
        'ELT': elts, 'ELT_SIZE': elt_size, 'FLAG': res["FLAG"], 'R_SIZE': res["R_SIZE"], 'R_OBS': res["R_OBS"],
        'R_INDEL': res["R_OBS"], 'MU': res["MU"], 'SIGMA': res["SIGMA"], 'MU_INDEL': res["MU"],
        'SIGMA_INDEL': res["SIGMA"], 'P_SUM': res["P"][:, 0], 'P_INDEL': p_indel})


# --------------------------------------------------------------------------------------------
# tiled model (reference :599-720): every tile is transferred from the ONE window that contains its start
# --------------------------------------------------------------------------------------------

def _index_transform(s):
    """'chr1:100-200' -> 'region_1_100_200' (reference :721-725)."""
    chrom = int(s.split(":")[0].lstrip("chr"))
    start = int(s.split(":")[-1].split('-')[0])
    end = int(s.split(":")[-1].split('-')[1])
    return "region_{}_{}_{}".format(chrom, start, end)


def tiled_model_arrays(elt_lst, L_table, region_model, win_counts64, d_pr):
    """tiled_nonc_model (reference :599-690) on in-memory inputs, one K6 launch for all tiles.  elt_lst: tile names
    'chr{c}:{s}-{e}'; L_table: DataFrame indexed by those names with the 192 substitution columns.

    The reference locates the window as floor(start / 10000) * 10000 with the 10 000 hard-coded (:636) while the
    window LENGTH comes from region_params: for any other window size its lookup raises KeyError, and so does this."""
    elt_lst = list(elt_lst)
    if int(region_model.window) != 10000 and len(elt_lst):
        raise KeyError("tiled_nonc_model looks windows up on a 10000-base grid (genic_driver_tools.py:636); "
                       "region_params uses window {}".format(region_model.window))
    chrom = np.array([int(e.split(":")[0].lstrip("chr")) for e in elt_lst], dtype=np.int64)
    start = np.array([int(e.split(":")[1].split("-")[0]) for e in elt_lst], dtype=np.int64)
    E = len(elt_lst)
    L = L_table.loc[elt_lst].values.astype(np.float64).reshape(E, 192)
    # a 1-base block at the tile's start overlaps exactly the window that contains it
    res = transfer_elements(region_model, win_counts64, d_pr, chrom, np.ones(E, dtype=np.int8), np.arange(E + 1),
                            start, start + 1, L_elt=L.reshape(E, 192, 1))
    elt_size = (L.sum(axis=1) / 3).astype(np.int64)
    with np.errstate(divide='ignore', invalid='ignore'):
        p_indel = elt_size / res["R_SIZE"].astype(np.float64)
    return pd.DataFrame({
        'ELT': [_index_transform(e) for e in elt_lst], 'ELT_SIZE': elt_size, 'FLAG': res["FLAG"],
        'R_SIZE': res["R_SIZE"], 'R_OBS': res["R_OBS"], 'R_INDEL': res["R_OBS"], 'MU': res["MU"], 'SIGMA': res["SIGMA"],
        'MU_INDEL': res["MU"], 'SIGMA_INDEL': res["SIGMA"], 'P_SUM': res["P"][:, 0], 'P_INDEL': p_indel})


def tiled_model_parallel(f_pretrained, f_nonc_data, save_key, N_procs=1):
    """Reference :692-719 on the directory/HDF5 stores (L_counts written by DigPreprocess.py preprocess_tiled)."""
    pre = storage.Store(f_pretrained, "r")
    rm = RegionModel(pre.read_table('region_params'))
    d_pr = sequence_tools.d_pr_from_model192(pre.read_table('sequence_model_192'))
    data = storage.Store(f_nonc_data, "r")
    wkey = 'window_{}'.format(rm.window)
    win_idx = data.read_array('{}/full_window_si_index'.format(wkey))
    win_vals = data.read_array('{}/full_window_si_values'.format(wkey))
    key = {tuple(r): i for i, r in enumerate(map(tuple, win_idx))}
    rows = np.array([key[(int(c), int(s), int(e))] for c, s, e in zip(rm.df.CHROM, rm.df.START, rm.df.END)])
    L_table = data.read_table("{}/L_counts".format(save_key))
    return tiled_model_arrays(L_table.index, L_table, rm, win_vals[rows].astype(np.int32), d_pr)


# --------------------------------------------------------------------------------------------
# the reference's per-chunk workers (what its multiprocessing.Pool ran on a slice of the element list)
# --------------------------------------------------------------------------------------------

def genic_model(genes_lst, f_pretrained_str, f_genic_str, counts_key, indels_direct, f_fasta=None):
    """Reference :31-203: the genic pretrain rows of the genes in ``genes_lst`` (X / Y genes are dropped).  f_fasta
    is only needed when f_pretrained holds no 'window_counts_64' (see genic_model_parallel)."""
    return genic_model_parallel(f_pretrained_str, f_genic_str, counts_key=counts_key, indels_direct=indels_direct,
                                f_fasta=f_fasta, genes_lst=list(genes_lst))


def tiled_nonc_model(elt_lst, f_pretrained, f_nonc_data, save_key):
    """Reference :599-690: the tiled pretrain rows of the tiles in ``elt_lst`` ('chr{c}:{s}-{e}' names)."""
    pre = storage.Store(f_pretrained, "r")
    rm = RegionModel(pre.read_table('region_params'))
    d_pr = sequence_tools.d_pr_from_model192(pre.read_table('sequence_model_192'))
    data = storage.Store(f_nonc_data, "r")
    wkey = 'window_{}'.format(rm.window)
    win_idx = data.read_array('{}/full_window_si_index'.format(wkey))
    win_vals = data.read_array('{}/full_window_si_values'.format(wkey))
    key = {tuple(r): i for i, r in enumerate(map(tuple, win_idx))}
    rows = np.array([key[(int(c), int(s), int(e))] for c, s, e in zip(rm.df.CHROM, rm.df.START, rm.df.END)])
    return tiled_model_arrays(list(elt_lst), data.read_table("{}/L_counts".format(save_key)), rm,
                              win_vals[rows].astype(np.int32), d_pr)


def _region_params_of_overlaps(region_model, overlaps):
    """get_region_params_direct (reference :258-272) for many elements in one K6 launch: every overlapped window is
    handed to the kernel as a one-base block, so its window set is exactly the persisted overlap list."""
    E = len(overlaps)
    ptr = np.zeros(E + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(o) for o in overlaps])
    chrom = np.array([int(o[0][0]) if len(o) else 1 for o in overlaps], dtype=np.int64)
    starts = np.array([w[1] for o in overlaps for w in o], dtype=np.int64)
    n = len(region_model.df)
    return transfer_elements(region_model, np.zeros((n, 64), dtype=np.int32), np.ones(192), chrom,
                             np.ones(E, dtype=np.int8), ptr, starts, starts + 1, L_elt=np.zeros((E, 192, 1)))


def nonc_model(elt_lst, f_pretrained, f_nonc_data, save_key, indels_direct):
    """Reference :300-431: element pretrain rows from the intermediates persisted by sequence_tools.preprocess_nonc /
    preprocess_sites (L_counts, region_counts and the overlap list of every element)."""
    pre = storage.Store(f_pretrained, "r")
    rm = RegionModel(pre.read_table('region_params'))
    d_pr = sequence_tools.d_pr_from_model192(pre.read_table('sequence_model_192'))
    data = storage.Store(f_nonc_data, "r")
    names, L, R, overlaps = data.read_element_groups('window_{}/{}'.format(rm.window, save_key), list(elt_lst))
    dev = torch.device("cuda", torch.cuda.current_device())
    p_sum = kernels.element_psum(L, R, d_pr, dev).cpu().numpy()
    res = _region_params_of_overlaps(rm, overlaps)
    if indels_direct:
        res_ind = _region_params_of_overlaps(RegionModel(pre.read_table('region_params_indels')), overlaps)
    else:
        res_ind = res
    r_size = (R.sum(axis=1) / 3).astype(np.int64)             # int(region_counts.sum() / 3)
    elt_size = (L.sum(axis=1) / 3).astype(np.int64)           # int(np.sum(L) / 3)
    with np.errstate(divide='ignore', invalid='ignore'):
        p_indel = elt_size / r_size.astype(np.float64)
    return pd.DataFrame({
        'ELT': names, 'ELT_SIZE': elt_size, 'FLAG': res["FLAG"], 'R_SIZE': r_size, 'R_OBS': res["R_OBS"],
        'R_INDEL': res_ind["R_OBS"], 'MU': res["MU"], 'SIGMA': res["SIGMA"], 'MU_INDEL': res_ind["MU"],
        'SIGMA_INDEL': res_ind["SIGMA"], 'P_SUM': p_sum, 'P_INDEL': p_indel})


def nonc_model_region(df_nonc, f_pretrained, f_nonc_data, nonc_L_key, f_sites=None, return_intermediates=False):
    """Reference :518-597 (its own version unpacks three of get_region_params' four return values and cannot run;
    this one computes what it set out to): R_OBS, MU, SIGMA, P_SUM per row of a bed12_boundaries-style frame, with the
    block context counts read from the table ``nonc_L_key`` and the window counts from the store's root
    'full_window_si_index' / 'full_window_si_values'.  With return_intermediates also (L, t_pi) frames."""
    df_nonc = df_nonc.copy().astype({'CHROM': int, 'ELT': str, 'STRAND': str})
    pre = storage.Store(f_pretrained, "r")
    rm = RegionModel(pre.read_table('region_params'))
    df192 = pre.read_table('sequence_model_192')
    d_pr = sequence_tools.d_pr_from_model192(df192)
    data = storage.Store(f_nonc_data, "r")
    L_contexts = data.read_table(nonc_L_key)
    win_idx = data.read_array('full_window_si_index')
    win_vals = data.read_array('full_window_si_values')
    key = {tuple(r): i for i, r in enumerate(map(tuple, win_idx))}
    rows = np.array([key[(int(c), int(s), int(e))] for c, s, e in zip(rm.df.CHROM, rm.df.START, rm.df.END)])
    win_counts = win_vals[rows].astype(np.int32)
    out = nonc_model_arrays(df_nonc, L_contexts, rm, win_counts, d_pr)
    res = df_nonc.drop(['BLOCK_STARTS', 'BLOCK_ENDS'], axis=1)
    for c in ('R_OBS', 'MU', 'SIGMA', 'P_SUM'):
        res[c] = out[c].values
    if not return_intermediates:
        return res
    idx = sequence_tools.mk_trans_idx(1, 1)
    E = len(df_nonc)
    ptr = np.zeros(E + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(b) for b in df_nonc.BLOCK_STARTS])
    bs = np.array([x for b in df_nonc.BLOCK_STARTS for x in b], dtype=np.int64)
    be = np.array([x for b in df_nonc.BLOCK_ENDS for x in b], dtype=np.int64)
    owner = np.repeat(np.arange(E), np.diff(ptr))
    chrom = df_nonc.CHROM.values.astype(np.int64)
    keys = ['chr{}:{}-{}'.format(c, s, e) for c, s, e in zip(chrom[owner], bs, be)]
    L = np.zeros((E, 192), dtype=np.float64)
    np.add.at(L, owner, L_contexts.loc[keys].values)
    dev = torch.device("cuda", torch.cuda.current_device())
    rc, _ = kernels.element_region_counts(chrom.astype(np.int32), _strand_code(df_nonc.STRAND.values), ptr, bs, be,
                                          rm.window, rm.win_map_off, rm.win_map, win_counts, dev)
    _, denom = kernels.element_psum(L, np.repeat(rc.cpu().numpy(), 3, axis=1), d_pr, dev, want_denom=True)
    t_pi = d_pr[None, :] / denom.cpu().numpy()[:, None]
    return res, pd.DataFrame(L, columns=idx, index=df_nonc.ELT), pd.DataFrame(t_pi, columns=idx, index=df_nonc.ELT)


def nonc_model_region_parallel(f_bed, f_pretrained, f_nonc_data, nonc_L_key, N_procs, f_sites=None):
    """Reference :463-516: nonc_model_region over every autosomal row of a bed12 file."""
    from ..data_tools import mutation_tools
    print('Parsing regions bed file')
    df_nonc = mutation_tools.bed12_boundaries(f_bed)
    print('Pretraining model')
    return nonc_model_region(df_nonc, f_pretrained, f_nonc_data, nonc_L_key, f_sites)
