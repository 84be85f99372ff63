"""Array-level API of the whole hot path: genome scan -> sequence model -> gene/element transfer ->
observed counts -> NB burden test.  The DataFrame / file level functions in sequence_model/,
driver_model/ and data_tools/ are thin shims over these.

Every stage cites the reference code it replaces; all heavy work is done by the sm_100a kernels of
libdigb200.so (see kernels.py).  torch is used for allocation and for O(E) glue arithmetic on the
per-gene table (20 k rows), never for the scan, the counting or the p-values.
"""
from dataclasses import dataclass, field

import numpy as np
import torch

from . import kernels

GENE_CLASSES = ("SYN", "MIS", "NONS", "SPL", "TRUNC", "NONSYN")
ANNOT_CLASS = {"Synonymous": 0, "Missense": 1, "Nonsense": 2, "Essential_Splice": 3, "INDEL": 4}


def scan_windows(genome, win_chrom_idx, win_start, win_end, n_up, n_down, out=None, totals=None, stream=None):
    """DigPreprocess.py countGenomeContext (:19-73): per-window counts + genome-wide totals."""
    if totals is not None:
        totals.zero_()
    return kernels.count_contexts(genome, win_chrom_idx, win_start, win_end, n_up, n_down, want_totals=True,
                                  out=out, totals=totals, stream=stream)


def sequence_model(genome, mut_chrom_idx, mut_start, mut_ref, mut_alt, genome_totals64, n_up=1, n_down=1):
    """train_sequence_model + mutation_freq_conditional (sequence_tools.py:321-373): FREQ of each of the
    3K substitutions = COUNT / S_genome[context].  Returned in sorted 'CTX>CTX2' order (the d_pr order of
    genic_driver_tools.py:321-325) together with the raw counts."""
    ctx = kernels.mutation_contexts(genome, mut_chrom_idx, mut_start, mut_ref, n_up, n_down)
    counts = kernels.substitution_counts(ctx, mut_alt, n_up, n_down)
    denom = genome_totals64.to(torch.float64).repeat_interleave(3)
    return counts.to(torch.float64) / denom, counts, ctx


@dataclass
class GeneTable:
    """Gene annotation in CSR form (f_genic: cds_intervals / chr / strands / L_data)."""
    chrom_idx: np.ndarray          # [E] index into the genome / window map
    strand: np.ndarray             # [E] int8
    blk_ptr: np.ndarray            # [E+1]
    blk_start: np.ndarray          # inclusive CDS interval starts
    blk_end: np.ndarray            # inclusive CDS interval ends
    L: np.ndarray                  # [E, 192, 4] silent, mis, nons, splice
    names: list = field(default_factory=list)
    tp53: int = -1                 # row of TP53 (excluded from the scale factor), -1 if absent
    cgc_mask: np.ndarray = None    # [E] bool, True for CGC genes (excluded from the indel scale factor)


def gene_pretrain(genes, window, win_map_off, win_map, win_counts64, y_pred, std, y_true, flag, d_pr,
                  device="cuda:0", cache=None):
    """genic_model (genic_driver_tools.py:31-203): MU, SIGMA, R_OBS, FLAG, R_SIZE, GENE_LENGTH and
    P_SILENT/P_MIS/P_NONS/P_SPLICE per gene.  Returns device tensors."""
    out = kernels.element_transfer(genes.chrom_idx, genes.strand, genes.blk_ptr, genes.blk_start, genes.blk_end,
                                   window, win_map_off, win_map, win_counts64, y_pred, std, y_true, flag, d_pr,
                                   L_elt=genes.L if cache is None else cache, device=device)
    return out


_CONST = {}


def _const_f64(value, device):
    """A cached one-element float64 device tensor (a torch.full per step is one more node in a captured stage)."""
    key = (float(value), str(device))
    t = _CONST.get(key)
    if t is None:
        t = _CONST[key] = torch.full((1,), float(value), dtype=torch.float64, device=device)
    return t


def gene_burden_test(pre, obs, nsamp, n_syn_non_tp53, tp53=-1, cgc_mask=None, scale_factor=None, cohort=0,
                     collectives=None):
    """run_gene_model's arithmetic (transfer_tools.py:809-861) fused on the device: one reduction kernel for the
    scale-factor sums, one kernel for the 13 NB tests of every gene, one for the Fisher combine.

    pre: output of gene_pretrain; obs [E,5] / nsamp [E,7] from kernels.tabulate_genes.  With `collectives`
    (sharding.Collectives) the sums and n_syn are all-reduced so every shard uses the cohort-wide scale factors.
    Returns a dict of [E] float64 device tensors with the reference's column names."""
    mu, sigma = pre["MU"][cohort].contiguous(), pre["SIGMA"][cohort].contiguous()
    P = pre["P"][cohort].contiguous()                        # silent, mis, nons, splice
    pi_indel = kernels.size_ratio(pre["ELT_SIZE"], pre["R_SIZE"])                     # genic_driver_tools.py:158-159
    sums = kernels.gene_scale_sums(mu, sigma, P, pi_indel, obs, cgc_mask, tp53)
    n_syn = float(n_syn_non_tp53)
    if collectives is not None and collectives.world > 1:
        # sums[3] carries this shard's synonymous count; after the all-reduce the kernel reads the cohort-wide value
        # from the device (no host read: the whole stage stays stream-ordered and CUDA-graph capturable)
        sums = torch.cat([sums, _const_f64(n_syn, sums.device)])
        collectives.all_reduce_sum(sums)
        n_syn = None
    out = kernels.gene_burden_test(mu, sigma, P, pi_indel, obs, nsamp, sums, n_syn, scale_factor)
    res = {name: out[i] for i, name in enumerate(kernels.GENE_OUT_ROWS)}
    o = obs.to(torch.float64)
    res.update({"MU": mu, "SIGMA": sigma, "Pi_SYN": P[:, 0], "Pi_MIS": P[:, 1], "Pi_NONS": P[:, 2], "Pi_SPL": P[:, 3],
                "OBS_SYN": o[:, 0], "OBS_MIS": o[:, 1], "OBS_NONS": o[:, 2], "OBS_SPL": o[:, 3], "OBS_INDEL": o[:, 4],
                "SUMS": sums})
    return res


def element_burden_test(pre, obs, cj, cj_indel, cohort=0, skip_pvals=False):
    """run_element_region_model's arithmetic (transfer_tools.py:1070-1087) on device tensors.
    pre: output of kernels.element_transfer (blk_counts mode); obs [E,3] from kernels.tabulate_elements."""
    dev = obs.device
    E = obs.shape[0]
    mu, sigma = pre["MU"][cohort], pre["SIGMA"][cohort]
    pi_sum = pre["P"][cohort][:, 0]
    pi_indel = pre["ELT_SIZE"].to(torch.float64) / pre["R_SIZE"].to(torch.float64)
    alpha = mu ** 2 / sigma ** 2
    theta0 = sigma ** 2 / mu
    theta = theta0 * cj
    theta_indel = theta0 * cj_indel
    o = obs.to(torch.float64)
    res = {"ELT_SIZE": pre["ELT_SIZE"], "FLAG": pre["FLAG"][cohort], "R_SIZE": pre["R_SIZE"],
           "R_OBS": pre["R_OBS"][cohort], "R_INDEL": pre["R_OBS"][cohort], "MU": mu, "SIGMA": sigma,
           "ALPHA": alpha, "THETA": theta, "MU_INDEL": mu, "SIGMA_INDEL": sigma, "ALPHA_INDEL": alpha,
           "THETA_INDEL": theta_indel, "Pi_SUM": pi_sum, "Pi_INDEL": pi_indel,
           "OBS_SAMPLES": o[:, 0], "OBS_SNV": o[:, 1], "OBS_INDEL": o[:, 2]}
    if skip_pvals:
        res["EXP_SNV"] = alpha * theta * pi_sum
        return res
    kk = torch.stack([o[:, 1], o[:, 0], o[:, 2]])
    pp = torch.stack([pi_sum, pi_sum, pi_indel])
    tt = torch.stack([theta, theta, theta_indel])
    exp, pval = kernels.nb_burden_test(kk.reshape(-1), alpha.expand(3, E).reshape(-1), tt.reshape(-1),
                                       pp.reshape(-1), dev)
    exp, pval = exp.reshape(3, E), pval.reshape(3, E)
    res["EXP_SNV"] = exp[0]
    res["PVAL_SNV_BURDEN"] = pval[0]
    res["PVAL_SAMPLE_BURDEN"] = pval[1]
    if float(o[:, 2].sum()) != 0:                             # transfer_tools.py:1081
        res["EXP_INDEL"] = exp[2]
        res["PVAL_INDEL_BURDEN"] = pval[2]
        res["PVAL_MUT_BURDEN"] = kernels.fisher_combine2(pval[0], pval[2], dev)
    return res


# --------------------------------------------------------------------------------------------
# synthetic workloads (BASELINE.json configs): deterministic, shaped like the real inputs
# --------------------------------------------------------------------------------------------

def synth_region_params(n_win, seed):
    """Y_PRED ~ Gamma(2,10), STD = Y_PRED*U(.05,.5), Y_TRUE ~ Poisson, FLAG ~ B(.1)."""
    rng = np.random.default_rng(seed)
    y_pred = rng.gamma(2.0, 10.0, n_win)
    std = y_pred * rng.uniform(0.05, 0.5, n_win)
    y_true = rng.poisson(y_pred).astype(np.float64)
    flag = rng.random(n_win) < 0.1
    return y_pred, std, y_true, flag


def synth_genes(n_genes, chrom_lengths, window, seed):
    """Multi-exon genes with the block statistics of the bundled CDS annotation (mean 9.7 blocks,
    ~170 bp exons, kb-scale introns), placed inside the tiled part of each chromosome."""
    rng = np.random.default_rng(seed)
    chrom_lengths = np.asarray(chrom_lengths, dtype=np.int64)
    usable = np.maximum((chrom_lengths - 1) // window * window - 1, 0)      # stay inside tiled windows
    p = usable / usable.sum()
    chrom = np.sort(rng.choice(len(chrom_lengths), size=n_genes, p=p)).astype(np.int32)
    nblk = np.minimum(1 + rng.geometric(1 / 8.7, n_genes), 300)
    ptr = np.zeros(n_genes + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(nblk)
    tot = int(ptr[-1])
    exon = np.maximum(rng.lognormal(np.log(130.0), 0.7, tot).astype(np.int64), 3)
    intron = rng.lognormal(np.log(1500.0), 1.2, tot).astype(np.int64) + 20
    step = exon + intron
    owner = np.repeat(np.arange(n_genes), nblk)
    first = ptr[:-1][owner]
    cs = np.cumsum(step) - step
    rel = cs - cs[first]                                  # block start relative to gene start
    span = np.zeros(n_genes, dtype=np.int64)
    np.maximum.at(span, owner, rel + exon)
    g0 = (rng.random(n_genes) * np.maximum(usable[chrom] - span - 10, 1)).astype(np.int64) + 5
    start = g0[owner] + rel
    end = start + exon - 1                                # inclusive CDS interval ends
    lim = usable[chrom][owner] - 1
    start = np.minimum(start, lim - 1)
    end = np.minimum(np.maximum(end, start), lim)
    strand = np.where(rng.random(n_genes) < 0.5, -1, 1).astype(np.int8)
    return chrom, strand, ptr, start.astype(np.int64), end.astype(np.int64)


def synth_gene_L(blk_counts64, blk_ptr, seed):
    """A consistent L[E,192,4]: the CDS trinucleotide content of each gene (K4 counts, x3 substitutions)
    split into silent / missense / nonsense / splice site counts with fixed per-substitution fractions."""
    rng = np.random.default_rng(seed)
    E = len(blk_ptr) - 1
    owner = np.repeat(np.arange(E), np.diff(blk_ptr))
    L64 = np.zeros((E, 64), dtype=np.int64)
    np.add.at(L64, owner, np.asarray(blk_counts64, dtype=np.int64))
    L192 = np.repeat(L64, 3, axis=1).astype(np.float64)
    frac = rng.dirichlet([2.3, 6.8, 0.4, 0.5], size=192)              # per substitution
    L = np.floor(L192[:, :, None] * frac[None, :, :])
    return L
