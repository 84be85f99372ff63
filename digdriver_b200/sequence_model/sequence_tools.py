"""Drop-in for DIGDriver/sequence_model/sequence_tools.py (context counting, mutation contexts,
sequence model, element preprocessing) with the per-base / per-mutation / per-element Python loops
replaced by the sm_100a kernels of libdigb200.so.

Same function names, argument meaning and return types (pandas objects with the reference's column and
index conventions).  ``f_fasta`` may be a FASTA path (plain or .gz), a ``genome.Genome`` or a
``genome.DeviceGenome``; the packed genome is cached per path, so the N_proc / N_chunk / n_procs
arguments are accepted and ignored (the reference used them for multiprocessing.Pool chunking).
There is no CPU fallback: without a GPU and libdigb200.so these functions raise.
"""
import itertools as it
import os

import numpy as np
import pandas as pd
import torch

from .. import kernels
from ..genome import DeviceGenome, Genome

DNA53 = 'NTCGA'
DNA35 = 'NAGCT'
trans = DNA53.maketrans(DNA53, DNA35)

_BASE_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
_GENOME_CACHE = {}


def default_device():
    return torch.device("cuda", torch.cuda.current_device())


def get_device_genome(f_fasta, device=None):
    """Resolve ``f_fasta`` (path | Genome | DeviceGenome) to a DeviceGenome, packing and caching on first use."""
    if isinstance(f_fasta, DeviceGenome):
        return f_fasta
    device = device or default_device()
    if isinstance(f_fasta, Genome):
        key = ("obj", id(f_fasta), str(device))
        if key not in _GENOME_CACHE:
            _GENOME_CACHE[key] = DeviceGenome.from_genome(f_fasta, device)
        return _GENOME_CACHE[key]
    path = os.path.abspath(str(f_fasta))
    key = (path, os.path.getmtime(path), str(device))
    if key not in _GENOME_CACHE:
        _GENOME_CACHE.clear()                       # one packed genome at a time is plenty
        _GENOME_CACHE[key] = DeviceGenome.from_genome(Genome.from_fasta(path), device)
    return _GENOME_CACHE[key]


def reverse_complement(seq):
    return seq[::-1].translate(trans)


def mk_context_sequences(n_up=2, n_down=2, collapse=False):
    """Ordered {k-mer: 0} (reference :31-40): lexicographic over ACGT, 5' base most significant --
    which is also the kernels' bin index."""
    DNA = 'ACGT'
    NUC = 'CT' if collapse else 'ACGT'
    prod_items = [DNA] * n_up + [NUC] + [DNA] * n_down
    return {''.join(tup): 0 for tup in it.product(*prod_items)}


def seq_to_context(seq, baseix=2, collapse=False):
    if 'N' in seq:
        return ''
    if collapse and seq[baseix] in 'GA':
        return reverse_complement(seq)
    return seq


def type_mutation(REF, ALT, collapse=False):
    if collapse and REF in ('G', 'A'):
        REF = REF.translate(trans)
        ALT = ALT.translate(trans)
    return "{}>{}".format(REF, ALT)


def _collapse_counts(counts, n_up, n_down):
    """Fold purine-centred k-mers onto their reverse complement (reference :51-53)."""
    full = list(mk_context_sequences(n_up, n_down))
    keep = list(mk_context_sequences(n_up, n_down, collapse=True))
    pos = {k: i for i, k in enumerate(full)}
    a = np.array([pos[k] for k in keep])
    b = np.array([pos[reverse_complement(k)] for k in keep])
    return counts[:, a] + counts[:, b], keep


def _count_regions(genome, chrom_idx, starts, ends, n_up, n_down, strand=None, collapse=False):
    starts = np.asarray(starts, dtype=np.int64)
    ends = np.asarray(ends, dtype=np.int64)
    if np.any((starts > 0) & (starts < n_up)):
        # pysam raises inside fetch() for a negative start (reference :28)
        raise ValueError("start out of range (%d)" % int((starts[(starts > 0) & (starts < n_up)] - n_up)[0]))
    if genome.n_other:
        raise KeyError("genome contains %d characters that are not A/C/G/T/N; the reference fails on them "
                       "at sequence_tools.py:76" % genome.n_other)
    counts, _ = kernels.count_contexts(genome, chrom_idx, starts, ends, n_up, n_down, strand=strand)
    counts = counts.cpu().numpy().astype(np.int64)
    cols = list(mk_context_sequences(n_up, n_down))
    if collapse:
        counts, cols = _collapse_counts(counts, n_up, n_down)
    return counts, cols


def count_sequence_context(seq, n_up=2, n_down=2, nuc_dict=None, collapse=False):
    """Count the nucleotide contexts present in a sequence string (reference :65-78)."""
    g = DeviceGenome.from_genome(Genome.from_dict({"s": seq.upper()}), default_device())
    L = len(seq)
    # the reference walks i in [n_up, len-n_down): region [n_up, L-n_down) of a "chromosome" of length L
    counts, cols = _count_regions(g, [0], [n_up], [max(L - n_down, n_up)], n_up, n_down, collapse=collapse)
    out = dict(nuc_dict) if nuc_dict else {k: 0 for k in cols}
    for k, v in zip(cols, counts[0]):
        out[k] = out.get(k, 0) + int(v)
    return out


def count_contexts_by_regions(f_fasta, chrom_lst, start_lst, end_lst, n_up=2, n_down=2, collapse=False):
    """Sequence context counts within a set of regions (reference :80-94).  ``chrom_lst`` holds FASTA
    sequence names ('chr1', ...)."""
    g = get_device_genome(f_fasta)
    chrom_lst = list(chrom_lst)
    cidx = g.chrom_indices(chrom_lst, prefix="")
    counts, cols = _count_regions(g, cidx, start_lst, end_lst, n_up, n_down, collapse=collapse)
    idx = ["{}:{}-{}".format(c, s, e) for c, s, e in zip(chrom_lst, start_lst, end_lst)]
    return pd.DataFrame(counts, index=idx, columns=cols)


def count_contexts_in_bed(f_fasta, df_bed, n_up=1, n_down=1, N_proc=1, N_chunk=10, collapse=False):
    """Count nucleotide contexts within regions of a bed-like dataframe (reference :96-128)."""
    chrom_lst = ['chr{}'.format(val) for val in df_bed.iloc[:, 0].values]
    return count_contexts_by_regions(f_fasta, chrom_lst, df_bed.iloc[:, 1].values, df_bed.iloc[:, 2].values,
                                     n_up=n_up, n_down=n_down, collapse=collapse)


def genome_context_totals(df_counts):
    """S_count = df.sum(axis=0) of DigPreprocess.py:59 (kept for API symmetry; the CLI uses the fused totals)."""
    return df_counts.sum(axis=0)


def _encode_bases(values):
    return np.array([_BASE_CODE.get(v, 255) if isinstance(v, str) else 255 for v in values], dtype=np.uint8)


def mutation_contexts_by_chrom(f_fasta, df, n_up=2, n_down=2, collapse=False):
    """Reference :130-178 for the rows of ONE chromosome: appends MUT_TYPE and CONTEXT and drops the rows the
    reference drops (REF mismatch, N context, inherited drop inside a same-START run)."""
    g = get_device_genome(f_fasta)
    CHROM = str(df.CHROM.iloc[0])
    if not CHROM.startswith('chr'):
        CHROM = "chr{}".format(CHROM)
    cidx = np.full(len(df), g.index(CHROM), dtype=np.int32)
    ref = _encode_bases(df.REF.values)
    ctx = kernels.mutation_contexts(g, cidx, df.START.values.astype(np.int64), ref, n_up, n_down).cpu().numpy()
    names = np.array(list(mk_context_sequences(n_up, n_down)) + [""])
    context = names[np.where(ctx >= 0, ctx, len(names) - 1)]
    if collapse:
        context = np.array([seq_to_context(c, baseix=n_up, collapse=True) if c else c for c in context])
    df = df.copy()
    df.insert(df.shape[1], 'MUT_TYPE', [type_mutation(r, a, collapse=collapse) for r, a in zip(df.REF.values, df.ALT.values)])
    df.insert(df.shape[1], 'CONTEXT', context)
    return df[df.CONTEXT != ""]


def add_context_to_mutations(f_fasta, df_mut, n_up=2, n_down=2, N_proc=1, collapse=False):
    """Add sequence context annotations to mutations (reference :180-222)."""
    df_indel = df_mut[df_mut.ANNOT.str.contains('INDEL')]
    df_mut = df_mut[~df_mut.ANNOT.str.contains('INDEL')]
    if len(df_mut) > 0:
        df_lst = []
        for chrom, df in df_mut.groupby('CHROM'):
            if 'MT' in str(chrom):
                continue
            df_lst.append(mutation_contexts_by_chrom(f_fasta, df, n_up=n_up, n_down=n_down, collapse=collapse))
        df_out = pd.concat(df_lst)
    if len(df_indel) > 0:
        df_indel = df_indel.rename({'ANNOT': 'MUT_TYPE'}, axis=1)
        df_indel.insert(df_indel.shape[1] - 1, 'ANNOT', 'INDEL')
        df_indel.insert(df_indel.shape[1], 'CONTEXT', '.')
        if len(df_mut) > 0:
            df_out = pd.concat([df_out, df_indel]).sort_values(['CHROM', 'START', 'END'])
        else:
            df_out = df_indel.sort_values(['CHROM', 'START', 'END'])
    return df_out


def mk_mutation_context(n_up=1, n_down=1, collapse=False, return_df=False):
    """(MUT_TYPE, CONTEXT) rows in the reference's order (:232-278)."""
    DNA = 'ACGT'

    def keys(centre):
        return [''.join(t) for t in it.product(*([DNA] * n_up + [centre] + [DNA] * n_down))]

    muts = {'A': ['A>T', 'A>C', 'A>G'], 'C': ['C>A', 'C>G', 'C>T'], 'G': ['G>T', 'G>C', 'G>A'],
            'T': ['T>A', 'T>G', 'T>C']}
    order = 'CT' if collapse else 'ACGT'
    tups = []
    for b in order:
        tups += [tup for tup in it.product(muts[b], keys(b))]
    if return_df:
        return pd.DataFrame(tups, columns=['MUT_TYPE', 'CONTEXT'])
    return {tup: 0 for tup in tups}


def mk_trans_idx(n_up=1, n_down=1, collapse=False):
    """All substitutions 'ATG>AGG', sorted (reference :282-289) -- the kernels' 192-bin substitution order."""
    d = mk_mutation_context(n_up=n_up, n_down=n_down, collapse=collapse)
    return sorted([k[1] + '>' + k[1][:n_up] + k[0][2] + k[1][n_up + 1:] for k in d.keys()])


def mutation_freq_conditional(df_freq, S_gen):
    """FREQ = COUNT / S_gen[CONTEXT] (reference :356-373)."""
    df_freq["FREQ"] = df_freq.COUNT.values / np.array([S_gen[c] for c in df_freq.CONTEXT.values], dtype=np.float64)
    return df_freq


def restrict_mutations_to_regions(df_mut, regions):
    """bedtools-intersect whitelist of train_sequence_model (reference :329 -> mutation_tools.py:8-30 with
    unique=True): keep mutation rows overlapping any region, drop exact duplicate rows."""
    regs = np.asarray(regions)
    if len(regs) == 0:
        return df_mut.iloc[:0]
    rk_s = (regs[:, 0].astype(np.int64) << 32) | regs[:, 1].astype(np.int64)
    rk_e = (regs[:, 0].astype(np.int64) << 32) | regs[:, 2].astype(np.int64)
    mk_s = (df_mut.CHROM.values.astype(np.int64) << 32) | df_mut.START.values.astype(np.int64)
    mk_e = (df_mut.CHROM.values.astype(np.int64) << 32) | df_mut.END.values.astype(np.int64)
    hit = kernels.overlap_counts(rk_s, rk_e, mk_s, mk_e, default_device()) > 0
    return df_mut[hit].drop_duplicates()


def train_sequence_model(regions, df_mut, genome_counts, n_up=1, n_down=1, key_prefix=None):
    """Sequence model from precalculated context frequencies and mutation counts (reference :321-354).
    Returns (df_freq_mut [MUT_TYPE, CONTEXT, COUNT, FREQ], df_freq_context [FREQ per context])."""
    df_white = restrict_mutations_to_regions(df_mut, regions)
    df_ct = mk_mutation_context(n_up=n_up, n_down=n_down, collapse=False, return_df=True)
    names = list(mk_context_sequences(n_up, n_down))
    cpos = {k: i for i, k in enumerate(names)}
    # device histogram over (context, alt) from the file's own MUT_TYPE / CONTEXT columns
    ctx = np.array([cpos.get(c, -1) for c in df_white.CONTEXT.values], dtype=np.int32)
    alt = _encode_bases([m[2] if isinstance(m, str) and len(m) == 3 and m[1] == '>' else None
                         for m in df_white.MUT_TYPE.values])
    refc = _encode_bases([m[0] if isinstance(m, str) and len(m) == 3 and m[1] == '>' else None
                          for m in df_white.MUT_TYPE.values])
    centre = np.where(ctx >= 0, (ctx >> (2 * n_down)) & 3, 255)
    ctx = np.where(refc == centre, ctx, -1).astype(np.int32)      # rows whose MUT_TYPE does not fit the context
    dev = default_device()
    hist = kernels.substitution_counts(torch.from_numpy(ctx).to(dev), alt, n_up, n_down).cpu().numpy()
    # hist bin = 3*ctx + rank(alt among non-ref, alphabetical)  ->  reference row order
    rank = {}
    for b in 'ACGT':
        for r, a in enumerate([x for x in 'ACGT' if x != b]):
            rank[(b, a)] = r
    count = np.array([hist[3 * cpos[c] + rank[(m[0], m[2])]] for m, c in zip(df_ct.MUT_TYPE, df_ct.CONTEXT)],
                     dtype=np.float64)
    df_ct['COUNT'] = count
    df_freq_mut = mutation_freq_conditional(df_ct, genome_counts)
    df_freq_context = df_freq_mut.pivot_table('FREQ', index=['CONTEXT'], aggfunc="sum")
    return df_freq_mut, df_freq_context


def d_pr_from_model192(df_model_192):
    """FREQ re-ordered to sorted substitution names (genic_driver_tools.py:321-325)."""
    idx = [r[1] + '>' + r[1][0] + r[0][2] + r[1][2] for r in zip(df_model_192.MUT_TYPE, df_model_192.CONTEXT)]
    return pd.DataFrame(df_model_192.FREQ.values, idx).sort_index()[0].values


# --------------------------------------------------------------------------------------------
# element preprocessing
# --------------------------------------------------------------------------------------------

def _bed6_blocks(f_nonc_bed):
    """bed12 -> bed6 blocks (pybedtools .bed6() of the reference :495-498), 'chr' prefix stripped."""
    df = pd.read_table(f_nonc_bed, header=None, low_memory=False)
    rows = []
    for r in df.itertuples(index=False):
        chrom = str(r[0])
        if chrom.startswith('chr'):
            chrom = chrom[3:]
        sizes = [int(x) for x in str(r[10]).strip(',').split(',')]
        starts = [int(x) for x in str(r[11]).strip(',').split(',')]
        for sz, st in zip(sizes, starts):
            rows.append((chrom, int(r[1]) + st, int(r[1]) + st + sz, r[3], r[4], r[5]))
    return pd.DataFrame(rows, columns=['CHROM', 'START', 'END', 'ELT', 'SCORE', 'STRAND'])


def nonc_elt_context_count(regions, trans_idx, f_fasta, n_up=1, n_down=1):
    """Strand-aware context counts of each region expanded to the 192 substitutions (reference :527-566).
    regions: iterable of (chrom, start, end, strand)."""
    g = get_device_genome(f_fasta)
    regions = list(regions)
    chroms = ['chr' + str(r[0]) for r in regions]
    starts = [r[1] for r in regions]
    ends = [r[2] for r in regions]
    strand = np.array([-1 if (r[3] == '-' or r[3] == -1) else 1 for r in regions], dtype=np.int8)
    counts, cols = _count_regions(g, g.chrom_indices(chroms, prefix=""), starts, ends, n_up, n_down, strand=strand)
    keys = sorted(set(trans_idx))
    col_of = np.array([cols.index(k.split('>')[0]) for k in keys])
    idx = ["{}:{}-{}".format(c, s, e) for c, s, e in zip(chroms, starts, ends)]
    return pd.DataFrame(counts[:, col_of].astype(np.float64), columns=keys, index=idx)


def precount_region_contexts_parallel(f_nonc_bed, f_fasta, n_procs, window, sub_elts=True, n_up=1, n_down=1):
    """Context counts within each sub-element of a bed12 file (reference :481-525); duplicated index rows
    are dropped keeping the first, as in the reference."""
    trans_idx = mk_trans_idx(n_up=1, n_down=1, collapse=False)
    if sub_elts:
        df6 = _bed6_blocks(f_nonc_bed)
        all_regions = list(zip(df6.CHROM, df6.START, df6.END, df6.STRAND))
    else:
        bed = pd.read_csv(f_nonc_bed, sep='\t', header=None, names=None, low_memory=False)
        chrom = bed[0].astype(str).map(lambda x: x[3:] if x.startswith('chr') else x)
        all_regions = list(zip(chrom, bed[1], bed[2], bed[5]))
    results = nonc_elt_context_count(all_regions, trans_idx, f_fasta)
    return results.loc[~results.index.duplicated()]


def _s_prob_table(S_prob, n_up, n_down, collapse=False):
    """Dense [4^k] probability table in k-mer index order from the reference's {k-mer: probability} mapping
    (dict or Series).  With collapse=True purine-centred k-mers take the value of their reverse complement
    (seq_to_context, reference :42-55)."""
    full = list(mk_context_sequences(n_up, n_down, collapse=False))
    get = S_prob.__getitem__
    out = np.empty(len(full), dtype=np.float64)
    for i, kmer in enumerate(full):
        key = reverse_complement(kmer) if (collapse and kmer[n_up] in 'GA') else kmer
        out[i] = float(get(key))
    return out


def base_probabilities_by_region(fasta, S_prob, CHROM, START, END, n_up=2, n_down=2, normed=True, collapse=False):
    """Probability of mutation at every position across a region (reference :292-317).  ``fasta`` is a FASTA path,
    Genome or DeviceGenome; returns (probs, positions) as numpy arrays."""
    g = get_device_genome(fasta)
    if 0 < START < n_up:
        raise ValueError("start out of range (%d)" % (START - n_up))
    cidx = g.chrom_indices([CHROM], prefix="")
    out = kernels.position_test(g, cidx, [START], [END], [1.0], [1.0], _s_prob_table(S_prob, n_up, n_down, collapse),
                                np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64), n_up=n_up, n_down=n_down,
                                binsize=1, normed=normed, want=("pt", "pos"))
    return out["pt"].cpu().numpy(), out["pos"].cpu().numpy().astype(np.int64)
