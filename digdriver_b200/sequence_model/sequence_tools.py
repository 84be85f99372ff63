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
        # pinned host copy (the packed cache next to the FASTA when it is valid, else the parsed file), uploaded and
        # packed chromosome by chromosome on two streams; a freshly packed genome is written back as the cache
        from .. import host_pipeline
        use_cache = os.environ.get("DIG_NO_GENOME_CACHE", "") == ""
        hg, hit = host_pipeline.host_genome_from_fasta(path, use_cache=use_cache)
        g = host_pipeline.HostScan(hg, np.zeros((0, 3), dtype=np.int64), device, tables=(1, 1)).run().genome
        if use_cache and not hit:
            host_pipeline.PackedGenomeCache.store(path, g)
        _GENOME_CACHE[key] = g
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


# --------------------------------------------------------------------------------------------
# remaining functions of the reference module: sequence fetch, 192-substitution counts over regions (si_*),
# and the persisted element / site-set intermediates (initialize_nonc_data, preprocess_nonc, preprocess_sites)
# --------------------------------------------------------------------------------------------

def fetch_sequence(fasta, CHROM, START, END, n_up=2, n_down=2):
    """A sequence expanded by the context size on either end (reference :21-29): (seq, START-n_up, END+n_down), with
    START == 0 silently becoming n_up.  ``fasta`` is a FASTA path or a genome.Genome (host strings; the kernels work
    on the packed genome and never call this)."""
    g = fasta if isinstance(fasta, Genome) else Genome.from_fasta(str(fasta))
    if START == 0:
        START = n_up
    if START - n_up < 0:
        raise ValueError("start out of range (%d)" % (START - n_up))
    return g.fetch(CHROM, START - n_up, END + n_down).upper(), START - n_up, END + n_down


def _parse_region_str(r):
    left, end = r.rsplit('-', 1)
    chrom, start = left.rsplit(':', 1)
    return chrom, int(start), int(end)


def si_by_regions(fasta, trans_idx, regions, strand=1, n_up=1, n_down=1, normed=True):
    """192-substitution context counts summed over a list of 'chr1:100-200' regions (reference :398-431): every
    counted k-mer adds one to each of its three substitutions; minus strand counts the reverse complement.  Returns a
    one-column DataFrame indexed by ``trans_idx`` (the reference's row order is that of a Python set)."""
    g = get_device_genome(fasta)
    keys = list(trans_idx)
    if len(regions) == 0:
        return pd.DataFrame(np.zeros(len(keys), dtype=np.int64), index=keys)
    parsed = [_parse_region_str(r) for r in regions]
    minus = strand == -1 or strand == '-'
    st = np.full(len(parsed), -1 if minus else 1, dtype=np.int8)
    counts, cols = _count_regions(g, g.chrom_indices([p[0] for p in parsed], prefix=""), [p[1] for p in parsed],
                                  [p[2] for p in parsed], n_up, n_down, strand=st)
    tot = counts.sum(axis=0)
    col_of = {c: i for i, c in enumerate(cols)}
    return pd.DataFrame(np.array([tot[col_of[k.split('>')[0]]] for k in keys], dtype=np.int64), index=keys)


def si_count_pretrain(gene_lst, f_genic_str, f_fasta, window):
    """192-substitution counts of the windows each gene overlaps (reference :375-396), one row per gene."""
    from . import genic_driver_tools
    from .. import storage
    st = storage.Store(f_genic_str, "r")
    trans_idx = (sorted(np.asarray(st.read_array('substitution_idx')).astype(str)) if st.has('substitution_idx')
                 else mk_trans_idx(1, 1))
    flat = None
    if st.has('genes') and st.has('cds_ptr'):        # this package's flattened f_genic (genic_model_parallel)
        meta = st.read_table('genes')
        flat = ({str(g): i for i, g in enumerate(meta.GENE.values)}, meta, st.read_array('cds_ptr'),
                st.read_array('cds_start'), st.read_array('cds_end'))
    rows = {}
    for gene in gene_lst:
        if flat is not None:
            i = flat[0][str(gene)]
            chrom, strd = str(flat[1].CHROM.values[i]), flat[1].STRAND.values[i]
            intervals = np.vstack((flat[3][flat[2][i]:flat[2][i + 1]], flat[4][flat[2][i]:flat[2][i + 1]]))
        else:
            chrom = st.read_array('chr/{}'.format(gene))[0]
            chrom = chrom.decode("utf-8") if isinstance(chrom, bytes) else str(chrom)
            strd = st.read_array('strands/{}'.format(gene))[0]
            intervals = st.read_array('cds_intervals/{}'.format(gene))
        regions = [genic_driver_tools.trip_to_str(r) for r in genic_driver_tools.get_ideal_overlaps(chrom, intervals, window)]
        rows[gene] = si_by_regions(f_fasta, trans_idx, regions, strand=strd)[0].values
    return pd.DataFrame.from_dict(rows, orient='index', columns=trans_idx)


def si_count_parallel(f_genic_str, f_fasta, window, n_procs):
    """Reference :434-449; the process pool is replaced by one pass over all genes."""
    from .. import storage
    st = storage.Store(f_genic_str, "r")
    genes = list(st.read_table('genes').GENE.values) if st.has('genes') else st.keys('cds_intervals')
    return si_count_pretrain(genes, f_genic_str, f_fasta, window)


def initialize_nonc_data(f_nonc_data_str, f_genome_counts, window, n_up=1, n_down=1):
    """Copy the substitution index and the per-window trinucleotide counts into the element data store
    (reference :451-478)."""
    from .. import storage
    window_key = 'window_{}'.format(window)
    dst = storage.Store(f_nonc_data_str, "a")
    if not dst.has('substitution_idx'):
        dst.write_array('substitution_idx', np.array(mk_trans_idx(n_up=n_up, n_down=n_down, collapse=False)))
    if not (dst.has(window_key + '/full_window_si_index') and dst.has(window_key + '/full_window_si_values')):
        src = storage.Store(f_genome_counts, "r")
        idx = src.read_array('idx')
        genome_df = src.read_table('all_window_genome_counts')
        assert int(str(genome_df.index[0]).split('-')[-1]) == window      # correct genome count window size
        dst.write_array(window_key + '/full_window_si_values', genome_df.values, dtype=np.int64)
        dst.write_array(window_key + '/full_window_si_index', idx)


def _strand_is_minus(s):
    return s == '-1' or s == '-' or s == -1


def _stored_window_map(store, window_key):
    """Window index / counts of the element data store as K6 inputs (chromosome number -> dense window map)."""
    idx = np.asarray(store.read_array(window_key + '/full_window_si_index')).astype(np.int64)
    vals = np.asarray(store.read_array(window_key + '/full_window_si_values'))
    window = int(window_key.split('_')[1])
    n_chrom = int(idx[:, 0].max()) + 1 if len(idx) else 1
    off, wmap = kernels.build_window_map(idx[:, 0], idx[:, 1], window, n_chrom)
    return off, wmap, vals.astype(np.int32)


def _region_counts_192(store, window_key, window, chrom, strand_minus, blk_ptr, blk_start, blk_end):
    """np.repeat(sum of window rows, 3) with the minus-strand re-ordering of reference :625-634, from the kernel's
    64-context sums (all three substitutions of a context share its count, so the x3 expansion commutes)."""
    off, wmap, vals = _stored_window_map(store, window_key)
    rc, _ = kernels.element_region_counts(np.asarray(chrom, dtype=np.int32),
                                          np.where(strand_minus, -1, 1).astype(np.int8), blk_ptr, blk_start, blk_end,
                                          window, off, wmap, vals, default_device())
    return np.repeat(rc.cpu().numpy(), 3, axis=1)


def preprocess_nonc(f_nonc_bed, f_nonc_data, f_pretrained, L_contexts, save_key, window):
    """Per element: L_counts (sum of its blocks' 192-substitution counts), region_counts (window counts over its
    overlapped windows, strand-aware) and the overlap list, stored under window_{W}/<save_key>/<ELT>
    (reference :596-644).  f_pretrained is only read for its substitution order in the reference; the order here is
    always the sorted one (mk_trans_idx)."""
    from . import genic_driver_tools
    from ..data_tools import mutation_tools
    from .. import storage
    window_key = 'window_{}'.format(window)
    store = storage.Store(f_nonc_data, "a")
    df_elts = mutation_tools.bed12_boundaries(f_nonc_bed)
    E = len(df_elts)
    ptr = np.zeros(E + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(b) for b in df_elts.BLOCK_STARTS])
    bs = np.array([x for b in df_elts.BLOCK_STARTS for x in b], dtype=np.int64)
    be = np.array([x for b in df_elts.BLOCK_ENDS for x in b], dtype=np.int64)
    owner = np.repeat(np.arange(E), np.diff(ptr))
    chrom = df_elts.CHROM.values.astype(np.int64)
    keys = ['chr{}:{}-{}'.format(c, s, e) for c, s, e in zip(chrom[owner], bs, be)]
    L = np.zeros((E, 192), dtype=np.float64)
    np.add.at(L, owner, L_contexts.loc[keys].values)
    minus = np.array([_strand_is_minus(s) for s in df_elts.STRAND.values], dtype=bool)
    R = _region_counts_192(store, window_key, window, chrom, minus, ptr, bs, be)
    overlaps = [genic_driver_tools.get_ideal_overlaps(c, np.vstack((s, e)), window)
                for c, s, e in zip(df_elts.CHROM, df_elts.BLOCK_STARTS, df_elts.BLOCK_ENDS)]
    store.write_element_groups('{}/{}'.format(window_key, save_key), df_elts.ELT.values, L, R, overlaps)


def preprocess_sites(f_sites, f_nonc_data, f_pretrained, save_key, window):
    """Per site set (the SAMPLE column of the sites file): L_counts[j] = number of its sites with substitution j
    (strand-flipped for minus-strand sets, 'nan' contexts skipped), region_counts and overlaps as preprocess_nonc
    (reference :647-711)."""
    from . import genic_driver_tools
    from ..data_tools import mutation_tools
    from .. import storage
    window_key = 'window_{}'.format(window)
    store = storage.Store(f_nonc_data, "a")
    subst_idx = mk_trans_idx(1, 1)
    pos = {s: i for i, s in enumerate(subst_idx)}
    df_sites = mutation_tools.read_mutation_file(f_sites)
    df_sites = df_sites.drop(columns=['GENE', 'ANNOT', 'REF', 'ALT']).rename(columns={'SAMPLE': 'GENE'})
    df_sites['CONTEXT'] = df_sites.CONTEXT.where(df_sites.CONTEXT.notna(), 'nan').astype(str)
    if 'STRAND' not in df_sites.columns:
        df_sites['STRAND'] = '.'
    names, first = np.unique(df_sites.GENE.values.astype(str), return_index=True)      # groupby('GENE'): sorted keys
    gid = {n: i for i, n in enumerate(names)}
    owner = np.array([gid[str(g)] for g in df_sites.GENE.values], dtype=np.int64)
    E = len(names)
    g_chrom = df_sites.CHROM.values[first].astype(np.int64)                             # list(group['CHROM'])[0]
    g_minus = np.array([_strand_is_minus(s) for s in df_sites.STRAND.values[first]], dtype=bool)
    # substitution index of every site, flipped for minus-strand sets; sites whose name contains 'nan' are skipped
    sub = np.full(len(df_sites), -1, dtype=np.int32)
    for i, (mt, cx, o) in enumerate(zip(df_sites.MUT_TYPE.values, df_sites.CONTEXT.values, owner)):
        name = cx + '>' + cx[0] + str(mt)[2] + cx[2] if len(cx) == 3 and len(str(mt)) == 3 else 'nan'
        if g_minus[o] and 'nan' not in name:
            a, b = name.split('>')
            name = reverse_complement(a) + '>' + reverse_complement(b)
        if 'nan' not in name:
            sub[i] = pos[name]                        # KeyError for an unknown substitution, like the reference's dict
    order = np.argsort(owner, kind="stable")
    keep = order[sub[order] >= 0]
    L = kernels.site_counts(owner[keep].astype(np.int32), sub[keep], E, 192, default_device()).cpu().numpy()
    ptr = np.zeros(E + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(np.bincount(owner, minlength=E))
    bs, be = df_sites.START.values[order].astype(np.int64), df_sites.END.values[order].astype(np.int64)
    R = _region_counts_192(store, window_key, window, g_chrom, g_minus, ptr, bs, be)
    overlaps = [genic_driver_tools.get_ideal_overlaps(int(g_chrom[e]), np.vstack((bs[ptr[e]:ptr[e + 1]], be[ptr[e]:ptr[e + 1]])),
                                                      window) for e in range(E)]
    store.write_element_groups('{}/{}'.format(window_key, save_key), names, L.astype(np.float64), R, overlaps)
