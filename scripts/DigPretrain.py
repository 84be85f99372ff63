#!/usr/bin/env python
"""CLI shim with the hot-path sub-commands of the reference's scripts/DigPretrain.py: countMutations,
sequenceModel, elementModel, genicModel, tiledModel.  (regionModel ingests the CNN+GP output and stays with the reference.)"""
import argparse
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from digdriver_b200 import kernels, storage  # noqa: E402
from digdriver_b200.data_tools import mutation_tools  # noqa: E402
from digdriver_b200.sequence_model import genic_driver_tools, sequence_tools  # noqa: E402


def get_cpus():
    return min(max(1, (os.cpu_count() or 3) - 2), 20)


PANELS = [('MSK_230', 'MSK_230'), ('MSK_341', 'MSK_341'), ('MSK_410', 'MSK_410'), ('MSK_468', 'MSK_468'),
          ('metabric_173', 'metabric_173'), ('ucla_1202', 'ucla_1202')]


def _distinct_gene_sample_pairs(df):
    """groupby(['GENE', 'SAMPLE']).size()...GENE.value_counts().sum() == number of distinct (GENE, SAMPLE) pairs:
    the K5 gene table (every row counted as one class) summed over genes."""
    if len(df) == 0:
        return 0
    genes, gid = np.unique(df.GENE.values.astype(str), return_inverse=True)
    _, sid = np.unique(df.SAMPLE.values.astype(str), return_inverse=True)
    _, nsamp = kernels.tabulate_genes(gid, sid, np.ones(len(df), dtype=np.uint8), len(genes))
    return int(nsamp[:, 1].sum().item())


def count_training_mutations(args):
    """Reference DigPretrain.py:100-177: cohort sizes of the training mutations as attributes of the pretrained
    model (read back by calc_scale_factor, scale_factor_by_cds / _by_samples and run_target_model), including the
    reference's N_MUT_SAMPLE_CDS = N_MUT_CDS copy (:161)."""
    from digdriver_b200.driver_model import transfer_tools
    st = storage.Store(args.outputFile, "a")
    df = st.read_table('region_params')
    attrs = {'N_MUT_TOTAL': df.loc[:, 'Y_TRUE'].sum(), 'N_MUT_TRAIN': df.loc[~df.FLAG.astype(bool), 'Y_TRUE'].sum()}
    df_mut = mutation_tools.read_mutation_file(args.fmut, drop_duplicates=True)
    df_mut_cds = df_mut[df_mut.ANNOT != 'Noncoding']          # counts Essential_Splice, which borders CDS regions
    attrs['N_SAMPLES'] = len(df_mut.SAMPLE.unique())
    attrs['N_MUT_CDS'] = len(df_mut_cds)
    attrs['N_MUT_SAMPLE_CDS'] = len(df_mut_cds)
    for key, panel in PANELS:
        genes = transfer_tools._panel(panel)
        if genes is None:
            print('WARNING: genes_{}.txt not found (set DIG_DATA_DIR); its counts are not written'.format(panel))
            continue
        sel = df_mut_cds[df_mut_cds.GENE.isin(genes) & ~df_mut_cds.ANNOT.isin(['Synonymous', 'Essential_Splice',
                                                                                 'Noncoding'])]
        attrs['N_MUT_' + key] = len(sel)
        attrs['N_MUT_SAMPLE_' + key] = _distinct_gene_sample_pairs(sel)
        if key == 'MSK_230':
            attrs['N_SAMPLE_MSK_230'] = len(sel.SAMPLE.unique())
    st.set_attrs(**{k: (int(v) if float(v).is_integer() else float(v)) for k, v in attrs.items()})


def pretrain_sequence_model(args):
    """Reference DigPretrain.py:179-208."""
    print('Loading genome-wide context counts')
    src = storage.Store(args.genome_counts, "r")
    df_genome = src.read_table('all_window_genome_counts')
    idx = src.read_array('idx')
    if src.has('mappability'):
        mapp = src.read_array('mappability')
        idx = idx[mapp > args.map_thresh]
        df_genome = df_genome[mapp > args.map_thresh]
    S_genome = df_genome.sum(axis=0)
    print('Loading mutation file')
    df_mut = mutation_tools.read_mutation_file(args.fmut, drop_duplicates=True)
    df_mut = df_mut[df_mut.ANNOT != 'INDEL']
    print('Training sequence model')
    df_freq_mut, df_freq_context = sequence_tools.train_sequence_model(idx, df_mut, S_genome)
    print('Saving sequence models to {}'.format(args.output_h5))
    out = storage.Store(args.output_h5, "a")
    out.write_table('sequence_model_192', df_freq_mut)
    out.write_table('sequence_model_64', df_freq_context)


def pretrain_nonc_model(args):
    """Reference DigPretrain.py:238-268."""
    print('Pretraining element model')
    df = genic_driver_tools.nonc_model_parallel(args.f_pretrained, args.f_element_data, args.save_key, args.N_procs,
                                                indels_direct=args.indels_direct)
    print("saving")
    storage.Store(args.output_h5 or args.f_pretrained, "a").write_table(args.save_key, df)


def pretrain_tiled(args):
    """Reference DigPretrain.py:271-278."""
    df = genic_driver_tools.tiled_model_parallel(args.f_pretrained, args.f_element_data, args.save_key, args.N_procs)
    print("saving")
    storage.Store(args.output_h5 or args.f_pretrained, "a").write_table(args.save_key, df)


def pretrain_genic_model(args):
    """Reference DigPretrain.py:225-236."""
    print('Running Genic model')
    df = genic_driver_tools.genic_model_parallel(args.f_pretrained, args.f_genic, args.N_procs,
                                                 counts_key=args.counts_key, indels_direct=args.indels_direct,
                                                 f_fasta=args.fasta)
    storage.Store(args.output_h5 or args.f_pretrained, "a").write_table('genic_model', df)


def parse_args(text=None):
    parser = argparse.ArgumentParser(description='Create a pre-trained Dig model (B200).')
    sub = parser.add_subparsers()
    a1 = sub.add_parser('countMutations', help='add mutation counts to a pretrained model')
    a1.add_argument('--outputFile', required=True)
    a1.add_argument('--mutation-file', required=True, type=str, dest='fmut')
    a1.set_defaults(func=count_training_mutations)
    b = sub.add_parser('sequenceModel', help='train the sequence-context model')
    b.add_argument('fmut')
    b.add_argument('genome_counts')
    b.add_argument('output_h5')
    b.add_argument('--map-thresh', default=0.5, type=float)
    b.set_defaults(func=pretrain_sequence_model)
    d = sub.add_parser('genicModel', help='pretrain the gene model')
    d.add_argument('f_pretrained')
    d.add_argument('f_genic')
    d.add_argument('--counts-key', default="window_10kb/counts")
    d.add_argument('--output_h5')
    d.add_argument('--indels-direct', action='store_true', default=False)
    d.add_argument('--n-procs', default=get_cpus(), type=int, dest='N_procs')
    d.add_argument('--fasta', default=None, help='genome FASTA (needed when f_pretrained has no window_counts_64)')
    d.set_defaults(func=pretrain_genic_model)
    e = sub.add_parser('elementModel', help='pretrain an element model')
    e.add_argument('f_pretrained')
    e.add_argument('f_element_data')
    e.add_argument('save_key')
    e.add_argument('--output_h5')
    e.add_argument('--indels-direct', action='store_true', default=False)
    e.add_argument('--n-procs', default=get_cpus(), type=int, dest='N_procs')
    e.set_defaults(func=pretrain_nonc_model)
    f = sub.add_parser('tiledModel', help='pretrain a model for the tiles of a tiled genome')
    f.add_argument('f_pretrained')
    f.add_argument('f_element_data')
    f.add_argument('save_key')
    f.add_argument('--output_h5')
    f.add_argument('--n-procs', default=get_cpus(), type=int, dest='N_procs')
    f.set_defaults(func=pretrain_tiled)
    return parser.parse_args(text.split()) if text else parser.parse_args()


if __name__ == "__main__":
    args = parse_args()
    args.func(args)
