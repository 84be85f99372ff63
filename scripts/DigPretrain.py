#!/usr/bin/env python
"""CLI shim with the hot-path sub-commands of the reference's scripts/DigPretrain.py: sequenceModel,
elementModel, genicModel.  (regionModel ingests the CNN+GP output and stays with the reference.)"""
import argparse
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from digdriver_b200 import storage  # noqa: E402
from digdriver_b200.data_tools import mutation_tools  # noqa: E402
from digdriver_b200.sequence_model import genic_driver_tools, sequence_tools  # noqa: E402


def get_cpus():
    return min(max(1, (os.cpu_count() or 3) - 2), 20)


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
