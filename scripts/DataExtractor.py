#!/usr/bin/env python
"""CLI shim for the one DataExtractor.py sub-command next to the hot path: ``addObjectives`` (per-window observed
mutation counts, reference scripts/DataExtractor.py:525-572, argparse :849-861).  Same positional arguments and
flags; the archive is the directory store of digdriver_b200.storage (or a real HDF5 file when h5py is present).
The other sub-commands of the reference script (epigenome tracks, mappability, chunking for CNN training) belong to
the region model and are out of scope."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def add_objectives(args):
    from digdriver_b200.data_tools import objectives
    print('Adding mutation counts from {} to {}'.format(args.mut_file, args.h5_file))
    name, counts = objectives.add_objectives(args.h5_file, args.mut_file, args.max_muts_per_sample,
                                             args.sample_filter_stdev, args.max_muts_per_elt_per_sample, args.suffix,
                                             cnv=args.cnv)
    print('Saving dataset as {}'.format(name))


def parse_args(text=None):
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    subparsers = parser.add_subparsers(help='DataExtractor sub-commands')
    parser_g = subparsers.add_parser('addObjectives',
                                     help='count mutations (SNVs) for a cancer type to an hd5 dataset.')
    parser_g.add_argument('h5_file', help='path to h5 file')
    parser_g.add_argument('mut_file', help='path to file of mutations')
    parser_g.add_argument('--max-muts-per-sample', type=int, default=None,
                          help='Maximum mutations allowed per sample. Samples with higher mutation counts are removed.')
    parser_g.add_argument('--sample-filter-stdev', type=float, default=None,
                          help='Remove samples with # mutations > filter-stdev * stdev of mutation counts across cohort.')
    parser_g.add_argument('--max-muts-per-elt-per-sample', type=int, default=None,
                          help='Cap the number of mutations a sample can contribute to any one window.')
    parser_g.add_argument('--suffix', type=str, default='',
                          help='suffix to add to end of cancer name when saving mutation counts to h5 archive.')
    parser_g.add_argument('--cnv', help='designates the mut file is of CNVs', action='store_true')
    parser_g.set_defaults(func=add_objectives)
    if text:
        args = parser.parse_args(text)
    else:
        args = parser.parse_args()
    return args


if __name__ == "__main__":
    args = parse_args()
    args.func(args)
