#!/usr/bin/env python
"""Drop-in for the reference's scripts/filter_hypermut.py: writes filter_hypermut/<name>.no_hypermut.annot.txt for
every '*<suffix>' mutation file of the current directory with the hypermutated samples removed.  As in the reference
the threshold actually applied is 3000 coding mutations per sample (its --max-muts-per-sample is parsed and ignored,
filter_hypermut.py:27-30); pass --honour-threshold to use the argument."""
import argparse
import os
import pathlib
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from digdriver_b200.data_tools import mutation_tools  # noqa: E402


def main(text=None):
    parser = argparse.ArgumentParser(description='Filter hypermutated samples.')
    parser.add_argument('--suffix', default='annot.txt', help='suffix of Dig mutation files to filter')
    parser.add_argument('--max-muts-per-sample', default=3000, type=int)
    parser.add_argument('--honour-threshold', action='store_true', default=False)
    args = parser.parse_args(text.split() if text is not None else None)
    os.makedirs("filter_hypermut", exist_ok=True)
    for f in sorted(pathlib.Path('.').glob('*' + args.suffix)):
        df = mutation_tools.read_mutation_file(str(f), drop_duplicates=True)
        df_mut = df[df.GENE != '.']
        thresh = args.max_muts_per_sample if args.honour_threshold else 3000
        _, sample_blacklist = mutation_tools.filter_hypermut_samples(df_mut, max_muts_per_sample=thresh,
                                                                     return_blacklist=True)
        df_out = df[~df.SAMPLE.isin(sample_blacklist)]
        print(f.name, df.shape, df_out.shape)
        f_out = os.path.join("filter_hypermut", f.name.split('.annot.txt')[0] + ".no_hypermut.annot.txt")
        df_out.to_csv(f_out, header=False, index=False, sep="\t")


if __name__ == "__main__":
    main()
