#!/usr/bin/env python
"""Remove hypermutated samples from annotated Dig mutation files (the job of the reference's
scripts/filter_hypermut.py): a sample is dropped from a file when it carries more coding mutations (GENE != '.')
than the limit.  Every `*<suffix>` file of --input-dir is rewritten as
`<output-dir>/<stem>.no_hypermut.annot.txt` (headerless TSV).

Reference quirk kept by default: its --max-muts-per-sample argument is parsed but the call hard-codes 3000
(filter_hypermut.py:27-30).  --honour-threshold applies the argument instead.
"""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from digdriver_b200.data_tools import mutation_tools  # noqa: E402

REFERENCE_LIMIT = 3000


def hypermutated_samples(df, limit):
    """Samples whose number of coding rows exceeds ``limit``."""
    coding = df[df.GENE != '.']
    return mutation_tools.filter_hypermut_samples(coding, max_muts_per_sample=limit, return_blacklist=True)[1]


def filter_file(src, dst_dir, limit):
    df = mutation_tools.read_mutation_file(str(src), drop_duplicates=True)
    kept = df[~df.SAMPLE.isin(hypermutated_samples(df, limit))]
    dst = Path(dst_dir) / (src.name.split('.annot.txt')[0] + ".no_hypermut.annot.txt")
    kept.to_csv(dst, header=False, index=False, sep="\t")
    return df.shape, kept.shape, dst


def main(text=None):
    ap = argparse.ArgumentParser(description='Filter hypermutated samples.')
    ap.add_argument('--suffix', default='annot.txt', help='suffix of the Dig mutation files to filter')
    ap.add_argument('--max-muts-per-sample', default=REFERENCE_LIMIT, type=int,
                    help='maximum number of coding mutations per sample (see --honour-threshold)')
    ap.add_argument('--honour-threshold', action='store_true', help='apply --max-muts-per-sample (the reference ignores it)')
    ap.add_argument('--input-dir', default='.')
    ap.add_argument('--output-dir', default='filter_hypermut')
    args = ap.parse_args(text.split() if text is not None else None)
    limit = args.max_muts_per_sample if args.honour_threshold else REFERENCE_LIMIT
    os.makedirs(args.output_dir, exist_ok=True)
    for src in sorted(Path(args.input_dir).glob('*' + args.suffix)):
        before, after, _ = filter_file(src, args.output_dir, limit)
        print(src.name, before, after)


if __name__ == "__main__":
    main()
