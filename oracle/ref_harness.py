"""TEST INFRASTRUCTURE -- never imported by the product path.

Loads the UNMODIFIED reference modules from /root/reference inside this
container, where the reference's I/O-only third-party imports (pysam, h5py,
pybedtools, statsmodels, ...) are absent.  Empty stub modules are registered
for those names and a duck-typed in-memory ``pysam.FastaFile`` is provided, so
the reference's *arithmetic* runs exactly as shipped.

Only ``tests/golden/make_golden.py`` uses this file, and only in the build
container: /root/reference does not exist on the GPU box, so nothing that runs
there may import it.  The frozen outputs live under ``tests/golden/``.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIG_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "pysam", "h5py", "pybedtools", "statsmodels", "statsmodels.stats",
    "statsmodels.stats.multitest", "bbi", "seaborn", "matplotlib",
    "matplotlib.pyplot", "tables", "pkg_resources",
]


class FakeFasta:
    """In-memory stand-in for ``pysam.FastaFile``.

    ``fetch(chrom, start, end)`` clips ``end`` at the chromosome length the way
    faidx does (reference relies on it: sequence_tools.py:28, quirk a1-ii).
    """

    registry = {}

    def __init__(self, name):
        self.seqs = FakeFasta.registry[name]

    def fetch(self, chrom, start=None, end=None):
        seq = self.seqs[chrom]
        if start is None:
            return seq
        if start < 0:
            raise ValueError("start out of range (%d)" % start)
        return seq[start:end]

    @property
    def references(self):
        return list(self.seqs.keys())

    def close(self):
        pass


def register_fasta(name, seqs):
    """Register ``{chrom_name: str}`` under a fake path ``name``."""
    FakeFasta.registry[name] = seqs
    return name


def load_reference():
    """Return a namespace with the reference's hot-path modules."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pysam"].FastaFile = FakeFasta
    pr = sys.modules["pkg_resources"]
    data_dir = os.path.join(REFERENCE_ROOT, "DIGDriver")
    pr.resource_stream = lambda pkg, rel: open(os.path.join(data_dir, rel), "rb")
    pr.resource_filename = lambda pkg, rel: os.path.join(data_dir, rel)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from DIGDriver.sequence_model import sequence_tools, genic_driver_tools, nb_model
    from DIGDriver.driver_model import transfer_tools
    from DIGDriver.data_tools import mutation_tools
    ns = types.SimpleNamespace(
        sequence_tools=sequence_tools,
        genic_driver_tools=genic_driver_tools,
        nb_model=nb_model,
        transfer_tools=transfer_tools,
        mutation_tools=mutation_tools,
    )
    return ns
