"""Range sharding of the hot path over the GPUs of one node.

Windows are independent given a halo of max(n_up, n_down) bases, elements are independent given the window
count table, p-values are independent per row.  So the genome is cut into contiguous genomic ranges on
window boundaries, one per rank; the only exchanges are
  * all-reduce(sum) of the genome-wide context totals (replaces df.sum(axis=0), DigPreprocess.py:59),
  * all-reduce(sum) of substitution counts / scale-factor sums,
  * gather of per-window / per-element result rows on rank 0 (replaces pd.concat, sequence_tools.py:125).
All of them go through torch.distributed (NCCL on GPUs, gloo in the CPU tests); payloads are KB-sized.
"""
import numpy as np
import torch
import torch.distributed as dist

from .genome import Genome


def partition_windows(win_start, win_end, world):
    """Cut the window list (sorted by chromosome, start) into `world` contiguous slices of nearly equal
    base count.  Returns a list of (lo, hi) index pairs."""
    size = (np.asarray(win_end, dtype=np.int64) - np.asarray(win_start, dtype=np.int64)).clip(min=0)
    cum = np.concatenate([[0], np.cumsum(size)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(len(size))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(size)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def slice_genome(genome, win_chrom, win_start, win_end, halo):
    """The part of `genome` one rank needs for its windows: for every chromosome touched, the segment from
    (first window start - halo) to (last window end + halo), clipped to the chromosome.  Returns
    (Genome of segments, chrom index per window into it, re-based starts, re-based ends).

    The halo keeps the reference's edge rules intact after re-basing: a re-based START is 0 only when the true
    START is 0 (sequence_tools.py:25-26), and a segment ends before its window's END + n_down only at the true
    chromosome end (the faidx clipping of :28)."""
    win_chrom = np.asarray(win_chrom)
    ws = np.asarray(win_start, dtype=np.int64)
    we = np.asarray(win_end, dtype=np.int64)
    names, seqs, seg_of = [], [], {}
    new_chrom = np.empty(len(ws), dtype=np.int32)
    new_s, new_e = ws.copy(), we.copy()
    for c in dict.fromkeys(win_chrom.tolist()):
        m = win_chrom == c
        L = len(genome.seqs[c])
        a = max(int(ws[m].min()) - halo, 0)
        b = min(int(we[m].max()) + halo, L)
        seg_of[c] = len(names)
        names.append(genome.names[c])
        seqs.append(genome.seqs[c][a:max(b, a)])
        new_chrom[m] = seg_of[c]
        new_s[m] -= a
        new_e[m] -= a
    return Genome(names, seqs), new_chrom, new_s, new_e


class Collectives:
    """Thin wrapper so the same sharded driver runs on NCCL (GPU) and gloo (CPU tests)."""

    def __init__(self):
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.on else 0
        self.world = dist.get_world_size() if self.on else 1

    def all_reduce_sum(self, t):
        if self.on and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def all_reduce_max(self, t):
        if self.on and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def gather_rows(self, t, sizes=None):
        """Concatenate per-rank row blocks (possibly of different length) on rank 0; None elsewhere.
        `sizes` (rows per rank), when the caller knows them, skips the size exchange and its host reads."""
        if not self.on or self.world == 1:
            return t
        if sizes is None:
            n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
            sizes = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(sizes, n)
            sizes = [int(s.item()) for s in sizes]
        m = max(sizes)
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(pad, bufs, dst=0)
        if self.rank != 0:
            return None
        return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def partition_elements(elt_chrom, elt_first_start, win_chrom, win_start, win_end, parts):
    """Owner rank of every element: the rank whose window slice holds the window that contains the element's FIRST
    block start (multi-exon elements may straddle a shard boundary, so they are assigned by
    their first block and resolved against the all-gathered window-count table).  `parts` = partition_windows(...)
    over windows sorted by (chromosome, start); chromosome labels are compared as given.  Elements whose first block
    lies in no window get rank -1 (the single-process path raises KeyError for them, K6 status 2)."""
    wc = np.asarray(win_chrom).astype(np.int64)
    ws = np.asarray(win_start, dtype=np.int64)
    we = np.asarray(win_end, dtype=np.int64)
    key = (wc << 40) | ws                                     # windows are sorted by (chromosome, start)
    ek = (np.asarray(elt_chrom).astype(np.int64) << 40) | np.asarray(elt_first_start, dtype=np.int64)
    w = np.searchsorted(key, ek, side="right") - 1
    ok = (w >= 0)
    wi = np.clip(w, 0, max(len(ws) - 1, 0))
    if len(ws):
        ok &= (wc[wi] == np.asarray(elt_chrom).astype(np.int64)) & (np.asarray(elt_first_start) < we[wi])
    else:
        ok &= False
    bounds = np.array([hi for _, hi in parts], dtype=np.int64)
    owner = np.searchsorted(bounds, wi, side="right").astype(np.int64)
    owner[~ok] = -1
    return owner


def all_gather_rows(coll, t, sizes):
    """Every rank's row block concatenated on EVERY rank (the all-gather of the [Nw_local, 64] window-count table:
    79 MB for hg19 at 10 kb, so any rank can resolve any element).  `sizes` = rows per rank."""
    if not coll.on or coll.world == 1:
        return t
    m = max(sizes)
    if t.shape[0] != m:                                   # ragged last slice: pad to the common block size
        pad = torch.empty((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)    # the padding rows are dropped below
        pad[: t.shape[0]] = t
        t = pad
    out = torch.empty((coll.world * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())      # one collective straight into the final buffer
    if all(s_ == m for s_ in sizes):
        return out
    return out.index_select(0, _ragged_keep(tuple(int(s_) for s_ in sizes), m, t.device))


_KEEP = {}


def _ragged_keep(sizes, m, device):
    """Row indices that drop the padding of a ragged all-gather.  Built once per (sizes, device): inside a stream-ordered
    or CUDA-graph-captured stage it would otherwise cost one arange per rank plus a concatenation at every step."""
    key = (sizes, m, str(device))
    keep = _KEEP.get(key)
    if keep is None:
        idx = np.concatenate([np.arange(r * m, r * m + s_, dtype=np.int64) for r, s_ in enumerate(sizes)])
        keep = _KEEP[key] = torch.from_numpy(idx).to(device)
    return keep


# --------------------------------------------------------------------------------------------
# ONE genome over all ranks (strong scaling): layout of the all-gathered trinucleotide table
# --------------------------------------------------------------------------------------------

TAIL_I64 = 1024 + 64 + 192   # partial genome totals (pentanucleotide | trinucleotide) and partial substitution counts
TOTALS_ROWS = 40             # TAIL_I64 int64 = 2560 int32 words = 40 rows of 64


class GatheredTable:
    """Layout of the [world, m + TOTALS_ROWS, 64] int32 buffer that ONE all_gather_into_tensor fills: block r holds the
    trinucleotide rows of rank r's window slice (padded to the longest slice, m rows) followed by that rank's partial
    genome-wide totals and substitution counts (pentanucleotide | trinucleotide | 192 substitutions, int64 viewed as 40
    int32 rows).  The element stage indexes the
    buffer in place through a window map that points at the gathered rows, so nothing is copied or compacted, and the
    totals ride in the same collective (replaces the df.sum(axis=0) of DigPreprocess.py:59 and the pd.concat of
    sequence_tools.py:125)."""

    def __init__(self, parts):
        self.parts = [(int(a), int(b)) for a, b in parts]
        self.world = len(self.parts)
        self.m = max(1, max(b - a for a, b in self.parts))
        self.block_rows = self.m + TOTALS_ROWS

    def row_of_window(self, n_win):
        """int64 [n_win]: row of global window w in buffer.view(-1, 64)."""
        row = np.full(n_win, -1, dtype=np.int64)
        for r, (a, b) in enumerate(self.parts):
            row[a:b] = r * self.block_rows + np.arange(b - a)
        return row

    def window_map(self, win_chrom_idx, win_start, window, n_chrom):
        """K6's dense (chromosome, window number) -> row map, pointing into the gathered buffer."""
        from .kernels import build_window_map
        off, wmap = build_window_map(win_chrom_idx, win_start, window, n_chrom)
        rows = self.row_of_window(len(win_chrom_idx))
        ok = wmap >= 0
        out = wmap.copy()
        out[ok] = rows[wmap[ok]].astype(np.int32)
        return off, out

    def local_views(self, local):
        """(rows of the own slice [m, 64], totals5 int64 [1024], totals3 int64 [64], substitution counts int64 [192]) as
        views of this rank's block.  The last one lets the ranks share the sequence model's numerator (the substitution
        counts of the mutations inside their own ranges, train_sequence_model sequence_tools.py:336-339) through the same
        exchange instead of an all-reduce."""
        tot = local[self.m:].view(torch.int64).reshape(-1)
        return local[: self.m], tot[:1024], tot[1024:1088], tot[1088:TAIL_I64]

    def summed_totals(self, gathered):
        """int64 [1280]: the ranks' partial totals (1024 | 64) and substitution counts (192) added up (gathered:
        [world, block_rows, 64] int32)."""
        # each rank's tail is one contiguous run of its block, so the int64 view needs no copy: one reduction kernel
        g = gathered.view(self.world, self.block_rows * 64)[:, self.m * 64:]
        return g.view(torch.int64).sum(dim=0)
