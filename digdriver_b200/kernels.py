"""Tensor-level wrappers over the C ABI (include/dig_b200.h).

PyTorch is used only to own device memory and streams; every computation below is one of
the hand-written sm_100a kernels in csrc/, reached through ctypes with raw device pointers.
"""
import numpy as np
import torch

from . import _lib


def _stream(device, stream=None):
    st = stream if stream is not None else torch.cuda.current_stream(device)
    return st.cuda_stream


def _dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x)).to(device=device, dtype=dtype)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def count_contexts(genome, reg_chrom, reg_start, reg_end, n_up=1, n_down=1, strand=None,
                   want_totals=False, out=None, totals=None, stream=None):
    """K2/K4: per-region context histogram.

    genome: DeviceGenome.  reg_chrom: chromosome indices into the genome (int32).
    Returns (counts int32 [n, K] on device, totals int64 [K] on device or None)."""
    dev = genome.device
    rc = _dev(reg_chrom, torch.int32, dev)
    rs = _dev(reg_start, torch.int64, dev)
    re = _dev(reg_end, torch.int64, dev)
    st = _dev(strand, torch.int8, dev) if strand is not None else None
    n = rc.numel()
    K = 4 ** (n_up + n_down + 1)
    if out is None:
        out = torch.empty((n, K), dtype=torch.int32, device=dev)
    if want_totals and totals is None:
        totals = torch.zeros(K, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dig_count_contexts", genome.packed2.data_ptr(), genome.nmask.data_ptr(), genome.n_bases,
                  genome.chrom_off_d.data_ptr(), genome.chrom_len_d.data_ptr(), rc.data_ptr(), rs.data_ptr(),
                  re.data_ptr(), _ptr(st), n, int(n_up), int(n_down), out.data_ptr(), _ptr(totals),
                  _stream(dev, stream))
    return out, totals
