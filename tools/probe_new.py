"""CUDA-event timings of the kernels added late in round 1, at the sizes of BASELINE.json's configs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from digdriver_b200 import _lib, kernels  # noqa: E402

dev = torch.device("cuda:0")
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)   # noqa: E731


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3          # us


rng = np.random.default_rng(5)
st = torch.cuda.current_stream(dev).cuda_stream
# ---- overlap join: 1 M mutations x 300 k blocks
n_blk, n_mut = 300_000, 1_000_000
cb = rng.integers(0, 22, n_blk).astype(np.int64)
bs = rng.integers(0, 100_000_000, n_blk)
be = bs + rng.choice([200, 2000, 50_000], n_blk, p=[0.45, 0.45, 0.1])
kbs, kbe = (cb << 32) | bs, (cb << 32) | be
order = np.argsort(kbs, kind="stable")
kbs, kbe = kbs[order], kbe[order]
cm = rng.integers(0, 22, n_mut).astype(np.int64)
ms = rng.integers(0, 100_000_000, n_mut)
b0, b1, b2 = t(kbs, torch.int64), t(kbe, torch.int64), t(np.maximum.accumulate(kbe), torch.int64)
m0, m1 = t((cm << 32) | ms, torch.int64), t((cm << 32) | (ms + 1), torch.int64)
cnt = torch.zeros(n_mut, dtype=torch.int64, device=dev)
us_count = timed(lambda: _lib.call("dig_overlap_count", b0.data_ptr(), b1.data_ptr(), b2.data_ptr(), n_blk, m0.data_ptr(),
                                   m1.data_ptr(), n_mut, cnt.data_ptr(), st))
off = torch.zeros(n_mut + 1, dtype=torch.int64, device=dev)
off[1:] = torch.cumsum(cnt, 0)
total = int(off[-1])
pm, pb = torch.empty(total, dtype=torch.int64, device=dev), torch.empty(total, dtype=torch.int64, device=dev)
us_fill = timed(lambda: _lib.call("dig_overlap_fill", b0.data_ptr(), b1.data_ptr(), b2.data_ptr(), n_blk, m0.data_ptr(),
                                  m1.data_ptr(), n_mut, off.data_ptr(), pm.data_ptr(), pb.data_ptr(), st))
print("overlap join  1M mutations x 300k blocks: count %.1f us, fill %.1f us (%d pairs)" % (us_count, us_fill, total))
# ---- region counts + psum: 100 k elements on a 310 k-window map
W, n_chrom, per = 10_000, 22, 14_000
chrom = np.repeat(np.arange(n_chrom), per)
start = np.tile(np.arange(per) * W, n_chrom)
woff, wmap = kernels.build_window_map(chrom, start, W, n_chrom)
wc = t(rng.integers(0, 400, (len(chrom), 64)), torch.int32)
E = 100_000
nb = rng.integers(1, 4, E)
ptr = np.concatenate([[0], np.cumsum(nb)])
first = rng.integers(0, (per - 5) * W, E)
owner = np.repeat(np.arange(E), nb)
step = rng.integers(300, 9000, len(owner))
rel = np.cumsum(step) - step
rel -= rel[ptr[:-1]][owner]
ebs = first[owner] + rel
ebe = ebs + rng.integers(200, 2000, len(ebs))
args = (t(rng.integers(0, n_chrom, E), torch.int32), t(rng.choice([-1, 1], E), torch.int8), t(ptr, torch.int64),
        t(ebs, torch.int64), t(ebe, torch.int64))
span = kernels.element_max_span(ptr, ebs, ebe, W)
rc = torch.empty((E, 64), dtype=torch.int64, device=dev)
nw = torch.empty(E, dtype=torch.int32, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
wo, wm = t(woff, torch.int64), t(wmap, torch.int32)
us_rc = timed(lambda: _lib.call("dig_element_region_counts", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(),
                                args[3].data_ptr(), args[4].data_ptr(), E, W, int(wo.numel()) - 1, wo.data_ptr(), wm.data_ptr(), wc.data_ptr(),
                                wc.shape[0], span, rc.data_ptr(), nw.data_ptr(), status.data_ptr(), st))
L = t(rng.integers(0, 30, (E, 192)), torch.float64)
R = torch.repeat_interleave(rc, 3, dim=1).contiguous()
dp = t(rng.lognormal(np.log(1e-6), 1.0, 192), torch.float64)
p = torch.empty(E, dtype=torch.float64, device=dev)
us_ps = timed(lambda: _lib.call("dig_element_psum", L.data_ptr(), R.data_ptr(), dp.data_ptr(), E, p.data_ptr(), None, st))
print("region counts 100k elements: %.1f us;  psum from persisted arrays (307 MB read): %.1f us = %.0f GB/s" %
      (us_rc, us_ps, (L.numel() * 8 + R.numel() * 8) / us_ps / 1e3))
# ---- p-value conventions, log-likelihoods, LLR tests: 1 M rows
n = 1_000_000
k = t(rng.poisson(3.0, n), torch.float64)
a = t(rng.gamma(2.0, 3.0, n) + 0.01, torch.float64)
pp = t(rng.uniform(0.01, 1.0, n), torch.float64)
out = torch.empty(n, dtype=torch.float64, device=dev)
for mode, code in kernels.NB_MODES.items():
    us = timed(lambda: _lib.call("dig_nb_pvalue_variant", code, k.data_ptr(), a.data_ptr(), pp.data_ptr(), None, n,
                                 out.data_ptr(), st))
    print("nb_pvalue_variant %-13s 1M: %.1f us = %.1f G p-values/s" % (mode, us, n / us / 1e3))
for kind, code in kernels.LL_KINDS.items():
    us = timed(lambda: _lib.call("dig_loglik", code, k.data_ptr(), a.data_ptr(), pp.data_ptr(), n, out.data_ptr(), st))
    print("loglik %-6s 1M: %.1f us" % (kind, us))
pi3, obs3 = t(rng.uniform(1e-4, 1e-2, (n, 3)), torch.float64), t(rng.poisson(2, (n, 3)), torch.float64)
mrf, ts = t(rng.uniform(0.5, 2, n), torch.float64), t(rng.uniform(0.1, 3, n), torch.float64)
o4 = torch.empty((4, n), dtype=torch.float64, device=dev)
for model in (0, 1):
    us = timed(lambda: _lib.call("dig_gene_llr_test", model, a.data_ptr(), pp.data_ptr(), pi3.data_ptr(), obs3.data_ptr(),
                                 mrf.data_ptr(), ts.data_ptr(), n, o4.data_ptr(), st))
    print("gene_llr_test model %d 1M rows: %.1f us = %.1f M rows/s" % (model, us, n / us))
