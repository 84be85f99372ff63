"""The kernels added late in round 1 (overlap join, region counts, persisted-array P_SUM, p-value variants, log-likelihood
and LLR tests) on small inputs for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import kernels
rng = np.random.default_rng(3)
n_blk, n_mut = 500, 4000
bs = rng.integers(0, 40_000, n_blk); be = bs + rng.choice([1, 50, 4000], n_blk)
ms = rng.integers(0, 42_000, n_mut); me = ms + rng.choice([1, 1, 7], n_mut)
im, ib = kernels.overlap_pairs(bs, be, ms, me)
cnt = kernels.overlap_counts(bs, be, ms, me)
assert cnt.sum() == len(im) and np.all((ms[im] < be[ib]) & (bs[ib] < me[im]))
kernels.overlap_pairs(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), ms, me)
# region counts + psum: 2 chromosomes x 30 windows of 1000, 60 elements with 1-40 blocks (some spanning 2 words)
W, n_win = 1000, 60
chrom = np.repeat([1, 2], 30); start = np.tile(np.arange(30) * W, 2)
off, wmap = kernels.build_window_map(chrom, start, W, 3)
wc = rng.integers(0, 50, (n_win, 64)).astype(np.int32)
nb = rng.integers(1, 41, 60); ptr = np.concatenate([[0], np.cumsum(nb)])
bst = np.concatenate([np.sort(rng.integers(0, 29_000, k)) for k in nb]); ben = bst + rng.integers(1, 900, len(bst))
ec = rng.integers(1, 3, 60).astype(np.int32); es = rng.choice([-1, 1], 60).astype(np.int8)
rc, nw = kernels.element_region_counts(ec, es, ptr, bst, ben, W, off, wmap, wc)
L = rng.integers(0, 9, (60, 192)).astype(np.float64)
p, den = kernels.element_psum(L, np.repeat(rc.cpu().numpy(), 3, axis=1), rng.lognormal(-13, 1, 192), want_denom=True)
k = rng.integers(0, 40, 3000).astype(float); a = rng.gamma(2, 3, 3000) + .01; pp = rng.uniform(.01, 1, 3000)
for mode in kernels.NB_MODES:
    kernels.nb_pvalue_variant(mode, k, a, pp, mu=a * (1 - pp) / pp * 1.3)
for kind in kernels.LL_KINDS:
    kernels.loglik(kind, k, a + 1, pp)
pi3 = rng.uniform(1e-4, 1e-2, (3000, 3)); obs3 = rng.poisson(2, (3000, 3)).astype(float)
for model in ("nb", "gamma_poisson"):
    kernels.gene_llr_test(model, a, pp * 50, pi3, obs3, rng.uniform(.5, 2, 3000), rng.uniform(.1, 3, 3000))
torch.cuda.synchronize()
print("sanitize workload (new kernels) ok", len(im), int(rc.sum()), float(p.sum()))
