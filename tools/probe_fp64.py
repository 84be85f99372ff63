"""FP64 evidence for the test stage: timings of the kernels where the FP64 pipe binds, at the sizes of
BASELINE configs 5 and 3, plus the worst-case rows of the golden p-value grid.

    python tools/probe_fp64.py            # CUDA-event timings (+ tools/micro_dfma for the DFMA peak)
    ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
        --csv --log-file gpurun_out/fp64_ncu.csv python tools/probe_fp64.py --once
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from digdriver_b200 import kernels

once = "--once" in sys.argv
dev = torch.device("cuda:0")
rng = np.random.default_rng(7)


def timeit(fn, n=5, warm=2):
    if once:
        fn(); torch.cuda.synchronize(); return float("nan")
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


peak = None
exe = os.path.join(ROOT, "tools", "micro_dfma")
if not once and os.path.exists(exe):
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True).stdout
    print(out, end="")
    for ln in out.splitlines():
        if ln.startswith("fp64_peak_dfma_per_s"):
            peak = float(ln.split()[1])

# ---- config 5: 37 cohorts x 20 k genes, 13 NB tests each: 9.6 M p-values in ONE launch of the burden-test kernel
C, E = 37, 20_000
n = C * E * 13
mu = rng.gamma(2.0, 20.0, n)
sigma = mu * rng.uniform(0.05, 0.5, n)
alpha, theta = mu ** 2 / sigma ** 2, sigma ** 2 / mu
pi = rng.uniform(1e-4, 0.05, n)
k = rng.poisson(mu * pi).astype(np.float64)
t = [torch.from_numpy(a).to(dev) for a in (k, alpha, theta, pi)]
ms = timeit(lambda: kernels.nb_burden_test(t[0], t[1], t[2], t[3], dev))
print("config 5  dig_nb_burden_test  %d p-values (37 cohorts x 20k genes x 13 tests): %.3f ms -> %.2f G p-values/s"
      % (n, ms, n / ms / 1e6))

# ---- config 3: 30 M one-site site sets
from digdriver_b200.genome import DeviceGenome, tile_windows
lengths = np.array([30_000_001, 20_000_500, 10_123_456], dtype=np.int64)
dg = DeviceGenome.synthetic(["chr1", "chr2", "chr3"], lengths, seed=44, device=dev)
wins = tile_windows([1, 2, 3], lengths, 10_000)
c64, _ = kernels.count_contexts(dg, wins[:, 0] - 1, wins[:, 1], wins[:, 2], 1, 1)
nw = len(wins)
yp = rng.gamma(2.0, 10.0, nw); sd = yp * rng.uniform(0.05, 0.5, nw)
off, wmap = kernels.build_window_map(wins[:, 0], wins[:, 1], 10_000, 4)
d_pr = np.exp(rng.normal(np.log(1e-6), 1.0, 192))
ns = 30_000_000
chrom = rng.integers(1, 4, ns).astype(np.int32)
usable = ((lengths - 1) // 10_000 * 10_000)[chrom - 1]
start = (rng.random(ns) * (usable - 1)).astype(np.int64)
sub = rng.integers(0, 192, ns).astype(np.uint8)
strand = np.where(rng.random(ns) < 0.5, -1, 1).astype(np.int8)
kk = rng.poisson(0.05, ns).astype(np.float64)
a_d = [torch.from_numpy(a).to(dev) for a in (chrom, start, sub, kk, strand)]
ms3 = timeit(lambda: kernels.site_test(a_d[0], a_d[1], a_d[2], a_d[3], 10_000, off, wmap, c64, yp, sd, d_pr, cj=1.37,
                                       site_strand=a_d[4], device=dev), n=3, warm=1)
print("config 3  dig_site_test (+ window denominators)  %d sites: %.3f ms -> %.2f G sites/s" % (ns, ms3, ns / ms3 / 1e6))

# ---- worst case of the continued fraction: the alpha >= 5e4 rows of the golden grid, replicated to 1 M values
z = np.load(os.path.join(ROOT, "tests", "golden", "nbtest.npz"))
gk, ga, gp = z["k"].astype(np.float64), z["alpha"].astype(np.float64), z["p"].astype(np.float64)
if gk is not None:
    big = (ga >= 5e4) & np.isfinite(ga) & np.isfinite(gk) & np.isfinite(gp)
    reps = max(1, 1_000_000 // max(int(big.sum()), 1))
    wk, wa, wp = (torch.from_numpy(np.tile(x[big], reps)).to(dev) for x in (gk, ga, gp))
    msw = timeit(lambda: kernels.nb_pvalue_greater_midp(wk, wa, wp, dev))
    allk, alla, allp = (torch.from_numpy(np.tile(x, max(1, 1_000_000 // len(gk)))).to(dev) for x in (gk, ga, gp))
    msa = timeit(lambda: kernels.nb_pvalue_greater_midp(allk, alla, allp, dev))
    print("worst case  dig_nb_pvalue_greater_midp on the %d grid rows with alpha >= 5e4 (x%d = %d values): %.3f ms -> %.3f G p-values/s"
          % (int(big.sum()), reps, wk.numel(), msw, wk.numel() / msw / 1e6))
    print("whole grid  (%d values): %.3f ms -> %.3f G p-values/s" % (allk.numel(), msa, allk.numel() / msa / 1e6))
if peak:
    print("DFMA peak %.3e thread-instr/s = %.1f TFLOP/s FP64; see the ncu pass for the FP64-pipe share of each kernel" % (peak, 2 * peak / 1e12))
