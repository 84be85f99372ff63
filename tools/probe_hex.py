"""Developer probe: hexamer-pair scan vs the per-base kernels (bit-exactness + timing at hg19 scale)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import genome as G, kernels, _lib

hook = ctypes.CDLL(_lib.LIB_PATH).dig_debug_set_scan_variant

def timeit(fn, n=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

total = int(float(sys.argv[1])) if len(sys.argv) > 1 and float(sys.argv[1]) > 0 else 3_100_000_000
lengths = G.hg19_like_lengths(total)
names = ["chr%d" % (i + 1) for i in range(22)]
dg = G.DeviceGenome.synthetic(names, lengths, seed=1)
for W in ((10_000,) if len(sys.argv) > 2 else (10_000, 1_000_000)):
    wins = G.tile_windows(np.arange(22), lengths, W)
    rc = torch.from_numpy(wins[:, 0].astype(np.int32)).cuda(); rs = torch.from_numpy(wins[:, 1]).cuda(); re = torch.from_numpy(wins[:, 2]).cuda()
    nb = float((wins[:, 2] - wins[:, 1]).sum())
    res = {}
    for variant in (1, 2, 0):
        hook(variant)
        out5 = torch.empty((len(wins), 1024), dtype=torch.int32, device="cuda")
        out3 = torch.empty((len(wins), 64), dtype=torch.int32, device="cuda")
        t5 = torch.zeros(1024, dtype=torch.int64, device="cuda"); t3 = torch.zeros(64, dtype=torch.int64, device="cuda")
        fn = lambda: kernels.count_contexts_fused53(dg, rc, rs, re, out5=out5, out3=out3, totals5=t5, totals3=t3)
        t5.zero_(); t3.zero_(); fn(); torch.cuda.synchronize()
        res[variant] = (out5.clone(), out3.clone(), t5.clone(), t3.clone())
        best, med = timeit(fn)
        bpb = 0.375 + 4.0 * (1024 + 64) / W
        print("fused53 W=%d variant=%d: best %.3f ms med %.3f ms -> %.1f GB/s algorithmic, frac %.3f"
              % (W, variant, best, med, nb * bpb / best / 1e6, nb * bpb / best / 1e6 / 6556.5), flush=True)
        o = torch.empty((len(wins), 1024), dtype=torch.int32, device="cuda")
        tt = torch.zeros(1024, dtype=torch.int64, device="cuda")
        fn2 = lambda: kernels.count_contexts(dg, rc, rs, re, 2, 2, out=o, totals=tt)
        best, med = timeit(fn2)
        bpb = 0.375 + 4.0 * 1024 / W
        print("penta   W=%d variant=%d: best %.3f ms med %.3f ms -> %.1f GB/s algorithmic, frac %.3f"
              % (W, variant, best, med, nb * bpb / best / 1e6, nb * bpb / best / 1e6 / 6556.5), flush=True)
        fn3 = lambda: kernels.count_contexts(dg, rc, rs, re, 2, 2, out=o)
        best, med = timeit(fn3)
        print("penta no totals W=%d variant=%d: best %.3f ms" % (W, variant, best), flush=True)
    for k, nm in enumerate(("counts5", "counts3", "totals5", "totals3")):
        print("  %s equal: %s %s" % (nm, bool(torch.equal(res[0][k], res[1][k])), bool(torch.equal(res[2][k], res[1][k]))), flush=True)
hook(0)
