"""Quick device-side timing probe of K1/K2 at hg19 scale (developer tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import genome as G, kernels

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
lengths = G.hg19_like_lengths(total)
names = ["chr%d" % (i + 1) for i in range(22)]
t0 = time.time()
dg, ascii_d = G.DeviceGenome.synthetic(names, lengths, seed=1, return_ascii=True)
torch.cuda.synchronize()
print("synth+pack %.2fs, n_bases %d" % (time.time() - t0, dg.n_bases))
best, med = timeit(lambda: G.DeviceGenome.pack_ascii(ascii_d))
print("pack: best %.3f ms  med %.3f ms -> %.1f GB/s (1.375 B/base)" % (best, med, dg.n_bases * 1.375 / best / 1e6))
del ascii_d
for W in (10_000, 1_000_000):
    wins = G.tile_windows(np.arange(22), lengths, W)
    rc = torch.from_numpy(wins[:, 0].astype(np.int32)).cuda(); rs = torch.from_numpy(wins[:, 1]).cuda(); re = torch.from_numpy(wins[:, 2]).cuda()
    for (u, d) in ((1, 1), (2, 2)):
        K = 4 ** (u + d + 1)
        out = torch.empty((len(wins), K), dtype=torch.int32, device="cuda")
        tot = torch.zeros(K, dtype=torch.int64, device="cuda")
        fn = lambda: kernels.count_contexts(dg, rc, rs, re, u, d, out=out, totals=tot)
        best, med = timeit(fn)
        nb = float((wins[:, 2] - wins[:, 1]).sum())
        bpb = 0.375 + 4.0 * K / W
        print("scan W=%d K=%d: best %.3f ms med %.3f ms -> %.2f Gbases/s, %.1f GB/s algorithmic (%.4f B/base), frac of 6556: %.3f"
              % (W, K, best, med, nb / best / 1e6, nb * bpb / best / 1e6, bpb, nb * bpb / best / 1e6 / 6556.5))
