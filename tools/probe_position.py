"""Developer probe: K8 per-position hotspot test at chromosome scale (p-values per second)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import genome as G, kernels

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
lengths = np.array([total // 2, total - total // 2], dtype=np.int64)
dg = G.DeviceGenome.synthetic(["chr1", "chr2"], lengths, seed=3)
wins = G.tile_windows(np.arange(2), lengths, 10_000)
rng = np.random.default_rng(0)
n_mut = 200_000
mc = rng.integers(0, 2, n_mut).astype(np.int32)
ms = (rng.random(n_mut) * lengths[mc]).astype(np.int64)
s_prob = rng.lognormal(np.log(1e-6), 1.0, 1024)
mu = rng.gamma(2.0, 10.0, len(wins)) + 0.1
sigma = mu * rng.uniform(0.05, 0.5, len(wins))
dev = dg.device
rc = torch.from_numpy(wins[:, 0].astype(np.int32)).to(dev); rs = torch.from_numpy(wins[:, 1]).to(dev); re = torch.from_numpy(wins[:, 2]).to(dev)
for binsize in (1, 50):
    for want in ((), ("pt", "exp", "pos")):
        for _ in range(2):
            out = kernels.position_test(dg, rc, rs, re, mu, sigma, s_prob, mc, ms, 2, 2, binsize, want=want)
        torch.cuda.synchronize()
        # time only the kernels: re-run the three launches on prepared inputs through the public wrapper
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        import time
        t0 = time.perf_counter(); a.record()
        out = kernels.position_test(dg, rc, rs, re, mu, sigma, s_prob, mc, ms, 2, 2, binsize, want=want)
        b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        n = out["pval"].numel()
        print("binsize %d want=%s: %d p-values, wrapper %.2f ms (events %.2f ms) -> %.2f G p-values/s, %.2f G positions/s"
              % (binsize, want, n, (t1 - t0) * 1e3, a.elapsed_time(b), n / a.elapsed_time(b) / 1e6,
                 float((wins[:, 2] - wins[:, 1]).sum()) / a.elapsed_time(b) / 1e6), flush=True)
        del out
