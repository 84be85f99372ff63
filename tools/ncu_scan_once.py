"""Runs the fused pentanucleotide + trinucleotide scan a few times at hg19 scale (target of `ncu -k regex:scan`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import genome as G, kernels
total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_100_000_000
lengths = G.hg19_like_lengths(total)
dg = G.DeviceGenome.synthetic(["chr%d" % (i + 1) for i in range(22)], lengths, seed=1)
wins = G.tile_windows(np.arange(22), lengths, 10_000)
rc = torch.from_numpy(wins[:, 0].astype(np.int32)).cuda(); rs = torch.from_numpy(wins[:, 1]).cuda(); re = torch.from_numpy(wins[:, 2]).cuda()
out5 = torch.empty((len(wins), 1024), dtype=torch.int32, device="cuda")
out3 = torch.empty((len(wins), 64), dtype=torch.int32, device="cuda")
t5 = torch.zeros(1024, dtype=torch.int64, device="cuda"); t3 = torch.zeros(64, dtype=torch.int64, device="cuda")
ws = kernels.scan_workspace(dg, len(wins))
for _ in range(3):
    kernels.count_contexts_fused53(dg, rc, rs, re, out5=out5, out3=out3, totals5=t5, totals3=t3, workspace=ws, tile_window=10_000)
torch.cuda.synchronize()
print("done")
