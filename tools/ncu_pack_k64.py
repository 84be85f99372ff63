"""One pack launch (K1) and one K=64-only scan (lane-private histograms) on an hg19-sized synthetic genome, for
`ncu --set full -k regex:pack_kernel|scan_sym_kernel -c 2` (profiles/r01_pack_k64_summary.txt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from digdriver_b200 import genome as G, kernels  # noqa: E402

lengths = G.hg19_like_lengths(3_100_000_000)
dg, ascii_d = G.DeviceGenome.synthetic(["chr%d" % (i + 1) for i in range(22)], lengths, seed=1, return_ascii=True)
wins = G.tile_windows(np.arange(22), lengths, 10_000)
w = [torch.from_numpy(np.ascontiguousarray(wins[:, i])).cuda() for i in range(3)]
torch.cuda.synchronize()
G.DeviceGenome.pack_ascii(ascii_d)                      # K1 on the whole 3.1 GB ASCII genome
kernels.count_contexts(dg, w[0].to(torch.int32), w[1], w[2], 1, 1, want_totals=True)
torch.cuda.synchronize()
print("ok")
