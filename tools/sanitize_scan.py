"""Small fused + plain scans and the test-stage kernels for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import genome as G, kernels
lengths = np.array([400_003, 99_990], dtype=np.int64)
dg = G.DeviceGenome.synthetic(["chr1", "chr2"], lengths, seed=5)
wins = np.concatenate([G.tile_windows(np.arange(2), lengths, 10_000), np.array([[0, 0, 400_003], [1, 3, 99_990], [0, 7, 8]])])
c5, c3, t5, t3 = kernels.count_contexts_fused53(dg, wins[:, 0], wins[:, 1], wins[:, 2], want_totals=True)
p5, pt = kernels.count_contexts(dg, wins[:, 0], wins[:, 1], wins[:, 2], 2, 2, want_totals=True)
k3, kt = kernels.count_contexts(dg, wins[:, 0], wins[:, 1], wins[:, 2], 1, 1, want_totals=True)
torch.cuda.synchronize()
assert torch.equal(c5, p5) and torch.equal(c3, k3) and torch.equal(t5, pt) and torch.equal(t3, kt)
# enough regions for the trinucleotide-only lane-bank kernel (>= 64 per SM), ragged ones included
many = np.concatenate([wins] * (64 * 160 // len(wins) + 1))
m3, mt = kernels.count_contexts(dg, many[:, 0], many[:, 1], many[:, 2], 1, 1, want_totals=True)
torch.cuda.synchronize()
assert torch.equal(m3[: len(wins)], k3) and torch.equal(m3[-len(wins):], k3)
rng = np.random.default_rng(0)
out = kernels.position_test(dg, wins[:5, 0], wins[:5, 1], wins[:5, 2], rng.gamma(2, 5, 5) + 1, rng.uniform(.5, 2, 5),
                            rng.random(1024) * 1e-6, np.zeros(500, dtype=np.int32), rng.integers(0, 50_000, 500), 2, 2, 1)
torch.cuda.synchronize()
print("sanitize workload ok", int(c5.sum()), float(out["pval"].sum()))
