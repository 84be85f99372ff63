"""Developer probe: K=64 scan at hg19 scale -- lane-bank 4-mer-pair kernel (scan_lb.cu, tri-only mode) vs the per-warp kernel."""
import sys, os
if "--timing" in sys.argv:      # python -m digdriver_b200.build --timing  first
    os.environ["DIG_LIB_PATH"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "digdriver_b200", "libdigb200_timing.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digdriver_b200 import genome as G, kernels, _lib


def timeit(fn, n=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


total = 3_100_000_000
lengths = G.hg19_like_lengths(total)
dg = G.DeviceGenome.synthetic(["chr%d" % (i + 1) for i in range(22)], lengths, seed=1)
for W in (10_000, 1_000_000):
    wins = G.tile_windows(np.arange(22), lengths, W)
    rc = torch.from_numpy(wins[:, 0].astype(np.int32)).cuda(); rs = torch.from_numpy(wins[:, 1]).cuda(); re = torch.from_numpy(wins[:, 2]).cuda()
    nb = float((wins[:, 2] - wins[:, 1]).sum())
    ws = kernels.scan_workspace(dg, len(wins))
    res = {}
    for variant, name in ((_lib.SCAN_PER_BASE, "per-warp lane-private"), (_lib.SCAN_AUTO, "lane-bank 4-mer pairs")):
        out = torch.empty((len(wins), 64), dtype=torch.int32, device="cuda")
        tt = torch.zeros(64, dtype=torch.int64, device="cuda")
        fn = lambda: kernels.count_contexts(dg, rc, rs, re, 1, 1, out=out, totals=tt, variant=variant, workspace=ws, tile_window=W)
        tt.zero_(); fn(); torch.cuda.synchronize()
        res[variant] = (out.clone(), tt.clone())
        if variant == _lib.SCAN_AUTO and "--timing" in sys.argv:
            off = (16 + 4 * len(wins) + 15) // 16 * 16
            t = ws[off:off + 128].view(torch.int64).cpu().numpy().astype(float)
            nwarp = 148 * 16
            nb_ = (len(wins) + 31) // 32 / 148.0
            print("  phase cycles per consumer warp and batch (%.1f batches per CTA): wait FULL %.0f, wait CLEAN %.0f, process %.0f, "
                  "barrier %.0f, wait DONE %.0f, write-out %.0f; whole loop %.0f; producer: copy latency %.0f, issue %.0f, idle %.0f per chunk"
                  % (nb_, t[0] / nwarp / nb_, t[1] / nwarp / nb_, t[2] / nwarp / nb_, t[3] / nwarp / nb_, t[4] / nwarp / nb_, t[6] / nwarp / nb_,
                     t[14] / nwarp / nb_, t[7] / (5 * nb_ * 148), t[12] / (5 * nb_ * 148), t[11] / (5 * nb_ * 148)), flush=True)
        best, med = timeit(fn)
        bpb = 0.375 + 4.0 * 64 / W
        print("W=%d K=64 %-24s best %.3f ms med %.3f ms -> %.1f GB/s algorithmic, frac %.3f"
              % (W, name, best, med, nb * bpb / best / 1e6, nb * bpb / best / 1e6 / 6556.5), flush=True)
    a, b = res[_lib.SCAN_AUTO], res[_lib.SCAN_PER_BASE]
    print("  counts equal: %s   totals equal: %s   redo list: %d" % (bool(torch.equal(a[0], b[0])), bool(torch.equal(a[1], b[1])),
                                                                  int(ws[:4].view(torch.int32).item())), flush=True)
    bad = 0
    o = torch.empty_like(a[0])
    for it in range(6):
        kernels.count_contexts(dg, rc, rs, re, 1, 1, out=o, variant=_lib.SCAN_AUTO, workspace=ws, tile_window=W)
        torch.cuda.synchronize()
        bad += int(not torch.equal(o, b[0]))
    print("  repeats that differ: %d of 6" % bad, flush=True)
