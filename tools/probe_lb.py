"""Developer probe: lane-bank scan (scan_lb.cu) vs the per-warp hexamer kernel (bit-exactness + timing at hg19 scale).
usage: probe_lb.py [total_bases] [--quick]"""
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


args = [a for a in sys.argv[1:] if not a.startswith("--")]
total = int(float(args[0])) if args else 3_100_000_000
lengths = G.hg19_like_lengths(total)
names = ["chr%d" % (i + 1) for i in range(22)]
print(torch.cuda.get_device_properties(0).name, torch.cuda.get_device_properties(0).uuid, flush=True)
dg = G.DeviceGenome.synthetic(names, lengths, seed=1)
W = 10_000
TW = 0 if "--no-hint" in sys.argv else W
wins = G.tile_windows(np.arange(22), lengths, W)
rc = torch.from_numpy(wins[:, 0].astype(np.int32)).cuda(); rs = torch.from_numpy(wins[:, 1]).cuda(); re = torch.from_numpy(wins[:, 2]).cuda()
nb = float((wins[:, 2] - wins[:, 1]).sum())
ws = kernels.scan_workspace(dg, len(wins))
res = {}
for variant, name in ((_lib.SCAN_HEX, "per-warp hexamer"), (_lib.SCAN_AUTO, "lane-bank")):
    out5 = torch.empty((len(wins), 1024), dtype=torch.int32, device="cuda")
    out3 = torch.empty((len(wins), 64), dtype=torch.int32, device="cuda")
    t5 = torch.zeros(1024, dtype=torch.int64, device="cuda"); t3 = torch.zeros(64, dtype=torch.int64, device="cuda")
    fn = lambda: kernels.count_contexts_fused53(dg, rc, rs, re, out5=out5, out3=out3, totals5=t5, totals3=t3,
                                                variant=variant, workspace=ws, tile_window=TW)
    ws.zero_(); fn(); torch.cuda.synchronize()
    res[variant] = (out5.clone(), out3.clone(), t5.clone(), t3.clone())
    if variant == _lib.SCAN_AUTO and "--timing" in sys.argv:
        off = (16 + 4 * len(wins) + 15) // 16 * 16
        t = ws[off:off + 128].view(torch.int64).cpu().numpy().astype(float)
        nwarp = 148 * 16
        names = ["wait FULL", "wait CLEAN", "process chunks", "barrier B1", "wait DONE", "wait OUTEMPTY", "write-out work", "-"]
        tot = t[:7].sum()
        print("  phase cycles per consumer warp (sum %.0f = %.3f ms at 1.955 GHz):" % (tot / nwarp, tot / nwarp / 1.955e6))
        for nm_, v in zip(names[:7], t):
            print("    %-16s %9.0f  %5.1f%%" % (nm_, v / nwarp, 100 * v / tot))
        n_chunks = 5 * ((len(wins) + 31) // 32)
        print("    copy latency (issued -> FULL complete, producer warp 0): %.0f cycles per chunk; issue %.0f; producer idle %.0f"
              % (t[7] / n_chunks, t[12] / n_chunks, t[11] / n_chunks))
        print("    geometry %.0f (%.1f%%); whole consumer loop %.0f cycles per warp" % (t[13] / nwarp, 100 * t[13] / t[14], t[14] / nwarp))
        nb_ = (len(wins) + 31) // 32
        print("    FULL wait per consumer warp and chunk: first %.0f, second %.0f, later %.0f cycles"
              % (t[8] / 16 / nb_, t[9] / 16 / nb_, t[10] / 16 / nb_ / 3))
    if variant == _lib.SCAN_AUTO:
        print("  redo list length: %d of %d windows" % (int(ws[:4].view(torch.int32).item()), len(wins)), flush=True)
    best, med = timeit(fn)
    bpb = 0.375 + 4.0 * (1024 + 64) / W
    print("fused53 %-18s best %.3f ms med %.3f ms -> %.1f GB/s algorithmic, frac %.3f"
          % (name, best, med, nb * bpb / best / 1e6, nb * bpb / best / 1e6 / 6556.5), flush=True)
    o = torch.empty((len(wins), 1024), dtype=torch.int32, device="cuda")
    tt = torch.zeros(1024, dtype=torch.int64, device="cuda")
    fn2 = lambda: kernels.count_contexts(dg, rc, rs, re, 2, 2, out=o, totals=tt, variant=variant, workspace=ws, tile_window=TW)
    best, med = timeit(fn2)
    bpb = 0.375 + 4.0 * 1024 / W
    print("penta   %-18s best %.3f ms med %.3f ms -> frac %.3f" % (name, best, med, nb * bpb / best / 1e6 / 6556.5), flush=True)
    fn3 = lambda: kernels.count_contexts(dg, rc, rs, re, 2, 2, out=o, variant=variant, workspace=ws, tile_window=TW)
    best, med = timeit(fn3)
    print("penta no totals %-10s best %.3f ms" % (name, best), flush=True)
if "--repeat" in sys.argv:          # race hunt: the lane-bank result must be identical every time
    n_rep = int(sys.argv[sys.argv.index("--repeat") + 1])
    want5, want3 = res[_lib.SCAN_HEX][0], res[_lib.SCAN_HEX][1]
    o5 = torch.empty_like(want5); o3 = torch.empty_like(want3)
    nbad = 0
    cold = torch.empty(1 << 30, dtype=torch.uint8, device="cuda") if "--cold" in sys.argv else None
    for it in range(n_rep):
        if cold is not None:
            cold.fill_(it & 255)          # evict L2 (and most TLB entries) before every run
            if it % 3 == 0:
                ws = kernels.scan_workspace(dg, len(wins))
                o5 = torch.empty_like(want5); o3 = torch.empty_like(want3)
        tt5 = torch.zeros(1024, dtype=torch.int64, device="cuda") if it % 2 == 0 else None
        tt3 = torch.zeros(64, dtype=torch.int64, device="cuda") if it % 2 == 0 else None
        kernels.count_contexts_fused53(dg, rc, rs, re, out5=o5, out3=o3, totals5=tt5, totals3=tt3, variant=_lib.SCAN_AUTO,
                                       workspace=ws, tile_window=TW)
        torch.cuda.synchronize()
        bad = torch.nonzero((o5 != want5).any(dim=1)).flatten()
        if bad.numel():
            nbad += 1
            print("  repeat %d (totals %s): %d bad rows %s (batch %s lane-window %s)" % (it, it % 2 == 0, bad.numel(), bad[:6].tolist(), (bad[:6] // 32).tolist(), (bad[:6] % 32).tolist()), flush=True)
    print("  repeats with differences: %d of %d" % (nbad, n_rep), flush=True)
for k, nm in enumerate(("counts5", "counts3", "totals5", "totals3")):
    a, b = res[_lib.SCAN_AUTO][k], res[_lib.SCAN_HEX][k]
    eq = bool(torch.equal(a, b))
    print("  %s equal: %s" % (nm, eq), flush=True)
    if not eq and a.dim() == 2 and k == 0:
        # which side is unstable?  run both again
        o5b = torch.empty_like(a); o3b = torch.empty((len(wins), 64), dtype=torch.int32, device="cuda")
        for vv, nm2, ref in ((_lib.SCAN_HEX, "per-warp", b), (_lib.SCAN_AUTO, "lane-bank", a)):
            kernels.count_contexts_fused53(dg, rc, rs, re, out5=o5b, out3=o3b, variant=vv, workspace=ws, tile_window=TW)
            torch.cuda.synchronize()
            print("    rerun %s: equals its first run: %s; equals the other kernel's first run: %s"
                  % (nm2, bool(torch.equal(o5b, ref)), bool(torch.equal(o5b, b if vv == _lib.SCAN_AUTO else a))), flush=True)
    if not eq and a.dim() == 2:
        badrows = torch.nonzero((a != b).any(dim=1)).flatten()
        print("    %d bad rows, first %s" % (badrows.numel(), badrows[:10].tolist()), flush=True)
        r0 = int(badrows[0]); cols = torch.nonzero(a[r0] != b[r0]).flatten()[:8]
        print("    row %d cols %s got %s want %s" % (r0, cols.tolist(), a[r0][cols].tolist(), b[r0][cols].tolist()), flush=True)
