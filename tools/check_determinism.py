import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from digdriver_b200 import kernels, _lib, genome as G
dev = torch.device("cuda:0")
dg, ascii_d, d = bench.build_workload(3.1e9, seed=1, device=dev)
di = bench.DeviceInputs(d, dev)
print("packed2 ptr %x nmask ptr %x counts5 ptr %x counts3 ptr %x" % (dg.packed2.data_ptr(), dg.nmask.data_ptr(), di.counts5.data_ptr(), di.counts3.data_ptr()))
ref5 = torch.empty_like(di.counts5); ref3 = torch.empty_like(di.counts3)
kernels.count_contexts_fused53(dg, di.win_chrom, di.win_start, di.win_end, out5=ref5, out3=ref3, variant=_lib.SCAN_HEX)
torch.cuda.synchronize()
def rounds(tag, n=8, **kw):
    bad = 0
    for it in range(n):
        bench.scan_stage(dg, di)
        torch.cuda.synchronize()
        b = torch.nonzero((di.counts5 != ref5).any(dim=1)).flatten()
        if b.numel():
            bad += 1
            print("  ", tag, it, b.numel(), "bad rows", b[:6].tolist(), "batch", (b[:6] // 32).tolist(), "cta", ((b[:6] // 32) % 148).tolist(), "lane-win", (b[:6] % 32).tolist())
    print(tag, "bad runs", bad, "of", n)
rounds("scan only")
for it in range(3):
    r = bench.hot_path_step(dg, di, d, None)
torch.cuda.synchronize()
rounds("after test stages")
del ascii_d
torch.cuda.empty_cache()
rounds("ascii freed")
# same call as the probe: fresh outputs
o5 = torch.empty_like(ref5); o3 = torch.empty_like(ref3)
ws = kernels.scan_workspace(dg, ref5.shape[0])
bad = 0
for it in range(8):
    kernels.count_contexts_fused53(dg, di.win_chrom, di.win_start, di.win_end, out5=o5, out3=o3, workspace=ws, tile_window=10000)
    torch.cuda.synchronize()
    bad += int(not torch.equal(o5, ref5))
print("probe-style bad runs", bad)
bad = 0
for it in range(8):
    kernels.count_contexts_fused53(dg, di.win_chrom, di.win_start, di.win_end, out5=o5, out3=o3, workspace=ws, tile_window=0)
    torch.cuda.synchronize()
    bad += int(not torch.equal(o5, ref5))
print("no-hint bad runs", bad)
