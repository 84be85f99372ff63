import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
dev = torch.device("cuda:0")
total = float(sys.argv[1]) if len(sys.argv) > 1 else 3.1e9
dg, ascii_d, d = bench.build_workload(total, seed=1, device=dev)
del ascii_d
di = bench.DeviceInputs(d, dev)
r1 = bench.hot_path_step(dg, di, d, None)
torch.cuda.synchronize()
c5a, c3a = di.counts5.clone(), di.counts3.clone()
t5a, t3a = di.totals5.clone(), di.totals3.clone()
print('totals3 vs column sums:', bool(torch.equal(t3a, c3a.sum(dim=0, dtype=torch.int64))), 'totals5:', bool(torch.equal(t5a, c5a.sum(dim=0, dtype=torch.int64))))
r1 = {k: v.clone() for k, v in r1.items()}
for rep in range(3):
    r2 = bench.hot_path_step(dg, di, d, None)
    torch.cuda.synchronize()
    print("rep", rep, "totals5 equal", bool(torch.equal(t5a, di.totals5)), "totals3 equal", bool(torch.equal(t3a, di.totals3)), "t3 vs colsum", bool(torch.equal(di.totals3, di.counts3.sum(dim=0, dtype=torch.int64))))
    print("rep", rep, "counts5 equal", bool(torch.equal(c5a, di.counts5)), "counts3 equal", bool(torch.equal(c3a, di.counts3)))
    for k in r1:
        a, b = r1[k], r2[k]
        if a.dtype.is_floating_point:
            same = torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
        else:
            same = torch.equal(a, b)
        if not same:
            diff = torch.nonzero((torch.nan_to_num(a.double(), nan=-1.0) != torch.nan_to_num(b.double(), nan=-1.0)).flatten()).flatten()
            print("   ", k, "differs at", diff.numel(), "entries, first", diff[:5].tolist(), a.flatten()[diff[:3]].tolist(), b.flatten()[diff[:3]].tolist())
st = bench.GraphedStep(dg, di, d, None, dev)
r3 = st.step()
torch.cuda.synchronize()
for k in r1:
    a, b = r1[k], r3[k]
    same = torch.equal(torch.nan_to_num(a.double(), nan=-1.0), torch.nan_to_num(b.double(), nan=-1.0))
    if not same:
        diff = torch.nonzero((torch.nan_to_num(a.double(), nan=-1.0) != torch.nan_to_num(b.double(), nan=-1.0)).flatten()).flatten()
        print("graph:", k, "differs at", diff.numel(), "first", diff[:5].tolist(), a.flatten()[diff[:3]].tolist(), b.flatten()[diff[:3]].tolist())
print("graph error:", st.error)
