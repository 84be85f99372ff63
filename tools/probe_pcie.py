"""Developer probe: host<->device copy rates of this box with pinned memory (alone and both directions at once)."""
import time, torch
dev = torch.device("cuda:0")
n = 700_000_000
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h, chunks=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = n // chunks
    for c in range(chunks):
        a, b = c * step, (c + 1) * step
        if h2d:
            with torch.cuda.stream(s1):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for name, a, b, ch in (("H2D alone", 1, 0, 1), ("D2H alone", 0, 1, 1), ("both at once", 1, 1, 1), ("both, 22 chunks", 1, 1, 22)):
    run(a, b, ch)
    ms = min(run(a, b, ch) for _ in range(3))
    print("%-18s %7.2f ms for %.2f GB per direction -> %.1f GB/s per direction" % (name, ms, n / 1e9, n / ms / 1e6))
