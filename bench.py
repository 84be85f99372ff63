#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200, measured as the driver contract asks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bases G]

Workload (BASELINE.json configs[1]): synthetic hg19-sized genome (3.1 Gb, 22 chromosomes, ~3 % N,
50 % lower case), 10 kb windows (310 k), pentanucleotide AND trinucleotide context maps with genome
totals, sequence model from 1 M SNVs, 20 k-gene CDS pretrain + observed counts + NB burden test
(13 p-values + Fisher per gene).  One "step" = one pass of that whole path.

  value  : genome bases scanned per second over the whole step, inputs resident in HBM;
  e2e    : the same step through the host-buffer API: ASCII genome, mutations and gene tables start
           in pinned host memory, counts and p-values end in host memory (copies inside the timing);
  roofline: the dominant kernel (pentanucleotide scan), algorithmic bytes / its CUDA-event time,
           against the measured HBM peak in MEASURED_PEAKS.json;
  cpu_baseline / --impl reference: the reference's own algorithm (pure-Python per-base loop under
           multiprocessing.Pool, SciPy p-values) timed on this box's host cores on a bounded sample.

Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank holds one hg19-sized shard of an
N x 3.1 Gb genome plus its own 20 k genes; NCCL all-reduces the context totals, substitution counts and
scale-factor sums and gathers the per-gene results on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WINDOW = 10_000
N_GENES = 20_000
N_MUT = 1_000_000
N_SAMPLES = 200
B_PER_BASE_K1024 = 0.25 + 0.125 + 4.0 * 1024 / WINDOW      # 0.7846 B/base
B_PER_BASE_FUSED = 0.25 + 0.125 + 4.0 * (1024 + 64) / WINDOW  # both tables written by one pass: 0.8102
B_PER_BASE_K64 = 0.25 + 0.125 + 4.0 * 64 / WINDOW


_JSON_OUT = None


def emit_json(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=float, default=3.1e9, help="genome size per GPU (debug override)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-sample", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# host placement
# ------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and therefore first-touch its pinned host buffers) on the CPUs of the NUMA node its
    GPU hangs off, so that the e2e leg's host<->device copies of eight ranks do not all cross the socket interconnect.
    Returns a short description for the JSON line; never raises."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "numa node unknown for %s" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no allowed cpu on numa node %d" % node
        os.sched_setaffinity(0, cpus)
        return "gpu %s -> numa node %d (%d cpus)" % (bdf, node, len(cpus))
    except Exception as exc:
        return "not bound: %r" % (exc,)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []          # (host time of arrival, csv line)
        self.t_begin = None
        self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for (t_arr, ln) in self.lines:
            if self.t_begin is not None and not (self.t_begin <= t_arr <= (self.t_end or t_arr) + 0.06):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload construction
# ------------------------------------------------------------------------------------------------

def build_workload(total_bases, seed, device):
    """All inputs of one rank's shard: device genome (+ASCII copy for the e2e leg), windows, region
    parameters, genes with L, mutations.  Returns a dict of HOST arrays plus the device genome."""
    import torch
    from digdriver_b200 import genome as G, kernels, pipeline
    lengths = G.hg19_like_lengths(int(total_bases))
    names = ["chr%d" % (i + 1) for i in range(len(lengths))]
    dg, ascii_d = G.DeviceGenome.synthetic(names, lengths, seed=seed, device=device, return_ascii=True)
    wins = G.tile_windows(np.arange(len(lengths)), lengths, WINDOW)
    n_win = len(wins)
    y_pred, std, y_true, flag = pipeline.synth_region_params(n_win, seed + 100)
    g_chrom, g_strand, g_ptr, g_bs, g_be = pipeline.synth_genes(N_GENES, lengths, WINDOW, seed + 200)
    owner = np.repeat(np.arange(N_GENES), np.diff(g_ptr))
    # L from the CDS content itself (K4, strand-aware; half-open block ends = inclusive end + 1)
    bc, _ = kernels.count_contexts(dg, g_chrom[owner], g_bs, g_be + 1, 1, 1, strand=g_strand[owner])
    L = pipeline.synth_gene_L(bc.cpu().numpy(), g_ptr, seed + 300)
    # mutations: 30 % coding SNVs, 63 % non-coding SNVs, 7 % indels; Zipf-weighted samples
    rng = np.random.default_rng(seed + 400)
    n_cds = int(0.30 * N_MUT)
    n_ind = int(0.07 * N_MUT)
    n_nc = N_MUT - n_cds - n_ind
    bsize = (g_be - g_bs + 1).astype(np.float64)
    bidx = rng.choice(len(g_bs), size=n_cds + n_ind // 2, p=bsize / bsize.sum())
    cds_pos = g_bs[bidx] + (rng.random(len(bidx)) * bsize[bidx]).astype(np.int64)
    cds_chrom = g_chrom[owner[bidx]]
    cds_gene = owner[bidx].astype(np.int32)
    nc_n = n_nc + (n_ind - n_ind // 2)
    nc_chrom = rng.choice(len(lengths), size=nc_n, p=lengths / lengths.sum()).astype(np.int32)
    nc_pos = (rng.random(nc_n) * (lengths[nc_chrom] - 10)).astype(np.int64) + 5
    chrom = np.concatenate([cds_chrom, nc_chrom]).astype(np.int32)
    pos = np.concatenate([cds_pos, nc_pos])
    gene = np.concatenate([cds_gene, np.full(nc_n, -1, dtype=np.int32)])
    is_indel = np.zeros(len(pos), dtype=bool)
    is_indel[n_cds:n_cds + n_ind // 2] = True
    is_indel[len(cds_pos) + n_nc:] = True
    cls = np.full(len(pos), 255, dtype=np.uint8)
    cls[:n_cds] = rng.choice(4, size=n_cds, p=[0.23, 0.68, 0.04, 0.05]).astype(np.uint8)
    cls[is_indel & (gene >= 0)] = 4
    w = 1.0 / np.arange(1, N_SAMPLES + 1)
    sample = rng.choice(N_SAMPLES, size=len(pos), p=w / w.sum()).astype(np.int32)
    order = np.lexsort((pos, chrom))
    chrom, pos, gene, is_indel, cls, sample = (a[order] for a in (chrom, pos, gene, is_indel, cls, sample))
    # REF = the genome base (device gather on the ASCII copy; setup only), ALT = another base
    gpos = torch.from_numpy(dg.chrom_off[chrom] + pos).to(device)
    up = (ascii_d[gpos] & 0xDF).cpu().numpy()
    ref = np.full(len(pos), 255, dtype=np.uint8)
    for code, ch in enumerate(b"ACGT"):
        ref[up == ch] = code
    ref[is_indel] = 255
    alt = ((ref.astype(np.int64) + 1 + rng.integers(0, 3, len(pos))) % 4).astype(np.uint8)
    alt[ref > 3] = 255
    d = dict(lengths=lengths, names=names, wins=wins, y_pred=y_pred, std=std, y_true=y_true, flag=flag,
             g_chrom=g_chrom, g_strand=g_strand, g_ptr=g_ptr, g_bs=g_bs, g_be=g_be, L=L,
             m_chrom=chrom, m_pos=pos, m_ref=ref, m_alt=alt, m_gene=gene, m_cls=cls, m_sample=sample,
             n_syn=int(((cls == 0) & (gene >= 0)).sum()))
    return dg, ascii_d, d


class DeviceInputs:
    """Device-resident copies of one shard's inputs (the `value` leg starts from these)."""

    def __init__(self, d, device):
        import torch
        from digdriver_b200 import kernels
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
        self.win_chrom = t(d["wins"][:, 0], torch.int32)
        self.win_start = t(d["wins"][:, 1], torch.int64)
        self.win_end = t(d["wins"][:, 2], torch.int64)
        self.y_pred, self.std, self.y_true = (t(d[k], torch.float64) for k in ("y_pred", "std", "y_true"))
        self.flag = t(d["flag"].astype(np.uint8), torch.uint8)
        off, wmap = kernels.build_window_map(d["wins"][:, 0], d["wins"][:, 1], WINDOW, len(d["lengths"]))
        self.wmap_off, self.wmap = t(off, torch.int64), t(wmap, torch.int32)
        self.g_chrom, self.g_strand = t(d["g_chrom"], torch.int32), t(d["g_strand"], torch.int8)
        self.g_ptr, self.g_bs, self.g_be = (t(d[k], torch.int64) for k in ("g_ptr", "g_bs", "g_be"))
        self.L = t(d["L"], torch.float64)
        self.m_chrom, self.m_pos = t(d["m_chrom"], torch.int32), t(d["m_pos"], torch.int64)
        self.m_ref, self.m_alt = t(d["m_ref"], torch.uint8), t(d["m_alt"], torch.uint8)
        self.m_gene, self.m_sample = t(d["m_gene"], torch.int32), t(d["m_sample"], torch.int32)
        self.m_cls = t(d["m_cls"], torch.uint8)
        n_win = len(d["wins"])
        self.max_span = kernels.element_max_span(d["g_ptr"], d["g_bs"], d["g_be"], WINDOW)
        self.counts5 = torch.empty((n_win, 1024), dtype=torch.int32, device=device)
        self.counts3 = torch.empty((n_win, 64), dtype=torch.int32, device=device)
        self.totals5 = torch.zeros(1024, dtype=torch.int64, device=device)
        self.totals3 = torch.zeros(64, dtype=torch.int64, device=device)
        self.side_stream = torch.cuda.Stream(device)


def scan_stage(dg, di, ev=None, lo=0, hi=None, zero=True):
    """Context maps + genome totals (K2) for windows [lo, hi): pentanucleotide and trinucleotide tables in one
    fused pass (dig_count_contexts_fused53)."""
    from digdriver_b200 import kernels
    hi = di.win_chrom.numel() if hi is None else hi
    if zero:
        di.totals5.zero_()
        di.totals3.zero_()
    if ev is not None:
        ev[0].record()
    if getattr(di, "scan_ws", None) is None or di.scan_ws_n < hi - lo:
        di.scan_ws, di.scan_ws_n = kernels.scan_workspace(dg, hi - lo), hi - lo
    kernels.count_contexts_fused53(dg, di.win_chrom[lo:hi], di.win_start[lo:hi], di.win_end[lo:hi],
                                   out5=di.counts5[lo:hi], out3=di.counts3[lo:hi], totals5=di.totals5,
                                   totals3=di.totals3, workspace=di.scan_ws, tile_window=WINDOW)
    if ev is not None:
        ev[1].record()


def test_stage(dg, di, d, dist_ctx, sink=None):
    """Sequence model (K3), gene pretrain (K6), observed counts (K5), burden test (K7).  With `sink` the device
    status words are collected instead of read (no host synchronisation inside the stage)."""
    import torch
    from digdriver_b200 import kernels, pipeline
    dev = dg.device
    # the observed counts (K5) depend on nothing else in the stage: they run on a second stream next to the
    # K3 -> sequence model -> K6 chain (both are latency-bound and leave most SMs idle); inside the CUDA graph
    # this becomes two parallel branches
    main = torch.cuda.current_stream(dev)
    side = di.side_stream
    side.wait_stream(main)
    with torch.cuda.stream(side):
        obs, nsamp = kernels.tabulate_genes(di.m_gene, di.m_sample, di.m_cls, N_GENES, device=dev, status_sink=sink)
    ctx = kernels.mutation_contexts(dg, di.m_chrom, di.m_pos, di.m_ref, 1, 1)
    sub = kernels.substitution_counts(ctx, di.m_alt, 1, 1)
    if dist_ctx is not None:
        # genome-wide totals and substitution counts are cohort-wide quantities: all-reduce across shards
        buf = torch.cat([di.totals5, di.totals3, sub])
        dist_ctx.all_reduce_sum(buf)
        tot3, sub = buf[1024:1088], buf[1088:]
    else:
        tot3 = di.totals3
    d_pr = kernels.sequence_freq(sub.contiguous(), tot3.contiguous())
    pre = kernels.element_transfer(di.g_chrom, di.g_strand, di.g_ptr, di.g_bs, di.g_be, WINDOW, di.wmap_off,
                                   di.wmap, di.counts3, di.y_pred, di.std, di.y_true, di.flag, d_pr,
                                   L_elt=di.L, device=dev, max_span=di.max_span, status_sink=sink)
    main.wait_stream(side)
    if not torch.cuda.is_current_stream_capturing():
        obs.record_stream(main)          # allocated on the side stream, consumed on the main one
        nsamp.record_stream(main)
    res = pipeline.gene_burden_test(pre, obs, nsamp, d["n_syn"], collectives=dist_ctx)
    res["D_PR"], res["CTX"] = d_pr, ctx                  # for the parity sample (outside the timed region)
    return res


def hot_path_step(dg, di, d, dist_ctx, ev=None):
    """One pass of the whole path on device-resident inputs.  Returns the per-gene result dict."""
    scan_stage(dg, di, ev)
    return test_stage(dg, di, d, dist_ctx)


def gather_results(coll, t):
    """Per-gene results of every shard on rank 0 (the reference's pd.concat of chunk results)."""
    return coll.gather_rows(t.unsqueeze(1), sizes=[N_GENES] * coll.world)


class GraphedStep:
    """The step as it is launched in production: the scan kernel, then ONE CUDA-graph launch holding the whole
    test stage (15 small kernels + the NCCL exchanges when sharded), so the step is not bound by launch gaps.
    Falls back to eager launches if capture is not possible (reported in the JSON line)."""

    def __init__(self, dg, di, d, dist_ctx, device):
        import torch
        from digdriver_b200 import kernels
        self.dg, self.di, self.d, self.dist_ctx = dg, di, d, dist_ctx
        self.graph, self.res, self.gathered, self.sink = None, None, None, []
        self.error = None
        try:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._stage()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            kernels.check_deferred(self.sink)
            g = torch.cuda.CUDAGraph()
            from digdriver_b200 import _lib
            n0 = _lib.launch_count
            with torch.cuda.graph(g):
                self._stage()
            self.launches_per_step = 1 + (_lib.launch_count - n0)      # the scan + the captured kernels
            self.graph = g
        except Exception as exc:          # report, never hide
            self.error = repr(exc)[:200]
            self.graph = None
            self.sink = []
            torch.cuda.synchronize(device)

    def _stage(self):
        self.res = test_stage(self.dg, self.di, self.d, self.dist_ctx, sink=self.sink)
        if self.dist_ctx is not None:
            self.gathered = gather_results(self.dist_ctx, self.res["PVAL_MUT_BURDEN"])

    def step(self, ev=None):
        scan_stage(self.dg, self.di, ev)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._stage()
        return self.res

    def check(self):
        """Reads the device status words of the stage (one host synchronisation, outside the timed region)."""
        from digdriver_b200 import kernels
        kernels.check_deferred(self.sink)


def parity_sample(dg, d, di, res, seed, n_windows=500, n_genes=200):
    """Post-timing self-check at full size: ~500 windows drawn over the whole 3.1 Gb genome (chromosomes beyond 2^31 in
    global coordinates included), the SNVs inside them and ~200 genes are recomputed by the CPU oracle from
    regenerated slices of the synthetic genome and compared with this run's results (oracle/parity_sample.py)."""
    import torch
    from oracle import parity_sample as ps
    dev = dg.device
    rows = lambda t: (lambda idx: t[torch.from_numpy(np.asarray(idx)).to(dev)].cpu().numpy())
    got = {"counts5": rows(di.counts5), "counts3": rows(di.counts3), "ctx": res["CTX"].cpu().numpy(),
           "d_pr": res["D_PR"].cpu().numpy(), "sums": res["SUMS"].cpu().numpy(), "n_syn": d["n_syn"]}
    for k in ("MU", "SIGMA", "Pi_SYN", "Pi_MIS", "Pi_NONS", "Pi_SPL", "Pi_TRUNC", "Pi_NONSYN", "ALPHA", "THETA", "OBS_SYN",
              "OBS_MIS", "OBS_NONS", "OBS_SPL", "PVAL_INDEL_BURDEN", "PVAL_MUT_BURDEN"):
        got[k] = res[k].cpu().numpy()
    for c in ps.CLASSES:
        got["PVAL_%s_BURDEN" % c] = res["PVAL_%s_BURDEN" % c].cpu().numpy()
    return ps.check_sample(d["lengths"], dg.chrom_off, seed, WINDOW, d, got, n_windows=n_windows, n_genes=n_genes)


# ------------------------------------------------------------------------------------------------
# e2e leg: host buffers in, host buffers out
# ------------------------------------------------------------------------------------------------

class HostPath:
    """The step through the host-buffer API: pinned ASCII genome -> H2D (per chromosome, overlapped with
    packing) -> scans -> D2H of both context tables; mutation/gene tables H2D; p-values D2H."""

    def __init__(self, ascii_d, dg, d, device):
        import torch
        self.device = device
        self.dg = dg
        self.d = d
        n = dg.n_bases
        self.host_ascii = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        self.host_ascii.copy_(ascii_d)
        self.dev_ascii = torch.empty(n, dtype=torch.uint8, device=device)
        n_win = len(d["wins"])
        self.host_counts5 = torch.empty((n_win, 1024), dtype=torch.int32, pin_memory=True)
        self.host_counts3 = torch.empty((n_win, 64), dtype=torch.int32, pin_memory=True)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        self.host_in = {k: pin(d[k]) for k in ("y_pred", "std", "y_true", "g_chrom", "g_strand", "g_ptr", "g_bs",
                                                "g_be", "L", "m_chrom", "m_pos", "m_ref", "m_alt", "m_gene",
                                                "m_cls", "m_sample")}
        self.host_in["flag"] = pin(d["flag"].astype(np.uint8))
        self.host_in["wins"] = pin(d["wins"])
        self.host_out = torch.empty((14, N_GENES), dtype=torch.float64, pin_memory=True)
        self.host_tot = torch.empty(1024 + 64, dtype=torch.int64, pin_memory=True)
        self.copy_stream = torch.cuda.Stream(device)
        self.out_stream = torch.cuda.Stream(device)
        self.h2d_bytes = n + sum(v.numel() * v.element_size() for v in self.host_in.values())
        self.d2h_bytes = (self.host_counts5.numel() + self.host_counts3.numel()) * 4 + self.host_out.numel() * 8 + \
            (1024 + 64) * 8

    def step(self, di):
        """H2D, pack, scan and D2H are pipelined per chromosome on three streams: while chromosome c is being
        scanned, chromosome c+1 is on its way in and the count rows of chromosome c-1 are on their way out."""
        import torch
        from digdriver_b200 import _lib, kernels, pipeline
        dg, dev, d = self.dg, self.device, self.d
        main = torch.cuda.current_stream(dev)
        bounds = list(dg.chrom_off) + [dg.n_bases]
        wins = d["wins"]
        wlo = np.searchsorted(wins[:, 0], np.arange(len(dg.chrom_off)), side="left")
        whi = np.searchsorted(wins[:, 0], np.arange(len(dg.chrom_off)), side="right")
        self.copy_stream.wait_stream(main)
        self.out_stream.wait_stream(main)
        # small tables first (they are needed only by the test stage)
        with torch.cuda.stream(self.copy_stream):
            t = {k: v.to(dev, non_blocking=True) for k, v in self.host_in.items()}
            ev_tables = torch.cuda.Event()
            ev_tables.record(self.copy_stream)
        main.wait_event(ev_tables)
        w = t["wins"]
        di.win_chrom, di.win_start, di.win_end = w[:, 0].to(torch.int32), w[:, 1].contiguous(), w[:, 2].contiguous()
        di.totals5.zero_()
        di.totals3.zero_()
        for c, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            with torch.cuda.stream(self.copy_stream):
                self.dev_ascii[a:b].copy_(self.host_ascii[a:b], non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(self.copy_stream)
            main.wait_event(e_in)
            _lib.call("dig_pack_genome", self.dev_ascii.data_ptr() + int(a), int(b - a),
                      dg.packed2.data_ptr() + int(a) // 16 * 4, dg.nmask.data_ptr() + int(a) // 32 * 4, None,
                      main.cuda_stream)
            lo, hi = int(wlo[c]), int(whi[c])
            if hi > lo:
                scan_stage(dg, di, None, lo, hi, zero=False)
                e_scan = torch.cuda.Event()
                e_scan.record(main)
                with torch.cuda.stream(self.out_stream):
                    self.out_stream.wait_event(e_scan)
                    self.host_counts5[lo:hi].copy_(di.counts5[lo:hi], non_blocking=True)
                    self.host_counts3[lo:hi].copy_(di.counts3[lo:hi], non_blocking=True)
        di.y_pred, di.std, di.y_true, di.flag = t["y_pred"], t["std"], t["y_true"], t["flag"]
        di.g_chrom, di.g_strand, di.g_ptr, di.g_bs, di.g_be, di.L = (t[k] for k in ("g_chrom", "g_strand", "g_ptr",
                                                                                    "g_bs", "g_be", "L"))
        di.m_chrom, di.m_pos, di.m_ref, di.m_alt, di.m_gene, di.m_cls, di.m_sample = (
            t[k] for k in ("m_chrom", "m_pos", "m_ref", "m_alt", "m_gene", "m_cls", "m_sample"))
        sink = []
        res = test_stage(dg, di, d, None, sink=sink)
        cols = [res["PVAL_%s_BURDEN" % c] for c in pipeline.GENE_CLASSES] + \
               [res["PVAL_%s_BURDEN_SAMPLE" % c] for c in pipeline.GENE_CLASSES] + \
               [res["PVAL_INDEL_BURDEN"], res["PVAL_MUT_BURDEN"]]
        self.host_out.copy_(torch.stack(cols), non_blocking=True)
        self.host_tot.copy_(torch.cat([di.totals5, di.totals3]), non_blocking=True)
        torch.cuda.synchronize(dev)
        kernels.check_deferred(sink)
        return self.host_tot


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own algorithm on the host cores
# ------------------------------------------------------------------------------------------------

def reference_sample_step(n_proc, windows_per_proc, genes_sample, seed, pool=None):
    """One bounded sample of the same workload with the reference's algorithm: per-base pure-Python
    context counting under multiprocessing.Pool (sequence_tools.py:65-128) for BOTH context sizes, and
    the SciPy burden test (transfer_tools.py:394-456, :484-592, :709-729) for `genes_sample` genes.
    Pool start-up is not timed.  Returns (seconds_scan, bases_scanned, seconds_test, genes_tested)."""
    from oracle import dig_oracle
    n_win = n_proc * windows_per_proc
    L = n_win * WINDOW + 10
    seq = dig_oracle.synth_genome(0, L, seed).tobytes().decode()
    starts = np.arange(n_win) * WINDOW
    chroms = ["chr1"] * n_win
    jobs_all = []
    for (u, d) in ((2, 2), (1, 1)):
        # each worker receives only its own windows' sequence (the reference re-opens the FASTA per worker)
        for p in range(n_proc):
            a, b = p * windows_per_proc, (p + 1) * windows_per_proc
            lo = max(int(starts[a]) - u, 0)
            sub = {"chr1": seq[lo:int(starts[b - 1]) + WINDOW + d]}
            st = starts[a:b] - lo
            st[st == 0] = u                     # the reference's START == 0 -> n_up rule happens upstream
            jobs_all.append((sub, chroms[a:b], st, starts[a:b] - lo + WINDOW, u, d))
    t0 = time.perf_counter()
    if pool is not None:
        pool.map(dig_oracle._py_chunk, jobs_all, chunksize=1)
    else:
        for j in jobs_all:
            dig_oracle._py_chunk(j)
    t_scan = time.perf_counter() - t0
    rng = np.random.default_rng(seed)
    E = genes_sample
    mu = rng.gamma(2.0, 20.0, E)
    sigma = mu * rng.uniform(0.05, 0.5, E)
    alpha, theta = dig_oracle.normal_params_to_gamma(mu, sigma)
    t0 = time.perf_counter()
    for _ in range(13):
        pi = rng.uniform(1e-4, 0.05, E)
        k = rng.poisson(mu * pi).astype(np.float64)
        dig_oracle.burden_test(k, alpha, theta, pi)
    dig_oracle.fisher2(rng.uniform(0, 1, E), rng.uniform(0, 1, E))
    t_test = time.perf_counter() - t0
    return t_scan, n_win * WINDOW, t_test, E


def reference_value(total_bases, n_proc, seed, pool, windows_per_proc):
    """bases/s of the whole step extrapolated linearly from the bounded sample."""
    t_scan, nb, t_test, ne = reference_sample_step(n_proc, windows_per_proc, N_GENES, seed, pool)
    t_full = t_scan * (total_bases / nb) + t_test * (N_GENES / ne)
    return total_bases / t_full, t_scan, nb, t_test


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm on this box's host cores (no GPU, no torch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    n_proc = max(1, min((os.cpu_count() or 2) - 2, 64))
    n_steps = args.warmup + args.steps
    # about 3 s of pool work per step (1000 windows = 10 Mb per worker and context size), the whole run bounded
    # to ~2.5 minutes
    wpp = 1000
    pool = mp.Pool(n_proc) if n_proc > 1 else None
    vals, t0, sample = [], time.perf_counter(), ""
    for i in range(n_steps):
        v, t_scan, nb, t_test = reference_value(args.bases, n_proc, 1000 + i, pool, wpp)
        if i >= args.warmup:
            vals.append(v)
        sample = ("%d x 10 kb windows (%.1f Mb) counted for K=1024 and K=64 with the pure-Python per-base loop "
                  "under multiprocessing.Pool(%d): %.2f s; 13 SciPy NB tests + Fisher on %d genes: %.3f s; "
                  "extrapolated linearly to %.3g bases" % (nb // WINDOW, nb / 1e6, n_proc, t_scan, N_GENES, t_test,
                                                          args.bases))
        if time.perf_counter() - t0 > 150.0 and vals:
            break
    if pool is not None:
        pool.close()
    v = float(np.median(vals))
    line = {"impl": "reference", "metric": "genome_bases_scanned_per_s", "value": v, "unit": "bases/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": args.bases / v * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 bases -> int64 counts; f64 p-values", "data": "synthetic",
            "config": workload_config(args.bases, args.gpus),
            "cpu_baseline": {"value": v, "unit": "bases/s", "cores": n_proc, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json(line)


def workload_config(bases, n_gpus):
    return {"workload": "synthetic hg19-sized genome (%.3g bases per GPU, 22 chromosomes), 10 kb windows, "
                        "pentanucleotide + trinucleotide context maps with genome totals, sequence model from "
                        "1M SNVs, 20k-gene CDS pretrain + observed counts + NB burden test (13 p-values + Fisher "
                        "per gene)" % bases,
            "window": WINDOW, "genes": N_GENES, "snvs": N_MUT, "samples": N_SAMPLES,
            "parallelism": "range-sharded x%d" % n_gpus,
            "l2_policy": "inputs (1.16 GB packed genome) and outputs (1.35 GB count tables) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# main
# ------------------------------------------------------------------------------------------------

def main():
    args = parse_args()
    # Only the JSON line may appear on stdout: library banners (NCCL prints its version to stdout under torchrun)
    # and the prints of helper code are sent to stderr for the whole run.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from digdriver_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    placement = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process, not bound"
    dist_ctx = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        from digdriver_b200.sharding import Collectives
        dist_ctx = Collectives()

    clocks = ClockSampler(local_rank)
    clocks.start()                       # nvidia-smi takes a while to start: launch it before the set-up
    dg, ascii_d, d = build_workload(args.bases, seed=1 + rank, device=device)
    di = DeviceInputs(d, device)
    n_scanned = float((d["wins"][:, 2] - d["wins"][:, 1]).sum())

    def barrier():
        torch.cuda.synchronize(device)
        if dist_ctx is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- value leg: inputs resident in HBM
    clocks.mark_begin()                  # samples kept: warm-up + timed region + e2e leg (all under load)
    res_eager = hot_path_step(dg, di, d, dist_ctx)       # eager pass: status words checked, reference result
    eager_p = res_eager["PVAL_MUT_BURDEN"].clone()
    stepper = GraphedStep(dg, di, d, dist_ctx, device)
    for _ in range(max(args.warmup, 3)):
        res = stepper.step()
    barrier()
    assert torch.equal(torch.nan_to_num(res["PVAL_MUT_BURDEN"], nan=-1.0), torch.nan_to_num(eager_p, nan=-1.0)), \
        "graph replay and eager step disagree"
    ev_all = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count
    barrier()
    start.record()
    for i in range(args.steps):
        res = stepper.step(ev=ev_all[i])
    end.record()
    barrier()
    stepper.check()
    launches = _lib.launch_count - launches0
    if stepper.graph is not None:
        # kernels replayed from the graph do not pass through _lib.call: count them from the capture pass
        launches = args.steps * stepper.launches_per_step
    elapsed_ms = start.elapsed_time(end)
    k5_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_all]))
    if dist_ctx is not None:
        t = torch.tensor([elapsed_ms, k5_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, k5_ms = float(t[0]), float(t[1])
    ms_per_step = elapsed_ms / args.steps
    value = n_scanned * world / (ms_per_step * 1e-3)

    # ---- parity at full size, outside every timed region (rank 0; the oracle regenerates the slices it needs)
    parity = None
    if rank == 0 and not args.no_parity_sample:
        try:
            parity = parity_sample(dg, d, di, res, seed=1 + rank)
        except Exception as exc:      # report, never hide
            parity = {"ok": False, "detail": "parity sample failed to run: %r" % (exc,)}

    # ---- e2e leg: host buffers in and out
    e2e = None
    if not args.no_e2e:
        hp = HostPath(ascii_d, dg, d, device)
        del ascii_d
        for _ in range(2):
            hp.step(di)
        barrier()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            hp.step(di)
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        if dist_ctx is not None:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        e2e = {"value": n_scanned * world / e2e_s, "unit": "bases/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(hp.h2d_bytes), "d2h_bytes_per_step": int(hp.d2h_bytes),
               "api": "HostPath.step: pinned host ASCII genome + tables in, count tables + p-values out"}

    clocks.mark_end()
    clock_info = clocks.stop()

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, which = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = n_scanned * B_PER_BASE_FUSED / (k5_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_hex_kernel<TRI, TOT> (pentanucleotide K=1024 through hexamer pairs + "
                                          "trinucleotide K=64 window tables and genome totals in one pass)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default size, from the committed
                # ncu --set full capture (profiles/r01_scan_hex_summary.txt): 1.169 GB + 1.302 GB
                "traffic": 2.4707e9 if abs(args.bases - 3.1e9) < 1 else None,
                "algorithmic_bytes": n_scanned * B_PER_BASE_FUSED,
                "peak_source": which, "kernel_ms": k5_ms,
                "algorithmic_bytes_per_base": B_PER_BASE_FUSED,
                "note": "HBM is the roofline the contract asks for; ncu shows the kernel is bound by the shared-memory "
                        "atomic data pipe (87 % busy after halving the atomics with hexamer pairs), see DESIGN.md section 4", "share_of_step": k5_ms / ms_per_step}

    def finish():
        """Leave without tearing NCCL down: destroy_process_group() after a CUDA graph that holds NCCL kernels was
        captured can block for ever (seen at N=2), and the process is exiting anyway."""
        sys.stdout.flush()
        sys.stderr.flush()
        if dist_ctx is not None:
            stepper.graph = None
            torch.cuda.synchronize(device)
            dist.barrier()
            torch.cuda.synchronize(device)
            os._exit(0)

    if rank != 0:
        finish()
        return

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the reference's algorithm
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        # run in a fresh process (no CUDA context to fork) through the same code as --impl reference
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                              "--warmup", "1", "--bases", str(args.bases)], stdout=subprocess.PIPE, text=True)
        try:
            cpu_baseline = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as exc:      # report, never hide
            cpu_baseline = {"error": "reference arm failed: %r" % (exc,)}

    line = {"metric": "genome_bases_scanned_per_s", "value": value, "unit": "bases/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 bases -> int32 counts; f64 p-values", "data": "synthetic",
            "config": workload_config(args.bases, world),
            "elements_tested_per_s": N_GENES * world / (ms_per_step * 1e-3),
            "cuda_graph": {"test_stage_captured": stepper.graph is not None, "error": stepper.error},
            "host_placement": placement,
            "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity_sample": parity}
    emit_json(line)
    finish()


if __name__ == "__main__":
    main()
