"""Parity at full size (test infrastructure, like everything under oracle/): a random sample of the 3.1 Gb
workload is recomputed on the CPU from regenerated slices of the synthetic genome -- the generator is a pure
function of the global position, so any window can be rebuilt without the other 3.1 Gb -- and compared with what
the GPU path produced for the whole genome.

Checked, each against the C oracle (dig_oracle.py, which is pinned to the reference's own outputs by the goldens):
  * pentanucleotide and trinucleotide rows of `n_windows` windows drawn uniformly over ALL windows, so chromosomes
    whose global offset exceeds 2^31 are hit (sequence_tools.py:65-128), bit-exact;
  * the mutation contexts (K3) of every SNV that falls into those windows (sequence_tools.py:130-178), exact;
  * for `n_genes` genes: MU / SIGMA and the four P columns of the gene pretrain (genic_driver_tools.py:86-168) from
    oracle-counted window rows, rel. 1e-9; ALPHA / THETA against the scale factor the run used; the 13 NB p-values
    and the Fisher combination from the run's own (k, ALPHA, THETA, Pi) (transfer_tools.py:394-456), 1e-6 in log10.
Only tests/, bench.py's self-check (outside the timed region) and smoke() may import this.
"""
import numpy as np

from . import dig_oracle as orc

CLASSES = ("SYN", "MIS", "NONS", "SPL", "TRUNC", "NONSYN")


def _mini_genome(lengths, chrom_off, seed, n_frac16, chrom, start, end, pad):
    """Concatenated slices [start - pad, end + pad) (clipped to the chromosome) of the synthetic genome, one
    mini-chromosome per region.  Returns (seq, off, len, new_start, new_end)."""
    seqs, off, ln, ns, ne = [], [], [], [], []
    pos = 0
    for c, s, e in zip(chrom, start, end):
        L = int(lengths[c])
        a, b = max(int(s) - pad, 0), min(int(e) + pad, L)
        b = max(b, a)
        seqs.append(orc.synth_genome(int(chrom_off[c]) + a, b - a, seed, n_frac16))
        off.append(pos)
        ln.append(b - a)
        ns.append(int(s) - a)
        ne.append(int(e) - a)
        pos += (b - a + 127) // 128 * 128
    seq = np.full(max(pos, 1), ord("N"), dtype=np.uint8)
    for o, s in zip(off, seqs):
        seq[o:o + len(s)] = s
    return seq, np.array(off, dtype=np.int64), np.array(ln, dtype=np.int64), np.array(ns, dtype=np.int64), \
        np.array(ne, dtype=np.int64)


def window_rows(lengths, chrom_off, seed, wins, idx, n_frac16=16):
    """Oracle pentanucleotide / trinucleotide rows of windows `idx`.  A slice that does not start at position 0 of
    its chromosome keeps a 2-base lead, so the reference's START == 0 rule applies exactly where it does in full."""
    w = wins[idx]
    seq, off, ln, ns, ne = _mini_genome(lengths, chrom_off, seed, n_frac16, w[:, 0], w[:, 1], w[:, 2], 2)
    rc = np.arange(len(idx), dtype=np.int32)
    c5, _ = orc.count_regions(seq, off, ln, rc, ns, ne, 2, 2)
    c3, _ = orc.count_regions(seq, off, ln, rc, ns, ne, 1, 1)
    return c5, c3, (seq, off, ln, ns, ne)


def check_sample(lengths, chrom_off, seed, window, d, got, n_windows=500, n_genes=200, rng_seed=0, n_frac16=16,
                 win_pool=None, gene_pool=None):
    """`d`: the host workload dict of bench.build_workload; `got`: host arrays of the GPU run --
    counts5 / counts3 as callables idx -> rows (so that only the sampled rows leave the device), ctx (all mutations),
    per-gene columns MU, SIGMA, Pi_SYN .., ALPHA, THETA, OBS_*, PVAL_*, d_pr, sums, n_syn.
    Returns {"windows": N, "genes": M, "mutations": K, "ok": bool, "detail": str}."""
    rng = np.random.default_rng(rng_seed)
    wins = d["wins"]
    lengths = np.asarray(lengths, dtype=np.int64)
    chrom_off = np.asarray(chrom_off, dtype=np.int64)
    detail = []
    # ---- windows, uniformly over the genome (the last chromosomes lie beyond 2^31 in global coordinates)
    # (win_pool / gene_pool: a range-sharded rank holds the rows of its own windows and genes only)
    pool = np.arange(len(wins)) if win_pool is None else np.asarray(win_pool, dtype=np.int64)
    idx = np.sort(rng.choice(pool, size=min(n_windows, len(pool)), replace=False))
    hi_off = int((chrom_off[wins[idx, 0]] + wins[idx, 1] > (1 << 31)).sum())
    c5, c3, (seq, off, ln, ns, ne) = window_rows(lengths, chrom_off, seed, wins, idx, n_frac16)
    g5, g3 = np.asarray(got["counts5"](idx), dtype=np.int64), np.asarray(got["counts3"](idx), dtype=np.int64)
    bad5 = np.flatnonzero((g5 != c5).any(axis=1))
    bad3 = np.flatnonzero((g3 != c3).any(axis=1))
    if bad5.size or bad3.size:
        detail.append("window rows differ: K=1024 %s, K=64 %s" % (idx[bad5[:5]].tolist(), idx[bad3[:5]].tolist()))
    # ---- mutation contexts inside the sampled windows
    n_mut_checked = 0
    key = (wins[:, 0].astype(np.int64) << 40) | wins[:, 1]
    mk = (d["m_chrom"].astype(np.int64) << 40) | d["m_pos"]
    mw = np.searchsorted(key, mk, side="right") - 1
    where = np.searchsorted(idx, mw)
    where = np.clip(where, 0, len(idx) - 1)
    inside = (idx[where] == mw) & (d["m_pos"] >= wins[mw, 1] + 1) & (d["m_pos"] < wins[mw, 2] - 1) & (d["m_pos"] >= 1)
    sel = np.flatnonzero(inside)
    if sel.size:
        j = where[sel]                                                    # mini-chromosome of each mutation
        rel = d["m_pos"][sel] - wins[idx[j], 1] + ns[j]
        order = np.argsort(j, kind="stable")                              # grouped by chromosome, file order inside
        want = orc.mutation_contexts(seq, off, ln, j[order].astype(np.int32), rel[order], d["m_ref"][sel][order], 1, 1)
        have = np.asarray(got["ctx"])[sel][order]
        # the same-START reuse quirk looks at the previous row of the chromosome group: rows whose predecessor in the
        # FULL file is outside the sample could differ only if they share START with it; compare where START is unique
        pos_s = d["m_pos"][sel][order]
        uniq = np.ones(len(pos_s), dtype=bool)
        uniq[1:] &= ~((pos_s[1:] == pos_s[:-1]) & (j[order][1:] == j[order][:-1]))
        full_prev = np.zeros(len(sel), dtype=bool)
        full_prev[:] = (sel > 0) & (d["m_pos"][np.maximum(sel - 1, 0)] == d["m_pos"][sel]) & \
            (d["m_chrom"][np.maximum(sel - 1, 0)] == d["m_chrom"][sel])
        uniq &= ~full_prev[order]
        n_mut_checked = int(uniq.sum())
        badm = np.flatnonzero((want != have) & uniq)
        if badm.size:
            detail.append("%d mutation contexts differ (first rows %s)" % (badm.size, sel[order][badm[:5]].tolist()))
    # ---- genes: pretrain columns from oracle-counted window rows, then the test itself
    E = len(d["g_ptr"]) - 1
    gpool = np.arange(E) if gene_pool is None else np.asarray(gene_pool, dtype=np.int64)
    gid = np.sort(rng.choice(gpool, size=min(n_genes, len(gpool)), replace=False))
    ptr, bs, be = d["g_ptr"], d["g_bs"], d["g_be"]
    need = set()
    for g in gid:
        c = int(d["g_chrom"][g])
        for w0 in orc.ideal_overlaps(bs[ptr[g]:ptr[g + 1]], be[ptr[g]:ptr[g + 1]], window):
            need.add((c, int(w0)))
    need = sorted(need)
    wkey = {(int(c), int(s)): i for i, (c, s, e) in enumerate(wins)}
    widx = np.array([wkey[k] for k in need], dtype=np.int64)
    _, t3, _ = window_rows(lengths, chrom_off, seed, wins, widx, n_frac16)
    local = {k: i for i, k in enumerate(need)}
    sub_ptr = np.concatenate([[0], np.cumsum(ptr[gid + 1] - ptr[gid])])
    take = np.concatenate([np.arange(ptr[g], ptr[g + 1]) for g in gid])
    ref = orc.gene_transfer(d["g_chrom"][gid], d["g_strand"][gid], sub_ptr, bs[take], be[take], d["L"][gid], window, local,
                            t3, d["y_pred"][widx], d["std"][widx], d["y_true"][widx], d["flag"][widx], got["d_pr"])

    def rel_err(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        r[(a == b) | (np.isnan(a) & np.isnan(b))] = 0.0
        return float(np.max(r)) if r.size else 0.0

    worst = 0.0
    for col, rcol in (("MU", "MU"), ("SIGMA", "SIGMA"), ("Pi_SYN", "P_SILENT"), ("Pi_MIS", "P_MIS"), ("Pi_NONS", "P_NONS"),
                      ("Pi_SPL", "P_SPLICE")):
        worst = max(worst, rel_err(np.asarray(got[col])[gid], ref[rcol]))
    if worst > 1e-9:
        detail.append("gene pretrain rel. error %.3g > 1e-9" % worst)
    mu, sg = np.asarray(got["MU"])[gid], np.asarray(got["SIGMA"])[gid]
    alpha, theta0 = orc.normal_params_to_gamma(mu, sg)
    cj = float(got["n_syn"]) / float(got["sums"][0])                      # transfer_tools.py:809-812
    e_par = max(rel_err(np.asarray(got["ALPHA"])[gid], alpha), rel_err(np.asarray(got["THETA"])[gid], theta0 * cj))
    if e_par > 1e-9:
        detail.append("ALPHA / THETA rel. error %.3g > 1e-9" % e_par)
    worst_p = 0.0
    al, th = np.asarray(got["ALPHA"])[gid], np.asarray(got["THETA"])[gid]
    pis = {"SYN": np.asarray(got["Pi_SYN"])[gid], "MIS": np.asarray(got["Pi_MIS"])[gid], "NONS": np.asarray(got["Pi_NONS"])[gid],
           "SPL": np.asarray(got["Pi_SPL"])[gid]}
    pis["TRUNC"] = np.asarray(got["Pi_TRUNC"])[gid]
    pis["NONSYN"] = np.asarray(got["Pi_NONSYN"])[gid]
    obs = {"SYN": got["OBS_SYN"], "MIS": got["OBS_MIS"], "NONS": got["OBS_NONS"], "SPL": got["OBS_SPL"]}
    obs = {k: np.asarray(v)[gid] for k, v in obs.items()}
    obs["TRUNC"] = obs["NONS"] + obs["SPL"]
    obs["NONSYN"] = obs["MIS"] + obs["TRUNC"]
    n_p = 0
    for c in CLASSES:
        _, wp = orc.burden_test(obs[c], al, th, pis[c])
        gp = np.asarray(got["PVAL_%s_BURDEN" % c])[gid]
        ok = np.isfinite(wp) & np.isfinite(gp) & (wp > 0) & (gp > 0)
        if ok.any():
            worst_p = max(worst_p, float(np.max(np.abs(np.log10(gp[ok]) - np.log10(wp[ok])))))
        n_p += int(ok.sum())
    fp = orc.fisher2(np.asarray(got["PVAL_TRUNC_BURDEN"])[gid], np.asarray(got["PVAL_INDEL_BURDEN"])[gid])   # transfer_tools.py:861
    gm = np.asarray(got["PVAL_MUT_BURDEN"])[gid]
    ok = np.isfinite(fp) & np.isfinite(gm) & (fp > 0) & (gm > 0)
    if ok.any():
        worst_p = max(worst_p, float(np.max(np.abs(np.log10(gm[ok]) - np.log10(fp[ok])))))
    if worst_p > 1e-6:
        detail.append("p-values differ by %.3g in log10 (> 1e-6)" % worst_p)
    return {"windows": int(len(idx)), "windows_beyond_2^31": hi_off, "mutations": n_mut_checked, "genes": int(len(gid)),
            "p_values": n_p, "max_rel_err_pretrain": worst, "max_dlog10_p": worst_p, "ok": not detail,
            "detail": "; ".join(detail)}


def check_all_windows(lengths, chrom_off, seed, wins, rows5, rows3, totals5=None, totals3=None, lo=0, hi=None, n_frac16=16):
    """EVERY window lo .. hi of the tiling (all of them by default) against the C oracle, one chromosome at a time: the
    chromosome is regenerated on the host from the global position alone, counted for both context sizes
    (sequence_tools.py:65-128), and compared bit for bit with the GPU rows that the callables rows5 / rows3
    (a, b) -> [b - a, K] host arrays return.  totals5 / totals3 (host arrays, optional): the genome-wide sums
    (DigPreprocess.py:59) of the same windows.  Returns {"windows": n, "ok": bool, "detail": str}."""
    wins = np.asarray(wins)
    hi = len(wins) if hi is None else hi
    lengths = np.asarray(lengths, dtype=np.int64)
    tot = {2: np.zeros(1024, dtype=np.int64), 1: np.zeros(64, dtype=np.int64)}
    detail, n_rows = [], 0
    for c in np.unique(wins[lo:hi, 0]):
        sel = lo + np.flatnonzero(wins[lo:hi, 0] == c)
        a, b = int(sel[0]), int(sel[-1]) + 1
        if b - a != len(sel):
            return {"windows": n_rows, "ok": False, "detail": "windows of chromosome %d are not contiguous" % c}
        n = int(lengths[c])
        seq = orc.synth_genome(int(chrom_off[c]), n, seed, n_frac16)
        off0, ln, rc = np.zeros(1, dtype=np.int64), np.array([n], dtype=np.int64), np.zeros(len(sel), dtype=np.int32)
        for nu, rows in ((2, rows5), (1, rows3)):
            if rows is None:
                continue
            want, _ = orc.count_regions(seq, off0, ln, rc, wins[sel, 1], wins[sel, 2], nu, nu)
            bad = np.flatnonzero((np.asarray(rows(a, b)) != want).any(axis=1))
            if bad.size:
                detail.append("chromosome %d, k = %d: %d rows differ, first window %d" % (c, 2 * nu + 1, bad.size, a + bad[0]))
            tot[nu] += want.sum(axis=0)
        n_rows += len(sel)
    if totals5 is not None and not np.array_equal(np.asarray(totals5, dtype=np.int64), tot[2]):
        detail.append("pentanucleotide totals differ")
    if totals3 is not None and not np.array_equal(np.asarray(totals3, dtype=np.int64), tot[1]):
        detail.append("trinucleotide totals differ")
    return {"windows": int(n_rows), "ok": not detail and n_rows == hi - lo, "detail": "; ".join(detail)}
