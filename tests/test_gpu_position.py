"""GPU parity tests for K8, the per-position / per-bin hotspot test, through the C ABI:
against the golden vectors of the unmodified reference (apply_nb_to_region, nb_pvalue_exact) and, at larger
scale, against the CPU oracle.  Tolerances: |dlog10 p| <= 1e-6, expectations / probabilities rel. 1e-9."""
import numpy as np
import pytest
import torch

from conftest import golden, assert_pvals_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _close(got, want, rtol=1e-9):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert np.allclose(got[m], want[m], rtol=rtol, atol=0.0), np.abs(got[m] / np.where(want[m] == 0, 1, want[m]) - 1).max()


def test_nb_pvalue_exact_golden_grid(dev):
    from digdriver_b200 import kernels
    z = golden("position")
    got = kernels.nb_pvalue_exact(z["ex_k"], z["ex_alpha"], z["ex_p"], dev).cpu().numpy()
    assert_pvals_close(got, z["ex_pval"])


def test_position_test_golden(dev):
    """Every golden region (N runs, chromosome end, START == 0 quirk, single-base regions, recurrent positions,
    bins of 1 / 7 / 50) in ONE launch per (context size, bin size), compared with the reference's own output."""
    from digdriver_b200 import kernels
    from digdriver_b200.genome import Genome, DeviceGenome
    z = golden("position")
    dg = DeviceGenome.from_genome(Genome(["chr1"], [z["seq"]]), dev)
    muts = z["mut_start"]
    cases = z["cases"]
    for (u, d) in ((2, 2), (1, 1)):
        for binsize in (1, 50, 7):
            sel = [i for i, c in enumerate(cases) if (int(c[0]), int(c[1]), int(c[2])) == (u, d, binsize)]
            c = cases[sel]
            out = kernels.position_test(dg, np.zeros(len(sel), dtype=np.int32), c[:, 3].astype(np.int64),
                                        c[:, 4].astype(np.int64), c[:, 5], c[:, 6], z["s_prob_%d_%d" % (u, d)],
                                        np.zeros(len(muts), dtype=np.int32), muts, n_up=u, n_down=d, binsize=binsize)
            ptr = out["bin_ptr"]
            for j, ci in enumerate(sel):
                a, b = int(ptr[j]), int(ptr[j + 1])
                assert b - a == int(c[j, 7]), (ci, b - a, c[j])
                assert np.array_equal(out["obs"][a:b].cpu().numpy(), z["obs_%d" % ci])
                assert np.array_equal(out["pos"][a:b].cpu().numpy(), z["pos_%d" % ci])
                _close(out["pt"][a:b].cpu().numpy(), z["pt_%d" % ci])
                _close(out["exp"][a:b].cpu().numpy(), z["exp_%d" % ci])
                assert_pvals_close(out["pval"][a:b].cpu().numpy(), z["pval_%d" % ci])


def test_position_test_windows_vs_oracle(dev, oracle):
    """A 2 Mb chromosome tiled in 10 kb windows (200 regions, 2 M per-position p-values, 40 k SNVs with hotspots)
    against the oracle, plus the 50-bp binned variant."""
    from digdriver_b200 import kernels
    from digdriver_b200.genome import DeviceGenome, tile_windows
    lengths = np.array([2_000_003], dtype=np.int64)
    dg = DeviceGenome.synthetic(["chr1"], lengths, seed=77, device=dev)
    seq = oracle.synth_genome(0, int(lengths[0]), 77)
    wins = tile_windows([0], lengths, 10_000)
    rng = np.random.default_rng(5)
    muts = np.concatenate([rng.integers(0, lengths[0], 40_000), np.repeat(rng.integers(0, lengths[0], 30), 12)])
    s_prob = rng.lognormal(np.log(1e-6), 1.0, 1024)
    mu = rng.gamma(2.0, 10.0, len(wins)) + 0.1
    sigma = mu * rng.uniform(0.05, 0.5, len(wins))
    for binsize, step in ((1, 23), (50, 1)):
        out = kernels.position_test(dg, wins[:, 0], wins[:, 1], wins[:, 2], mu, sigma, s_prob,
                                    np.zeros(len(muts), dtype=np.int32), muts, n_up=2, n_down=2, binsize=binsize)
        ptr = out["bin_ptr"]
        pv, pt, ob = out["pval"].cpu().numpy(), out["pt"].cpu().numpy(), out["obs"].cpu().numpy()
        assert int(ob.sum()) == int(((muts >= 2) & (muts < wins[-1, 2])).sum())
        for r in range(0, len(wins), step):
            w = wins[r]
            sel = muts[(muts >= w[1] - 5) & (muts < w[2] + 5)]
            wpv, wpos, wobs, wexp, wpt = oracle.position_test(seq, int(w[1]), int(w[2]), mu[r], sigma[r], s_prob, sel,
                                                              2, 2, binsize)
            a, b = int(ptr[r]), int(ptr[r + 1])
            assert b - a == len(wpv)
            assert np.array_equal(ob[a:b], wobs)
            _close(pt[a:b], wpt)
            assert_pvals_close(pv[a:b], wpv)


def test_nb_model_api_matches_reference_golden(dev):
    """The reference-facing functions (nb_model.nb_model / apply_nb_to_region / nb_pvalue_exact and
    sequence_tools.base_probabilities_by_region) with the reference's own argument conventions."""
    import pandas as pd
    from digdriver_b200.genome import Genome
    from digdriver_b200.sequence_model import nb_model as nb, sequence_tools as st
    z = golden("position")
    g = Genome(["chr1"], [z["seq"]])
    df_mut = pd.DataFrame({"CHROM": "1", "START": z["mut_start"]})
    cases = z["cases"]
    for u, d in ((2, 2), (1, 1)):
        S = dict(zip(st.mk_context_sequences(u, d).keys(), z["s_prob_%d_%d" % (u, d)]))
        sel = [i for i, c in enumerate(cases) if (int(c[0]), int(c[1]), int(c[2])) == (u, d, 50)]
        c = cases[sel]
        idx = np.stack([np.ones(len(sel)), c[:, 3], c[:, 4]], axis=1).astype(np.int64)
        df = nb.nb_model(S, idx, c[:, 5], c[:, 6], df_mut, g, n_up=u, n_down=d, binsize=50)
        assert list(df.columns) == ['CHROM', 'POS', 'OBS', 'EXP', 'PVAL', 'Pi', 'MU', 'SIGMA', 'REGION']
        want_p = np.concatenate([z["pval_%d" % ci] for ci in sel])
        want_pos = np.concatenate([z["pos_%d" % ci] for ci in sel])
        assert np.array_equal(df.POS.values, want_pos)
        assert_pvals_close(df.PVAL.values, want_p)
        assert df.REGION.iloc[0] == "1:%d-%d" % (idx[0, 1], idx[0, 2])
        # one region, per position
        ci = [i for i, cc in enumerate(cases) if (int(cc[0]), int(cc[1]), int(cc[2])) == (u, d, 1)][2]
        cc = cases[ci]
        pv, pos, obs, exps, pts = nb.apply_nb_to_region(1, int(cc[3]), int(cc[4]), cc[5], cc[6], S, df_mut, g,
                                                        n_up=u, n_down=d, binsize=1)
        assert np.array_equal(obs, z["obs_%d" % ci]) and np.array_equal(pos, z["pos_%d" % ci])
        assert_pvals_close(pv, z["pval_%d" % ci])
        _close(np.array(pts), z["pt_%d" % ci])
        probs, poss = st.base_probabilities_by_region(g, S, "chr1", int(cc[3]), int(cc[4]), n_up=u, n_down=d)
        _close(probs, z["pt_%d" % ci])
        assert np.array_equal(poss, z["pos_%d" % ci].astype(np.int64))
    got = nb.nb_pvalue_exact(z["ex_k"], z["ex_alpha"], z["ex_p"])
    assert_pvals_close(got, z["ex_pval"])
    assert isinstance(nb.nb_pvalue_exact(3, 2.0, 0.4), float)
