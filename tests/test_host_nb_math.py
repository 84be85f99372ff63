"""CPU tier: the FP64 device arithmetic of csrc/nb_math.cuh, compiled for the host by g++ (tests/host_nb_check.cpp),
against the reference-generated golden vectors.  The same checks run against the kernels themselves in the -m gpu
tests; this tier catches arithmetic mistakes without a GPU.  Tolerance: |dlog10 p| <= 1e-6 (BASELINE.json)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_pvals_close, golden

MODES = {"greater": 0, "greater_midp": 1, "less": 2, "less_midp": 3, "exact": 4, "midp": 5}


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hc") / "libhc.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_nb_check.cpp")])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _variant(hc, mode, k, a, p, mu=None):
    k, a, p = _c(k), _c(a), _c(p)
    mu = _c(mu) if mu is not None else None
    out = np.empty_like(k)
    hc.hc_nb_variant(MODES[mode], _p(k), _p(a), _p(p), _p(mu), ctypes.c_int64(len(k)), _p(out))
    return out


def test_midp_and_exact_on_the_existing_goldens(hc):
    z = golden("nbtest")
    k, a, p = _c(z["k"]), _c(z["alpha"]), _c(z["p"])
    out = np.empty_like(k)
    hc.hc_nb_midp(_p(k), _p(a), _p(p), ctypes.c_int64(len(k)), _p(out))
    assert_pvals_close(out, z["pval"])
    z = golden("position")
    k, a, p = _c(z["ex_k"]), _c(z["ex_alpha"]), _c(z["ex_p"])
    out = np.empty_like(k)
    hc.hc_nb_exact(_p(k), _p(a), _p(p), ctypes.c_int64(len(k)), _p(out))
    assert_pvals_close(out, z["ex_pval"])


@pytest.mark.parametrize("mode,key,with_mu", [
    ("greater", "v_greater", False), ("greater_midp", "v_greater_midp_deprecated", False), ("less", "v_less", False),
    ("less_midp", "v_less_midp", False), ("midp", "v_midp", False), ("midp", "v_midp_mu", True),
    ("exact", "v_exact_mu", True)])
def test_pvalue_conventions(hc, mode, key, with_mu):
    z = golden("variants")
    got = _variant(hc, mode, z["v_k"], z["v_alpha"], z["v_p"], z["v_mu"] if with_mu else None)
    assert_pvals_close(got, z[key])


def test_exact_variant_equals_dedicated_exact(hc):
    z = golden("variants")
    k, a, p = _c(z["v_k"]), _c(z["v_alpha"]), _c(z["v_p"])
    out = np.empty_like(k)
    hc.hc_nb_exact(_p(k), _p(a), _p(p), ctypes.c_int64(len(k)), _p(out))
    assert np.array_equal(_variant(hc, "exact", k, a, p), out, equal_nan=True)


def _secondary_frame():
    s = golden("secondary")
    d = {c[3:]: s[c] for c in s.files if c.startswith("in_")}
    for c in ("T_SYN", "MRFOLD"):
        d[c] = s["out_" + c]
    return d


def test_loglik_terms(hc):
    z, d = golden("variants"), _secondary_frame()

    def ll(kind, x, a, b=None):
        x, a = _c(x), _c(a)
        b = _c(b) if b is not None else None
        out = np.empty_like(x)
        hc.hc_loglik(kind, _p(x), _p(a), _p(b), ctypes.c_int64(len(x)), _p(out))
        return out
    cases = [(ll(0, d["OBS_MIS"], d["ALPHA"], d["THETA"] * d["Pi_MIS"]), z["ll_nb"]),
             (ll(1, d["OBS_MIS"], d["ALPHA"] * d["THETA"] * d["Pi_MIS"]), z["ll_pois"]),
             (ll(1, d["OBS_NONS"], d["OBS_NONS"]), z["ll_pois_self"]),
             (ll(2, d["T_SYN"], d["ALPHA"], d["THETA"] * d["Pi_SYN"] * d["MRFOLD"]), z["ll_gamma"])]
    for got, want in cases:
        assert np.array_equal(np.isnan(got), np.isnan(want))
        m = np.isfinite(want)
        assert np.array_equal(got[~m & ~np.isnan(want)], want[~m & ~np.isnan(want)])          # +-inf
        np.testing.assert_allclose(got[m], want[m], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("model,third,key", [(0, "TRUNC", "llr_nb_rows"), (1, "NONS", "llr_pg_rows")])
def test_llr_rows(hc, model, third, key):
    z, d = golden("variants"), _secondary_frame()
    cls = ("SYN", "MIS", third)
    pi3 = _c(np.stack([d["Pi_" + c] for c in cls], axis=1))
    obs3 = _c(np.stack([d["OBS_" + c] for c in cls], axis=1))
    n = len(d["ALPHA"])
    out = np.empty((4, n))
    a, t, m, ts = _c(d["ALPHA"]), _c(d["THETA"]), _c(d["MRFOLD"]), _c(d["T_SYN"])
    hc.hc_llr(model, _p(a), _p(t), _p(pi3), _p(obs3), _p(m), _p(ts), ctypes.c_int64(n), _p(out))
    for j in range(4):
        assert_pvals_close(out[j], z[key][:, j])
