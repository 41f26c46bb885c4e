"""CPU: the numpy restatement of the reference's convergence diagnostics (oracle/convergence_oracle.py)
against the fixture generated from the live reference (tests/golden/convergence.npz)."""
import numpy as np
import pytest

from helpers import load_npz
from oracle import convergence_oracle as CO, ref_harness as RH


@pytest.fixture(scope="module")
def gold():
    return load_npz("convergence.npz")


def test_autocorrelation_auto_bitexact(gold):
    ac, tau, sy, sx = CO.autocorr_method(gold["traces"], "auto", 1, 100)
    assert np.array_equal(ac, gold["auto_autocorr"])
    assert np.array_equal(tau, gold["auto_tau"])
    assert sy is None and sx is None


def test_autocorrelation_block_average_bitexact(gold):
    ac, tau, sy, sx = CO.autocorr_method(gold["traces"], "block-avg-auto", 5, 50)
    assert np.array_equal(ac, gold["block_autocorr"])
    assert np.array_equal(tau, gold["block_tau"])
    assert sy.shape == (3, 51) and sx.shape == (3,)


def test_jsd_bootstrap_same_mt_stream_bitexact(gold):
    np.random.seed(7)
    a, b = CO.process(gold["indices"], gold["auto_tau"], gold["n_bins"], nfold=10, nround=20)
    assert np.array_equal(a, gold["all_JSD"])
    assert np.array_equal(b, gold["all_JSDs"])


def test_g_drops_last_sample_and_handles_long_lags():
    f = np.array([1.0, 4.0, 2.0, 8.0, 5.0])
    z = f - f.mean()
    r = CO.g(f, max_tau=6, normalize=False)
    assert r[0] == np.dot(z[:4], z[:4]) / 5            # the last sample never enters (convergence.py:108)
    assert r[3] == (z[0] * z[3]) / 2
    assert r[4] == 0.0 and np.isnan(r[5]) and r[6] == 0.0 and np.signbit(r[6])


def test_jsd_properties():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 50, 4000)
    assert abs(CO.jsd(x[:2000], x[:2000], np.concatenate([x[:2000], x[:2000]]), 50)) < 1e-14   # identical halves
    a, b = np.zeros(100, np.int64), np.ones(100, np.int64)
    assert abs(CO.jsd(a, b, np.concatenate([a, b]), 2) - np.log(2.0)) < 1e-15                 # disjoint halves


def test_philox_half_is_a_half():
    m = CO.philox_first_half(1001, 500, seed=3, stream=11)
    assert m.shape == (500,) and np.unique(m).shape == (500,) and m.min() >= 0 and m.max() < 1001
    assert not np.array_equal(m, CO.philox_first_half(1001, 500, seed=3, stream=12))
    # statistically the same bootstrap distribution as the reference's MT draw
    g = load_npz("convergence.npz")
    _, b = CO.process(g["indices"], g["auto_tau"], g["n_bins"], nfold=10, nround=20, rng="philox", seed=5)
    assert abs(b.mean() - g["all_JSDs"].mean()) < 0.02


@pytest.mark.skipif(not RH.reference_available(), reason="reference tree not present")
def test_oracle_vs_live_reference(tmp_path, gold):
    biceps = RH.import_reference()
    fn = RH.REFERENCE_ROOT + "/docs/examples/Tutorials/Prep_Rest_Post_Ana/results/traj_lambda0.00.npz"
    with RH.quiet():
        C = biceps.Convergence(filename=fn, outdir=str(tmp_path))
        C.get_autocorrelation_curves(method="block-avg-auto", nblocks=4, maxtau=80)
        traces, idx, nb = CO.columnar(C.traj)
        ac, tau, _, _ = CO.autocorr_method(traces, "block-avg-auto", 4, 80)
        assert np.array_equal(ac, C.autocorr) and np.array_equal(tau, C.tau_c)
        np.random.seed(11)
        C.process(nblock=5, nfold=5, nround=7, savefile=True, block_avg=False)
    np.random.seed(11)
    a, b = CO.process(idx, tau, nb, nfold=5, nround=7)
    assert np.array_equal(a, np.load(tmp_path / "all_JSD.npy"))
    assert np.array_equal(b, np.load(tmp_path / "all_JSDs.npy"))
