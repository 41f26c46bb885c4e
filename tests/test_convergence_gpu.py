"""GPU: biceps_b200.convergence (csrc/convergence.cu through bcp_autocorr / bcp_jsd / bcp_param_hist) against
the oracle and the fixture generated from the live reference.

Tolerances: the autocorrelation differs from the reference's np.dot only by summation order -> 1e-12 of g(0)
(observed ~1e-15); JSD histograms are exact integers, the entropy uses CUDA's log (<= 1 ulp) -> 1e-13 absolute."""
import numpy as np
import pytest

from helpers import load_npz
from oracle import convergence_oracle as CO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return load_npz("convergence.npz")


@pytest.fixture(scope="module")
def conv(gold):
    from biceps_b200.convergence import Convergence
    allowed = [gold["allowed%d" % p] for p in range(3)]
    return Convergence.from_arrays(gold["traces"], gold["indices"], allowed, [str(x) for x in gold["rest_type"]],
                                   int(gold["freq_save_traj"]), outdir=None)


def test_auto_matches_reference(conv, gold):
    conv.get_autocorrelation_curves(method="auto", maxtau=100)
    assert np.allclose(conv.autocorr, gold["auto_autocorr"], rtol=0, atol=1e-12)
    assert np.allclose(conv.tau_c, gold["auto_tau"], rtol=1e-12, atol=1e-11)


def test_block_average_matches_reference(conv, gold):
    conv.get_autocorrelation_curves(method="block-avg-auto", nblocks=5, maxtau=50)
    assert np.allclose(conv.autocorr, gold["block_autocorr"], rtol=0, atol=1e-12)
    assert np.allclose(conv.tau_c, gold["block_tau"], rtol=1e-12, atol=1e-11)
    assert conv.std_y.shape == (3, 51)


def test_exp_fit_matches_reference(conv, gold):
    conv.get_autocorrelation_curves(method="exp", maxtau=100)
    # scipy's Levenberg-Marquardt stops at a relative step of ~1e-8 in the cost: curves that differ in the
    # last bits end the iteration a few 1e-5 apart in tau
    assert np.allclose(conv.tau_c, gold["exp_tau"], rtol=1e-3)


def test_jsd_bootstrap_same_mt_stream(conv, gold, tmp_path):
    conv.get_autocorrelation_curves(method="auto", maxtau=100)
    conv.outdir = str(tmp_path)
    np.random.seed(7)
    conv.process(nblock=5, nfold=10, nround=20, savefile=True, block_avg=True)
    assert np.allclose(conv.all_JSD, gold["all_JSD"], rtol=0, atol=1e-13)
    assert np.allclose(conv.all_JSDs, gold["all_JSDs"], rtol=0, atol=1e-13)
    assert np.array_equal(np.load(tmp_path / "all_JSD.npy"), conv.all_JSD)
    rt, rm = CO.block_avg(gold["indices"], conv.tau_c, conv.allowed_parameters, nblock=5, nfold=10)
    for i in range(3):
        assert np.array_equal(np.array(rt[i]), np.array(conv.r_total[i]))
        assert np.array_equal(np.array(rm[i]), np.array(conv.r_max[i]))


def test_jsd_philox_split_bitexact_vs_oracle(conv, gold):
    conv.get_autocorrelation_curves(method="auto", maxtau=100)
    conv.process(nfold=10, nround=12, savefile=False, rng="philox", seed=99)
    a, b = CO.process(gold["indices"], conv.tau_c, gold["n_bins"], nfold=10, nround=12, rng="philox", seed=99)
    assert np.allclose(conv.all_JSD, a, rtol=0, atol=1e-13)
    assert np.allclose(conv.all_JSDs, b, rtol=0, atol=1e-13)


def _ar1(rng, n, T, phi):
    x = np.zeros((n, T))
    e = rng.standard_normal((n, T))
    for t in range(1, T):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    return x + 3.0


@pytest.mark.parametrize("n,T,maxtau", [(1, 2, 0), (2, 5, 7), (3, 1023, 1100), (2, 1025, 1024), (4, 4097, 2500),
                                        (2, 20000, 3000)])
def test_autocorrelation_sizes_and_ragged_edges(n, T, maxtau):
    from biceps_b200 import convergence as BC
    x = _ar1(np.random.default_rng(T), n, T, 0.9)
    for normalize in (True, False):
        got, tau = BC.autocorrelation(x, maxtau, normalize)
        want = CO.autocorrelation_curves(x, maxtau, normalize)
        scale = 1.0 if normalize else np.nanmax(np.abs(want))
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.allclose(got, want, rtol=0, atol=1e-12 * max(scale, 1e-300), equal_nan=True)
        if not np.isnan(want).any():
            assert np.allclose(tau, CO.autocorrelation_time(want), rtol=1e-11, atol=1e-9 * max(scale, 1e-300))
    # lags with an empty overlap: 0 / (T - tau) keeps the reference's sign (nan at tau == T, -0.0 beyond)
    if maxtau > T:
        assert np.all(np.signbit(got[:, T + 1:])) and np.all(got[:, T + 1:] == 0)


def test_autocorrelation_exponential_decay_property():
    """AR(1) with coefficient phi: g(tau) ~ phi^tau (size-independent check at a length the oracle does not run)."""
    from biceps_b200 import convergence as BC
    x = _ar1(np.random.default_rng(5), 8, 400000, 0.95)
    got, tau = BC.autocorrelation(x, 40, True)
    assert np.allclose(got[:, 0], 1.0)
    assert np.abs(got.mean(axis=0) - 0.95 ** np.arange(41)).max() < 0.01


def test_jsd_modes_edges_and_errors():
    from biceps_b200 import convergence as BC
    rng = np.random.default_rng(2)
    x = rng.integers(0, 300, 5000).astype(np.int32)
    y = rng.integers(0, 7, 333).astype(np.int32)
    idx = np.concatenate([x, y])
    sel = np.sort(rng.choice(5000, 2500, replace=False)).astype(np.int32)
    rows = [(0, 5000, 2500, 0, 0, 300, 0), (0, 5000, 2500, 0, 0, 300, 1), (0, 5000, 2500, 0, 17, 300, 2),
            (5000, 333, 166, 0, 18, 7, 2), (5000, 333, 0, 0, 0, 7, 0), (5000, 333, 333, 0, 0, 7, 2),
            (5000, 1, 0, 0, 0, 7, 0)]
    got = BC.jsd_tasks(idx, rows, sel, seed=4)
    m = np.zeros(5000, bool); m[sel] = True
    h2 = CO.philox_first_half(5000, 2500, 4, 17); m2 = np.zeros(5000, bool); m2[h2] = True
    h3 = CO.philox_first_half(333, 166, 4, 18); m3 = np.zeros(333, bool); m3[h3] = True
    want = [CO.jsd(x[:2500], x[2500:], x, 300), CO.jsd(x[m], x[~m], x, 300), CO.jsd(x[m2], x[~m2], x, 300),
            CO.jsd(y[m3], y[~m3], y, 7)]
    # an empty part: the reference's 0/0 terms are zeroed by nan_to_num (convergence.py:178), JSD = H - H = 0
    want += [CO.jsd(y[:0], y, y, 7), CO.jsd(y, y[:0], y, 7), CO.jsd(y[:0], y[:1], y[:1], 7)]
    assert np.allclose(got, want, rtol=0, atol=1e-13) and not np.isnan(got).any()
    with pytest.raises(RuntimeError):
        BC.jsd_tasks(idx, [(0, 5000, 2500, 0, 0, 100, 0)])          # index >= n_bins
    with pytest.raises(RuntimeError):
        BC.jsd_tasks(idx, [(0, 6000, 2500, 0, 0, 300, 0)])          # reads outside idx


def test_convergence_from_sampler_and_file(tmp_path):
    """Convergence straight from an in-process sampler equals Convergence on the files it wrote."""
    import biceps_b200 as biceps
    from biceps_b200 import synthetic
    from biceps_b200.convergence import Convergence
    ens = synthetic.make_ensemble("C1s", seed=1)
    rs = [dict(cls="J" if r["name"] == "J" else "noe", model=r["model"], exp=r["exp"], weight=r["weight"],
               options=dict(ref=r["ref"], sigma=r["sigma"], **({"gamma": r["gamma"]} if "gamma" in r else {})))
          for r in ens["restraints"]]
    E = biceps.Ensemble.from_arrays(1.0, ens["energies"], rs)
    s = biceps.PosteriorSampler(E, freq_save_traj=10.)
    s.sample(200000)
    fn = str(tmp_path / "traj_lambda1.00.npz")
    s.traj.process_results(fn)
    a = Convergence(filename=fn, outdir=str(tmp_path))
    b = Convergence.from_sampler(s, outdir=str(tmp_path))
    c = Convergence(traj=s.traj, outdir=str(tmp_path))
    for C in (a, b, c):
        C.get_autocorrelation_curves(method="auto", maxtau=200)
        C.process(nfold=10, nround=5, savefile=False, rng="philox", seed=1)
    assert np.array_equal(a.autocorr, b.autocorr) and np.array_equal(a.autocorr, c.autocorr)
    assert np.array_equal(a.all_JSDs, b.all_JSDs) and np.array_equal(a.all_JSDs, c.all_JSDs)
    want = CO.autocorrelation_curves(a.sampled_parameters, 200, True)
    assert np.allclose(a.autocorr, want, rtol=0, atol=1e-12)


def test_all_chains_in_one_launch():
    import biceps_b200 as biceps
    from biceps_b200 import synthetic
    from biceps_b200.convergence import Convergence, chain_autocorrelation_times
    ens = synthetic.make_ensemble("C1s", seed=2)
    rs = [dict(cls="J" if r["name"] == "J" else "noe", model=r["model"], exp=r["exp"], weight=r["weight"],
               options=dict(ref=r["ref"], sigma=r["sigma"], **({"gamma": r["gamma"]} if "gamma" in r else {})))
          for r in ens["restraints"]]
    E = biceps.Ensemble.from_arrays(1.0, ens["energies"], rs)
    s = biceps.PosteriorSampler(E, freq_save_traj=10., nchains=64, seed=4)
    s.sample(30000)
    ac, tau = chain_autocorrelation_times(s, maxtau=150)          # device-resident snapshots (bcp_chains_autocorr)
    assert ac.shape == (64, 3, 151) and tau.shape == (64, 3)
    blk, s._chains = s._chains, None                               # host route: the fetched snapshot arrays
    ac_h, tau_h = chain_autocorrelation_times(s, maxtau=150)
    s._chains = blk
    assert np.array_equal(ac_h, ac) and np.array_equal(tau_h, tau)
    for c in (0, 17, 63):
        C = Convergence.from_sampler(s, chain=c, outdir=None)
        C.get_autocorrelation_curves(method="auto", maxtau=150)
        assert np.array_equal(C.autocorr, ac[c]) and np.array_equal(C.tau_c, tau[c])


def test_bad_arguments_are_reported():
    import ctypes as C
    from biceps_b200 import _cabi as A, convergence as BC
    lib = A.lib()
    x = np.zeros((2, 10))
    out = np.zeros((2, 4))
    assert lib.bcp_autocorr(None, 2, 10, 3, 1, 0, A.dptr(out), None) == -1            # NULL series
    assert lib.bcp_autocorr(A.dptr(x), 2, 1, 3, 1, 0, A.dptr(out), None) == -1         # T < 2
    assert lib.bcp_autocorr(A.dptr(x), 2, 10, -1, 1, 0, A.dptr(out), None) == -1       # negative max_tau
    assert b"bcp_autocorr" in lib.bcp_last_error()
    with pytest.raises(RuntimeError):
        BC.jsd_tasks(np.zeros(10, np.int32), [(0, 10, 5, 0, 0, 5000, 0)])              # more bins than the kernel holds
    with pytest.raises(RuntimeError):
        BC.jsd_tasks(np.zeros(10, np.int32), [(0, 10, 11, 0, 0, 4, 0)])                # n_first > n_total
    with pytest.raises(RuntimeError):
        BC.jsd_tasks(np.zeros(10, np.int32), [(0, 10, 5, 0, 0, 4, 1)], sel=np.arange(3, dtype=np.int32))   # sel too short
    # a constant series has zero variance: g(0) = 0 and the normalised curve is 0/0 = nan, like the reference
    c, t = BC.autocorrelation(np.full((1, 50), 2.5), 5, True)
    assert np.isnan(c).all()
