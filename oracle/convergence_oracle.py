"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (biceps_b200/).

CPU restatement (numpy) of the reference's convergence diagnostics, SURVEY.md section 8 row f4:
autocorrelation curves / times of the nuisance-parameter traces and the Jensen-Shannon-divergence
bootstrap.  Reference: /root/reference/biceps/convergence.py
    g                                   :93-111
    compute_autocorrelation_curves      :78-90
    compute_autocorrelation_time        :114-125
    get_blocks                          :128-141
    compute_JSD                         :145-186
    Convergence.get_autocorrelation_curves  :380-440  (methods "auto", "block-avg-auto")
    Convergence.process                 :445-523
Pinned against the live reference (tests/test_convergence_oracle.py, reference present) and against the
fixture tests/golden/convergence.npz generated from it (tests/golden/make_golden_convergence.py).

The trajectory is taken in columnar form: `traces[P][T]` (parameter values per snapshot) and
`indices[P][T]` (parameter indices per snapshot) instead of the reference's list of per-snapshot lists.

Reference quirks that are part of the specification:
  * g() drops the LAST sample from both factors (`f[0:-1-tau] . f[tau:-1]`, :108) but divides by T - tau;
    lags with an empty overlap give 0 / (T - tau) (nan at tau == T, -0.0 beyond).
  * the block loop of process(block_avg=True) cuts blocks of int(nsnaps / nfold) snapshots (:470), not nblock.
  * the bootstrap draws np.random.choice(n, int(n/2), replace=False) from numpy's global MT19937 stream
    (:509), one draw per (parameter, fold, round) in that loop order.
"""
import numpy as np

from . import philox


def g(f, max_tau=10000, normalize=True):
    """convergence.py:93-111."""
    f = np.asarray(f, dtype=np.float64)
    T = f.shape[0]
    z = f - f.mean()
    out = np.zeros(max_tau + 1)
    for tau in range(max_tau + 1):
        n = max(T - 1 - tau, 0)
        out[tau] = np.dot(z[:n], z[tau:tau + n]) / (T - tau) if T != tau else np.nan
    return out / out[0] if normalize else out


def autocorrelation_curves(series, max_tau, normalize=True):
    """convergence.py:78-90."""
    return np.array([g(s, max_tau, normalize) for s in series])


def autocorrelation_time(curves):
    """convergence.py:114-125: plain left-to-right sum of each curve."""
    return np.array([sum(c) for c in curves])


def blocks(series, nblocks):
    """convergence.py:128-141: nblocks consecutive blocks of int(T / nblocks) samples; the tail is dropped."""
    out = []
    for v in series:
        dx = int(len(v) / nblocks)
        out.append([v[dx * n:dx * (n + 1)] for n in range(nblocks)])
    return out


def autocorr_method(traces, method="auto", nblocks=5, maxtau=10000):
    """The "auto" / "block-avg-auto" branch of get_autocorrelation_curves (:410-428) ->
    (autocorr[P][maxtau+1], tau_c[P], std_y, std_x)."""
    if method == "auto":
        nblocks = 1
    b = blocks(traces, nblocks)
    y = [autocorrelation_curves(bi, maxtau, True) for bi in b]
    x = [autocorrelation_time(yi) for yi in y]
    autocorr = np.average(y, axis=1)
    tau_c = np.average(x, axis=1)
    if nblocks == 1:
        return autocorr, tau_c, None, None
    return autocorr, tau_c, np.std(y, axis=1), np.std(x, axis=1)


def _entropy(r):
    N = sum(r)
    with np.errstate(divide="ignore", invalid="ignore"):
        h = -1. * r / N * np.log(r / N)
    return sum(np.nan_to_num(h)), N


def jsd_from_counts(r1, r2, rt):
    """convergence.py:174-186 from the three histograms."""
    H1, N1 = _entropy(np.asarray(r1, dtype=np.float64))
    H2, N2 = _entropy(np.asarray(r2, dtype=np.float64))
    H, N = _entropy(np.asarray(rt, dtype=np.float64))
    return H - (N1 / N) * H1 - (N2 / N) * H2


def jsd(idx1, idx2, idx_total, nbins):
    """compute_JSD (:145-186) on index arrays of one parameter."""
    r1 = np.bincount(idx1, minlength=nbins)
    r2 = np.bincount(idx2, minlength=nbins)
    rt = np.bincount(idx_total, minlength=nbins)
    return jsd_from_counts(r1, r2, rt)


def philox_first_half(n_total, n_first, seed, stream):
    """Device-RNG split (biceps_b200 extension, used when rng="philox"): element e gets the key
    (w0 << 32) | e with w0 the first word of Philox4x32-10(counter = (e_lo, e_hi, stream_lo, stream_hi),
    key = seed); the n_first elements with the smallest keys form the first part."""
    e = np.arange(n_total, dtype=np.uint64)
    w0 = philox.philox4x32_np(e & np.uint64(philox.MASK), e >> np.uint64(32), stream & philox.MASK,
                              (stream >> 32) & philox.MASK, seed & philox.MASK, (seed >> 32) & philox.MASK)[0]
    key = (w0.astype(np.uint64) << np.uint64(32)) | e
    order = np.argsort(key, kind="stable")
    return np.sort(order[:n_first])


def process(indices, tau_c, nbins, nfold=10, nround=100, rng="numpy", seed=0):
    """Convergence.process (:482-517), JSD part: -> all_JSD[P][nfold], all_JSDs[P][nfold][nround].
    rng="numpy" consumes np.random's global stream exactly like the reference."""
    P = len(tau_c)
    all_JSD = np.zeros((P, nfold))
    all_JSDs = np.zeros((P, nfold, nround))
    for i in range(P):
        tau = int(1 + 2 * tau_c[i])
        sub = np.asarray(indices[i])[::tau]
        nsnaps = len(sub)
        if nsnaps < 2 * nfold:
            raise ValueError("not enough data left after subsampling")
        dx = int(nsnaps / nfold)
        for subset in range(nfold):
            n = dx * (subset + 1)
            half = int(n / 2)
            tot = sub[:n]
            all_JSD[i, subset] = jsd(tot[:half], tot[half:], tot, nbins[i])
            for r in range(nround):
                if rng == "numpy":
                    m1 = np.random.choice(n, int(n / 2), replace=False)
                else:
                    m1 = philox_first_half(n, int(n / 2), seed, (i * nfold + subset) * nround + r)
                mask = np.zeros(n, dtype=bool)
                mask[m1] = True
                all_JSDs[i, subset, r] = jsd(tot[mask], tot[~mask], tot, nbins[i])
    return all_JSD, all_JSDs


def block_avg(indices, tau_c, allowed, nblock=5, nfold=10):
    """Convergence.process(block_avg=True) (:462-480): -> r_total[P][nblock][G_p], r_max[P][nblock]."""
    r_total, r_max = [], []
    for i in range(len(tau_c)):
        tau = int(1 + 2 * tau_c[i])
        sub = np.asarray(indices[i])[::tau]
        dx = int(len(sub) / nfold)
        rt, rm = [], []
        for subset in range(nblock):
            grid = np.bincount(sub[dx * subset:dx * (subset + 1)], minlength=len(allowed[i])).astype(np.float64)
            rt.append(grid)
            rm.append(allowed[i][int(np.argmax(grid))])
        r_total.append(rt)
        r_max.append(rm)
    return r_total, r_max


def columnar(traj):
    """traces[P][T], indices[P][T], allowed lengths from a reference `results` dict
    (PosteriorSampler.py:444-466; Convergence.__init__ / get_sampled_parameters :213-240)."""
    traces = np.ascontiguousarray(np.array(traj["traces"], dtype=np.float64).T)
    idx = np.array([np.concatenate(fr[4]) for fr in traj["trajectory"]], dtype=np.int32).T
    return traces, np.ascontiguousarray(idx), [len(a) for a in traj["allowed_parameters"]]
