"""biceps.Convergence on the GPU: autocorrelation curves / times of the
nuisance-parameter traces and the Jensen-Shannon-divergence bootstrap of a sampled trajectory.

Mirror of /root/reference/biceps/convergence.py (class Convergence :192-653 and the module functions
:12-186): same constructor, method names, arguments, attributes (`autocorr`, `tau_c`, `sampled_parameters`,
`labels`, `maxtau`, ...) and output files (`all_JSD.npy`, `all_JSDs.npy`).  The O(T * maxtau) lag loop and
the nfold x nround bootstrap run in csrc/convergence.cu through bcp_autocorr / bcp_jsd / bcp_param_hist;
there is no CPU fallback.  Figures are drawn only when matplotlib is importable (it is an optional
dependency here; the numbers do not depend on it).

Extensions: the results of process() are kept as attributes (`all_JSD`, `all_JSDs`, `r_total`, `r_max`);
`process(rng="philox")` draws the bootstrap halves on the device instead of from np.random (for long
trajectories); `Convergence.from_arrays` takes the columnar traces directly.
"""
import os

import numpy as np

from . import _cabi as A


# --------------------------------------------------------------------------------------------------
# flat-array layer
def autocorrelation(series, max_tau, normalize=True, device=0):
    """curves[n][max_tau+1], tau[n] for the rows of `series` (equal lengths): g() of convergence.py:93-111
    for every row plus compute_autocorrelation_time (:114-125)."""
    s = A.f64(np.atleast_2d(series))
    n, T = s.shape
    curves = np.zeros((n, int(max_tau) + 1))
    tau = np.zeros(n)
    A.check(A.lib().bcp_autocorr(A.dptr(s), n, T, int(max_tau), 1 if normalize else 0, int(device), A.dptr(curves),
                                 A.dptr(tau)), "bcp_autocorr")
    return curves, tau


def _tasks(rows):
    arr = (A.bcp_jsd_task * len(rows))()
    for t, (off, n, n1, sel_off, stream, nbins, mode) in zip(arr, rows):
        t.offset, t.n_total, t.n_first, t.sel_offset = int(off), int(n), int(n1), int(sel_off)
        t.stream, t.n_bins, t.mode = int(stream), int(nbins), int(mode)
    return arr


def jsd_tasks(idx, rows, sel=None, seed=0, device=0):
    """One JSD per row (offset, n_total, n_first, sel_offset, stream, n_bins, mode) over the concatenated
    index series `idx` (compute_JSD, convergence.py:145-186)."""
    idx = A.i32(idx)
    sel = A.i32(sel) if sel is not None and len(sel) else None
    out = np.zeros(len(rows))
    A.check(A.lib().bcp_jsd(A.i32ptr(idx), idx.shape[0], A.i32ptr(sel), 0 if sel is None else sel.shape[0],
                            _tasks(rows), len(rows), int(seed) & (2 ** 64 - 1), int(device), A.dptr(out)), "bcp_jsd")
    return out


def param_hist(idx, rows, max_bins, device=0):
    idx = A.i32(idx)
    out = np.zeros((len(rows), int(max_bins)), np.int64)
    A.check(A.lib().bcp_param_hist(A.i32ptr(idx), idx.shape[0], _tasks(rows), len(rows), int(max_bins), int(device),
                                   A.i64ptr(out)), "bcp_param_hist")
    return out


def chain_autocorrelation_times(sampler, maxtau=1000, device=0):
    """Autocorrelation curve and time of every nuisance-parameter trace of EVERY chain of a many-chain
    PosteriorSampler in one launch: -> autocorr[n_chains][P][maxtau+1], tau_c[n_chains][P].  Chain c, parameter
    p equals Convergence.from_sampler(sampler, chain=c).get_autocorrelation_curves("auto", maxtau=maxtau)."""
    t = sampler.traj
    if t.chains is None:
        raise ValueError("the sampler has not sampled yet")
    if getattr(sampler, "_chains", None) is not None:           # snapshots still resident in HBM: no host round trip
        return sampler._chains.autocorrelation(t.allowed_parameters, maxtau, True)
    idx = t.chains["indices"]                                   # [T][P][n_chains]
    T, P, n = idx.shape
    allowed = [np.asarray(a, dtype=np.float64) for a in t.allowed_parameters]
    series = np.empty((n, P, T))
    for p in range(P):
        series[:, p, :] = allowed[p][idx[:, p, :]].T
    curves, tau = autocorrelation(series.reshape(n * P, T), maxtau, True, device)
    return curves.reshape(n, P, int(maxtau) + 1), tau.reshape(n, P)


# --------------------------------------------------------------------------------------------------
# module functions of the reference (same names)
def single_exp_decay(x, a0, a1, tau1):
    return a0 + a1 * np.exp(-(x / tau1))


def double_exp_decay(x, a0, a1, a2, tau1, tau2):
    return a0 + a1 * np.exp(-(x / tau1)) + a2 * np.exp(-(x / tau2))


def exponential_fit(autocorrelation, exp_function='single', v0=None, verbose=False):
    """Least-squares fit of one or two exponentials to an autocorrelation curve (convergence.py:40-75);
    host side (scipy), a few parameters per curve."""
    from scipy.optimize import curve_fit
    x = np.arange(autocorrelation.shape[0])
    if exp_function == 'single':
        if v0 is None:
            v0 = [0.0, 1.0, 4000.]
        popt, pcov = curve_fit(single_exp_decay, x, autocorrelation, p0=v0, maxfev=10000)
        if verbose:
            print(('Best-fit tau1 = %s +/- %s' % (popt[2], pcov[2][2])))
        return single_exp_decay(x, *popt[:3]), popt[2]
    if v0 is None:
        v0 = [0.0, 0.9, 0.1, 4000., 200.0]
    popt, pcov = curve_fit(double_exp_decay, x, autocorrelation, p0=v0, maxfev=10000)
    if verbose:
        print(('Best-fit tau1 = %s +/- %s' % (popt[3], pcov[3][3])))
        print(('Best-fit tau2 = %s +/- %s' % (popt[4], pcov[4][4])))
    return double_exp_decay(x, *popt[:5]), max(popt[3], popt[4])


def g(f, max_tau=10000, normalize=True):
    return autocorrelation(np.asarray(f, dtype=np.float64)[None, :], max_tau, normalize)[0][0]


def compute_autocorrelation_curves(data, max_tau, normalize=True):
    data = [np.asarray(d, dtype=np.float64) for d in data]
    if len({d.shape[0] for d in data}) == 1:
        return autocorrelation(np.stack(data), max_tau, normalize)[0]
    return np.array([g(d, max_tau, normalize) for d in data])


def compute_autocorrelation_time(autocorrelations):
    return np.array([sum(autocorrelations[i]) for i in range(len(autocorrelations))])


def get_blocks(data, nblocks=5):
    blocks = []
    for vec in data:
        dx = int(len(vec) / nblocks)
        blocks.append([vec[dx * n:dx * (n + 1)] for n in range(nblocks)])
    return blocks


def compute_JSD(T1, T2, T_total, ind, allowed_parameters):
    """Reference signature (lists of trajectory frames); evaluated on the GPU."""
    col = lambda T: np.array([np.concatenate(fr[4])[ind] for fr in T], dtype=np.int32)
    a, b = col(T1), col(T2)
    # JSD depends only on the histograms of part 1 and of part 1 + part 2
    idx = np.concatenate([a, b])
    return jsd_tasks(idx, [(0, idx.shape[0], a.shape[0], 0, 0, len(allowed_parameters), 0)])[0]


class Convergence(object):

    def __init__(self, traj=None, filename=None, outdir="./", verbose=False, device=0):
        """traj: a PosteriorSamplingTrajectory, its `results`-style dict, or None with `filename` =
        path of a trajectory npz (convergence.py:194-230)."""
        if (traj is None) and (filename is None):
            raise ValueError("Must provide a trajectory or a relative path to a trajectory file.")
        self.verbose = verbose
        self.device = int(device)
        if filename:
            self.traj = np.load(filename, allow_pickle=True)['arr_0'].item()
        if traj is not None:
            self.traj = traj if isinstance(traj, dict) else traj.__dict__
        trajectory = self.traj["trajectory"]
        self.freq_save_traj = int(trajectory[1][0] - trajectory[0][0])
        self.rest_type = self.traj['rest_type']
        self.allowed_parameters = self.traj['allowed_parameters']
        self.sampled_parameters = self.get_sampled_parameters()
        # columnar parameter indices [P][T]: what compute_JSD reads out of every frame (:163-171)
        self._indices = np.ascontiguousarray(
            np.array([np.concatenate(fr[4]) for fr in trajectory], dtype=np.int32).T)
        self.labels = self.get_labels()
        self.exp_function = "single"
        self.outdir = os.getcwd() if outdir is None else outdir

    @classmethod
    def from_arrays(cls, traces, indices, allowed_parameters, rest_type, freq_save_traj=1, outdir="./",
                    verbose=False, device=0):
        """File-free constructor: traces[P][T] (parameter values per snapshot), indices[P][T]."""
        self = cls.__new__(cls)
        self.verbose, self.device = verbose, int(device)
        self.traj = None
        self.freq_save_traj = int(freq_save_traj)
        self.rest_type = list(rest_type)
        self.allowed_parameters = [np.asarray(a) for a in allowed_parameters]
        self.sampled_parameters = np.ascontiguousarray(traces, dtype=np.float64)
        self._indices = np.ascontiguousarray(indices, dtype=np.int32)
        if self.sampled_parameters.shape != self._indices.shape or len(self.rest_type) != self._indices.shape[0]:
            raise ValueError("traces and indices must both be [len(rest_type)][T]")
        self.labels = self.get_labels()
        self.exp_function = "single"
        self.outdir = os.getcwd() if outdir is None else outdir
        return self

    @classmethod
    def from_sampler(cls, sampler, chain=0, outdir="./", verbose=False, device=0):
        """Straight from an in-process PosteriorSampler: the columnar snapshots of chain `chain`
        (every chain of a many-chain sampler is a trajectory of its own)."""
        t = sampler.traj
        if t.chains is None:
            raise ValueError("the sampler has not sampled yet")
        idx = np.ascontiguousarray(t.chains["indices"][:, :, int(chain)].T)           # [P][T]
        allowed = [np.asarray(a) for a in t.allowed_parameters]
        traces = np.stack([allowed[p][idx[p]] for p in range(idx.shape[0])])
        return cls.from_arrays(traces, idx, allowed, t.rest_type, sampler.traj_every, outdir, verbose, device)

    def get_sampled_parameters(self):
        t = np.array(self.traj['traces'])
        return np.array([t[:, i] for i in range(len(self.rest_type))])

    def get_labels(self):
        labels = []
        for r in self.rest_type:
            c = r.count('_')
            if c == 0:
                labels.append('$\\%s$' % r if r == 'gamma' else '$%s$' % r)
            elif c == 1:
                labels.append("$\\%s_{%s}$" % tuple(r.split('_')[:2]))
            elif c == 2:
                labels.append("$\\%s_{{%s}_{%s}}$" % tuple(r.split('_')[:3]))
        return labels

    # ---- numbers -----------------------------------------------------------------------------------
    def get_autocorrelation_curves(self, method="auto", nblocks=5, maxtau=10000, plot_traces=False):
        """convergence.py:380-440.  method: "auto" | "block-avg-auto" | "exp"."""
        sp = np.asarray(self.sampled_parameters, dtype=np.float64)
        self.maxtau = maxtau
        if method == "exp":
            self.autocorr, self.tau_c = autocorrelation(sp, maxtau, True, self.device)
        if method in ["block-avg-auto", "auto"]:
            if method == "auto":
                nblocks = 1
            P, T = sp.shape
            dx = int(T / nblocks)
            # every block of every parameter is one series of the same launch
            blocks = sp[:, :dx * nblocks].reshape(P * nblocks, dx)
            y, x = autocorrelation(blocks, maxtau, True, self.device)
            y = y.reshape(P, nblocks, maxtau + 1)
            x = x.reshape(P, nblocks)
            self.autocorr = np.average(y, axis=1)
            self.tau_c = np.average(x, axis=1)
            if any(i < 0 for i in self.tau_c):
                print("NOTE: Found a negative autocorrelation time...")
            if nblocks == 1:
                std_y, std_x = None, None
            else:
                std_y, std_x = np.std(y, axis=1), np.std(x, axis=1)
            self.std_y, self.std_x = std_y, std_x
        elif method == "exp":
            self.yFits, self.popts = [], []
            for i in range(len(self.autocorr)):
                yFit, popt = exponential_fit(self.autocorr[i], exp_function=self.exp_function)
                self.popts.append(popt)
                self.yFits.append(yFit)
            self.tau_c = np.array(self.popts)
        else:
            raise KeyError(f"Method must be 'block-avg-auto' or 'exp' or 'auto'. You provided: {method}")

    def process(self, nblock=5, nfold=10, nround=100, savefile=True, block_avg=False, normalize=True,
                rng="numpy", seed=0):
        """JSD of the two halves of growing prefixes of the tau-subsampled trajectory and its bootstrap
        distribution (convergence.py:445-523).  rng="numpy" consumes np.random's global stream exactly
        like the reference (one choice(n, n//2, replace=False) per parameter, fold and round);
        rng="philox" draws the halves on the device from (seed, task number)."""
        if rng not in ("numpy", "philox"):
            raise ValueError("rng must be 'numpy' or 'philox'")
        P = len(self.tau_c)
        subs, offs, off = [], [], 0
        for i in range(P):
            tau = int(1 + 2 * self.tau_c[i])
            s = self._indices[i][::tau]
            subs.append(s)
            offs.append(off)
            off += s.shape[0]
        idx = np.concatenate(subs)
        nb = [len(a) for a in self.allowed_parameters]

        if block_avg:
            rows = []
            for i in range(P):
                dx = int(subs[i].shape[0] / nfold)          # the reference's block length (:470)
                rows += [(offs[i] + dx * b, dx, 0, 0, 0, nb[i], 0) for b in range(nblock)]
            h = param_hist(idx, rows, max(nb), self.device).reshape(P, nblock, max(nb))
            self.r_total = [[h[i, b, :nb[i]].astype(np.float64) for b in range(nblock)] for i in range(P)]
            self.r_max = [[self.allowed_parameters[i][int(np.argmax(r))] for r in self.r_total[i]] for i in range(P)]

        rows, sel, sel_off = [], [], 0
        for i in range(P):
            nsnaps = subs[i].shape[0]
            if nsnaps < 2 * nfold:
                raise ValueError('not enough data left after subsampling using auto-correlation time '
                                 'with the given nfold')
            dx = int(nsnaps / nfold)
            for subset in range(nfold):
                n = dx * (subset + 1)
                half = int(n / 2)
                rows.append((offs[i], n, half, 0, 0, nb[i], 0))
                for r in range(nround):
                    if rng == "numpy":
                        m1 = np.random.choice(n, half, replace=False)
                        sel.append(m1)
                        rows.append((offs[i], n, half, sel_off, 0, nb[i], 1))
                        sel_off += half
                    else:
                        rows.append((offs[i], n, half, 0, (i * nfold + subset) * nround + r, nb[i], 2))
        out = jsd_tasks(idx, rows, np.concatenate(sel) if sel else None, seed, self.device)
        out = out.reshape(P, nfold, nround + 1)
        self.all_JSD = np.ascontiguousarray(out[:, :, 0])
        self.all_JSDs = np.ascontiguousarray(out[:, :, 1:])
        if savefile:
            np.save(os.path.join(self.outdir, "all_JSD.npy"), self.all_JSD)
            np.save(os.path.join(self.outdir, "all_JSDs.npy"), self.all_JSDs)
