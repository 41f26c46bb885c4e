"""Analysis: BICePs scores and reweighted populations by MBAR on the GPU.

Mirrors /root/reference/biceps/Analysis.py:38-237 (constructor arguments, attributes f_df, P_dP,
u_kln, u_kn, N_k, states_kn, Deltaf_ij, dDeltaf_ij, lam, K, the BS.dat / populations.dat files).
The K*K*nsnaps Python calls into sampler.neglogP (Analysis.py:165-183) become one bcp_ukn kernel
over all snapshots, and pymbar (Analysis.py:192-223) becomes the streaming solver of csrc/mbar.cu.
Plotting (Analysis.py:269-436) is out of scope.
"""
import os
import pickle

import numpy as np

from . import engine, mbar
from .toolbox import get_files


def find_all_state_sampled_time(trace, nstates, verbose=True):
    """First step at which every state has been visited (Analysis.py:16-35), vectorised."""
    trace = np.asarray(trace, dtype=np.int64)
    first = np.full(nstates, -1, dtype=np.int64)
    uniq, idx = np.unique(trace, return_index=True)
    first[uniq] = idx
    order = np.sort(idx)
    frac_steps = np.zeros(len(trace), dtype=np.float64)
    if len(trace):
        marks = np.zeros(len(trace), dtype=np.int64)
        marks[order] = 1
        frac_steps = np.cumsum(marks) / float(nstates)
    if np.any(first < 0):
        if verbose:
            print('not all state sampled, these states', np.where(first < 0)[0], 'are not sampled')
        return 'null', list(frac_steps)
    init = int(first.max()) + 1
    return init, list(frac_steps[:init])


class Analysis(object):
    def __init__(self, outdir, nstates=0, precheck=True, BSdir='BS.dat', popdir='populations.dat',
                 picfile='BICePs.pdf', verbose=False, device=0):
        self.verbose = verbose
        self.resultdir = outdir
        self.trajs = os.path.join(outdir, "*.npz")
        self.BSdir = os.path.join(self.resultdir, BSdir)
        self.popdir = os.path.join(self.resultdir, popdir)
        self.picfile = os.path.join(self.resultdir, picfile)
        self.scheme = None
        self.traj = []
        self.sampler = []
        self.lam = None
        self.f_df = None
        self.P_dp = None
        self.precheck = precheck
        self.nstates = nstates
        self.device = int(device)
        if self.nstates == 0:
            raise ValueError("State number cannot be zero.")
        self.MBAR_analysis()

    @classmethod
    def from_samplers(cls, samplers, nstates, device=0, verbose=False):
        """File-free route: BICePs scores and populations straight from sampler
        objects that have been sampled in this process, one per lambda in ascending lambda order.  The
        columnar snapshot arrays of EVERY chain of a sampler (`traj.chains`, nchains seeds) are the samples of
        its lambda run (chain-major = Analysis.py:214 kln_to_kn order); no traj_lambda*.npz / .pkl round trip
        and no Python list-of-lists.  Fills lam, K, N_k, u_kn, f_df, P_dP, Deltaf_ij, dDeltaf_ij like
        MBAR_analysis (u_kln / states_kn, whose [K][K][Nmax] layout only makes sense for the file route, are
        not built)."""
        self = cls.__new__(cls)
        self.verbose, self.device, self.nstates = verbose, int(device), int(nstates)
        if self.nstates == 0:
            raise ValueError("State number cannot be zero.")
        self.sampler = list(samplers)
        self.traj = [s.traj for s in self.sampler]
        self.lam = [float(s.lam) for s in self.sampler]
        self.K = self.nlambda = K = len(self.sampler)
        self.scheme = self.sampler[0].rest_type
        tabs = dict(self.sampler[0].tables)
        tabs["energy"] = np.array([s.tables["energy"][0] for s in self.sampler])
        tabs["logZ"] = np.array([s.tables["logZ"][0] for s in self.sampler])
        T = engine.DeviceTables(tabs, device=self.device)
        state, idx, N_k = [], [], []
        for s in self.sampler:
            ch = getattr(s.traj, "chains", None)
            if ch is None or ch["state"].shape[0] == 0:
                raise ValueError("from_samplers: every sampler must have been sampled (traj.chains is empty)")
            ns, n = ch["state"].shape
            state.append(np.ascontiguousarray(ch["state"].T).reshape(-1))                    # chain-major
            idx.append(np.ascontiguousarray(ch["indices"].transpose(2, 0, 1)).reshape(ns * n, -1))
            N_k.append(ns * n)
        state = np.concatenate(state).astype(np.int32)
        idx = np.concatenate(idx).astype(np.int32)
        self.N_k = np.array(N_k)
        self.u_kn = mbar.ukn(T, state, idx)                                                  # K6
        r = mbar.mbar(self.u_kn, self.N_k, state, self.nstates, device=self.device)          # K7
        T.close()
        self.Deltaf_ij, self.dDeltaf_ij = r["Delta_f"], r["dDelta_f"]
        self.mbar_iterations = r["iters"]
        self.f_df = np.zeros((K, 2))
        self.f_df[:, 0] = self.Deltaf_ij[0, :]
        self.f_df[:, 1] = self.dDeltaf_ij[0, :]
        self.P_dP = np.zeros((self.nstates, 2 * K))
        self.P_dP[:, 0:K] = r["P"]
        self.P_dP[:, K:2 * K] = r["dP"]
        return self

    def load_data(self):
        """Trajectories (*.npz), their samplers (*.pkl) and the lambda values parsed from the
        file names (Analysis.py:74-116)."""
        files = get_files(self.trajs)
        for file in files:
            if self.verbose:
                print('Loading %s ...' % file)
            self.traj.append(np.load(file, allow_pickle=True)['arr_0'].item())
        self.nreplicas = len(self.traj[0]['trajectory'][0][3])
        if self.precheck:
            self.sampled_fractions = [find_all_state_sampled_time(t['state_trace'], self.nstates, self.verbose)
                                      for t in self.traj]
        for file in files:
            with open(file.replace('.npz', '.pkl'), 'rb') as f:
                self.sampler.append(pickle.load(f))
        self.nlambda = len(files)
        self.lam = [float((s.split('lambda')[1]).replace('.npz', '')) for s in files]
        self.scheme = self.traj[0]['rest_type']

    def MBAR_analysis(self, debug=False):
        self.load_data()
        self.K = K = self.nlambda
        N_k = np.array([len(t['trajectory']) for t in self.traj])
        nsnaps = int(N_k.max())
        # one table set for all lambdas: the restraint tables are shared, energy / logZ are per lambda
        tabs = dict(self.sampler[0].tables)
        tabs["energy"] = np.array([s.tables["energy"][0] for s in self.sampler])
        tabs["logZ"] = np.array([s.tables["logZ"][0] for s in self.sampler])
        T = engine.DeviceTables(tabs, device=self.device)
        state = np.concatenate([[t[3][0] for t in tr['trajectory']] for tr in self.traj]).astype(np.int32)
        idx = np.concatenate([[np.concatenate(t[4]) for t in tr['trajectory']] for tr in self.traj]).astype(np.int32)
        u_kn = mbar.ukn(T, state, idx)                                     # [K][N], K6
        # the reference keeps the stored E on the diagonal (Analysis.py:170)
        u_kln = np.zeros((K, K, nsnaps))
        states_kn = np.zeros((K, nsnaps, self.nreplicas))
        o = 0
        for k in range(K):
            u_kln[k, :, :N_k[k]] = u_kn[:, o:o + N_k[k]]
            u_kln[k, k, :N_k[k]] = [t[1] for t in self.traj[k]['trajectory']]
            u_kn[k, o:o + N_k[k]] = u_kln[k, k, :N_k[k]]
            if K > 1:                                                      # (:172-174) only set when k != l
                states_kn[k, :N_k[k], 0] = state[o:o + N_k[k]]
            o += N_k[k]
        self.u_kln, self.N_k, self.states_kn, self.u_kn = u_kln, N_k, states_kn, u_kn
        states_n = np.concatenate([states_kn[k, :N_k[k], 0] for k in range(K)]).astype(np.int32)
        r = mbar.mbar(u_kn, N_k, states_n, self.nstates, device=self.device)   # K7
        self.Deltaf_ij, self.dDeltaf_ij = r["Delta_f"], r["dDelta_f"]
        self.mbar_iterations = r["iters"]
        self.f_df = np.zeros((self.nlambda, 2))
        self.f_df[:, 0] = self.Deltaf_ij[0, :]      # BICePs score
        self.f_df[:, 1] = self.dDeltaf_ij[0, :]     # and its standard deviation
        self.P_dP = np.zeros((self.nstates, 2 * K))
        self.P_dP[:, 0:K] = r["P"]
        self.P_dP[:, K:2 * K] = r["dP"]
        self.save_MBAR()

    @staticmethod
    def _field(t, key):
        return t[key] if isinstance(t, dict) else getattr(t, key)

    def get_max_likelihood_parameters(self, model=0, sigma_only=False):
        """Most-sampled value of every nuisance parameter of trajectory `model` (Analysis.py:119-133)."""
        import pandas as pd
        t1 = self.traj[model]
        out = {}
        for k, rest in enumerate(self.scheme):
            if sigma_only and "sigma" not in rest:
                continue
            x, y = self._field(t1, 'allowed_parameters')[k], self._field(t1, 'sampled_parameters')[k]
            out[rest] = [x[np.argmax(y)]]
        return pd.DataFrame(out)

    def get_traces(self, traj_index=-1):
        """DataFrame of the parameter traces, columns `<rest_type><occurrence>` (Analysis.py:240-267)."""
        import pandas as pd
        t = self.traj[traj_index]
        seen, columns = {}, []
        for rest in self._field(t, 'rest_type'):
            columns.append("%s%d" % (rest, seen.get(rest, 0)))
            seen[rest] = seen.get(rest, 0) + 1
        return pd.DataFrame(np.array(self._field(t, 'traces')).transpose(), columns).transpose()

    def save_MBAR(self):
        np.savetxt(self.BSdir, self.f_df)
        np.savetxt(self.popdir, self.P_dP)

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is out of scope of biceps_b200 (matplotlib figures, Analysis.py:269-436)")
