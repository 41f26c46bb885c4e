"""PosteriorSampler / PosteriorSamplingTrajectory: the reference's sampler API on the CUDA engine.

Mirrors /root/reference/biceps/PosteriorSampler.py:9-488 (constructor arguments, sample()
signature, attributes, trajectory / results schema, file outputs).  Construction flattens the
ensemble into dense tables (reference potentials and logZ from the precompute kernels), sample()
runs `nchains` independent Metropolis chains in one kernel and unflattens the device buffers
into the reference's trajectory layout.

Differences a user can see, all deliberate:
  * the RNG is counter-based Philox4x32-10 (oracle/philox.py is its specification), not numpy's
    global MT19937; `seed=None` draws the seed from numpy's global stream, so np.random.seed(x)
    before constructing the sampler still makes a run reproducible;
  * `rng="numpy"` (one chain; the default when neither `nchains` nor `seed` is given, i.e. for an unchanged
    reference script) instead consumes np.random's global MT19937 stream draw for draw like the
    reference: after np.random.seed(x) the trajectory, traces and histograms are the reference's own, bit for
    bit (csrc/sampler_mt.cu; a compatibility path: 8.5e5 steps/s, 17x the reference itself);
  * `nchains` (default 1) runs that many independent chains; histograms (state_counts,
    sampled_parameters) are summed over them, the trajectory / traces / state_trace attributes are
    those of chain 0 and `traj.chains` holds every chain's thinned snapshots as arrays.
"""
import numpy as np

from . import engine, precompute
from .toolbox import save_object


def build_tables(ensembles, device=0):
    """Flat sampler tables for one or several lambda ensembles that share their restraints
    (energy [K][S], logZ [K]).  Reference potentials are computed here, as the reference does in
    the sampler constructor (PosteriorSampler.py:38-59)."""
    ens0 = ensembles[0]
    en = np.array([np.asarray(e.energies, dtype=np.float64) for e in ensembles])
    tabs = dict(n_states=en.shape[1], lam=[float(e.lam) for e in ensembles], energy=en,
                logZ=precompute.logz(en, device), restraints=[])
    for col in ens0.columns:
        if col["ref"] not in ("uniform", "exponential", "gaussian"):
            raise ValueError('Please choose a reference potential of the following:\n \
                    {%s,%s,%s}' % ('uniform', 'exponential', 'gaussian'))
        nlsig, two, cnorm = precompute.sigma_vectors(col["ndof"], col["sigma"])
        if col["kind"] == "pfgrid":
            if col["ref"] == "gaussian":      # the reference fails here too (PosteriorSampler.py:52 vs :181)
                raise TypeError("build_gaussian_ref_pf() got an unexpected keyword argument 'use_global_ref_sigma'")
            tabs["restraints"].append(dict(kind="pfgrid", name="pf", extension=col["extension"], ref=col["ref"],
                                           ndof=col["ndof"], sigma=col["sigma"], nlsig=nlsig, two_sig2=two, cnorm=cnorm,
                                           grids=col["grids"], prior=col["prior"], pf=col["pf"], sse=None, refsum=None))
            continue
        d = dict(kind=col["kind"], name=col["name"], extension=col["extension"], ref=col["ref"], ndof=col["ndof"],
                 sigma=col["sigma"], nlsig=nlsig, two_sig2=two, cnorm=cnorm, sse=col["sse"],
                 grids=[col["gamma"]] if col["kind"] == "gamma" else [], refsum=None)
        if col["ref"] != "uniform":
            out = precompute.restraint_tables(col["kind"], col["model"], col["exp"], col["weight"], ref=col["ref"],
                                              gamma=col["gamma"], log_normal=col["log_normal"],
                                              use_global_ref_sigma=col["use_global_ref_sigma"], device=device)
            d["refsum"] = out["refsum"]
            for key in ("betas", "ref_mean", "ref_sigma"):
                if key in out:
                    d[key] = out[key]
        tabs["restraints"].append(d)
    return tabs


class PosteriorSampler(object):

    def __init__(self, ensemble, freq_write_traj=100., freq_save_traj=100., verbose=False, nchains=1, seed=None,
                 device=0, rng=None):
        self.lam = ensemble.lam
        self.ensemble = ensemble.to_list()
        self.nreplicas = 1
        self.write_traj = freq_write_traj
        self.traj_every = freq_save_traj
        self.nstates = len(self.ensemble)
        self.nchains = int(nchains)
        self.device = int(device)
        if rng is None:
            # drop-in default: ONE chain and no explicit seed -> numpy's global MT19937 stream, draw for draw like the
            # reference (np.random.seed(s) then gives the reference's own trajectory); otherwise Philox
            rng = "numpy" if (self.nchains == 1 and seed is None) else "philox"
        if rng not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' or 'numpy'")
        if rng == "numpy" and self.nchains != 1:
            raise ValueError("rng='numpy' replays the reference's single chain on np.random's stream: nchains must be 1")
        self.rng = rng
        self._init_state = None
        if rng == "numpy":
            # the reference's own draw of the initial state (PosteriorSampler.py:29)
            self._init_state = np.random.randint(low=0, high=self.nstates, size=self.nreplicas).astype(np.int32)
            self.seed = 0
        else:
            self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.E = 1.0e99
        self.accepted = 0
        self.total = 0
        self.verbose = verbose

        self._columns = ensemble.columns
        self.tables = build_tables([ensemble], device=self.device)
        self.logZ = float(self.tables["logZ"][0])
        # reference-potential parameters back onto the restraint objects (PosteriorSampler.py:102-104, 147-150)
        for k, rt in enumerate(self.tables["restraints"]):
            if rt["kind"] == "pfgrid":      # its reference term is evaluated on the fly by the sampler
                continue
            for i, s in enumerate(self.ensemble):
                R = s[k]
                if rt["ref"] == "exponential":
                    R.betas = rt["betas"]
                    R.sum_neglog_exp_ref = rt["refsum"][i]
                elif rt["ref"] == "gaussian":
                    R.ref_mean, R.ref_sigma = rt["ref_mean"], rt["ref_sigma"]
                    R.sum_neglog_gaussian_ref = rt["refsum"][i]
        self._device_tables = None
        self._chains = None
        # initial state of chain 0 (reference: np.random.randint at PosteriorSampler.py:29; here the
        # Philox block of (seed, chain 0, step 2^64-1)); known once the chains exist
        self._ensure_chains()
        first = self._chains.fetch()
        self.state = np.array([int(first["state"][0])])
        self.indices = [int(x) for x in first["indices"][0]]
        self.traj = PosteriorSamplingTrajectory(ensemble=ensemble, sampler=self, nreplicas=self.nreplicas)
        for i, rt in enumerate(self.tables["restraints"]):
            if rt["ref"] == "uniform":
                self.traj.ref[i].append('Nan')
            elif rt["ref"] == "exponential":
                self.traj.ref[i].append(rt.get("betas"))
            else:
                self.traj.ref[i].append(rt["ref_mean"])
                self.traj.ref[i].append(rt["ref_sigma"])
        self._calls = 0

    # -- reference potentials of one restraint type, recomputed on the GPU (PosteriorSampler.py:74-150) --------
    def _build_ref(self, rest_index, ref, use_global_ref_sigma=True):
        col = None
        k = 0
        if getattr(self, "_columns", None) is None:
            raise RuntimeError("build_*_ref needs the raw observables: not available on an unpickled sampler")
        if getattr(self, "_calls", 0) > 0:
            # the rebuild replaces the device tables and with them the chains' state, energy and histograms, which the
            # reference keeps across this call (PosteriorSampler.py:74-150 only touches the restraints)
            raise RuntimeError("build_*_ref after sample(): rebuild the reference potentials before the first sample() "
                               "call (or on a new PosteriorSampler)")
        for c in self._columns:
            if k == rest_index:
                col = c
            k += 1
        if col is None or col["kind"] == "pfgrid":
            raise ValueError("restraint %r has no tabulated reference potential" % (rest_index,))
        out = precompute.restraint_tables(col["kind"], col["model"], col["exp"], col["weight"], ref=ref,
                                          gamma=col["gamma"], log_normal=col["log_normal"],
                                          use_global_ref_sigma=use_global_ref_sigma, device=self.device)
        rt = self.tables["restraints"][rest_index]
        rt["ref"], rt["refsum"] = ref, out["refsum"]
        for key in ("betas", "ref_mean", "ref_sigma"):
            if key in out:
                rt[key] = out[key]
        for i, s_ in enumerate(self.ensemble):
            R = s_[rest_index]
            R.ref = ref
            if ref == "exponential":
                R.betas, R.sum_neglog_exp_ref = out["betas"], out["refsum"][i]
            else:
                R.ref_mean, R.ref_sigma, R.sum_neglog_gaussian_ref = out["ref_mean"], out["ref_sigma"], out["refsum"][i]
        # the device copy of the tables is stale now
        if self._chains is not None:
            self._chains.close()
        if self._device_tables is not None:
            self._device_tables.close()
        self._chains = self._device_tables = None

    def build_exp_ref(self, rest_index, verbose=False):
        """beta_j = sum_i r_j(X_i) / (S + 1) and the per-state sums of -ln P_ref (PosteriorSampler.py:74-104)."""
        self._build_ref(rest_index, "exponential")

    def build_gaussian_ref(self, rest_index, use_global_ref_sigma=True, verbose=False):
        """mean / sigma per observable and the per-state sums of -ln P_ref (PosteriorSampler.py:107-150)."""
        self._build_ref(rest_index, "gaussian", use_global_ref_sigma)

    # -- device objects are rebuilt on demand (a pickled sampler carries only host tables) -----
    def _ensure_chains(self):
        if self._device_tables is None:
            self._device_tables = engine.DeviceTables(self.tables, device=self.device)
        if self._chains is None:
            self._chains = self._device_tables.chains(self.nchains, self.seed, n_groups=1,
                                                      chain_group=np.zeros(self.nchains, np.int32),
                                                      init_state=getattr(self, "_init_state", None),
                                                      rng_mode=1 if getattr(self, "rng", "philox") == "numpy" else 0)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_device_tables"] = None
        d["_chains"] = None
        d["_columns"] = None          # raw observables stay out of the pickle (build_*_ref needs a live sampler)
        return d

    def compute_logZ(self):
        self.logZ = float(precompute.logz(self.tables["energy"], self.device)[0])

    def compile_nuisance_parameters(self, verbose=False):
        """Allowed values of every nuisance parameter, restraint by restraint (PosteriorSampler.py:212-230)."""
        para = []
        for R in self.ensemble[0]:
            para.extend(np.array(g) for g in R._param_grids())
        self.nuisance_para = np.array(para, dtype=object)
        return self.nuisance_para

    def _layout(self):
        rest_index, rest_type = [], []
        for i, R in enumerate(self.ensemble[0]):
            for name in R._param_names():
                rest_type.append(name + "_" + R._label())
                rest_index.append(i)
        return rest_index, rest_type

    def neglogP(self, states, parameters, parameter_indices):
        """-ln P of a configuration (PosteriorSampler.py:233-248), evaluated by bcp_neglogp.
        `parameter_indices` is the per-restraint nested index list; `parameters` (the values) is
        accepted for signature compatibility -- the energy depends on the grid indices only."""
        self._ensure_chains()
        flat = [int(i) for sub in parameter_indices for i in sub]
        result = 0
        for state in states:
            result += float(self._device_tables.neglogp([int(state)], [flat])[0])
        return result

    def sample(self, nsteps, burn=0, print_freq=1000, verbose=False, progress=True):
        """nsteps of Metropolis sampling (+ burn) for every chain (PosteriorSampler.py:251-374)."""
        self._ensure_chains()
        allowed = self.compile_nuisance_parameters()
        rest_index, self.rest_type = self._layout()
        n_rest = max(rest_index) + 1
        P = len(rest_index)
        blk = self._chains
        if self._calls > 0:
            # what the reference does on a second sample() call: indices restart at the restraints'
            # never-updated *_index attributes while state and E persist (PosteriorSampler.py:271-278)
            start = [int(i) for R in self.ensemble[0] for i in R._param_start()]
            blk.set_state(indices=np.tile(np.array(start, np.int32), (self.nchains, 1)))
        every = int(self.traj_every)
        # the reference keeps one state trace per sampler: chain 0 stands for it
        run = blk.sample_numpy_stream if getattr(self, "rng", "philox") == "numpy" else blk.sample
        o = run(int(nsteps), burn=int(burn), traj_every=every, record_state_trace=True, trace_chains=1)
        self._calls += 1

        # chain 0 -> the reference's attribute layout
        self.E = float(o["energy"][0])
        self.state = np.array([int(o["state"][0])])
        self.indices = [int(x) for x in o["indices"][0]]
        self.values = [allowed[p][self.indices[p]] for p in range(P)]
        self.accepted = float(o["accepted"].sum()) if self.nchains > 1 else float(o["accepted"][0])
        self.total = float(o["total"].sum()) if self.nchains > 1 else float(o["total"][0])
        self.accept = bool(o["traj_accept"][-1, 0]) if o["traj_accept"].shape[0] else False

        T = self._device_tables
        t = self.traj
        t.state_counts = np.ones(self.nstates) + o["state_counts"][0].astype(np.float64)
        off = T.hist_offsets
        t.sampled_parameters = [o["param_counts"][0][off[p]:off[p] + T.grid_len[p]].astype(np.float64) for p in range(P)]
        t.state_trace.extend(int(x) for x in o["state_trace"][:, 0])
        nsnap = o["traj_energy"].shape[0]
        idx0 = o["traj_indices"][:, :, 0]
        for m in range(nsnap):
            nested = [[] for _ in range(n_rest)]
            for p, r in enumerate(rest_index):
                nested[r].append(int(idx0[m, p]))
            t.trajectory.append([int(m * every), float(o["traj_energy"][m, 0]), int(o["traj_accept"][m, 0]),
                                 [int(o["traj_state"][m, 0])], nested])
            t.traces.append([allowed[p][idx0[m, p]] for p in range(P)])
        t.chains = dict(energy=o["traj_energy"], state=o["traj_state"], accept=o["traj_accept"],
                        indices=o["traj_indices"], accepted=o["accepted"], sep_accepted=o["sep_accepted"])
        total = o["total"].astype(np.float64)
        sep = o["sep_accepted"].astype(np.float64)
        if self.nchains > 1:
            sep_pct = sep.sum(axis=0) / total.sum() * 100.
            acc_pct = o["accepted"].sum() / total.sum() * 100.
        else:
            sep_pct = sep[0] / total[0] * 100.
            acc_pct = o["accepted"][0] / total[0] * 100.
        if verbose or self.verbose:
            print('\nAccepted %s %% \n' % acc_pct)
            print('\nAccepted [ ...Nuisance paramters..., state] %')
            print('Accepted %s %% \n' % sep_pct)
        t.sep_accept.append(sep_pct)
        t.sep_accept.append(acc_pct)


class PosteriorSamplingTrajectory(object):
    """Container of what a sampling run produced (PosteriorSampler.py:378-488)."""

    def __init__(self, ensemble, sampler, nreplicas):
        self.lam = ensemble.lam
        self.sampler = sampler
        self.ensemble = ensemble.to_list()
        self.nreplicas = nreplicas
        self.nstates = len(self.ensemble)
        self.state_counts = np.ones(self.nstates)      # pseudo-count, as the reference (:393)
        self.sampled_parameters = []
        self.allowed_parameters = []
        self.ref = [[] for _ in range(len(self.ensemble[0]))]
        self.model = [[] for _ in range(len(self.ensemble[0]))]
        self.sep_accept = []
        self.state_trace = []
        parameter_indices = []
        self.rest_type = []
        for R in self.ensemble[0]:
            for g in R._param_grids():
                self.allowed_parameters.append(g)
                self.sampled_parameters.append(np.zeros(len(g)))
            parameter_indices.extend(R._param_start())
            for name in R._param_names():
                self.rest_type.append(name + "_" + R._label())
        self.trajectory_headers = ["step", "E", "accept", "state", "para_index = %s" % parameter_indices]
        self.trajectory = []
        self.traces = []
        self.results = {}
        self.chains = None

    def process_results(self, filename=None):
        """Write the trajectory dict as npz and the sampler as a pickle next to it
        (PosteriorSampler.py:422-470); file naming is part of the contract (Analysis parses lambda)."""
        if filename is None:
            filename = f"traj_lambda{self.lam}.npz"
        for rest_index in range(len(self.ensemble[0])):
            if not self.model[rest_index] and self.ensemble[0][rest_index]._model is not None:
                model = np.stack([s[rest_index]._model for s in self.ensemble])     # [S][J]
                self.model[rest_index] = [list(model[:, n]) for n in range(model.shape[1])]
        for key in ('rest_type', 'trajectory_headers', 'trajectory', 'sep_accept', 'allowed_parameters',
                    'sampled_parameters', 'model', 'ref', 'traces', 'state_trace'):
            self.results[key] = getattr(self, key)
        self.write(filename, self.results)
        save_object(self.sampler, filename.replace(".npz", ".pkl"))

    def write(self, filename='traj.npz', *args, **kwds):
        np.savez_compressed(filename, *args, **kwds)
