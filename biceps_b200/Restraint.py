"""Ensemble and Restraint_{cs,J,noe,pf}: the reference's restraint API over columnar storage.

Mirrors /root/reference/biceps/Restraint.py:81-894 (same class names, constructor / init_restraint
arguments, option routing, error messages and attributes), but an ensemble is held as one dense
fp64 table per restraint type (model[S][J], exp[J], weight[J]) and every S-sized quantity -- the
sum of squared errors over the gamma grid, the reference-potential sums -- is computed by the
CUDA kernels of csrc/precompute.cu in one call per restraint type instead of S*J*G Python
iterations.  `ensemble.to_list()[state][k]` still returns one restraint object per (state, type);
its `.restraints` list of dicts is materialised lazily from the columns.
"""
import inspect
import sys

import numpy as np
import pandas as pd

from . import precompute, toolbox


def get_restraint_options(input_data=None):
    """List of default option dicts, one per restraint class (or per restraint of input_data),
    as biceps.get_restraint_options (Restraint.py:10-75)."""
    current_module = sys.modules[__name__]
    restraints = toolbox.list_possible_restraints()
    extensions = None
    if input_data is not None:
        scheme = toolbox.list_res(input_data)
        extensions = [res.split("_")[-1] for res in scheme]
        restraints = [r for rest in scheme for r in restraints if r.split("_")[-1] in rest]
    options = []
    for k, name in enumerate(restraints):
        R = getattr(current_module, name)
        opt = {}
        for key, par in inspect.signature(R).parameters.items():
            opt[key] = [] if par.default is inspect._empty else par.default
        for key, par in inspect.signature(R.init_restraint).parameters.items():
            opt[key] = np.nan if par.default is inspect._empty else par.default
        for key in ("data", "energy", "self", "verbose"):
            opt.pop(key, None)
        if extensions is not None:
            opt["extension"] = extensions[k]
        options.append(opt)
    return options


class Restraint(object):
    """Parent restraint class (Restraint.py:201-298): sigma grid and reference-potential slots."""

    kind = "scalar"
    _columns = ("exp", "model")

    def __init__(self, ref="uniform", sigma=[0.05, 20.0, 1.02], use_global_ref_sigma=True, verbose=False):
        self.betas = None
        self.neglog_exp_ref = None
        self.sum_neglog_exp_ref = 0.0
        self.ref_sigma = None
        self.ref_mean = None
        self.neglog_gaussian_ref = None
        self.sum_neglog_gaussian_ref = 0.0
        self.use_global_ref_sigma = use_global_ref_sigma
        self._sse = None
        self.ref = ref
        # log-spaced sigma grid, start index len/2 (Restraint.py:235-241)
        self.dlogsigma = np.log(sigma[2])
        self.sigma_min = sigma[0]
        self.sigma_max = sigma[1]
        self.allowed_sigma = np.exp(np.arange(np.log(self.sigma_min), np.log(self.sigma_max), self.dlogsigma))
        self.sigma_index = int(len(self.allowed_sigma) / 2)
        self.sigma = self.allowed_sigma[self.sigma_index]
        self.verbose = verbose
        self._frame = None
        self._restraints = None
        self._weights = None
        self.Ndof = None

    # -- data ------------------------------------------------------------------------------
    def load_data(self, filename, As="pickle"):
        if self.verbose:
            print('Loading %s as %s...' % (filename, As))
        return getattr(pd, "read_%s" % As)(filename)

    def _ingest(self, data, energy, extension, weight, file_fmt):
        self.extension = extension
        self.energy = energy
        frame = self.load_data(data, As=file_fmt)
        self.n = len(frame.values)
        self._frame = frame[list(self._columns_present(frame))]
        self._exp = np.asarray(self._frame["exp"], dtype=np.float64)
        self._model = np.asarray(self._frame["model"], dtype=np.float64) if "model" in self._frame else None
        self._weight_option = weight

    def _columns_present(self, frame):
        return [c for c in self._columns]

    def _compute_weights(self):
        return np.full(self.n, float(self._weight_option), dtype=np.float64)

    @property
    def weights(self):
        if self._weights is None:
            self._weights = self._compute_weights()
            self.Ndof = precompute.ndof(self._weights)
        return self._weights

    @property
    def restraints(self):
        """List of per-observable dicts {exp, model, weight, ...} like the reference's
        self.restraints (Restraint.py:336-341, 414-416, 523-525); built on first access."""
        if self._restraints is None:
            cols = {k: self._frame[k].tolist() for k in self._frame.columns}
            w = self.weights
            self._restraints = [dict({k: v[j] for k, v in cols.items()}, weight=float(w[j])) for j in range(self.n)]
        return self._restraints

    def add_restraint(self, restraint):
        self.restraints.append(restraint)

    # -- sum of squared errors -------------------------------------------------------------
    @property
    def sse(self):
        if self._sse is None:      # stand-alone restraint (not part of an Ensemble): one-state GPU call
            self._sse = self._gpu_sse(self._model[None, :])[0]
        return self._sse

    @sse.setter
    def sse(self, v):
        self._sse = v

    def _gpu_sse(self, model):
        out = precompute.restraint_tables(self.kind, model, self._exp, self.weights, ref="uniform",
                                          gamma=getattr(self, "allowed_gamma", None),
                                          log_normal=getattr(self, "log_normal", False))
        return out["sse"]

    def compute_sse(self, f):
        """(Weighted) sum of squared errors of a list of {model, exp, weight} dicts
        (Restraint.py:344-353, 445-454, 552-569), evaluated on the GPU."""
        model = np.array([[d["model"] for d in f]], dtype=np.float64)
        exp = np.array([d["exp"] for d in f], dtype=np.float64)
        w = np.array([d["weight"] for d in f], dtype=np.float64)
        self.Ndof = precompute.ndof(w)
        out = precompute.restraint_tables(self.kind, model, exp, w, gamma=getattr(self, "allowed_gamma", None),
                                          log_normal=getattr(self, "log_normal", False))
        return out["sse"][0]

    def compute_neglogP(self, parameters, parameter_indices, sse):
        """-ln P of this restraint for one state given its sum of squared errors (Restraint.py:356-383, 457-477,
        881): the reference's scalar expression (`Ndof`, not `Ndof + 1`, :376).  sse and the reference-potential
        sums come from the precompute kernels; the sampler evaluates the same expression on the device."""
        result = 0
        result += (self.Ndof) * np.log(parameters[0])
        result += sse / (2.0 * float(parameters[0]) ** 2.0)
        result += (self.Ndof) / 2.0 * np.log(2.0 * np.pi)
        if self.ref == "exponential":
            result -= self.sum_neglog_exp_ref
        if self.ref == "gaussian":
            result -= self.sum_neglog_gaussian_ref
        return result

    # label used in rest_type, reproducing the reference's parsing of str(R.__repr__)
    # (PosteriorSampler.py:280, 414): 'sigma_J', 'sigma_noe', 'gamma_noe', and for cs 'sigma_H>>'
    def _label(self):
        return type(self).__name__.split("_")[-1]

    def _param_names(self):
        return ["sigma"]

    def _param_grids(self):
        return [self.allowed_sigma]

    def _param_start(self):
        return [self.sigma_index]


class Restraint_cs(Restraint):
    """Chemical shifts (Restraint.py:301-383)."""

    _ext = ["H", "Ca", "N"]
    _columns = ("atom_index1", "exp", "model")

    def __repr__(self):
        if getattr(self, "extension", None) is not None:
            return "<%s.Restraint_cs_%s>" % (str(__name__), str(self.extension))

    def init_restraint(self, data, energy, extension="H", weight=1, file_fmt="pickle", verbose=False):
        self._ingest(data, energy, extension, weight, file_fmt)

    def _label(self):
        return "%s>>" % self.extension


class _Grouped(Restraint):
    """J and NOE: equivalency groups by restraint_index get weight 1/n (Restraint.py:418-442)."""

    def _compute_weights(self):
        ri = self._frame["restraint_index"].tolist()
        w = np.zeros(self.n, dtype=np.float64)
        groups = {}
        for i, r in enumerate(ri):
            if r is not None:
                groups.setdefault(r, []).append(i)
        for g in groups.values():
            w[g] = 1.0 / float(len(g))
        self.equivalency_groups = groups
        return w


class Restraint_J(_Grouped):
    """J coupling constants (Restraint.py:386-477)."""

    _ext = ["J"]
    _columns = ("atom_index1", "atom_index2", "atom_index3", "atom_index4", "exp", "model", "restraint_index")

    def init_restraint(self, data, energy, extension="J", weight=1, file_fmt="pickle", verbose=False):
        self._ingest(data, energy, extension, weight, file_fmt)


class Restraint_noe(_Grouped):
    """NOE distances with the gamma scaling grid (Restraint.py:481-592)."""

    _ext = ["noe"]
    _columns = ("exp", "model", "restraint_index")
    kind = "gamma"

    def init_restraint(self, data, energy, extension="noe", weight=1, file_fmt="pickle", verbose=False,
                       log_normal=False, gamma=[0.2, 10.0, 1.01]):
        self.dloggamma = np.log(gamma[2])
        self.gamma_min = gamma[0]
        self.gamma_max = gamma[1]
        self.allowed_gamma = np.exp(np.arange(np.log(self.gamma_min), np.log(self.gamma_max), self.dloggamma))
        self.gamma_index = int(len(self.allowed_gamma) / 2)
        self.gamma = self.allowed_gamma[self.gamma_index]
        self.log_normal = log_normal
        self._ingest(data, energy, extension, weight, file_fmt)

    def compute_neglogP(self, parameters, parameter_indices, sse):
        """NOE: sse is the row over the gamma grid, indexed by the gamma index (Restraint.py:572-592)."""
        return Restraint.compute_neglogP(self, parameters, parameter_indices, sse[int(parameter_indices[1])])

    def _param_names(self):
        return ["sigma", "gamma"]

    def _param_grids(self):
        return [self.allowed_sigma, self.allowed_gamma]

    def _param_start(self):
        return [self.sigma_index, self.gamma_index]


class Restraint_pf(Restraint):
    """HDX protection factors (Restraint.py:595-894).

    precomputed=True : model ln PF values come with the data file, one sigma nuisance parameter
                       (Restraint.py:736-741, 881) -- a scalar-SSE restraint.
    precomputed=False: <ln PF> = beta_c <N_c> + beta_h <N_h> + beta_0 on the 6-D nuisance grid
                       (beta_c, beta_h, beta_0, xcs, xhs, bs), Restraint.py:654-703, 765-801.  The dense
                       6-D SSE / reference arrays of the reference (11.6 MB per state by default) are
                       never built: the CUDA sampler evaluates the term from the raw <N_c>/<N_h>
                       counts (csrc/sampler_pf.cu).
    <N_c>/<N_h> files: '<Ncs_fi>/Nc_x%0.1f_b%d_state%03d.npy', '<Nhs_fi>/Nh_x...' (Restraint.py:637-644).
    With `states=[...]` the reference keeps the LAST listed state's arrays for every state
    (Restraint.py:634-644); that behaviour is reproduced.  `states=None` (not possible in the
    reference) loads the state whose index is the data file's stem (`<i>.pf`)."""

    _ext = ["pf"]
    _columns = ("atom_index1", "exp", "model")

    def init_restraint(self, data, energy, precomputed=False, pf_prior=None, Ncs_fi=None, Nhs_fi=None,
                       beta_c=(0.05, 0.25, 0.01), beta_h=(0.0, 5.2, 0.2), beta_0=(-10.0, 0.0, 0.2),
                       xcs=(5.0, 8.5, 0.5), xhs=(2.0, 2.7, 0.1), bs=(15.0, 16.0, 1.0), extension="pf", weight=1,
                       file_fmt="pickle", states=None, verbose=False):
        self.precomputed = precomputed
        if precomputed:
            self._ingest(data, energy, extension, weight, file_fmt)
            return
        self.kind = "pfgrid"
        self._columns = ("atom_index1", "exp")
        self._ingest(data, energy, extension, weight, file_fmt)
        for name, g in (("beta_c", beta_c), ("beta_h", beta_h), ("beta_0", beta_0), ("xcs", xcs), ("xhs", xhs), ("bs", bs)):
            grid = np.arange(g[0], g[1], g[2])
            setattr(self, "allowed_" + name, grid)
            setattr(self, name + "_index", int(len(grid) / 2))
            setattr(self, name, grid[int(len(grid) / 2)])
        if states is None:
            states = [int(__import__("os").path.basename(data).split(".")[0])]
        st = states[-1]
        nxc, nxh, nb = len(self.allowed_xcs), len(self.allowed_xhs), len(self.allowed_bs)
        self.Ncs = np.zeros((nxc, nb, self.n))
        self.Nhs = np.zeros((nxh, nb, self.n))
        for o in range(nxc):
            for q in range(nb):
                self.Ncs[o, q, :] = np.load('%s/Nc_x%0.1f_b%d_state%03d.npy' % (Ncs_fi, self.allowed_xcs[o], self.allowed_bs[q], st))
        for p_ in range(nxh):
            for q in range(nb):
                self.Nhs[p_, q, :] = np.load('%s/Nh_x%0.1f_b%d_state%03d.npy' % (Nhs_fi, self.allowed_xhs[p_], self.allowed_bs[q], st))
        self.pf_prior = np.load(pf_prior) if pf_prior is not None else None

    def _pf_grids(self):
        return [self.allowed_beta_c, self.allowed_beta_h, self.allowed_beta_0, self.allowed_xcs, self.allowed_xhs,
                self.allowed_bs]

    @property
    def sse(self):
        if self.kind == "pfgrid":
            raise NotImplementedError("the dense 6-D SSE of a PF-grid restraint is never materialised; "
                                      "energies come from PosteriorSampler.neglogP")
        return Restraint.sse.fget(self)

    @sse.setter
    def sse(self, v):
        self._sse = v

    def _param_names(self):
        return ["sigma"] if self.kind != "pfgrid" else ["sigma", "c", "h", "0", "xcs", "xhs", "bs"]

    def _param_grids(self):
        return [self.allowed_sigma] if self.kind != "pfgrid" else [self.allowed_sigma] + self._pf_grids()

    def _param_start(self):
        if self.kind != "pfgrid":
            return [self.sigma_index]
        return [self.sigma_index, self.beta_c_index, self.beta_h_index, self.beta_0_index, self.xcs_index,
                self.xhs_index, self.bs_index]


class Preparation(object):
    """Writes the per-state input files of Ensemble.initialize_restraints (Restraint.py:896-1152): same
    constructor, prepare_{cs,noe,J,pf} arguments, error messages, file names (`<state>.<ext>`) and DataFrame
    columns as the reference.  All S frames of a restraint type are built from one [S][J] array instead of
    S*J list appends.

    Extension: `write_as="columnar"` writes ONE file per restraint type,
    `<outdir>/ensemble.<ext>` (an npz holding model[S][J], exp[J], restraint_index[J] and the atom indices),
    which `Ensemble.from_columnar` feeds straight to the precompute kernels: no S*R small files."""

    _N_ATOMS = {"cs": 1, "noe": 2, "J": 4, "pf": 1}

    def __init__(self, nstates=0, top_file=None, outdir=None):
        self.nstates = nstates
        if top_file is not None:
            try:
                import mdtraj as md
            except ImportError as e:       # residue / atom-name columns need a topology reader
                raise ImportError("Preparation(top_file=...) needs mdtraj to read the topology") from e
            self.topology = md.load(top_file).topology
        else:
            self.topology = None
        self.outdir = outdir

    def to_sorted_list(self):
        return toolbox.sort_data(self.outdir)

    def write_DataFrame(self, filename, As="pickle", verbose=False):
        if verbose:
            print('Writing %s as %s...' % (filename, As))
        getattr(pd.DataFrame(self.biceps_df), "to_%s" % As)(filename)

    def _prepare(self, kind, ext, exp_data, model_data, indices, write_as, verbose, with_model=True):
        import os
        n_at = self._N_ATOMS[kind]
        self.ind = toolbox.check_indices(indices)
        self.exp_data = toolbox.check_exp_data(exp_data)
        self.model_data = toolbox.check_model_data(model_data)
        if int(len(self.model_data)) != int(self.nstates):
            raise ValueError("The number of states doesn't equal to file numbers")
        exp_data = np.asarray(self.exp_data)
        if self.ind.shape[0] != exp_data.shape[0]:
            raise ValueError('The number of atom pairs (%d) does not match the number of restraints (%d)! Exiting.'
                             % (self.ind.shape[0], exp_data.shape[0]))
        J = self.ind.shape[0]
        atoms = self.ind.reshape(J, -1)[:, :n_at] if (kind != "cs") else self.ind.reshape(J, 1)
        model = np.stack([np.asarray(np.loadtxt(m) if isinstance(m, str) else m, dtype=np.float64).reshape(-1)[:J]
                          for m in self.model_data]) if with_model else None
        ridx = exp_data[:, 0].astype(int)
        exp = exp_data[:, 1].astype(np.float64)
        self.header = ("exp",) + (("model",) if with_model else ()) + ("restraint_index",)
        names = {}
        if self.topology is not None:
            by_index = {a.index: a for a in self.topology.atoms}
        cols_tail = []
        for a in range(n_at):
            cols_tail.append(("atom_index%d" % (a + 1), atoms[:, a].astype(int)))
            if self.topology is not None:
                cols_tail.append(("res%d" % (a + 1), [str(by_index[int(x)].residue) for x in atoms[:, a]]))
                if kind != "pf":
                    cols_tail.append(("atom_name%d" % (a + 1), [str(by_index[int(x)].name) for x in atoms[:, a]]))
        self.header += tuple(c for c, _ in cols_tail)
        if write_as == "columnar":
            if self.outdir:
                out = dict(exp=exp, restraint_index=ridx, atoms=atoms.astype(np.int64), kind=np.array(kind),
                           extension=np.array(ext))
                if with_model:
                    out["model"] = model
                fn = os.path.join(self.outdir, "ensemble.%s" % ext)
                with open(fn, "wb") as f:
                    np.savez(f, **out)
            return
        for j in range(len(self.model_data)):
            dd = {"exp": exp}
            if with_model:
                dd["model"] = model[j]
            dd["restraint_index"] = ridx
            for c, v in cols_tail:
                dd[c] = v
            self.biceps_df = pd.DataFrame(dd)
            if verbose:
                print(self.biceps_df)
            if self.outdir:
                self.write_DataFrame(os.path.join(self.outdir, "%s.%s" % (j, ext)), As=write_as, verbose=verbose)

    def prepare_cs(self, exp_data, model_data, indices, extension, write_as="pickle", verbose=False):
        """Chemical shifts; extension = nucleus "H" | "Ca" | "N" -> files `<state>.cs_<extension>` (:937-983)."""
        self._prepare("cs", "cs_%s" % extension, exp_data, model_data, indices, write_as, verbose)

    def prepare_noe(self, exp_data, model_data, indices, extension=None, write_as="pickle", verbose=False):
        """NOE distances -> `<state>.noe` (:987-1033)."""
        self._prepare("noe", "noe", exp_data, model_data, indices, write_as, verbose)

    def prepare_J(self, exp_data, model_data, indices, extension=None, write_as="pickle", verbose=False):
        """Scalar couplings -> `<state>.J` (:1036-1094)."""
        self._prepare("J", "J", exp_data, model_data, indices, write_as, verbose)

    def prepare_pf(self, exp_data, model_data=None, indices=None, extension=None, write_as="pickle", verbose=False):
        """HDX protection factors -> `<state>.pf` (:1098-1152).  The reference's own method cannot run
        (`check_model_data(None)` raises, and with model data its `if model_data:` on an array raises); here
        model_data=None writes the experimental columns for `nstates` states (what Restraint_pf with
        precomputed=False reads) and a model array adds the `model` column (precomputed=True)."""
        if model_data is None:
            model_data = [None] * int(self.nstates)
            self._prepare("pf", "pf", exp_data, model_data, indices, write_as, verbose, with_model=False)
        else:
            self._prepare("pf", "pf", exp_data, model_data, indices, write_as, verbose)


class Ensemble(object):
    """Container of restraint objects for every conformational state (Restraint.py:81-197)."""

    def __init__(self, lam, energies, debug=False):
        self.ensemble = []
        if not isinstance(lam, float):
            raise ValueError("lambda should be a single number with type of 'float'")
        self.lam = lam
        if np.array(energies).dtype != float:
            raise ValueError("Energies should be array with type of 'float'")
        self.energies = self.lam * energies      # scale the energies (Restraint.py:99)
        self.debug = debug
        self.columns = []                         # one dict per restraint type (device-ready tables)
        self.options = None

    def to_list(self):
        return self.ensemble

    @staticmethod
    def _route_options(R, opts, input_data):
        """Split an options dict into constructor and init_restraint keyword arguments, rejecting
        unknown keys with the reference's message (Restraint.py:169-189)."""
        ctor = [k for k in inspect.signature(R).parameters]
        init = [k for k in inspect.signature(R.init_restraint).parameters if k != "self"]
        a1, a2 = {}, {}
        for key, val in opts.items():
            if key in ctor:
                a1[key] = val
            elif key in init:
                a2[key] = val
            else:
                possible_args = np.unique(np.concatenate([ctor, init]))
                possible_args = np.delete(possible_args, np.where(possible_args == "data"))
                raise TypeError(f"{key} is an invalid keyword argument for {R}\n\n\
Please check your options... \n\
The input data provided suggests the ordering of dictionaries should be: {toolbox.list_extensions(input_data)}\n\
Dictionary keys for {R.__name__} are any of: {possible_args}")
        return a1, a2

    def initialize_restraints(self, input_data, options=None, device=0):
        """Create a restraint object per (state, data file), then compute every sum of squared
        errors of the ensemble on the GPU (one bcp_precompute call per restraint type)."""
        self.options = options
        if options is None:
            options = [dict() for _ in range(len(input_data[0]))]
        current_module = sys.modules[__name__]
        routed = {}
        for i in range(self.energies.shape[0]):
            self.ensemble.append([])
            energy = self.energies[i]
            for k in range(len(input_data[0])):
                data = input_data[i][k]
                ext = data.split(".")[-1]
                restraint, extension = ext.split("_")[0], ext.split("_")[-1]
                R = getattr(current_module, "Restraint_%s" % restraint)
                if k not in routed:
                    routed[k] = self._route_options(R, options[k], input_data)
                a1, a2 = routed[k]
                obj = R(**a1)
                obj.init_restraint(data=data, energy=energy, **dict(dict(extension=extension), **a2))
                self.ensemble[-1].append(obj)
        self._finalize(device)

    @classmethod
    def from_arrays(cls, lam, energies, restraints, device=0):
        """Bulk constructor (no per-state files): restraints = list of dicts
        {cls: 'cs'|'J'|'noe', extension, model[S][J], exp[J], weight[J] or restraint_index[J], options}."""
        self = cls(lam, energies)
        S = self.energies.shape[0]
        mod = sys.modules[__name__]
        objs_per_k = []
        for rd in restraints:
            R = getattr(mod, "Restraint_%s" % rd["cls"])
            a1, a2 = cls._route_options(R, rd.get("options", {}), [["x.%s" % rd.get("extension", rd["cls"])]])
            model = np.asarray(rd["model"], np.float64)
            exp = np.asarray(rd["exp"], np.float64)
            J = exp.shape[0]
            ri = rd.get("restraint_index", list(range(J)))
            if "weight" in rd:                                  # shared by every state: summed once (reference order)
                w_given = np.asarray(rd["weight"], np.float64)
                ndof_given = precompute.ndof(w_given)
            objs = []
            for i in range(S):
                o = R(**a1)
                cols = {"exp": exp, "model": model[i]}
                for c in R._columns:
                    if c not in cols:
                        cols[c] = ri if c == "restraint_index" else list(range(J))
                o._frame = pd.DataFrame(cols) if i == 0 else objs[0]._frame
                o.extension = rd.get("extension", rd["cls"])
                o.energy = self.energies[i]
                o.n = J
                o._exp, o._model = exp, model[i]
                o._weight_option = rd.get("options", {}).get("weight", 1)
                if "weight" in rd:
                    o._weights, o.Ndof = w_given, ndof_given
                if R is Restraint_noe:
                    g = a2.get("gamma", [0.2, 10.0, 1.01])
                    o.allowed_gamma = np.exp(np.arange(np.log(g[0]), np.log(g[1]), np.log(g[2])))
                    o.gamma_index = int(len(o.allowed_gamma) / 2)
                    o.gamma = o.allowed_gamma[o.gamma_index]
                    o.log_normal = a2.get("log_normal", False)
                objs.append(o)
            objs_per_k.append(objs)
        self.ensemble = [[objs_per_k[k][i] for k in range(len(restraints))] for i in range(S)]
        for k in range(len(restraints)):                    # share weights of state 0 with every state
            w = objs_per_k[k][0].weights
            for o in objs_per_k[k][1:]:
                o._weights, o.Ndof = w, objs_per_k[k][0].Ndof
        self._finalize(device)
        return self

    @classmethod
    def from_columnar(cls, lam, energies, files, options=None, device=0):
        """Ensemble from the one-file-per-restraint-type format `Preparation.prepare_*(write_as="columnar")`
        writes: files in the order of `options`, e.g. ["d/ensemble.J", "d/ensemble.noe"]."""
        options = options if options is not None else [dict() for _ in files]
        rs = []
        for fn, opt in zip(files, options):
            z = np.load(fn, allow_pickle=False)
            kind, ext = str(z["kind"]), str(z["extension"])
            if "model" not in z:
                raise ValueError("%s holds no model data" % fn)
            rs.append(dict(cls=kind, extension=ext.split("_")[-1], model=z["model"], exp=z["exp"],
                           restraint_index=z["restraint_index"].tolist(), options=opt))
        return cls.from_arrays(lam, energies, rs, device=device)

    def _finalize(self, device):
        """Columnar tables + GPU SSE for each restraint type."""
        S = len(self.ensemble)
        self.columns = []
        for k, R0 in enumerate(self.ensemble[0]):
            objs = [s[k] for s in self.ensemble]
            w = R0.weights
            # the tables take exp / weights / Ndof from state 0: the reference reads every state's own file
            # (Restraint.py:344-353), so states whose experimental columns differ must not pass silently
            for i, o in enumerate(objs):
                if not (np.array_equal(np.asarray(o._exp), np.asarray(R0._exp)) and np.array_equal(np.asarray(o.weights), np.asarray(w))):
                    raise ValueError("state %d: the experimental values / restraint_index groups of its %s file differ from "
                                     "state 0's -- every state must carry the same experimental data" % (i, R0.extension))
            if R0.kind == "pfgrid":
                col = dict(kind="pfgrid", name="pf", extension=R0.extension, ref=R0.ref, ndof=R0.Ndof, sigma=R0.allowed_sigma,
                           grids=R0._pf_grids(), prior=R0.pf_prior, exp=R0._exp, weight=w,
                           pf=dict(Nc=np.ascontiguousarray(np.stack([o.Ncs for o in objs])),
                                   Nh=np.ascontiguousarray(np.stack([o.Nhs for o in objs])), exp=R0._exp, weight=w))
                for o in objs:
                    o._weights, o.Ndof = w, R0.Ndof
                self.columns.append(col)
                continue
            model = np.ascontiguousarray(np.stack([o._model for o in objs]), dtype=np.float64)
            col = dict(kind=R0.kind, name=type(R0).__name__.split("_")[-1], extension=R0.extension, ref=R0.ref,
                       model=model, exp=R0._exp, weight=w, ndof=R0.Ndof, sigma=R0.allowed_sigma,
                       gamma=getattr(R0, "allowed_gamma", None), log_normal=getattr(R0, "log_normal", False),
                       use_global_ref_sigma=R0.use_global_ref_sigma)
            out = precompute.restraint_tables(R0.kind, model, R0._exp, w, ref="uniform", gamma=col["gamma"],
                                              log_normal=col["log_normal"], device=device)
            col["sse"] = out["sse"]
            for i, o in enumerate(objs):
                o._sse = out["sse"][i]
                o._weights, o.Ndof = w, R0.Ndof
            self.columns.append(col)
        self.device = device
        assert S == self.energies.shape[0]
