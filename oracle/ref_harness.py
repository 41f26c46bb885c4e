"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (biceps_b200/).

Harness that imports the *unmodified* reference package from /root/reference in THIS
container (it does not exist on the GPU box), with import shims for the four absent
third-party modules (oracle/stubs/), and flattens reference objects into the plain
"tables" dict that every other oracle module and the CUDA C-ABI consume.

Used only by tests/golden/make_golden.py (fixture generation) and by the CPU-side
tests that pin oracle/*.py against the live reference.

Reference call sites this harness drives (all under /root/reference/biceps/):
  Ensemble / initialize_restraints      Restraint.py:81-197
  Restraint_{cs,J,noe,pf}               Restraint.py:301-894
  PosteriorSampler ctor / sample        PosteriorSampler.py:11-71, 251-374
  Analysis.MBAR_analysis                Analysis.py:137-227
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("BICEPS_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "biceps"))


def import_reference():
    """Import the reference `biceps` package (v2.0) with shims; returns the module."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):      # swallow the import banner
        import biceps  # noqa
    return biceps


@contextlib.contextmanager
def quiet():
    """The reference prints acceptance summaries and draws a tqdm bar; silence both."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def restraint_kind(R):
    name = type(R).__name__
    if name == "Restraint_noe":
        return "gamma"
    if name == "Restraint_pf" and not R.precomputed:
        return "pfgrid"
    return "scalar"


PF_GRID_ATTRS = ("allowed_beta_c", "allowed_beta_h", "allowed_beta_0",
                 "allowed_xcs", "allowed_xhs", "allowed_bs")


def extract_tables(sampler):
    """Flatten a constructed reference PosteriorSampler into the tables of SURVEY A.4.

    Everything `sample()` reads is captured: energy[S] (already lambda-scaled,
    Restraint.py:99), logZ (PosteriorSampler.py:64-71) and per restraint Ndof, the
    sigma grid, optional gamma / PF grids, sse, the reference-potential sum that
    compute_neglogP subtracts (Restraint.py:379-382, 588-591, 884-893) and pf_prior.
    The sigma-derived vectors are evaluated with the reference's own scalar
    expressions (Restraint.py:376-378) so that energies are reproducible bit for bit.
    """
    ens = sampler.ensemble
    S = len(ens)
    tabs = dict(n_states=S, lam=float(sampler.lam),
                energy=np.array([float(s[0].energy) for s in ens], dtype=np.float64),
                logZ=float(sampler.logZ), restraints=[])
    for r, R0 in enumerate(ens[0]):
        kind = restraint_kind(R0)
        sig = np.array(R0.allowed_sigma, dtype=np.float64)
        ndof = float(R0.Ndof)
        use_float = type(R0).__name__ in ("Restraint_cs", "Restraint_J") or kind == "scalar"
        d = dict(kind=kind, name=type(R0).__name__.split("_")[-1], extension=str(R0.extension),
                 ref=str(R0.ref), ndof=ndof, sigma=sig,
                 nlsig=np.array([ndof * np.log(x) for x in sig], dtype=np.float64),
                 two_sig2=np.array([(2.0 * float(x) ** 2.0) if use_float else (2.0 * x ** 2.0)
                                    for x in sig], dtype=np.float64),
                 cnorm=float(ndof / 2.0 * np.log(2.0 * np.pi)))
        if R0.ref == "exponential":
            refattr = "sum_neglog_exp_ref"
        elif R0.ref == "gaussian":
            refattr = "sum_neglog_gaussian_ref"
        else:
            refattr = None
        if kind == "scalar":
            d["sse"] = np.array([float(s[r].sse) for s in ens], dtype=np.float64)
            d["grids"] = []
        elif kind == "gamma":
            d["sse"] = np.array([np.asarray(s[r].sse, dtype=np.float64) for s in ens])
            d["grids"] = [np.array(R0.allowed_gamma, dtype=np.float64)]
            d["log_normal"] = bool(R0.log_normal)
        else:
            d["sse"] = np.array([np.asarray(s[r].sse, dtype=np.float64) for s in ens])
            d["grids"] = [np.array(getattr(R0, a), dtype=np.float64) for a in PF_GRID_ATTRS]
            prior = getattr(R0, "pf_prior", None)
            d["prior"] = None if prior is None else np.asarray(prior, dtype=np.float64)
        if refattr is None:
            d["refsum"] = None
        else:
            d["refsum"] = np.array([np.asarray(getattr(s[r], refattr), dtype=np.float64)
                                    for s in ens])
        tabs["restraints"].append(d)
    return tabs


def extract_observables(sampler):
    """model[S,J], exp[J], weight[J] per restraint (inputs of the precompute kernels)."""
    ens = sampler.ensemble if hasattr(sampler, "ensemble") else sampler
    out = []
    for r, R0 in enumerate(ens[0]):
        if restraint_kind(R0) == "pfgrid":
            out.append(None)
            continue
        model = np.array([[float(s[r].restraints[j]["model"]) for j in range(R0.n)] for s in ens])
        exp = np.array([float(R0.restraints[j]["exp"]) for j in range(R0.n)])
        w = np.array([float(R0.restraints[j]["weight"]) for j in range(R0.n)])
        out.append(dict(model=model, exp=exp, weight=w))
    return out


def cineromycin_inputs(biceps):
    """The C1 workload of BASELINE.json: docs/examples/datasets/cineromycin_B (J + NOE)."""
    D = os.path.join(REFERENCE_ROOT, "docs/examples/datasets/cineromycin_B/")
    energies = np.loadtxt(D + "cineromycinB_QMenergies.dat") * 627.509 / 0.5959
    energies = energies - energies.min()
    input_data = biceps.toolbox.sort_data(D + "J_NOE")
    return energies, input_data


def build_reference_sampler(biceps, lam, energies, input_data, options, seed):
    """np.random.seed(seed) immediately before the ctor (which draws the initial state,
    PosteriorSampler.py:29), as SURVEY 8c prescribes."""
    ens = biceps.Ensemble(float(lam), energies)
    ens.initialize_restraints(input_data, options)
    np.random.seed(seed)
    return biceps.PosteriorSampler(ens)
