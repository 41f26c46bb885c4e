"""Restraint-energy precompute on the GPU (K1/K2/K4 of DESIGN.md): raw observables ->
the flat tables the sampler consumes.  Host code only builds grids and tiny J-length vectors;
every S-sized quantity comes from the CUDA kernels in csrc/precompute.cu."""
import ctypes as C
import math

import numpy as np

from . import _cabi as A

REF_KINDS = {"uniform": 0, "exponential": 1, "gaussian": 2}


def log_grid(lo, hi, step):
    """Log-spaced nuisance grid, as the reference builds it (Restraint.py:235-239, 502-505)."""
    return np.exp(np.arange(np.log(lo), np.log(hi), np.log(step)))


def group_weights(restraint_index):
    """1/n weight per equivalency group (Restraint.py:418-442, 527-549)."""
    counts = {}
    for r in restraint_index:
        counts[r] = counts.get(r, 0) + 1
    return np.array([1.0 / float(counts[r]) for r in restraint_index], dtype=np.float64)


def ndof(weight):
    """Ndof = left-to-right fp64 sum of the weights (Restraint.py:351-352)."""
    n = 0.0
    for w in weight:
        n += float(w)
    return n


def sigma_vectors(ndof_, sigma):
    """Ndof*ln(sigma), 2*sigma^2, Ndof/2*ln(2 pi) with the reference's scalar expressions
    (Restraint.py:376-378), so energies are reproducible bit for bit."""
    nlsig = np.array([ndof_ * np.log(x) for x in sigma], dtype=np.float64)
    two = np.array([2.0 * float(x) ** 2.0 for x in sigma], dtype=np.float64)
    return nlsig, two, float(ndof_ / 2.0 * np.log(2.0 * np.pi))


def restraint_tables(kind, model, exp, weight, ref="uniform", gamma=None, log_normal=False,
                     use_global_ref_sigma=True, device=0):
    """One bcp_precompute call: returns dict(sse, refsum, ref parameters)."""
    if ref not in REF_KINDS:
        raise ValueError('Please choose a reference potential of the following:\n \
                    {%s,%s,%s}' % ('uniform', 'exponential', 'gaussian'))
    model = A.f64(model)
    if model.ndim != 2:
        raise ValueError("model must be [n_states][n_observables]")
    S, J = model.shape
    exp, weight = A.f64(exp), A.f64(weight)
    if exp.shape != (J,) or weight.shape != (J,):
        raise ValueError("exp and weight must have one entry per observable")
    is_gamma = kind == "gamma"
    g = A.f64(gamma) if is_gamma else None
    G = g.shape[0] if is_gamma else 1
    obs = A.bcp_obs(A.BCP_KIND_GAMMA if is_gamma else A.BCP_KIND_SCALAR, J, REF_KINDS[ref], int(bool(log_normal)),
                    int(bool(use_global_ref_sigma)), G, A.dptr(model), A.dptr(exp), A.dptr(weight), A.dptr(g))
    sse = np.zeros((S, G) if is_gamma else (S,), np.float64)
    refsum = np.zeros(S, np.float64) if REF_KINDS[ref] else None
    p0 = np.zeros(J, np.float64) if REF_KINDS[ref] else None
    p1 = np.zeros(J, np.float64) if REF_KINDS[ref] == 2 else None
    A.check(A.lib().bcp_precompute(C.byref(obs), S, int(device), A.dptr(sse), A.dptr(refsum), A.dptr(p0), A.dptr(p1)),
            "bcp_precompute")
    out = dict(sse=sse, refsum=refsum)
    if ref == "exponential":
        out["betas"] = p0
    elif ref == "gaussian":
        out["ref_mean"], out["ref_sigma"] = p0, p1
    return out


def logz(energy, device=0):
    energy = np.atleast_2d(A.f64(energy))
    K, S = energy.shape
    out = np.zeros(K, np.float64)
    A.check(A.lib().bcp_logz(A.dptr(energy), K, S, int(device), A.dptr(out)), "bcp_logz")
    return out


def build_tables(energies, lambdas, restraints, device=0):
    """Flat sampler tables from raw observables (same dict layout the oracle produces).

    restraints: list of dicts {kind: 'scalar'|'gamma', model[S][J], exp[J], weight[J],
    sigma: (min, max, step), gamma: (min, max, step), ref, log_normal, use_global_ref_sigma, name}."""
    energies = A.f64(energies)
    lam = [float(l) for l in np.atleast_1d(lambdas)]
    en = np.array([l * energies for l in lam])                   # Restraint.py:99
    out = dict(n_states=energies.shape[0], lam=lam, energy=en, logZ=logz(en, device), restraints=[])
    for rd in restraints:
        w = A.f64(rd["weight"])
        nd = ndof(w)
        sigma = log_grid(*rd["sigma"])
        nlsig, two, cnorm = sigma_vectors(nd, sigma)
        d = dict(kind=rd["kind"], name=rd.get("name", rd["kind"]), ref=rd.get("ref", "uniform"), ndof=nd,
                 sigma=sigma, nlsig=nlsig, two_sig2=two, cnorm=cnorm, grids=[])
        gamma = None
        if rd["kind"] == "gamma":
            gamma = log_grid(*rd["gamma"])
            d["grids"] = [gamma]
        d.update(restraint_tables(rd["kind"], rd["model"], rd["exp"], w, d["ref"], gamma, rd.get("log_normal", False),
                                  rd.get("use_global_ref_sigma", True), device))
        out["restraints"].append(d)
    return out
