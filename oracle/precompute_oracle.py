"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (biceps_b200/).

numpy restatement of the per-state restraint-energy precompute (SURVEY.md section 8 rows
a1, a2, a4, a5, a7), vectorised over states but sequential over observables j so that the
left-to-right fp64 summation order of the reference's Python loops is preserved.

  compute_sse (cs / J / pf precomputed)     /root/reference/biceps/Restraint.py:344-353, 445-454, 736-741
  Restraint_noe.compute_sse                 /root/reference/biceps/Restraint.py:552-569
  build_exp_ref + compute_neglog_exp_ref    /root/reference/biceps/PosteriorSampler.py:74-104, Restraint.py:275-286
  build_gaussian_ref + compute_neglog_...   /root/reference/biceps/PosteriorSampler.py:107-150, Restraint.py:288-298
  compute_logZ                              /root/reference/biceps/PosteriorSampler.py:64-71
  sigma / gamma grids                       /root/reference/biceps/Restraint.py:235-241, 502-507
  equivalency-group weights                 /root/reference/biceps/Restraint.py:418-442, 527-549

Pinned by tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py against tables the live reference produced
(tests/golden/cineromycin_tables.npz, apomyoglobin_tables.npz).
"""
import math

import numpy as np


def log_grid(lo, hi, step):
    """allowed_sigma / allowed_gamma (Restraint.py:238-239, 505)."""
    return np.exp(np.arange(np.log(lo), np.log(hi), np.log(step)))


def group_weights(restraint_index):
    """1/n weight per equivalency group (Restraint.py:418-442)."""
    ri = list(restraint_index)
    counts = {}
    for r in ri:
        counts[r] = counts.get(r, 0) + 1
    return np.array([1.0 / float(counts[r]) for r in ri], dtype=np.float64)


_libm_pow2 = np.frompyfunc(lambda x: math.pow(x, 2.0), 1, 1)


def square(err, exact_pow):
    """`err**2.0` of the reference is a scalar libm pow(); glibc's pow(x, 2) is not always the
    correctly rounded x*x (it differs in the last bit for ~1 % of inputs), so the bit-faithful
    mode calls libm per element.  exact_pow=False uses x*x (for sizes where a Python-level call
    per element is too slow; the difference is <= 1 ulp per term)."""
    if exact_pow:
        return _libm_pow2(err).astype(np.float64)
    return err * err


def ndof(weight):
    n = 0.0
    for w in weight:
        n += w
    return n


def sse_scalar(model, exp, weight, exact_pow=True):
    """sse[S] = sum_j w_j (model_ij - exp_j)^2, summed left to right over j."""
    model = np.asarray(model, np.float64)
    sse = np.zeros(model.shape[0])
    for j in range(model.shape[1]):
        err = model[:, j] - exp[j]
        sse += weight[j] * square(err, exact_pow)
    return sse


def sse_gamma(model, exp, weight, gamma, log_normal=False, exact_pow=True):
    """sse[S][G] (Restraint.py:552-569): err = gamma*exp - model, or ln(model/(gamma*exp))."""
    model = np.asarray(model, np.float64)
    gamma = np.asarray(gamma, np.float64)
    sse = np.zeros((model.shape[0], gamma.shape[0]))
    for j in range(model.shape[1]):
        if log_normal:
            err = np.log(model[:, j][:, None] / (gamma[None, :] * exp[j]))
        else:
            err = gamma[None, :] * exp[j] - model[:, j][:, None]
        sse += weight[j] * square(err, exact_pow)
    return sse


def exp_ref(model, weight):
    """betas[J] and refsum[S] of the exponential reference potential."""
    model = np.asarray(model, np.float64)
    S, J = model.shape
    betas = np.array([model[:, j].sum() / (S + 1.0) for j in range(J)])      # PosteriorSampler.py:100
    ref = np.zeros(S)
    for j in range(J):
        ref += weight[j] * (np.log(betas[j]) + model[:, j] / betas[j])       # Restraint.py:284-286
    return betas, ref


def gaussian_ref(model, weight, use_global_ref_sigma=True):
    """ref_mean[J], ref_sigma[J], refsum[S] of the Gaussian reference potential."""
    model = np.asarray(model, np.float64)
    S, J = model.shape
    mean = np.array([model[:, j].mean() for j in range(J)])                  # :133
    sig = np.array([np.sqrt(((model[:, j] - mean[j]) ** 2.0).sum() / (S + 1.0)) for j in range(J)])  # :134-135
    if use_global_ref_sigma:
        g = (np.array([sig[j] ** (-2.0) for j in range(J)]).mean()) ** -0.5  # :141
        sig = np.full(J, g)
    ref = np.zeros(S)
    for j in range(J):
        ref += weight[j] * (np.log(np.sqrt(2.0 * np.pi)) + np.log(sig[j])
                            + (model[:, j] - mean[j]) ** 2.0 / (2.0 * sig[j] ** 2.0))  # Restraint.py:295-298
    return mean, sig, ref


def logZ(energy):
    """ln sum_i exp(-energy_i), sequential sum, no max shift (PosteriorSampler.py:64-71)."""
    Z = 0.0
    for e in np.asarray(energy, np.float64):
        Z += np.exp(-e)
    return float(np.log(Z))


def sigma_vectors(ndof_, sigma):
    """nlsig, two_sig2, cnorm with the reference's scalar expressions (Restraint.py:376-378)."""
    nlsig = np.array([ndof_ * np.log(x) for x in sigma], dtype=np.float64)
    two = np.array([2.0 * float(x) ** 2.0 for x in sigma], dtype=np.float64)
    return nlsig, two, float(ndof_ / 2.0 * np.log(2.0 * np.pi))


def build_tables(energies, lambdas, restraints, exact_pow=True):
    """Flat tables from raw observables.

    restraints: list of dicts {kind: 'scalar'|'gamma', model[S,J], exp[J], weight[J],
    sigma: (min,max,step), gamma: (min,max,step) (gamma only), ref: 'uniform'|'exponential'|
    'gaussian', log_normal, use_global_ref_sigma, name}.  Returns the tables dict of
    ref_harness.extract_tables with energy [K][S] and logZ [K]."""
    energies = np.asarray(energies, np.float64)
    lam = [float(l) for l in np.atleast_1d(lambdas)]
    en = np.array([l * energies for l in lam])                               # Restraint.py:99
    out = dict(n_states=energies.shape[0], lam=lam, energy=en,
               logZ=np.array([logZ(e) for e in en]), restraints=[])
    for rd in restraints:
        w = np.asarray(rd["weight"], np.float64)
        nd = ndof(w)
        sigma = log_grid(*rd["sigma"])
        nlsig, two, cnorm = sigma_vectors(nd, sigma)
        d = dict(kind=rd["kind"], name=rd.get("name", rd["kind"]), ref=rd.get("ref", "uniform"), ndof=nd,
                 sigma=sigma, nlsig=nlsig, two_sig2=two, cnorm=cnorm, grids=[])
        if rd["kind"] == "gamma":
            gamma = log_grid(*rd["gamma"])
            d["grids"] = [gamma]
            d["sse"] = sse_gamma(rd["model"], rd["exp"], w, gamma, rd.get("log_normal", False), exact_pow)
        else:
            d["sse"] = sse_scalar(rd["model"], rd["exp"], w, exact_pow)
        if d["ref"] == "exponential":
            d["betas"], d["refsum"] = exp_ref(rd["model"], w)
        elif d["ref"] == "gaussian":
            d["ref_mean"], d["ref_sigma"], d["refsum"] = gaussian_ref(rd["model"], w, rd.get("use_global_ref_sigma", True))
        else:
            d["refsum"] = None
        out["restraints"].append(d)
    return out
