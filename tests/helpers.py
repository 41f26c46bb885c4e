"""Shared test helpers: golden-fixture loading and dict comparison."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def tables_from_npz(z, prefix):
    """Inverse of tests/golden/make_golden.py:pack_tables."""
    t = dict(n_states=int(z[prefix + "energy"].shape[0]), lam=float(z[prefix + "lam"]),
             energy=z[prefix + "energy"], logZ=float(z[prefix + "logZ"]), restraints=[])
    for q in range(int(z[prefix + "n_restraints"])):
        p = "%sr%d_" % (prefix, q)
        rt = dict(kind=str(z[p + "kind"]), name=str(z[p + "name"]), extension=str(z[p + "extension"]),
                  ref=str(z[p + "ref"]), ndof=float(z[p + "ndof"]), cnorm=float(z[p + "cnorm"]),
                  sigma=z[p + "sigma"], nlsig=z[p + "nlsig"], two_sig2=z[p + "two_sig2"], sse=z[p + "sse"],
                  grids=[z[p + "grid%d" % g] for g in range(int(z[p + "n_grids"]))],
                  refsum=z[p + "refsum"] if (p + "refsum") in z else None)
        if (p + "log_normal") in z:
            rt["log_normal"] = bool(z[p + "log_normal"])
        if (p + "prior") in z:
            rt["prior"] = z[p + "prior"]
        t["restraints"].append(rt)
    return t


def stack_lambdas(tabs_list):
    """Tables of several lambda ensembles that share restraint tables -> energy [K][S], logZ [K]."""
    t = dict(tabs_list[0])
    t["energy"] = np.array([x["energy"] for x in tabs_list])
    t["logZ"] = np.array([x["logZ"] for x in tabs_list])
    t["lam"] = [x["lam"] for x in tabs_list]
    return t


def cineromycin_tables():
    z = load_npz("cineromycin_tables.npz")
    return z, tables_from_npz(z, "lam0_"), tables_from_npz(z, "lam1_")


def assert_outputs_equal(a, b, keys=None):
    """Bit-exact comparison of two sampler output dicts (engine.alloc_outputs layout)."""
    for k in (keys or a.keys()):
        assert a[k].shape == b[k].shape, k
        if a[k].dtype.kind == "f":
            assert np.array_equal(a[k].view(np.int64), b[k].view(np.int64)), "%s differs (bitwise)" % k
        else:
            assert np.array_equal(a[k], b[k]), "%s differs" % k


def pf_raw_tables(g, dense):
    """On-the-fly PF tables (raw <N_c>/<N_h>) matching the dense reference tables of pf_grid.npz."""
    rt = dense["restraints"][0]
    raw = dict(kind="pfgrid", name="pf", ref=rt["ref"], ndof=rt["ndof"], sigma=rt["sigma"], nlsig=rt["nlsig"],
               two_sig2=rt["two_sig2"], cnorm=rt["cnorm"], grids=rt["grids"], prior=rt.get("prior"),
               pf=dict(Nc=g["Nc"], Nh=g["Nh"], exp=g["exp"], weight=g["weight"]), refsum=None, sse=None)
    return dict(n_states=dense["n_states"], energy=dense["energy"], logZ=dense["logZ"], restraints=[raw])
