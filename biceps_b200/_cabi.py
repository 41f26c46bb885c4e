"""ctypes mirror of include/biceps_b200.h and the loader of the CUDA library.

There is no CPU fallback: if lib/libbiceps_b200.so is missing or fails to load, every
entry point of the package raises.  Build it with `python __graft_entry__.py build`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BCP_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("BCP_LIB") or os.path.join(_HERE, "lib", "libbiceps_b200.so")

BCP_KIND_SCALAR, BCP_KIND_GAMMA, BCP_KIND_PFGRID = 0, 1, 2
BCP_MAX_RESTRAINTS = 8
BCP_PF_NGRID = 6

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class bcp_restraint(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_sigma", C.c_int32), ("n_gamma", C.c_int32),
                ("has_ref", C.c_int32), ("cnorm", C.c_double),
                ("nlsig", c_double_p), ("two_sig2", c_double_p), ("sse", c_double_p),
                ("refsum", c_double_p),
                ("pf_n", C.c_int32 * BCP_PF_NGRID), ("pf_n_obs", C.c_int32),
                ("pf_ref_kind", C.c_int32),
                ("pf_grid", c_double_p), ("pf_Nc", c_double_p), ("pf_Nh", c_double_p),
                ("pf_exp", c_double_p), ("pf_weight", c_double_p), ("pf_prior", c_double_p)]


class bcp_tables(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_lambda", C.c_int32), ("n_restraints", C.c_int32),
                ("reserved", C.c_int32), ("energy", c_double_p), ("logZ", c_double_p),
                ("restraints", C.POINTER(bcp_restraint))]


class bcp_chain_cfg(C.Structure):
    _fields_ = [("n_chains", C.c_int64), ("seed", C.c_uint64), ("chain_id0", C.c_uint64),
                ("n_groups", C.c_int32), ("rng_mode", C.c_int32),
                ("chain_lambda", c_int32_p), ("chain_group", c_int32_p), ("init_state", c_int32_p)]


class bcp_sample_cfg(C.Structure):
    _fields_ = [("n_steps", C.c_int64), ("burn", C.c_int64), ("traj_every", C.c_int64),
                ("record_state_trace", C.c_int32), ("trace_chains", C.c_int32)]


class bcp_sample_out(C.Structure):
    _fields_ = [("state_counts", c_int64_p), ("param_counts", c_int64_p),
                ("accepted", c_int64_p), ("total", c_int64_p), ("sep_accepted", c_int64_p),
                ("state", c_int32_p), ("indices", c_int32_p), ("energy", c_double_p),
                ("traj_energy", c_double_p), ("traj_state", c_int32_p),
                ("traj_accept", c_uint8_p), ("traj_indices", c_int32_p),
                ("state_trace", c_int32_p)]


class bcp_obs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_obs", C.c_int32), ("ref_kind", C.c_int32),
                ("log_normal", C.c_int32), ("use_global_ref_sigma", C.c_int32), ("n_gamma", C.c_int32),
                ("model", c_double_p), ("exp", c_double_p), ("weight", c_double_p), ("gamma", c_double_p)]


class bcp_jsd_task(C.Structure):
    _fields_ = [("offset", C.c_int64), ("n_total", C.c_int64), ("n_first", C.c_int64), ("sel_offset", C.c_int64),
                ("stream", C.c_uint64), ("n_bins", C.c_int32), ("mode", C.c_int32)]


def _ptr(a, ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


def dptr(a):
    return _ptr(a, C.c_double)


def i32ptr(a):
    return _ptr(a, C.c_int32)


def i64ptr(a):
    return _ptr(a, C.c_int64)


def u8ptr(a):
    return _ptr(a, C.c_uint8)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class PackedTables(object):
    """Owns contiguous numpy copies of a `tables` dict (see flatten.py) and exposes the
    bcp_tables struct that points into them.  `tables["energy"]` may be [S] or [K][S]."""

    def __init__(self, tables, pinned=False):
        self._keep = []
        self._pool = PinnedPool() if pinned else None

        def conv(a):
            """contiguous fp64 copy; with pinned=True the big tables live in page-locked memory, so that
            bcp_create uploads them at the full PCIe rate"""
            a = f64(a)
            if self._pool is None or a.nbytes < 65536:
                return a
            b = self._pool.zeros(a.shape, np.float64)
            b[...] = a
            return b
        energy = np.atleast_2d(conv(tables["energy"]))
        logZ = np.atleast_1d(conv(tables["logZ"]))
        K, S = energy.shape
        if logZ.shape[0] != K:
            raise ValueError("logZ must have one entry per lambda (got %d for %d)" % (logZ.shape[0], K))
        rs = tables["restraints"]
        if not (1 <= len(rs) <= BCP_MAX_RESTRAINTS):
            raise ValueError("number of restraints must be in 1..%d" % BCP_MAX_RESTRAINTS)
        arr = (bcp_restraint * len(rs))()
        for q, rt in enumerate(rs):
            r = arr[q]
            kind = {"scalar": BCP_KIND_SCALAR, "gamma": BCP_KIND_GAMMA, "pfgrid": BCP_KIND_PFGRID}[rt["kind"]]
            r.kind = kind
            nlsig, two = conv(rt["nlsig"]), conv(rt["two_sig2"])
            r.n_sigma = nlsig.shape[0]
            r.cnorm = float(rt["cnorm"])
            r.nlsig, r.two_sig2 = dptr(nlsig), dptr(two)
            self._keep += [nlsig, two]
            r.n_gamma = 1
            if kind != BCP_KIND_PFGRID:
                sse = conv(rt["sse"])
                if kind == BCP_KIND_GAMMA:
                    if sse.ndim != 2 or sse.shape[0] != S:
                        raise ValueError("restraint %d: sse must be [S][G]" % q)
                    r.n_gamma = sse.shape[1]
                elif sse.shape != (S,):
                    raise ValueError("restraint %d: sse must be [S]" % q)
                r.sse = dptr(sse)
                self._keep.append(sse)
                if rt.get("refsum") is not None:
                    ref = conv(rt["refsum"])
                    if ref.shape != (S,):
                        raise ValueError("restraint %d: refsum must be [S]" % q)
                    r.has_ref, r.refsum = 1, dptr(ref)
                    self._keep.append(ref)
            else:
                pf = rt["pf"]
                grids = [conv(g) for g in rt["grids"]]
                for k in range(BCP_PF_NGRID):
                    r.pf_n[k] = grids[k].shape[0]
                gcat = conv(np.concatenate(grids))
                Nc, Nh = conv(pf["Nc"]), conv(pf["Nh"])
                ex, w = conv(pf["exp"]), conv(pf["weight"])
                r.pf_n_obs = ex.shape[0]
                r.pf_ref_kind = 1 if rt.get("ref") == "exponential" else 0
                r.pf_grid, r.pf_Nc, r.pf_Nh, r.pf_exp, r.pf_weight = dptr(gcat), dptr(Nc), dptr(Nh), dptr(ex), dptr(w)
                self._keep += [gcat, Nc, Nh, ex, w]
                if rt.get("prior") is not None:
                    pr = conv(rt["prior"])
                    r.pf_prior = dptr(pr)
                    self._keep.append(pr)
        self._keep += [energy, logZ, arr]
        self.struct = bcp_tables(S, K, len(rs), 0, dptr(energy), dptr(logZ), arr)
        self.n_states, self.n_lambda, self.n_restraints = S, K, len(rs)


_SIGS = {
    "bcp_abi_version": (C.c_int, []),
    "bcp_last_error": (C.c_char_p, []),
    "bcp_kernel_launches": (C.c_uint64, []),
    "bcp_host_alloc": (C.c_void_p, [C.c_size_t]),
    "bcp_host_free": (None, [C.c_void_p]),
    "bcp_create": (C.c_int, [C.POINTER(bcp_tables), C.c_int, C.POINTER(C.c_void_p)]),
    "bcp_destroy": (None, [C.c_void_p]),
    "bcp_n_params": (C.c_int, [C.c_void_p]),
    "bcp_param_bins": (C.c_int, [C.c_void_p]),
    "bcp_param_layout": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p]),
    "bcp_chains_create": (C.c_int, [C.c_void_p, C.POINTER(bcp_chain_cfg), C.POINTER(C.c_void_p)]),
    "bcp_chains_destroy": (None, [C.c_void_p]),
    "bcp_chains_set_mt_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.c_int64]),
    "bcp_chains_mt_consumed": (C.c_int, [C.c_void_p, c_int64_p]),
    "bcp_sample": (C.c_int, [C.c_void_p, C.POINTER(bcp_sample_cfg), C.POINTER(bcp_sample_out)]),
    "bcp_sample_launch": (C.c_int, [C.c_void_p, C.POINTER(bcp_sample_cfg), C.c_void_p]),
    "bcp_chains_sync": (C.c_int, [C.c_void_p]),
    "bcp_chains_set_state": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_double_p]),
    "bcp_chains_fetch": (C.c_int, [C.c_void_p, C.POINTER(bcp_sample_out)]),
    "bcp_chains_device_hist": (C.c_int, [C.c_void_p, C.POINTER(c_int64_p), C.POINTER(c_int64_p)]),
    "bcp_chains_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), c_int32_p]),
    "bcp_chains_last_kernel": (C.c_char_p, [C.c_void_p]),
    "bcp_neglogp": (C.c_int, [C.c_void_p, C.c_int64, c_int32_p, c_int32_p, c_int32_p, c_double_p]),
    "bcp_precompute": (C.c_int, [C.POINTER(bcp_obs), C.c_int32, C.c_int, c_double_p, c_double_p, c_double_p,
                                 c_double_p]),
    "bcp_precompute_dev": (C.c_int, [C.POINTER(bcp_obs), C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "bcp_logz": (C.c_int, [c_double_p, C.c_int32, C.c_int32, C.c_int, c_double_p]),
    "bcp_ukn": (C.c_int, [C.c_void_p, C.c_int64, c_int32_p, c_int32_p, c_double_p]),
    "bcp_ukn_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bcp_mbar": (C.c_int, [c_double_p, c_int64_p, C.c_int32, C.c_int64, c_int32_p, C.c_int32, C.c_int, C.c_double,
                           C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "bcp_mbar_dev": (C.c_int, [C.c_void_p, c_int64_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int, C.c_void_p,
                               C.c_double, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "bcp_chains_score": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, c_int64_p, c_double_p, c_double_p,
                                   c_double_p, c_double_p, c_int32_p]),
    "bcp_autocorr": (C.c_int, [c_double_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int, c_double_p, c_double_p]),
    "bcp_autocorr_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "bcp_chains_autocorr": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, C.c_int32, c_double_p, c_double_p]),
    "bcp_jsd": (C.c_int, [c_int32_p, C.c_int64, c_int32_p, C.c_int64, C.POINTER(bcp_jsd_task), C.c_int64, C.c_uint64,
                          C.c_int, c_double_p]),
    "bcp_param_hist": (C.c_int, [c_int32_p, C.c_int64, C.POINTER(bcp_jsd_task), C.c_int64, C.c_int32, C.c_int,
                                 c_int64_p]),
    "bcp_mbar_pass_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, c_int64_p, c_double_p, C.c_int32, C.c_int,
                                    C.c_void_p, c_double_p, c_double_p]),
    "bcp_chains_n_samples": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "bcp_chains_ukn_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bcp_mbar_pops_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, c_int64_p, c_double_p, C.c_void_p, C.c_int32,
                                    C.c_int, C.c_void_p, c_double_p, c_double_p]),
}

_lib = None


def declare(lib, sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    """The loaded CUDA library; raises if it has not been built (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("biceps_b200: %s is missing -- build it with "
                               "`python __graft_entry__.py build`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        declare(L, _SIGS)
        _lib = L
    return _lib


class PinnedPool(object):
    """numpy arrays backed by page-locked host memory (bcp_host_alloc); freed with the pool."""

    def __init__(self):
        self._ptrs = []

    def zeros(self, shape, dtype):
        shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = int(np.prod(shape)) if shape else 1
        nbytes = max(1, n * np.dtype(dtype).itemsize)
        p = lib().bcp_host_alloc(nbytes)
        if not p:
            raise MemoryError("bcp_host_alloc(%d) failed" % nbytes)
        self._ptrs.append(p)
        buf = (C.c_uint8 * nbytes).from_address(p)
        a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        a[...] = 0
        return a

    def close(self):
        for p in self._ptrs:
            lib().bcp_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().bcp_last_error()
        raise RuntimeError("biceps_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
