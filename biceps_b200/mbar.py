"""u_kn rescoring and MBAR on the GPU (K6/K7 of DESIGN.md) -- flat-array layer under Analysis."""
import ctypes as C

import numpy as np

from . import _cabi as A


def ukn(tables, state, indices):
    """u_kn[l][n] = -log P of snapshot n under lambda ensemble l (Analysis.py:165-183 for all l)."""
    state = A.i32(np.atleast_1d(state))
    n = state.shape[0]
    indices = A.i32(np.asarray(indices).reshape(n, tables.n_params))
    out = np.zeros((tables.n_lambda, n), np.float64)
    A.check(A.lib().bcp_ukn(tables._h, n, A.i32ptr(state), A.i32ptr(indices), A.dptr(out)), "bcp_ukn")
    return out


def mbar(u_kn, N_k, state_n=None, n_states=0, device=0, tol=0.0, max_iter=0):
    """MBAR free energies, Theta = W^T W and state populations (pymbar calls of Analysis.py:192-223).
    Returns dict(f, theta, Delta_f, dDelta_f, P, dP, iters)."""
    u_kn = A.f64(u_kn)
    K, N = u_kn.shape
    N_k = np.ascontiguousarray(N_k, dtype=np.int64)
    if N_k.shape != (K,):
        raise ValueError("N_k must have one entry per row of u_kn")
    st = None if state_n is None else A.i32(state_n)
    if st is not None and st.shape != (N,):
        raise ValueError("state_n must have one entry per sample")
    S = int(n_states) if st is not None else 0
    f = np.zeros(K); theta = np.zeros((K, K))
    P = np.zeros((S, K)) if S else None
    dP = np.zeros((S, K)) if S else None
    it = C.c_int32(0)
    A.check(A.lib().bcp_mbar(A.dptr(u_kn), A.i64ptr(N_k), K, N, A.i32ptr(st), S, int(device), float(tol), int(max_iter),
                             A.dptr(f), A.dptr(theta), A.dptr(P), A.dptr(dP), C.byref(it)), "bcp_mbar")
    d = np.diag(theta)
    return dict(f=f, theta=theta, Delta_f=f[None, :] - f[:, None],
                dDelta_f=np.sqrt(np.abs(d[:, None] + d[None, :] - 2.0 * theta)), P=P, dP=dP, iters=int(it.value))
