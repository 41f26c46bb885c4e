"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (biceps_b200/).

numpy restatement of the analysis half of the hot path:

  u_kln construction          /root/reference/biceps/Analysis.py:145-185
  MBAR(u_kln, N_k)            third-party pymbar (>= 4.0.1, /root/reference/requirements.txt:27),
                              called at Analysis.py:192; NOT present in the reference tree nor
                              installable here, so the published MBAR equations are restated
                              (Shirts & Chodera, J. Chem. Phys. 129, 124105 (2008), eqs. 11, C1-C3)
  Delta_f / dDelta_f          Analysis.py:201-210, uncertainty_method='approximate' (Theta = W^T W)
  populations P_i(l), dP      Analysis.py:217-223, compute_expectations(indicator)

Pinning: tests/test_mbar_oracle.py reproduces the reference's own stored BS.dat and
populations.dat (docs/examples/Tutorials/Prep_Rest_Post_Ana/results, K = 2) from its stored
trajectories (tests/golden/tutorial_mbar.npz).  Beyond that single K = 2 case the MBAR parity
is UNPINNED (no other valid golden exists); K > 2 is checked only through
solver-independent properties (fixed point, gauge, agreement with exact reweighting).
"""
import numpy as np

from . import sampler_oracle as O


def logsumexp(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(a - m), axis=axis))


def u_kln_from_trajectories(tables_per_lambda, trajs):
    """Analysis.py:145-185.  trajs[k] = dict(E[n], state[n], indices[n][P]) of run k.
    u_kln[k,k,n] = stored E (:170); u_kln[k,l,n] = sampler[l].neglogP(snapshot) (:182)."""
    K = len(trajs)
    N_k = np.array([len(t["E"]) for t in trajs])
    u_kln = np.zeros((K, K, N_k.max()))
    states_kn = np.zeros((K, N_k.max()), dtype=np.int64)
    for k in range(K):
        for n in range(N_k[k]):
            st = int(trajs[k]["state"][n])
            idx = [int(x) for x in trajs[k]["indices"][n]]
            states_kn[k, n] = st
            for l in range(K):
                u_kln[k, l, n] = trajs[k]["E"][n] if k == l else O.neglogP(tables_per_lambda[l], st, idx)
    return u_kln, N_k, states_kn


def kln_to_kn(u_kln, N_k):
    """pymbar.utils.kln_to_kn: concatenate the first N_k[k] samples of each run along n."""
    return np.concatenate([u_kln[k, :, :N_k[k]] for k in range(len(N_k))], axis=1)


def mbar_solve(u_kn, N_k, tol=1e-13, max_iter=100000):
    """f_k with f_0 = 0 from the MBAR self-consistent equations; Newton-accelerated."""
    u_kn = np.asarray(u_kn, np.float64)
    N_k = np.asarray(N_k, np.float64)
    K = u_kn.shape[0]
    logN = np.log(N_k)
    f = np.zeros(K)

    def weights(f):
        logden = logsumexp((f + logN)[:, None] - u_kn, axis=0)
        return np.exp(f[:, None] - u_kn - logden[None, :])       # W[k][n]

    for it in range(max_iter):
        W = weights(f)
        s = W.sum(axis=1)
        g = N_k * (s - 1.0)
        fn = None
        if it >= 2 and K > 1:
            H = np.diag(N_k * s) - (N_k[:, None] * N_k[None, :]) * (W @ W.T)
            try:
                d = np.linalg.solve(H[1:, 1:], -g[1:])
                cand = f.copy(); cand[1:] += d
                s2 = weights(cand).sum(axis=1)
                if np.linalg.norm(N_k * (s2 - 1.0)) < np.linalg.norm(g):
                    fn = cand
            except np.linalg.LinAlgError:
                pass
        if fn is None:
            fn = f - np.log(s)
            fn -= fn[0]
        done = np.max(np.abs(fn - f)) <= tol * max(1.0, np.max(np.abs(fn)))
        f = fn
        if done:
            break
    return f, weights(f).T, it + 1                               # W as [N][K]


def free_energy_differences(f, W):
    """Delta_f[i,j] = f_j - f_i; dDelta_f from Theta = W^T W ('approximate')."""
    theta = W.T @ W
    d = np.diag(theta)
    d2 = d[:, None] + d[None, :] - 2.0 * theta
    return f[None, :] - f[:, None], np.sqrt(np.abs(d2)), theta


def populations(W, theta, states_n, n_states):
    """P_i(l) = sum_n W_nl 1[s_n = i]; dP with W_nA = W_nl A_n / P (no shift of A);
    never-sampled states give P = 0, dP = nan."""
    N, K = W.shape
    P = np.zeros((n_states, K)); dP = np.full((n_states, K), np.nan)
    for i in range(n_states):
        A = (states_n == i).astype(np.float64)
        for l in range(K):
            mu = W[:, l] @ A
            P[i, l] = mu
            if mu == 0:
                continue
            WA = W[:, l] * A / mu
            dP[i, l] = abs(mu) * np.sqrt(abs(WA @ WA + theta[l, l] - 2.0 * (WA @ W[:, l])))
    return P, dP


def analysis(tables_per_lambda, trajs, n_states):
    """The whole of Analysis.MBAR_analysis: returns f_df [K][2], P_dP [S][2K], u_kln, N_k."""
    u_kln, N_k, states_kn = u_kln_from_trajectories(tables_per_lambda, trajs)
    u_kn = kln_to_kn(u_kln, N_k)
    states_n = np.concatenate([states_kn[k, :N_k[k]] for k in range(len(N_k))])
    f, W, iters = mbar_solve(u_kn, N_k)
    Df, dDf, theta = free_energy_differences(f, W)
    K = len(N_k)
    f_df = np.zeros((K, 2)); f_df[:, 0] = Df[0, :]; f_df[:, 1] = dDf[0, :]
    P, dP = populations(W, theta, states_n, n_states)
    return dict(f_df=f_df, P_dP=np.concatenate([P, dP], axis=1), u_kln=u_kln, u_kn=u_kn, N_k=N_k,
                states_n=states_n, f=f, theta=theta, iters=iters)
