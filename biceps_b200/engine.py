"""Host-side engine over the C ABI: DeviceTables (an ensemble resident in HBM) and ChainBlock
(a block of independent Metropolis chains).  This is the layer PosteriorSampler / Analysis
sit on; it deals only in flat numpy arrays."""
import ctypes as C

import numpy as np

from . import _cabi as A


def n_snapshots(n_steps, traj_every):
    return (n_steps + traj_every - 1) // traj_every if traj_every > 0 else 0


def alloc_outputs(n_chains, n_groups, S, P, HB, n_steps, traj_every, record_state_trace, pool=None, reuse=None,
                  trace_chains=0):
    """numpy buffers + the bcp_sample_out struct pointing at them.  With a PinnedPool the big
    trajectory buffers are page-locked (the returned dict keeps the pool alive under "_pool");
    `reuse` = a dict returned earlier for the same sizes: its buffers are used again."""
    ns = n_snapshots(n_steps, traj_every)
    big = pool.zeros if pool is not None else np.zeros
    o = reuse if reuse is not None else dict(state_counts=np.zeros((n_groups, S), np.int64), param_counts=np.zeros((n_groups, HB), np.int64),
             accepted=np.zeros(n_chains, np.int64), total=np.zeros(n_chains, np.int64),
             sep_accepted=np.zeros((n_chains, P + 1), np.int64),
             state=np.zeros(n_chains, np.int32), indices=np.zeros((n_chains, P), np.int32),
             energy=np.zeros(n_chains, np.float64),
             traj_energy=big((ns, n_chains), np.float64), traj_state=big((ns, n_chains), np.int32),
             traj_accept=big((ns, n_chains), np.uint8), traj_indices=big((ns, P, n_chains), np.int32),
             state_trace=big((n_steps if record_state_trace else 0,
                              trace_chains if 0 < trace_chains < n_chains else n_chains), np.int32))
    s = A.bcp_sample_out(A.i64ptr(o["state_counts"]), A.i64ptr(o["param_counts"]), A.i64ptr(o["accepted"]),
                         A.i64ptr(o["total"]), A.i64ptr(o["sep_accepted"]), A.i32ptr(o["state"]),
                         A.i32ptr(o["indices"]), A.dptr(o["energy"]),
                         A.dptr(o["traj_energy"]) if ns else None, A.i32ptr(o["traj_state"]) if ns else None,
                         A.u8ptr(o["traj_accept"]) if ns else None, A.i32ptr(o["traj_indices"]) if ns else None,
                         A.i32ptr(o["state_trace"]) if record_state_trace else None)
    return o, s


def chain_cfg(n_chains, seed, chain_id0=0, chain_lambda=None, chain_group=None, n_groups=0, init_state=None,
              rng_mode=0):
    keep = [None if a is None else A.i32(a) for a in (chain_lambda, chain_group, init_state)]
    for a in keep:
        if a is not None and a.shape != (n_chains,):
            raise ValueError("per-chain arrays must have shape (n_chains,)")
    cfg = A.bcp_chain_cfg(int(n_chains), int(seed) & (2 ** 64 - 1), int(chain_id0), int(n_groups), int(rng_mode),
                          A.i32ptr(keep[0]), A.i32ptr(keep[1]), A.i32ptr(keep[2]))
    return cfg, keep


class DeviceTables(object):
    """bcp_handle: the flattened ensemble (all lambdas) resident on one GPU."""

    def __init__(self, tables, device=0):
        self._lib = A.lib()
        self.packed = tables if isinstance(tables, A.PackedTables) else A.PackedTables(tables)
        h = C.c_void_p()
        A.check(self._lib.bcp_create(C.byref(self.packed.struct), int(device), C.byref(h)), "bcp_create")
        self._h = h
        self.device = int(device)
        self.n_states, self.n_lambda = self.packed.n_states, self.packed.n_lambda
        self.n_params = self._lib.bcp_n_params(h)
        self.param_bins = self._lib.bcp_param_bins(h)
        ri = np.zeros(self.n_params, np.int32); gl = np.zeros(self.n_params, np.int32)
        A.check(self._lib.bcp_param_layout(h, A.i32ptr(ri), A.i32ptr(gl)), "bcp_param_layout")
        self.rest_index, self.grid_len = ri, gl
        self.hist_offsets = np.concatenate([[0], np.cumsum(gl)[:-1]]).astype(np.int64)

    def neglogp(self, state, indices, lam=None):
        state = A.i32(np.atleast_1d(state))
        n = state.shape[0]
        indices = A.i32(np.asarray(indices).reshape(n, self.n_params))
        lam_a = None if lam is None else A.i32(np.broadcast_to(np.asarray(lam), (n,)))
        out = np.zeros(n, np.float64)
        A.check(self._lib.bcp_neglogp(self._h, n, A.i32ptr(lam_a), A.i32ptr(state), A.i32ptr(indices),
                                      A.dptr(out)), "bcp_neglogp")
        return out

    def chains(self, n_chains, seed, **kw):
        return ChainBlock(self, n_chains, seed, **kw)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bcp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ChainBlock(object):
    """bcp_chains: n independent chains on one DeviceTables."""

    def __init__(self, tables, n_chains, seed, chain_id0=0, chain_lambda=None, chain_group=None,
                 n_groups=0, init_state=None, rng_mode=0):
        self._lib = A.lib()
        self.tables = tables
        self.n_chains = int(n_chains)
        self.n_groups = int(n_groups) if n_groups else tables.n_lambda
        cfg, keep = chain_cfg(n_chains, seed, chain_id0, chain_lambda, chain_group, n_groups, init_state, rng_mode)
        self.rng_mode = int(rng_mode)
        c = C.c_void_p()
        A.check(self._lib.bcp_chains_create(tables._h, C.byref(cfg), C.byref(c)), "bcp_chains_create")
        self._c = c

    def sample(self, n_steps, burn=0, traj_every=0, record_state_trace=False, pinned=False, out=None, trace_chains=0):
        """Synchronous bcp_sample with host output buffers; returns a dict of numpy arrays
        (trace_chains=k: only chains 0..k-1 are traced; pinned=True: page-locked trajectory buffers, owned by the returned dict's "_pool";
        out=<dict of an earlier call with the same sizes>: reuse its buffers)."""
        T = self.tables
        pool = A.PinnedPool() if (pinned and out is None) else None
        if out is not None:
            ns = n_snapshots(int(n_steps), int(traj_every))
            if out["traj_energy"].shape != (ns, self.n_chains) or out["state_counts"].shape != (self.n_groups, T.n_states):
                raise ValueError("out= buffers do not match this call's sizes")
        o, s = alloc_outputs(self.n_chains, self.n_groups, T.n_states, T.n_params, T.param_bins, int(n_steps),
                             int(traj_every), bool(record_state_trace), pool, reuse=out, trace_chains=int(trace_chains))
        if pool is not None:
            o["_pool"] = pool
        cfg = A.bcp_sample_cfg(int(n_steps), int(burn), int(traj_every), int(bool(record_state_trace)), int(trace_chains))
        A.check(self._lib.bcp_sample(self._c, C.byref(cfg), C.byref(s)), "bcp_sample")
        return o

    def launch(self, n_steps, burn=0, traj_every=0, record_state_trace=False, stream=None, trace_chains=0):
        cfg = A.bcp_sample_cfg(int(n_steps), int(burn), int(traj_every), int(bool(record_state_trace)), int(trace_chains))
        self._last = (int(n_steps), int(traj_every), bool(record_state_trace), int(trace_chains))
        A.check(self._lib.bcp_sample_launch(self._c, C.byref(cfg), C.c_void_p(stream or 0)), "bcp_sample_launch")

    def sync(self):
        A.check(self._lib.bcp_chains_sync(self._c), "bcp_chains_sync")

    def fetch(self):
        T = self.tables
        n_steps, traj_every, rec, tc = getattr(self, "_last", (0, 0, False, 0))
        o, s = alloc_outputs(self.n_chains, self.n_groups, T.n_states, T.n_params, T.param_bins, n_steps,
                             traj_every, rec, trace_chains=tc)
        A.check(self._lib.bcp_chains_fetch(self._c, C.byref(s)), "bcp_chains_fetch")
        return o

    def score(self, populations=True, tol=0.0, max_iter=0):
        """BICePs scores straight from the device-resident snapshots of the last sample()/launch() call
        (bcp_chains_score): u_kn by rescoring under every lambda, then MBAR; no file or host round trip.
        Several chains per lambda (seeds) are concatenated in chain order.  Returns dict(f, theta, Delta_f,
        dDelta_f, P, dP, iters, N_k) like mbar.mbar."""
        T = self.tables
        K, S = T.n_lambda, T.n_states
        f = np.zeros(K); theta = np.zeros((K, K)); N_k = np.zeros(K, np.int64)
        P = np.zeros((S, K)) if populations else None
        dP = np.zeros((S, K)) if populations else None
        it = C.c_int32(0)
        A.check(self._lib.bcp_chains_score(self._c, int(bool(populations)), float(tol), int(max_iter), A.i64ptr(N_k),
                                           A.dptr(f), A.dptr(theta), A.dptr(P), A.dptr(dP), C.byref(it)), "bcp_chains_score")
        d = np.diag(theta)
        return dict(f=f, theta=theta, Delta_f=f[None, :] - f[:, None],
                    dDelta_f=np.sqrt(np.abs(d[:, None] + d[None, :] - 2.0 * theta)), P=P, dP=dP, iters=int(it.value), N_k=N_k)

    def autocorrelation(self, grids, max_tau, normalize=True):
        """Autocorrelation curves / times of every parameter trace of every chain from the device-resident snapshots
        of the last sample() call (bcp_chains_autocorr): -> curves[n][P][max_tau+1], tau[n][P]."""
        T = self.tables
        g = A.f64(np.concatenate([np.asarray(x, dtype=np.float64) for x in grids]))
        if g.shape[0] != T.param_bins:
            raise ValueError("grids must be the %d parameter grids of the ensemble" % T.n_params)
        curves = np.zeros((self.n_chains, T.n_params, int(max_tau) + 1))
        tau = np.zeros((self.n_chains, T.n_params))
        A.check(self._lib.bcp_chains_autocorr(self._c, A.dptr(g), int(max_tau), int(bool(normalize)), A.dptr(curves),
                                              A.dptr(tau)), "bcp_chains_autocorr")
        return curves, tau

    def set_mt_stream(self, raw):
        """Reference-stream mode (rng_mode=1): raw 32-bit MT19937 outputs from the generator's current position."""
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        A.check(self._lib.bcp_chains_set_mt_stream(self._c, raw.ctypes.data_as(C.POINTER(C.c_uint32)), raw.shape[0]),
                "bcp_chains_set_mt_stream")

    def mt_consumed(self):
        n = C.c_int64(0)
        A.check(self._lib.bcp_chains_mt_consumed(self._c, C.byref(n)), "bcp_chains_mt_consumed")
        return int(n.value)

    def sample_numpy_stream(self, n_steps, burn=0, **kw):
        """sample() driven by np.random's global MT19937 generator exactly like the reference's sampler: the
        generator's next outputs are handed to the device and the global state is advanced by what the chain
        consumed, so any later np.random call continues as it would after the reference's sample()."""
        if int(n_steps) + int(burn) > 50000000:
            raise ValueError("the numpy-stream mode hands the generator's outputs over in one piece (12 per step): "
                             "split runs longer than 5e7 steps into several sample() calls or use the Philox mode")
        st0 = np.random.get_state()
        rs = np.random.RandomState()
        rs.set_state(st0)
        raw = rs._bit_generator.random_raw(12 * (int(n_steps) + int(burn)) + 4096).astype(np.uint32)
        self.set_mt_stream(raw)
        o = self.sample(n_steps, burn=burn, **kw)
        used = self.mt_consumed()
        rs.set_state(st0)
        if used:
            rs._bit_generator.random_raw(used, output=False)       # advance only
        np.random.set_state(rs.get_state())
        return o

    def set_state(self, state=None, indices=None, energy=None):
        """Overwrite chain state / indices [n][P] / energy (None = keep): checkpoint restart."""
        st = None if state is None else A.i32(state)
        ix = None if indices is None else A.i32(np.asarray(indices).reshape(self.n_chains, self.tables.n_params))
        en = None if energy is None else A.f64(energy)
        A.check(self._lib.bcp_chains_set_state(self._c, A.i32ptr(st), A.i32ptr(ix), A.dptr(en)), "bcp_chains_set_state")

    def last_kernel_ms(self):
        ms = C.c_float(); nl = C.c_int32()
        A.check(self._lib.bcp_chains_last_kernel_ms(self._c, C.byref(ms), C.byref(nl)), "bcp_chains_last_kernel_ms")
        return float(ms.value), int(nl.value)

    def last_kernel(self):
        """Name of the Metropolis kernel the last launch ran (the choice is automatic and invisible in the results)."""
        return self._lib.bcp_chains_last_kernel(self._c).decode()

    def device_hist_ptrs(self):
        a = A.c_int64_p(); b = A.c_int64_p()
        A.check(self._lib.bcp_chains_device_hist(self._c, C.byref(a), C.byref(b)), "bcp_chains_device_hist")
        return C.cast(a, C.c_void_p).value, C.cast(b, C.c_void_p).value

    def close(self):
        if getattr(self, "_c", None):
            self._lib.bcp_chains_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
