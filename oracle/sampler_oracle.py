"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (biceps_b200/).

CPU restatement of the reference's Metropolis sampler, driven purely by the flat tables of
oracle/ref_harness.extract_tables and a pluggable per-step RNG (oracle/philox.py).

Follows, line by line:
  PosteriorSampler.sample        /root/reference/biceps/PosteriorSampler.py:251-374
  PosteriorSampler.neglogP       /root/reference/biceps/PosteriorSampler.py:233-248
  Restraint_*.compute_neglogP    /root/reference/biceps/Restraint.py:356-383, 457-477,
                                 572-592, 862-894
  trajectory bookkeeping         /root/reference/biceps/PosteriorSampler.py:339-358, 393

Pinning (tests/test_oracle_vs_reference.py, tests/golden/): with MTGlobalRNG + np.exp this
restatement reproduces the live reference bit for bit (snapshots incl. E, state_trace,
state_counts, sampled_parameters, accepted, sep_accept).  With PhiloxStepRNG + bcp_exp it is
the bit-exact specification of the CUDA kernel (two-hop chain of SURVEY.md section 7.2).

Floating-point order is kept exactly:
  t_r = ((0 + Ndof*ln(sigma)) + sse/(2*sigma^2)) [+ prior] + Ndof/2*ln(2 pi)) [- refsum]
  E   = ((0 + (energy_i + logZ)) + t_0) + t_1 + ...
"""
import math

import numpy as np

# ---------------------------------------------------------------------------------------
# bcp_exp: the deterministic exp used in the Metropolis test of the Philox/CUDA hop.
# Only +,-,* on IEEE doubles, no fused multiply-add, so C (-ffp-contract=off), Python and
# CUDA (__dmul_rn/__dadd_rn) produce identical bits.  Valid for -40 <= x <= 0.
# ---------------------------------------------------------------------------------------
LOG2E = 1.4426950408889634
LN2_HI = 6.93147180369123816490e-01      # low 21 mantissa bits are zero -> k*LN2_HI exact
LN2_LO = 1.90821492927058770002e-10
RND = 6755399441055744.0                 # 1.5 * 2^52 : add/sub rounds to nearest integer
EXP_COEF = [1.0 / math.factorial(n) for n in range(14)]    # Taylor, degree 13
ACCEPT_CUTOFF = -40.0                    # exp(-40) < 2^-53 <= u2  -> always reject


def bcp_exp(x):
    t = x * LOG2E + RND
    k = t - RND
    r = (x - k * LN2_HI) - k * LN2_LO
    p = EXP_COEF[13]
    for n in range(12, -1, -1):
        p = p * r + EXP_COEF[n]
    return p * (2.0 ** int(k))


def accept_np_exp(E_old, E_new, rng):
    """Reference rule (PosteriorSampler.py:322-326): second uniform only if E_new >= E_old."""
    if E_new < E_old:
        return True
    return bool(rng.u2() < np.exp(E_old - E_new))


def accept_bcp(E_old, E_new, rng):
    """Same rule with the deterministic exp; d < -40 can never pass since u2 >= 2^-53."""
    if E_new < E_old:
        return True
    d = E_old - E_new
    if d < ACCEPT_CUTOFF:
        return False
    return bool(rng.u2() < bcp_exp(d))


# ---------------------------------------------------------------------------------------
def n_params(rt):
    return 1 + len(rt["grids"])


def param_layout(tables):
    """rest_index[p], grid length per parameter, start index int(len/2)
    (Restraint.py:240, 506, 687-703; PosteriorSampler.py:275-281)."""
    rest_index, glen = [], []
    for r, rt in enumerate(tables["restraints"]):
        rest_index.append(r); glen.append(len(rt["sigma"]))
        for g in rt["grids"]:
            rest_index.append(r); glen.append(len(g))
    return rest_index, glen


def allowed_parameters(tables):
    out = []
    for rt in tables["restraints"]:
        out.append(rt["sigma"])
        out.extend(rt["grids"])
    return out


def restraint_term(rt, state, idx):
    """One Restraint_*.compute_neglogP call; idx = this restraint's parameter indices."""
    t = 0
    t += rt["nlsig"][idx[0]]                                  # Ndof*ln(sigma)      :376/:470/:585/:875
    if rt["kind"] == "scalar":
        t += rt["sse"][state] / rt["two_sig2"][idx[0]]        # :377/:471/:881
    elif rt["kind"] == "gamma":
        t += rt["sse"][state][idx[1]] / rt["two_sig2"][idx[0]]  # :586
    else:
        cell = tuple(idx[1:7])
        t += rt["sse"][state][cell] / rt["two_sig2"][idx[0]]  # :877
        if rt.get("prior") is not None:
            t += rt["prior"][cell]                            # :879
    t += rt["cnorm"]                                          # Ndof/2*ln(2 pi)     :378
    if rt["refsum"] is not None:
        ref = rt["refsum"][state]
        if np.ndim(ref) > 0:
            ref = ref[tuple(idx[1:7])]                        # :888/:893
        t -= ref                                              # :380-382
    return t


def neglogP(tables, state, indices):
    """PosteriorSampler.neglogP for one replica; indices = flat parameter index list."""
    result = 0
    result += tables["energy"][state] + tables["logZ"]        # :245
    p = 0
    for rt in tables["restraints"]:
        m = n_params(rt)
        result += restraint_term(rt, state, indices[p:p + m])  # :247
        p += m
    return result


class ChainResult(object):
    pass


def run_chain(tables, rng, nsteps, burn=0, traj_every=100, accept_fn=accept_np_exp,
              chain=None, step0=0, record_state_trace=True, reset_indices=False):
    """One sample() call.  `chain` (dict: state, indices, E, accepted, total) lets a second
    call continue where the first stopped; `step0` is the Philox step counter offset.

    reset_indices=True reproduces what the reference really does when sample() is called twice
    on one sampler: state, E, accepted and total persist (PosteriorSampler.py:294, 336-337) but
    self.indices is rebuilt from the restraints' never-updated `*_index` attributes, i.e. the
    parameter indices jump back to int(len/2) while E stays that of the old configuration
    (PosteriorSampler.py:271-278)."""
    S = tables["n_states"]
    R = len(tables["restraints"])
    rest_index, glen = param_layout(tables)
    P = len(rest_index)
    allowed = allowed_parameters(tables)
    if chain is None:
        chain = dict(state=int(rng.initial_state(S)), indices=[int(g / 2) for g in glen],
                     E=1.0e99, accepted=0, total=0)
    cur_state, cur_idx, cur_E = chain["state"], list(chain["indices"]), chain["E"]
    if reset_indices:
        cur_idx = [int(g / 2) for g in glen]
    accepted, total = chain["accepted"], chain["total"]

    res = ChainResult()
    res.state_counts = np.ones(S)                              # pseudo-count  :393
    res.sampled_parameters = [np.zeros(g) for g in glen]
    res.sep_accepted = np.zeros(P + 1)
    res.state_trace = []
    res.traj_step, res.traj_E, res.traj_accept, res.traj_state, res.traj_indices = [], [], [], [], []
    res.traces = []
    values = None

    for step in range(nsteps + burn):
        rng.begin_step(step0 + step)
        state, idx = cur_state, list(cur_idx)
        if rng.is_nuisance_move(R):                            # :298-300
            pick = rng.pick_restraint(R)                       # :303
            moved = []
            for k in range(P):
                if rest_index[k] == pick:
                    idx[k] = (idx[k] + rng.delta()) % glen[k]  # :304
                    moved.append(k)
        else:
            state = int(rng.pick_state(S))                     # :307
            moved = [P]
        E = neglogP(tables, state, idx)                        # :319
        acc = accept_fn(cur_E, E, rng)                         # :322-326
        if acc:                                                # :329-336
            cur_E, cur_state, cur_idx = E, state, idx
            values = [allowed[k][idx[k]] for k in range(P)]
            for k in moved:
                res.sep_accepted[k] += 1.0
            accepted += 1
        total += 1                                             # :337
        if step >= burn:                                       # :339
            _step = step - burn
            res.state_counts[cur_state] += 1                   # :344
            if record_state_trace:
                res.state_trace.append(cur_state)              # :345
            for k in range(P):
                res.sampled_parameters[k][cur_idx[k]] += 1     # :348
            if _step % traj_every == 0:                        # :354
                res.traj_step.append(_step); res.traj_E.append(float(cur_E))
                res.traj_accept.append(int(acc)); res.traj_state.append(cur_state)
                res.traj_indices.append(list(cur_idx))
                res.traces.append(None if values is None else list(values))
    res.accepted, res.total = accepted, total
    res.chain = dict(state=cur_state, indices=cur_idx, E=cur_E, accepted=accepted, total=total)
    res.sep_accept = [res.sep_accepted / total * 100., accepted / total * 100.]   # :373-374
    return res
