"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Philox4x32-10 counter-based RNG (Salmon et al., SC'11, "Parallel random numbers: as easy
as 1, 2, 3") and the fixed per-step draw budget the CUDA sampler uses.  The reference
draws a data-dependent number of MT19937 variates per step from numpy's global stream
(/root/reference/biceps/PosteriorSampler.py:299-307, 325); a counter-based design instead
reserves ONE Philox block per (chain, step) and derives every draw of the step from it
(last paragraph).  This file is the executable specification of
that mapping; biceps_b200/csrc/philox.cuh must agree with it bit for bit.

Block layout for counter = (step_lo, step_hi, chain_lo, chain_hi), key = (seed_lo, seed_hi):
  w0 : move type.  nuisance move  iff  w0 < floor(R * 2^32 / (R+1))      [dice < 1-1/(R+1)]
  w1 : state move    -> proposed state  = (w1 * S) >> 32
       nuisance move -> p = w1 * R ; restraint = p >> 32 ; f = p & 0xffffffff ;
                        for each parameter of that restraint in order:
                            q = f * 3 ; delta = (q >> 32) - 1 ; f = q & 0xffffffff
  w2,w3 : Metropolis uniform  u2 = (2*k + 1) * 2^-53,  k = (w2 << 20) | (w3 >> 12)  (52 bits;
       u2 is never 0 or 1, every value is an exact double).
The initial state of a chain (reference: randint at PosteriorSampler.py:29) comes from the
block at step = 2^64-1:  state0 = (w1 * S) >> 32.
"""
import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF
INIT_STEP = (1 << 64) - 1


def philox4x32(ctr, key, rounds=10):
    """Scalar Philox4x32; ctr = 4 uint32, key = 2 uint32 -> 4 uint32 (python ints)."""
    c0, c1, c2, c3 = [int(x) & MASK for x in ctr]
    k0, k1 = [int(x) & MASK for x in key]
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_np(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32 over uint64-held 32-bit lanes (numpy arrays or scalars)."""
    c0 = np.asarray(c0, dtype=np.uint64); c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64); c3 = np.asarray(c3, dtype=np.uint64)
    k0 = np.uint64(k0); k1 = np.uint64(k1)
    m = np.uint64(MASK); s32 = np.uint64(32)
    for _ in range(rounds):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = ((p1 >> s32) ^ c1 ^ k0) & m
        n2 = ((p0 >> s32) ^ c3 ^ k1) & m
        c1 = p1 & m
        c3 = p0 & m
        c0, c2 = n0, n2
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def step_block(seed, chain_id, step):
    return philox4x32((step & MASK, (step >> 32) & MASK, chain_id & MASK, (chain_id >> 32) & MASK),
                      (seed & MASK, (seed >> 32) & MASK))


def nuisance_threshold(n_rest):
    return (n_rest << 32) // (n_rest + 1)


def u2_from_words(w2, w3):
    k = (int(w2) << 20) | (int(w3) >> 12)
    return float(2 * k + 1) * 2.0 ** -53


class PhiloxStepRNG(object):
    """Draw protocol adapter used by oracle.sampler_oracle (one block per step)."""

    def __init__(self, seed, chain_id):
        self.seed = int(seed)
        self.chain_id = int(chain_id)
        self.w = None
        self.f = 0

    def initial_state(self, n_states):
        w = step_block(self.seed, self.chain_id, INIT_STEP)
        return (w[1] * n_states) >> 32

    def begin_step(self, step):
        self.w = step_block(self.seed, self.chain_id, int(step))

    def is_nuisance_move(self, n_rest):
        return self.w[0] < nuisance_threshold(n_rest)

    def pick_restraint(self, n_rest):
        p = self.w[1] * n_rest
        self.f = p & MASK
        return p >> 32

    def delta(self):
        q = self.f * 3
        self.f = q & MASK
        return (q >> 32) - 1

    def pick_state(self, n_states):
        return (self.w[1] * n_states) >> 32

    def u2(self):
        return u2_from_words(self.w[2], self.w[3])


class MTGlobalRNG(object):
    """Draw protocol adapter that replays the reference's own calls on numpy's legacy
    global MT19937 stream, in the reference's order (PosteriorSampler.py:29, 299-307, 325)."""

    def __init__(self, seed=None):
        if seed is not None:
            np.random.seed(seed)

    def initial_state(self, n_states):
        return int(np.random.randint(low=0, high=n_states, size=1)[0])

    def begin_step(self, step):
        self._dice = np.random.random()

    def is_nuisance_move(self, n_rest):
        return self._dice < 1. - 1. / (n_rest + 1.)

    def pick_restraint(self, n_rest):
        return int(np.random.randint(n_rest))

    def delta(self):
        return int(np.random.randint(3)) - 1

    def pick_state(self, n_states):
        return int(np.random.randint(low=0, high=n_states, size=1)[0])

    def u2(self):
        return np.random.random()
