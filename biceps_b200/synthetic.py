"""Deterministic synthetic ensembles of the shapes BASELINE.json names.

Pure input synthesis (raw observables, no energies of the sampler are computed here):
forward-model values model[S][J], experimental values exp[J], equivalency-group weights
and prior energies.  Both the CUDA path and the oracle consume these same arrays.
"""
import numpy as np


def _weights_from_groups(restraint_index):
    # 1/n per equivalency group (reference: Restraint.py:418-442)
    _, inv, counts = np.unique(restraint_index, return_inverse=True, return_counts=True)
    return 1.0 / counts[inv].astype(np.float64)


def _restraint(rng, S, J, flavour):
    if flavour == "noe":
        exp = rng.uniform(2.0, 6.0, J)
        model = np.clip(exp[None, :] * (1.0 + 0.15 * rng.standard_normal((S, J))), 0.5, None)
        ridx = np.arange(J)
        third = np.arange(2, J, 3)
        ridx[third] = ridx[third - 1]          # every 3rd observable shares an index with its predecessor
        return dict(kind="gamma", name="noe", model=model, exp=exp, weight=_weights_from_groups(ridx),
                    sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02), ref="exponential", log_normal=False)
    scale = {"J": 1.5, "cs_H": 0.3, "cs_Ca": 1.0, "cs_N": 1.0}[flavour]
    if flavour == "J":
        exp = rng.uniform(1.0, 10.0, J)
    elif flavour == "cs_H":
        exp = rng.normal(4.0, 1.0, J)
    elif flavour == "cs_Ca":
        exp = rng.normal(55.0, 5.0, J)
    else:
        exp = rng.normal(120.0, 5.0, J)
    model = exp[None, :] + scale * rng.standard_normal((S, J))
    return dict(kind="scalar", name=flavour, model=model, exp=exp, weight=np.ones(J),
                sigma=(0.05, 20.0, 1.02), ref="uniform")


def make_ensemble(config="C3", seed=0, n_states=None, n_lambda=None):
    """Raw observables of a named configuration.

    C3: S=10k, R=4 in the reference's file-extension order: cs_H 200, cs_Ca 100, J 300, NOE 400
        = 1000 observables (BASELINE configs[2])
    C2: S=1k,  R=5 in the reference's fixed order cs_H, cs_Ca, cs_N, J, noe      (BASELINE configs[1])
    C1s: a cineromycin-B-shaped synthetic stand-in, S=100, J 10 + NOE 33          (BASELINE configs[0])
    """
    rng = np.random.default_rng(seed)
    if config == "C3":
        S, K = n_states or 10000, n_lambda or 1
        layout = [("cs_H", 200), ("cs_Ca", 100), ("J", 300), ("noe", 400)]
    elif config == "C2":
        S, K = n_states or 1000, n_lambda or 8
        layout = [("cs_H", 60), ("cs_Ca", 60), ("cs_N", 60), ("J", 40), ("noe", 120)]
    elif config == "C1s":
        S, K = n_states or 100, n_lambda or 2
        layout = [("J", 10), ("noe", 33)]
    else:
        raise ValueError("unknown synthetic configuration %r" % (config,))
    energies = 3.0 * np.abs(rng.standard_normal(S))
    energies -= energies.min()
    restraints = [_restraint(rng, S, J, fl) for fl, J in layout]
    lambdas = np.linspace(0.0, 1.0, K) if K > 1 else np.array([1.0])
    return dict(config=config, n_states=S, energies=energies, lambdas=lambdas, restraints=restraints)
