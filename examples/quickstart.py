"""End-to-end example with the reference's own workflow (cf. docs/examples/Tutorials/Prep_Rest_Post_Ana of
vvoelz/biceps), on a small synthetic system: prepare the input files, build the ensembles, sample the posterior for
two lambda values, compute BICePs scores and populations, check convergence.

    python examples/quickstart.py [workdir]

The only change against a script written for the reference is the import line."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import biceps_b200 as biceps                                   # instead of: import biceps

work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="biceps_b200_")
data_dir, out_dir = os.path.join(work, "J_NOE"), os.path.join(work, "results")
biceps.toolbox.mkdir(data_dir)
biceps.toolbox.mkdir(out_dir)

# ---- a synthetic system: 100 conformational states, 10 J couplings, 33 NOE distances --------------------
rng = np.random.default_rng(0)
S, nJ, nNOE = 100, 10, 33
energies = 3.0 * np.abs(rng.standard_normal(S))
energies -= energies.min()
truth = int(np.argmin(energies + rng.normal(0, 0.5, S)))         # the state the "experiment" was measured on
J_model = rng.uniform(1.0, 10.0, (S, nJ))
noe_model = rng.uniform(2.0, 6.0, (S, nNOE))
J_exp = np.column_stack([np.arange(nJ), J_model[truth] + rng.normal(0, 0.5, nJ)])
noe_exp = np.column_stack([np.arange(nNOE), noe_model[truth] * 1.1 + rng.normal(0, 0.2, nNOE)])

# ---- Preparation: one <state>.J and <state>.noe file per state (biceps.Preparation) ------------------------
prep = biceps.Preparation(nstates=S, outdir=data_dir)
prep.prepare_J(J_exp, list(J_model), rng.integers(0, 200, (nJ, 4)))
prep.prepare_noe(noe_exp, list(noe_model), rng.integers(0, 200, (nNOE, 2)))
input_data = prep.to_sorted_list()

# ---- Ensemble + PosteriorSampler for lambda = 0 and 1 -------------------------------------------------------
options = [dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
           dict(ref="exponential", sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02))]
nsteps = 200000
for lam in (0.0, 1.0):
    ensemble = biceps.Ensemble(lam, energies)
    ensemble.initialize_restraints(input_data, options)
    sampler = biceps.PosteriorSampler(ensemble, nchains=64, seed=1)       # nchains / seed: biceps_b200 additions
    sampler.sample(nsteps, burn=1000)
    sampler.traj.process_results(os.path.join(out_dir, "traj_lambda%2.2f.npz" % lam))
    print("lambda %.1f: accepted %.1f %% of %d steps x %d chains" % (
        lam, 100.0 * sampler.accepted / sampler.total, nsteps + 1000, sampler.nchains))

# ---- Analysis: MBAR -> BICePs score and populations ---------------------------------------------------------
A = biceps.Analysis(out_dir, nstates=S, precheck=False)
print("BICePs score f(lambda=1) - f(lambda=0) = %.3f +/- %.3f" % (A.f_df[1, 0], A.f_df[1, 1]))
pops = A.P_dP[:, 1]
print("most populated states at lambda = 1:", np.argsort(pops)[::-1][:5], "(experiment generated from state %d)" % truth)
print(A.get_max_likelihood_parameters(model=1))

# ---- Convergence of the lambda = 1 trajectory ---------------------------------------------------------------
C = biceps.Convergence(filename=os.path.join(out_dir, "traj_lambda1.00.npz"), outdir=out_dir)
C.get_autocorrelation_curves(method="auto", maxtau=200)
print("autocorrelation times (snapshots):", dict(zip(C.rest_type, np.round(C.tau_c, 2))))
C.process(nfold=10, nround=50, savefile=True)
print("JSD of the two halves of the full trajectory:", dict(zip(C.rest_type, np.round(C.all_JSD[:, -1], 4))))
print("files in", out_dir, ":", sorted(os.listdir(out_dir)))
