"""Generates tests/golden/convergence.npz by running the UNMODIFIED reference's biceps.Convergence
(/root/reference/biceps/convergence.py) in the build container on the trajectory the reference ships with
its tutorial (docs/examples/Tutorials/Prep_Rest_Post_Ana/results/traj_lambda1.00.npz: 1000 snapshots,
parameters sigma_J, sigma_noe, gamma_noe).

    python tests/golden/make_golden_convergence.py

Stored: the columnar inputs (traces[P][T], indices[P][T], grid lengths) and what the reference computed:
  auto_*     get_autocorrelation_curves(method="auto", maxtau=100)            -> autocorr, tau_c
  block_*    get_autocorrelation_curves(method="block-avg-auto", nblocks=5, maxtau=50)
  exp_tau    get_autocorrelation_curves(method="exp", maxtau=100)             -> tau_c (scipy fit)
  all_JSD, all_JSDs   process(nfold=10, nround=20) after "auto", with np.random.seed(7) set right before
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as H  # noqa: E402
from oracle import convergence_oracle as CO  # noqa: E402

TRAJ = "docs/examples/Tutorials/Prep_Rest_Post_Ana/results/traj_lambda1.00.npz"


def main():
    biceps = H.import_reference()
    fn = os.path.join(H.REFERENCE_ROOT, TRAJ)
    out = {}
    with tempfile.TemporaryDirectory() as tmp, H.quiet():
        C = biceps.Convergence(filename=fn, outdir=tmp)
        traces, idx, nb = CO.columnar(C.traj)
        out.update(traces=traces, indices=idx, n_bins=np.array(nb, np.int64),
                   rest_type=np.array(C.rest_type), freq_save_traj=np.int64(C.freq_save_traj))
        for p, a in enumerate(C.allowed_parameters):
            out["allowed%d" % p] = np.asarray(a, dtype=np.float64)
        C.get_autocorrelation_curves(method="block-avg-auto", nblocks=5, maxtau=50)
        out.update(block_autocorr=C.autocorr, block_tau=C.tau_c)
        C.get_autocorrelation_curves(method="exp", maxtau=100)
        out.update(exp_tau=C.tau_c)
        C.get_autocorrelation_curves(method="auto", maxtau=100)
        out.update(auto_autocorr=C.autocorr, auto_tau=C.tau_c)
        np.random.seed(7)
        C.process(nblock=5, nfold=10, nround=20, savefile=True, block_avg=True)
        out.update(all_JSD=np.load(os.path.join(tmp, "all_JSD.npy")), all_JSDs=np.load(os.path.join(tmp, "all_JSDs.npy")))
    np.savez_compressed(os.path.join(HERE, "convergence.npz"), **out)
    print("wrote convergence.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
