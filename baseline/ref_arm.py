"""Reference arm of bench.py (`--impl reference`): the UNMODIFIED reference package (installed from /root/reference
into baseline/_ref by baseline/install_ref.sh) driven through its own public API --
biceps.Ensemble(...).initialize_restraints(...), biceps.PosteriorSampler(ensemble).sample(nsteps) -- one process per
host core, the way the reference parallelises itself (biceps/decorators.py:27-37, one sampler per process).
None of this repository's kernels, engine or oracle is on this path; numpy/pandas only write the synthetic input
files (the reference's `<state>.<ext>` DataFrame pickles, Restraint.py:246-263) the reference then reads itself."""
import contextlib
import io
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

# reference options for the C3 layout (file-extension order cs_Ca, cs_H, J, noe as biceps.toolbox.sort_data returns it)
OPTIONS = {"cs_H": dict(ref="uniform", sigma=(0.05, 20.0, 1.02)), "cs_Ca": dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
           "cs_N": dict(ref="uniform", sigma=(0.05, 20.0, 1.02)), "J": dict(ref="uniform", sigma=(0.05, 20.0, 1.02)),
           "noe": dict(ref="exponential", sigma=(0.05, 5.0, 1.02), gamma=(0.2, 5.0, 1.02))}


def available():
    return os.path.isdir(os.path.join(REF, "biceps"))


def import_reference():
    if not available():
        raise ImportError("baseline/_ref/biceps is missing: run baseline/install_ref.sh where /root/reference exists")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(io.StringIO()):      # the import banner
        import biceps
    assert os.path.realpath(os.path.dirname(biceps.__file__)).startswith(os.path.realpath(REF)), biceps.__file__
    return biceps


def write_inputs(ens, outdir):
    """The synthetic observables as the reference's per-state input files; returns input_data (one sorted
    file list per state) and the per-file options."""
    import pandas as pd
    S = ens["n_states"]
    names = []
    for rt in ens["restraints"]:
        names.append(rt["name"] if rt["name"] in ("J", "noe") else rt["name"])
    order = sorted(range(len(names)), key=lambda k: names[k])            # toolbox.sort_data order: by extension
    input_data = []
    for i in range(S):
        files = []
        for k in order:
            rt, nm = ens["restraints"][k], names[k]
            J = len(rt["exp"])
            w = rt["weight"]
            ridx = np.arange(J)
            if nm == "noe":                                              # equivalency groups <- weights 1/n
                ridx = np.zeros(J, dtype=int)
                g, j = 0, 0
                while j < J:
                    m = int(round(1.0 / w[j]))
                    ridx[j:j + m] = g
                    g += 1
                    j += m
            df = pd.DataFrame({"restraint_index": ridx, "atom_index1": np.arange(J), "atom_index2": np.arange(J) + 1,
                               "atom_index3": np.arange(J) + 2, "atom_index4": np.arange(J) + 3,
                               "exp": rt["exp"], "model": rt["model"][i]})
            f = os.path.join(outdir, "%d.%s" % (i, nm))
            df.to_pickle(f)
            files.append(f)
        input_data.append(files)
    return input_data, [OPTIONS[names[k]] for k in order]


def _worker(rank, sampler, nsteps, reps, warm, barrier, q):
    np.random.seed(1000 + rank)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        for _ in range(warm):
            sampler.sample(max(100, nsteps // 10))
        times = []
        for _ in range(reps):
            barrier.wait()
            t0 = time.perf_counter()
            sampler.sample(nsteps)
            times.append(time.perf_counter() - t0)
    q.put((rank, times))


def build(ens, lam):
    """ONE reference sampler through the reference's own API (construction timed); returns (sampler, seconds)."""
    biceps = import_reference()
    with tempfile.TemporaryDirectory() as d:
        input_data, options = write_inputs(ens, d)
        t0 = time.perf_counter()
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            e = biceps.Ensemble(float(lam), ens["energies"].astype(float))
            e.initialize_restraints(input_data, options)
            np.random.seed(0)
            sampler = biceps.PosteriorSampler(e)
        return sampler, time.perf_counter() - t0


def time_sample(sampler, n_procs, nsteps, reps, warm):
    """Forks n_procs processes that each run sampler.sample(nsteps) `reps` times on their own copy of the sampler
    (own np.random seed); returns the wall time of every repetition (max over the processes; a barrier starts them
    together)."""
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(n_procs)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, sampler, nsteps, reps, warm, barrier, q)) for r in range(n_procs)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    return [max(t[k] for _, t in res) for k in range(reps)]
