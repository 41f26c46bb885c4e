"""The handful of biceps.toolbox helpers the sampling path's callers use
(reference: /root/reference/biceps/toolbox.py:16-149, 965-975).  Forward-model generation,
plotting and the legacy v1 helpers of that file are out of scope."""
import glob
import os
import pickle
import re

import numpy as np

# file extension -> restraint class, in the reference's fixed order (toolbox.py:137-149)
_EXTENSIONS = [("H", "Restraint_cs"), ("Ca", "Restraint_cs"), ("N", "Restraint_cs"),
               ("J", "Restraint_J"), ("noe", "Restraint_noe"), ("pf", "Restraint_pf")]


def list_possible_restraints():
    return ["Restraint_cs", "Restraint_J", "Restraint_noe", "Restraint_pf"]


def list_possible_extensions():
    return [e for e, _ in _EXTENSIONS]


def _natural_key(x):
    return [int(s) if s.isdigit() else s for s in re.split("([0-9]+)", x)]


def sort_data(dataFiles):
    """Sort the per-state data files of a directory (or comma-separated directories) by extension
    into input_data[state][restraint] (toolbox.py:16-63)."""
    if not os.path.exists(dataFiles):
        raise ValueError("data directory doesn't exist")
    dirs = dataFiles.split(",") if "," in dataFiles else [dataFiles]
    patterns = [d + "*" if d[-1] == "/" else d + "/*" for d in dirs]
    exts = list_possible_extensions()
    data = [[] for _ in exts]
    for pat in patterns:
        for j in sorted(glob.glob(pat), key=_natural_key):
            if not any(j.endswith(ext) for ext in exts):
                raise ValueError(f"Incompatible File extension. Use:{exts}")
            for k, ext in enumerate(exts):
                if j.endswith(ext):
                    data[k].append(j)
    data = np.array([f for f in data if f])
    return np.stack(data, axis=-1).tolist()


def get_files(path):
    """Globbed files in natsort's default order (toolbox.py:66-80 calls natsorted(glob)): the name is split into
    text and UNSIGNED INTEGER runs -- a decimal point is text -- so `lambda0.5` sorts before `lambda0.25`
    (5 < 25), exactly like the reference; Analysis takes the order of its lambdas from this list."""
    return sorted(glob.glob(path), key=_natural_key)


def list_res(input_data):
    scheme = []
    for i in input_data[0]:
        if not any(i.endswith(ext) for ext in list_possible_extensions()):
            raise ValueError(f"Incompatible File extension. Use:{list_possible_extensions()}")
        scheme.append(i.split(".")[-1])
    return scheme


def list_extensions(input_data):
    return [res.split("_")[-1] for res in list_res(input_data)]


def _file_or_array(x, what, loader):
    """A path is loaded with `loader`; a list / ndarray passes through; anything else is a TypeError
    (same acceptance rule as the reference's check_* helpers, toolbox.py:152-176)."""
    if isinstance(x, str):
        return loader(x)
    if isinstance(x, (list, np.ndarray)):
        return x
    raise TypeError("%s: expected a file path (str) or a list / numpy array, got %s" % (what, type(x).__name__))


def check_indices(indices):
    return np.array(_file_or_array(indices, "indices", np.loadtxt)).astype(int)


def check_exp_data(exp_data):
    return _file_or_array(exp_data, "exp_data", np.loadtxt)


def check_model_data(model_data):
    return _file_or_array(model_data, "model_data", get_files)


def get_state_populations(file):
    """toolbox.py:268-275"""
    return np.loadtxt(file)[:][0]


def print_scores(file):
    """BICePs score out of a BS.dat file (toolbox.py:287-294)"""
    return np.loadtxt(file)[1, 0]


def npz_to_DataFrame(file, out_filename="traj_lambda0.00.pkl", verbose=False):
    """Trajectory npz -> pandas DataFrame pickle (toolbox.py:297-318)."""
    import pandas as pd
    npz = np.load(file, allow_pickle=True)["arr_0"].item()
    if verbose:
        print(npz.keys())
    traj = npz["trajectory"]
    headers = [h.split()[0] for h in npz["trajectory_headers"]]
    df = pd.DataFrame({h: [fr[k] for fr in traj] for k, h in enumerate(headers)}, columns=headers)
    df.to_pickle(out_filename)
    return df


def convert_pop_to_energy(pop_filename, out_filename=None):
    """Populations -> reduced energies U = -ln(P / sum P), NaN populations count as 0.001 (toolbox.py:322-348)."""
    if pop_filename.endswith('txt') or pop_filename.endswith('dat'):
        pop = np.loadtxt(pop_filename)
    elif pop_filename.endswith('npy'):
        pop = np.load(pop_filename)
    else:
        raise ValueError('Incompatible file extention. Use:{.txt,.dat,.npy}')
    pop[np.isnan(pop)] = 0.001
    total = float(sum(pop))
    energy = [-np.log((i / total)) for i in pop]
    np.savetxt('energy.txt' if out_filename is None else out_filename, energy)
    return energy


def save_object(obj, filename):
    with open(filename, "wb") as output:
        pickle.dump(obj, output, pickle.HIGHEST_PROTOCOL)


def mkdir(path):
    if not os.path.exists(path):
        os.makedirs(path)
