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
    """Globbed files in natural (numeric-aware) order (toolbox.py:66-80)."""
    def key(s):
        out = []
        for p in re.split(r"(\d+(?:\.\d+)?)", s):
            if p == "":
                continue
            out.append((0, float(p), "") if re.fullmatch(r"\d+(?:\.\d+)?", p) else (1, 0.0, p))
        return out
    return sorted(glob.glob(path), key=key)


def list_res(input_data):
    scheme = []
    for i in input_data[0]:
        if not any(i.endswith(ext) for ext in list_possible_extensions()):
            raise ValueError(f"Incompatible File extension. Use:{list_possible_extensions()}")
        scheme.append(i.split(".")[-1])
    return scheme


def list_extensions(input_data):
    return [res.split("_")[-1] for res in list_res(input_data)]


def save_object(obj, filename):
    with open(filename, "wb") as output:
        pickle.dump(obj, output, pickle.HIGHEST_PROTOCOL)


def mkdir(path):
    if not os.path.exists(path):
        os.makedirs(path)
