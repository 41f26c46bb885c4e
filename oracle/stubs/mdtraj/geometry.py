def compute_phi(*a, **k):
    raise ImportError("mdtraj is not installed (oracle import shim)")
