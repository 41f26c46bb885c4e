"""Import shim: mdtraj is absent here and the sampling path never calls it."""
def load(*a, **k):
    raise ImportError("mdtraj is not installed (oracle import shim)")
