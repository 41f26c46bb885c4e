def kln_to_kn(*a, **k):
    raise ImportError("pymbar is not installed (oracle import shim)")
