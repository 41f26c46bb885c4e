"""Import shim: pymbar is absent; oracle/mbar_oracle.py injects a numpy MBAR when needed."""
class MBAR:
    def __init__(self, *a, **k):
        raise ImportError("pymbar is not installed (oracle import shim)")
