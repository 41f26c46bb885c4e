"""Import shim: numeric-aware sort standing in for natsort.natsorted."""
import re

_tok = re.compile(r'(\d+(?:\.\d+)?)')

def _key(s):
    out = []
    for p in _tok.split(str(s)):
        if p == '':
            continue
        out.append((0, float(p), '') if _tok.fullmatch(p) else (1, 0.0, p))
    return out

def natsorted(seq):
    return sorted(seq, key=_key)
