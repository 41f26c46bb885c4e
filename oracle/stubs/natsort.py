"""Import shim standing in for natsort.natsorted with natsort's DEFAULT algorithm (ns.INT, unsigned): a name is
split into text and unsigned-integer runs, a decimal point is text ('lambda0.5' < 'lambda0.25' because 5 < 25)."""
import re


def _key(s):
    return [int(p) if i % 2 else p for i, p in enumerate(re.split(r'(\d+)', str(s)))]


def natsorted(seq):
    return sorted(seq, key=_key)
