"""Import shim: plotting is out of scope for the oracle.  Every attribute is a do-nothing object
that can be called, indexed, iterated and unpacked (`fig, ax = plt.subplots()`), so the reference's
plotting code runs through without drawing anything."""


class _Null(object):
    def __call__(self, *a, **k):
        return _Null()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Null()

    def __getitem__(self, k):
        return _Null()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter((_Null(), _Null()))

    def __len__(self):
        return 2

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Null()
