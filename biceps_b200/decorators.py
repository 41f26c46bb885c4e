"""biceps.multiprocess (reference: /root/reference/biceps/decorators.py:11-55): decorator that calls a function once
per element of `iterable` (the reference's scripts use it to run one lambda value per forked CPU process).

Here every call already fills the GPU with chains and a CUDA context must not be forked, so the calls run one after
the other in this process, in the order of `iterable`; like the reference, decorating runs them immediately and the
decorated name is bound to None."""


def multiprocess(*args, **kwargs):
    def wrapper(function):
        for item in kwargs['iterable']:
            function(item)
    return wrapper
