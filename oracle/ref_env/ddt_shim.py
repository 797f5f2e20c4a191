"""Minimal stand-in for the ``ddt`` package (TEST INFRASTRUCTURE ONLY): the reference's circuit / algorithm
tests that drive BasicAer are parametrised with ``@ddt`` / ``@data`` / ``@unpack`` / ``@idata``
(e.g. test/python/circuit/test_diagonal_gate.py); ddt is not installed here.  ``ddt(cls)`` expands every
decorated test into one method per datum, like the real package."""
import re

_DATA = "_ddt_shim_data"
_UNPACK = "_ddt_shim_unpack"


def data(*values):
    def wrap(func):
        setattr(func, _DATA, list(values))
        return func

    return wrap


def idata(iterable):
    return data(*list(iterable))


def unpack(func):
    setattr(func, _UNPACK, True)
    return func


def _simple(value):
    if isinstance(value, (str, int, float, bool, type(None))):
        return True
    return isinstance(value, (tuple, list)) and all(_simple(v) for v in value)


def _label(value):
    """Name suffix; only for plain values, so that test ids do not depend on randomly generated data."""
    if not _simple(value):
        return ""
    text = re.sub(r"\W+", "_", str(value))
    return text[:40]


def _bind(func, value, unpacked):
    def test(self):
        if unpacked and isinstance(value, (tuple, list)):
            return func(self, *value)
        if unpacked and isinstance(value, dict):
            return func(self, **value)
        return func(self, value)

    test.__doc__ = func.__doc__
    return test


def ddt(cls):
    for name, func in list(vars(cls).items()):
        values = getattr(func, _DATA, None)
        if values is None:
            continue
        for i, value in enumerate(values):
            test = _bind(func, value, getattr(func, _UNPACK, False))
            test.__name__ = "%s_%d_%s" % (name, i + 1, _label(value))
            setattr(cls, test.__name__, test)
        delattr(cls, name)
    return cls
