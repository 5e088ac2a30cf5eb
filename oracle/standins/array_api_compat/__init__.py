"""TEST INFRASTRUCTURE ONLY - numpy-only stand-in for the third-party package
``array_api_compat`` so that the UNMODIFIED reference (bilby, imported from
/root/reference) can be imported in the build container to validate the oracle
and to generate golden vectors.  Surface = the symbols bilby touches
(SURVEY.md Appendix B).  Never imported by the product package."""
import numpy as _np


def _is_np(x):
    return isinstance(x, (_np.ndarray, _np.generic))


def array_namespace(*xs, **_kw):
    seen = False
    for x in xs:
        if x is None or isinstance(x, (bool, int, float, complex)):
            continue
        if _is_np(x):
            seen = True
            continue
        raise TypeError(f"not an array API object: {type(x)}")
    if not seen:
        raise TypeError("no array inputs")
    return _np


get_namespace = array_namespace


def is_numpy_namespace(xp):
    return xp is _np


def is_jax_namespace(xp):
    return False


def is_torch_namespace(xp):
    return False


def is_jax_array(x):
    return False


def is_torch_array(x):
    return False


def is_numpy_array(x):
    return _is_np(x)


def is_array_api_obj(x):
    return _is_np(x)


def device(x):
    return "cpu"


def to_device(x, device, **_kw):
    return x
