"""TEST INFRASTRUCTURE ONLY - numpy-only stand-in for ``array_api_extra``
(only ``at(arr, idx).set/add/multiply``), see oracle/standins/array_api_compat."""
import numpy as _np


class _At:
    def __init__(self, arr, idx=None):
        self._arr = arr
        self._idx = idx

    def __getitem__(self, idx):
        return _At(self._arr, idx)

    def _copy(self):
        return _np.array(self._arr, copy=True)

    def set(self, value, **_kw):
        out = self._copy()
        out[self._idx] = value
        return out

    def add(self, value, **_kw):
        out = self._copy()
        out[self._idx] += value
        return out

    def multiply(self, value, **_kw):
        out = self._copy()
        out[self._idx] *= value
        return out


def at(arr, idx=None):
    return _At(arr, idx)
