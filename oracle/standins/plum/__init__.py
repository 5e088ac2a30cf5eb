"""TEST INFRASTRUCTURE ONLY - minimal stand-in for ``plum.dispatch``: overloads
of the same qualified name are distinguished by arity and, where annotated, by
isinstance checks on the annotations.  See oracle/standins/array_api_compat."""
import inspect
import typing
import types

_REGISTRY = {}


def _matches(ann, value):
    if ann is inspect.Parameter.empty:
        return True
    origin = typing.get_origin(ann)
    if origin in (typing.Union, types.UnionType):
        return any(_matches(a, value) for a in typing.get_args(ann))
    if isinstance(ann, type):
        return isinstance(value, ann)
    return True


class _Function:
    def __init__(self, name):
        self.__name__ = name
        self._methods = []

    def register(self, fn, precedence=0):
        sig = inspect.signature(fn)
        try:
            hints = typing.get_type_hints(fn)
        except Exception:
            hints = {}
        self._methods.append((precedence, len(self._methods), fn, sig, hints))
        self.__doc__ = fn.__doc__
        self.__wrapped__ = fn
        return self

    def dispatch(self, fn=None, precedence=0):
        if fn is None:
            return lambda f: self.register(f, precedence)
        return self.register(fn, precedence)

    def __call__(self, *args, **kwargs):
        best = None
        for prec, order, fn, sig, hints in self._methods:
            try:
                bound = sig.bind(*args, **kwargs)
            except TypeError:
                continue
            ok = True
            n_typed = 0
            for pname, value in bound.arguments.items():
                ann = hints.get(pname, inspect.Parameter.empty)
                if ann is not inspect.Parameter.empty:
                    n_typed += 1
                if not _matches(ann, value):
                    ok = False
                    break
            if not ok:
                continue
            key = (prec, n_typed, order)
            if best is None or key > best[0]:
                best = (key, fn)
        if best is None:
            raise TypeError(f"no overload of {self.__name__} matches")
        return best[1](*args, **kwargs)

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        return types.MethodType(self, obj)


def dispatch(fn=None, precedence=0):
    if fn is None:
        return lambda f: dispatch(f, precedence=precedence)
    key = (fn.__module__, fn.__qualname__)
    func = _REGISTRY.get(key)
    if func is None:
        func = _Function(fn.__name__)
        func.__module__ = fn.__module__
        func.__qualname__ = fn.__qualname__
        _REGISTRY[key] = func
    return func.register(fn, precedence)
