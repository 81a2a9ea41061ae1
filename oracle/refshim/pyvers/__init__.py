"""pyvers.implement_for stand-in: version dispatch pinned to the shimmed module's __version__."""
import importlib


def _ver(s):
    return tuple(int(x) for x in str(s).split(".")[:3])


class implement_for:
    def __init__(self, module_name, from_version=None, to_version=None):
        self.module_name, self.lo, self.hi = module_name, from_version, to_version

    def _matches(self, lo, hi):
        v = _ver(importlib.import_module(self.module_name).__version__)
        return (lo is None or v >= _ver(lo)) and (hi is None or v < _ver(hi))

    def __call__(self, fn):
        outer = self
        state = {"impl": fn if (self.lo or self.hi) and self._matches(self.lo, self.hi) else None, "base": fn}

        def dispatcher(*a, **k):
            impl = state["impl"] or state["base"]
            return impl(*a, **k)

        def register(from_version=None, to_version=None):
            def deco(f):
                if outer._matches(from_version, to_version):
                    state["impl"] = f
                return dispatcher

            return deco

        dispatcher.register = register
        dispatcher.__name__ = getattr(fn, "__name__", "dispatcher")
        if (self.lo or self.hi) and not self._matches(self.lo, self.hi):
            # redefinition pattern: a later @implement_for with a matching range rebinds the same module-level name
            def unavailable(*a, **k):
                raise ModuleNotFoundError(f"{fn.__name__}: no implementation for this {outer.module_name} version")

            unavailable.register = register
            return unavailable
        return dispatcher
