"""tensordict stand-in (TEST INFRASTRUCTURE): dict-backed TensorClass + UnbatchedTensor, enough for the reference's
single-env eager path (attribute access, replace/update via ``_tensordict._tensordict``, clone, indexing, torch.cat /
torch.stack, and pytree flattening so instances can cross ``torch.vmap``)."""
import dataclasses

import torch
import torch.utils._pytree as pytree


class UnbatchedTensor(torch.Tensor):
    @staticmethod
    def __new__(cls, data=None, **kw):
        t = data if isinstance(data, torch.Tensor) else torch.as_tensor(data)
        return torch.Tensor._make_subclass(cls, t.detach(), False)

    def __init__(self, data=None, **kw):
        pass

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # behave like a plain tensor in every computation
        def unwrap(x):
            return x.as_subclass(torch.Tensor) if isinstance(x, UnbatchedTensor) else x

        args = pytree.tree_map(unwrap, args)
        kwargs = pytree.tree_map(unwrap, kwargs or {})
        return func(*args, **kwargs)


class _Store:
    def __init__(self, d=None):
        self._tensordict = dict(d or {})
        self.leaf_keys = None

    def __getitem__(self, k):
        return self._tensordict[k]

    def __setitem__(self, k, v):
        self._tensordict[k] = v


class _TensorClassMeta(type):
    def __new__(mcs, name, bases, namespace, **kwargs):
        cls = super().__new__(mcs, name, bases, namespace)
        if namespace.get("__annotations__"):
            cls = dataclasses.dataclass(init=False, eq=False, repr=False)(cls)
            defaults = {}
            for f in dataclasses.fields(cls):
                if f.default is not dataclasses.MISSING:
                    defaults[f.name] = f.default
                # field values live in the instance store, never on the class
                if f.name in cls.__dict__ and not isinstance(cls.__dict__[f.name], (property, classmethod, staticmethod)):
                    try:
                        delattr(cls, f.name)
                    except AttributeError:
                        pass
            cls._tc_defaults = {**getattr(cls, "_tc_defaults", {}), **defaults}
            cls._tc_names = tuple(f.name for f in dataclasses.fields(cls))
            _register(cls)
        return cls

    def __getitem__(cls, item):  # TensorClass["nocast", "shadow"]
        return cls


class TensorClass(metaclass=_TensorClassMeta):
    _tc_names = ()
    _tc_defaults = {}

    def __init__(self, *args, batch_size=None, device=None, **kw):
        vals = dict(self._tc_defaults)
        for n, a in zip(self._tc_names, args):
            vals[n] = a
        vals.update(kw)
        object.__setattr__(self, "_tensordict", _Store(vals))
        object.__setattr__(self, "_batch_size", torch.Size(batch_size or []))

    def __getattr__(self, name):
        try:
            store = object.__getattribute__(self, "_tensordict")._tensordict
        except AttributeError:
            raise AttributeError(name) from None
        if name in store:
            return store[name]
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    def __setattr__(self, name, value):
        if name in self._tc_names:
            self._tensordict._tensordict[name] = value
        else:
            object.__setattr__(self, name, value)

    @property
    def batch_size(self):
        return self._batch_size

    def auto_batch_size_(self):
        for v in self._tensordict._tensordict.values():
            if isinstance(v, torch.Tensor) and v.dim() > 0:
                object.__setattr__(self, "_batch_size", torch.Size([v.shape[0]]))
                break
        return self

    def _new(self, d, batch_size=None):
        out = object.__new__(type(self))
        object.__setattr__(out, "_tensordict", _Store(d))
        object.__setattr__(out, "_batch_size", self._batch_size if batch_size is None else torch.Size(batch_size))
        return out

    def clone(self, recurse=True):
        d = {}
        for k, v in self._tensordict._tensordict.items():
            if recurse and isinstance(v, torch.Tensor):
                d[k] = v.clone()
            elif isinstance(v, TensorClass):
                d[k] = v.clone(recurse)
            else:
                d[k] = v
        return self._new(d)

    def to(self, *a, **k):
        d = {kk: (v.to(*a, **k) if isinstance(v, (torch.Tensor, TensorClass)) else v) for kk, v in self._tensordict._tensordict.items()}
        return self._new(d)

    def __getitem__(self, idx):
        d = {}
        for k, v in self._tensordict._tensordict.items():
            d[k] = v[idx] if isinstance(v, (torch.Tensor, TensorClass)) and not isinstance(v, UnbatchedTensor) else v
        out = self._new(d)
        return out.auto_batch_size_() if len(self._batch_size) else out

    def __len__(self):
        return self._batch_size[0]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.cat, torch.stack):
            items = list(args[0])
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            first = items[0]
            d = {}
            for k, v in first._tensordict._tensordict.items():
                if isinstance(v, torch.Tensor):
                    d[k] = func([it._tensordict._tensordict[k] for it in items], dim=dim)
                else:
                    d[k] = v
            out = first._new(d)
            return out.auto_batch_size_()
        return NotImplemented


def _register(cls):
    # Model / Data / Option / Statistic only ever cross torch.vmap with in_dim=None (closed-over, unbatched):
    # keep them opaque so functorch passes them through untouched, like real tensordict does.
    if cls.__name__ in ("Model", "Data", "Option", "Statistic"):
        return

    def flatten(obj):
        # only tensors / nested tensorclasses are leaves; everything else (None, ints, tuples ...) rides in the context
        store = obj._tensordict._tensordict
        forced = getattr(obj._tensordict, "leaf_keys", None)  # set by unflatten: the split is structural, not by value type
        is_leaf = (lambda k, v: k in forced) if forced is not None else (lambda k, v: isinstance(v, (torch.Tensor, TensorClass)))
        keys = tuple(k for k, v in store.items() if is_leaf(k, v))
        static = tuple((k, v) for k, v in store.items() if not is_leaf(k, v))
        return [store[k] for k in keys], (keys, static, tuple(store.keys()), tuple(obj._batch_size))

    def unflatten(values, ctx):
        keys, static, order, bs = ctx
        d = dict(zip(keys, values))
        d.update(dict(static))
        out = object.__new__(cls)
        st = _Store({k: d[k] for k in order})
        st.leaf_keys = frozenset(keys)
        object.__setattr__(out, "_tensordict", st)
        object.__setattr__(out, "_batch_size", torch.Size(bs))
        return out

    try:
        pytree.register_pytree_node(cls, flatten, unflatten)
    except ValueError:
        pass
