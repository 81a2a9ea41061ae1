"""``Data`` / ``Contact``: per-environment simulation state with the reference's container surface.

The reference builds these on ``tensordict.TensorClass`` (reference types.py:1036-1261, dataclasses.py:85-134).
Here they are small attribute containers over plain tensors whose leading dims are the batch dims; after a
``step`` the fields are zero-copy strided views of the env-minor (W, B) buffer the kernels wrote, so
``d.qpos[b]`` semantics hold while the device layout stays coalesced.  Supported call forms:
``replace``, ``update_``, ``tree_replace``, ``clone``, ``to``, ``expand``, ``d[i]``, ``d[mask] = other``,
``torch.stack([...])`` and use under ``torch.vmap`` (both classes are registered pytrees).
"""

from __future__ import annotations

import torch
import torch.utils._pytree as pytree


class _TensorBag:
    """Base: ordered tensor fields + static metadata, batch dims leading."""

    _ndim_key: str = ""       # field used to infer batch_size
    _ndim_event: int = 1      # number of trailing (event) dims of that field

    def __init__(self, fields: dict, meta: dict | None = None):
        object.__setattr__(self, "_fields", dict(fields))
        object.__setattr__(self, "_meta", dict(meta or {}))

    # -- attribute access
    def __getattr__(self, name):
        f = object.__getattribute__(self, "_fields")
        if name in f:
            return f[name]
        mt = object.__getattribute__(self, "_meta")
        if name in mt:
            return mt[name]
        raise AttributeError(f"{type(self).__name__} has no field {name!r}")

    def __setattr__(self, name, value):
        if name in self._fields or isinstance(value, (torch.Tensor, _TensorBag)):
            self._fields[name] = value
        else:
            self._meta[name] = value

    def __contains__(self, name):
        return name in self._fields

    def keys(self):
        return self._fields.keys()

    @classmethod
    def fields(cls):
        return tuple(cls._FIELD_ORDER) if hasattr(cls, "_FIELD_ORDER") else ()

    # -- batch
    @property
    def batch_size(self) -> torch.Size:
        t = self._fields[self._ndim_key]
        return t.shape[: t.dim() - self._ndim_event]

    @property
    def device(self):
        return self._fields[self._ndim_key].device

    @property
    def dtype(self):
        return self._fields[self._ndim_key].dtype

    def _map(self, fn):
        out = {}
        for k, v in self._fields.items():
            out[k] = v._map(fn) if isinstance(v, _TensorBag) else fn(v)
        return type(self)(out, self._meta)

    def _map_batched(self, fn):
        """Apply fn(tensor, n_event_dims) to every field."""
        nb = len(self.batch_size)
        out = {}
        for k, v in self._fields.items():
            if isinstance(v, _TensorBag):
                out[k] = v._map_batched(fn)
            else:
                out[k] = fn(v, v.dim() - nb)
        return type(self)(out, self._meta)

    # -- reference container API
    def replace(self, **kw):
        """Shallow copy with fields replaced (reference dataclasses.py:101-110)."""
        new = type(self)(self._fields, self._meta)
        new._fields.update(kw)
        return new

    def update_(self, **kw):
        """In-place field update (reference dataclasses.py:112-120)."""
        self._fields.update(kw)
        return self

    def tree_replace(self, params: dict):
        """Replace nested fields addressed by dotted paths (reference dataclasses.py:122-130)."""
        new = self
        for path, val in params.items():
            head, _, rest = path.partition(".")
            if rest:
                new = new.replace(**{head: getattr(new, head).tree_replace({rest: val})})
            else:
                new = new.replace(**{head: val})
        return new

    def clone(self, recurse: bool = True):
        if not recurse:
            return type(self)({k: (v.clone(False) if isinstance(v, _TensorBag) else v) for k, v in self._fields.items()}, self._meta)
        nb = len(self.batch_size)

        def env_minor_clone(t, ne):
            # A batched container is cloned into env-minor memory (batch index fastest; same logical shape and values): the
            # layout the kernels read, so stepping a freshly cloned batch needs no per-step transposes of its inputs.
            if nb == 1 and ne >= 1 and t.shape[0] > 1:
                return t.movedim(0, -1).contiguous().movedim(-1, 0)
            return t.clone()

        return self._map_batched(env_minor_clone)

    def to(self, *args, **kw):
        device = kw.get("device")
        dtype = kw.get("dtype")
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            elif isinstance(a, (str, torch.device, int)):
                device = a
        nb = bool(kw.get("non_blocking", False))

        def f(t):
            return t.to(device=device, dtype=dtype if t.is_floating_point() else None, non_blocking=nb)

        return self._map(f)

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def expand(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        return self._map_batched(lambda t, ne: t.expand(*shape, *t.shape[t.dim() - ne:]))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        return self._map_batched(lambda t, ne: t.reshape(*shape, *t.shape[t.dim() - ne:]))

    def contiguous(self):
        return self._map(lambda t: t.contiguous())

    def __getitem__(self, idx):
        return self._map(lambda t: t[idx])

    def __setitem__(self, idx, other):
        for k, v in self._fields.items():
            o = other._fields[k] if isinstance(other, _TensorBag) else other
            if isinstance(v, _TensorBag):
                v[idx] = o
            else:
                # model-constant fields come back from step() as stride-0 expansions of one per-model tensor (Contact.friction,
                # geom ids ...): a write must neither alias across envs nor reach the other Data objects, so such a field is
                # materialised on its first write (the reference's tensorclass owns all of its memory, types.py:1091)
                if any(st == 0 and sz > 1 for st, sz in zip(v.stride(), v.shape)):
                    v = v.clone(memory_format=torch.contiguous_format)
                    self._fields[k] = v
                v[idx] = o

    def __len__(self):
        bs = self.batch_size
        if not bs:
            raise TypeError("len() of an unbatched container")
        return bs[0]

    def __repr__(self):
        return f"{type(self).__name__}(batch_size={tuple(self.batch_size)}, fields={len(self._fields)})"

    # -- torch.stack / torch.cat over containers
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.stack, torch.cat):
            items = list(args[0])
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            first = items[0]
            out = {}
            for k, v in first._fields.items():
                vals = [it._fields[k] for it in items]
                out[k] = func(vals, dim=dim) if not isinstance(v, _TensorBag) else type(v).__torch_function__(
                    func, (type(v),), (vals,), {"dim": dim})
            return type(first)(out, first._meta)
        return NotImplemented


class Contact(_TensorBag):
    """reference types.Contact :1036-1088."""

    _ndim_key, _ndim_event = "dist", 1
    _FIELD_ORDER = ("dist", "pos", "frame", "includemargin", "friction", "solref", "solreffriction", "solimp", "contact_dim",
                    "geom1", "geom2", "geom", "efc_address")


class Data(_TensorBag):
    """reference types.Data :1091-1261."""

    _ndim_key, _ndim_event = "qpos", 1


def _flatten(bag):
    keys = tuple(bag._fields.keys())
    return [bag._fields[k] for k in keys], (keys, tuple(sorted(bag._meta.items(), key=lambda kv: kv[0])))


def _make_unflatten(cls):
    def unflatten(values, ctx):
        keys, meta = ctx
        return cls(dict(zip(keys, values)), dict(meta))

    return unflatten


for _cls in (Contact, Data):
    pytree.register_pytree_node(_cls, _flatten, _make_unflatten(_cls), serialized_type_name=f"mujoco_torch_b200.{_cls.__name__}")
