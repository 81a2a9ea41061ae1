"""``device_get_into``: copy a (batched) ``Data`` off the device into MjData-like host structs.

Mirror of the reference's ``mujoco_torch.device_get_into`` (``mujoco_torch/_src/device.py:1119-1205``): same argument meaning,
same field rules, same errors.  ``result`` is anything with MjData's attribute names -- a ``mujoco.MjData`` when the ``mujoco``
package is present, or any plain object (``types.SimpleNamespace``) otherwise; nested ``contact`` goes into ``result.contact``.

Field rules kept from the reference:
  * rotation matrices travel flattened, ``(..., 3, 3) -> (..., 9)`` (``_INVERSE_TRANSFORMS`` :90-107), ``contact.frame`` too;
  * ``efc_J``, ``actuator_moment``, ``qM``, ``qLD`` are skipped: their MuJoCo formats differ (sparse / diagonal) (``_DERIVED`` :112-127);
  * ``contact.contact_dim`` is written to MuJoCo's name ``dim`` (``_FIELD_TARGET_MAP`` :110);
  * a target attribute whose shape differs from the value's is left untouched (:1192-1200);
  * a list ``result`` must have one entry per env and receives the envs one by one with the batch dimension stripped.
"""

from __future__ import annotations

import numpy as np
import torch

from .data import Contact, Data

_FLATTEN_33 = {"ximat", "xmat", "geom_xmat", "site_xmat", "cam_xmat"}
_SKIPPED = {"efc_J", "actuator_moment", "qM", "qLD"}
_RENAMED = {(Contact, "contact_dim"): "dim"}


def _realloc_con_efc(result, ncon: int, nefc: int):
    """``mujoco._functions._realloc_con_efc`` when the target is a real MjData (reference :1161-1169); duck-typed targets have
    no arena to resize."""
    try:
        import mujoco  # noqa: PLC0415
    except ImportError:
        return
    if isinstance(result, mujoco.MjData):
        mujoco._functions._realloc_con_efc(result, ncon=ncon, nefc=nefc)  # pylint: disable=protected-access


def _put(result, name, value: np.ndarray):
    cur = getattr(result, name, None)
    if cur is not None and hasattr(cur, "shape") and cur.shape != value.shape:
        return   # incompatible layout on the MuJoCo side: skipped like the reference
    try:
        setattr(result, name, value)
    except (AttributeError, ValueError):
        getattr(result, name)[:] = value


def _get_into_one(result, value):
    if isinstance(value, Data):
        _realloc_con_efc(result, int(value._meta.get("ncon", 0)), int(value._meta.get("nefc", 0)))
    for name in value.keys():
        if isinstance(value, Data) and name in _SKIPPED:
            continue
        v = getattr(value, name)
        if isinstance(v, Contact):
            _get_into_one(getattr(result, name), v)
            continue
        if not isinstance(v, torch.Tensor):
            continue
        v = v.detach().cpu()
        if (isinstance(value, Data) and name in _FLATTEN_33) or (isinstance(value, Contact) and name == "frame"):
            v = v.reshape(v.shape[:-2] + (9,)) if v.shape[-2:] == (3, 3) else v
        _put(result, _RENAMED.get((type(value), name), name), v.numpy())
    if isinstance(value, Data):
        for k in ("ne", "nf", "nl", "ncon", "nefc"):
            if k in value._meta:
                try:
                    setattr(result, k, int(value._meta[k]))
                except (AttributeError, ValueError, TypeError):
                    pass   # read-only on a real MjData (set through _realloc_con_efc)


def device_get_into(result, value: Data):
    """Transfers ``value`` off the device into ``result`` (one MjData-like struct: arrays keep the batch dimension; or a list of
    them, one per env)."""
    if not isinstance(value, Data):
        raise NotImplementedError(f"device_get_into: {type(value)} is not supported (Data only).")
    if isinstance(result, list):
        bs = tuple(value.batch_size)
        if len(bs) < 1:
            raise ValueError("unrecognizable batch dimension in value")
        if len(result) != bs[0]:
            raise ValueError(f"result length ({len(result)}) doesn't match value batch size ({bs[0]})")
        host = value.to("cpu")
        for i, r in enumerate(result):
            _get_into_one(r, host[i])
        return
    _get_into_one(result, value)
