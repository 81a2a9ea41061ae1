"""ctypes binding of libmjt_b200 (include/mjt.h) and its registration as ``torch.library`` custom ops.

``torch.ops.mjt.step`` / ``torch.ops.mjt.forward`` take the state tensors of a (batched) Data and return the
env-minor output buffer viewed batch-first.  The ops are CUDA-only by construction: there is no CPU
implementation and no fallback -- a missing extension or a CPU tensor raises.
"""

from __future__ import annotations

import ctypes
import os
import weakref

import torch

# MJT_LIB selects another build of the same library (tuning variants, scripts/build_variants.sh); there is still no fallback
_LIB_PATH = os.environ.get("MJT_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libmjt_b200.so")
_lib = None

MODE_STEP, MODE_FORWARD = 1, 2


class StateIn(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "qpos", "qvel", "act", "time", "qacc", "qacc_warmstart", "ctrl", "qfrc_applied", "xfrc_applied", "mocap_pos",
        "mocap_quat", "sensordata", "qfrc_constraint", "eq_active")]


def lib():
    """Load the shared library (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = ctypes.CDLL(_LIB_PATH)
    L.mjt_version.restype = ctypes.c_int
    L.mjt_last_error.restype = ctypes.c_char_p
    L.mjt_model_create.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.mjt_model_create.restype = ctypes.c_int
    L.mjt_model_destroy.argtypes = [ctypes.c_void_p]
    L.mjt_model_destroy.restype = None
    for fn in ("mjt_model_data_words", "mjt_model_scratch_words", "mjt_model_dtype_size", "mjt_step_launches"):
        getattr(L, fn).argtypes = [ctypes.c_void_p]
        getattr(L, fn).restype = ctypes.c_int
    L.mjt_model_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.mjt_model_field.restype = ctypes.c_int
    L.mjt_step.argtypes = [ctypes.c_void_p, ctypes.POINTER(StateIn), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.mjt_step.restype = ctypes.c_int
    L.mjt_forward.argtypes = [ctypes.c_void_p, ctypes.POINTER(StateIn), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.mjt_forward.restype = ctypes.c_int
    L.mjt_profile_begin.restype = None
    L.mjt_profile_end.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    L.mjt_profile_end.restype = ctypes.c_int
    L.mjt_solve_config.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.mjt_solve_config.restype = ctypes.c_int
    L.mjt_fma_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
    L.mjt_fma_peak.restype = ctypes.c_int
    _lib = L
    return L


def fma_peak(f64: bool) -> float:
    """Measured FMA-pipe throughput of the current CUDA device in TFLOP/s (csrc/mjt_peak.cu)."""
    v = ctypes.c_double()
    _check(lib().mjt_fma_peak(1 if f64 else 0, ctypes.byref(v), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return v.value


def profile_begin():
    lib().mjt_profile_begin()


def profile_end() -> dict:
    """{kernel name: (total ms, launches)} since profile_begin()."""
    buf = ctypes.create_string_buffer(4096)
    _check(lib().mjt_profile_end(buf, 4096))
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            name, ms, n = item.split(":")
            out[name] = (float(ms), int(n))
    return out


EXPORTS = ("mjt_version", "mjt_last_error", "mjt_model_create", "mjt_model_destroy", "mjt_model_data_words",
           "mjt_model_scratch_words", "mjt_model_dtype_size", "mjt_model_field", "mjt_step", "mjt_forward", "mjt_step_launches",
           "mjt_profile_begin", "mjt_profile_end", "mjt_solve_config", "mjt_fma_peak")


def _check(rc):
    if rc != 0:
        raise RuntimeError("libmjt_b200: " + lib().mjt_last_error().decode())


# ------------------------------------------------------------------------------------------ model registry

_MODELS: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_next_id = [1]


def register_model(model) -> int:
    mid = _next_id[0]
    _next_id[0] += 1
    _MODELS[mid] = model
    return mid


def model_handle(model):
    """Device handle of a Model (created lazily on the model's CUDA device)."""
    if model._handle is not None:
        return model._handle
    if model.device.type != "cuda":
        raise RuntimeError("mujoco_torch_b200 runs on CUDA only: move the model with .to('cuda') (there is no CPU path)")
    L = lib()
    blob = model.blob_np
    h = ctypes.c_void_p()
    dev = model.device.index if model.device.index is not None else torch.cuda.current_device()
    _check(L.mjt_model_create(blob.ctypes.data_as(ctypes.c_void_p), blob.nbytes, dev, ctypes.byref(h)))
    model._handle = h
    weakref.finalize(model, L.mjt_model_destroy, h)
    assert L.mjt_model_data_words(h) == model.data_words
    return h


# ------------------------------------------------------------------------------------------ custom ops

_STATE_ARGS = ("qpos", "qvel", "act", "time", "qacc_warmstart", "ctrl", "qfrc_applied", "xfrc_applied", "mocap_pos", "mocap_quat",
               "sensordata", "qfrc_constraint", "eq_active")


def _env_minor(t: torch.Tensor, batched: bool, dtype) -> torch.Tensor:
    """(B, n) / (n,) logical -> contiguous (n, B) physical (no copy when the tensor already is an env-minor view)."""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not batched:
        t = t.reshape(1, -1)
    else:
        t = t.reshape(t.shape[0], -1)
    return t.t().contiguous()


# Inputs that usually stay the same tensor from step to step (user forces, mocap targets, equality switches): their env-minor
# copy is kept per model and reused while the source is untouched (same storage, shape, strides and version counter), which
# removes two strided-copy kernels from every Humanoid step.
_CACHED_ARGS = ("qfrc_applied", "xfrc_applied", "mocap_pos", "mocap_quat", "eq_active")
_MINOR_CACHE_SLOTS, _MINOR_CACHE_BYTES = 8, 512 << 20


def _env_minor_cached(model, name, t: torch.Tensor, batched: bool, dtype) -> torch.Tensor:
    if name not in _CACHED_ARGS or t.numel() == 0:
        return _env_minor(t, batched, dtype)
    try:
        key = (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype, str(t.device), batched)
    except RuntimeError:   # inference tensors do not track versions: no reuse
        return _env_minor(t, batched, dtype)
    cache = model.__dict__.setdefault("_minor_cache", {})   # (name, storage address) -> (key, source, env-minor copy)
    slot = (name, key[0])
    hit = cache.get(slot)
    if hit is not None and hit[0] == key:
        return hit[2]
    e = _env_minor(t, batched, dtype)
    if e.data_ptr() != t.data_ptr():
        cache.pop(slot, None)
        cache[slot] = (key, t, e)   # the source is kept alive with its copy, so its address cannot be reused meanwhile
        # HostPipeline steps several chunk templates in turn, so a few slots per input; bounded in count and in bytes
        while len(cache) > 1 and (len(cache) > _MINOR_CACHE_SLOTS
                                  or sum(v[2].numel() * v[2].element_size() for v in cache.values()) > _MINOR_CACHE_BYTES):
            cache.pop(next(iter(cache)))
    return e


def _launch(model_id: int, mode: int, tensors) -> torch.Tensor:
    model = _MODELS.get(model_id)
    if model is None:
        raise RuntimeError("mjt: unknown model id (model was garbage collected)")
    qpos = tensors[0]
    if not qpos.is_cuda:
        raise RuntimeError("mjt::step is CUDA-only; got a CPU tensor (no CPU fallback exists)")
    batched = qpos.dim() == 2
    B = qpos.shape[0] if batched else 1
    h = model_handle(model)
    L = lib()
    keep = []
    st = StateIn()
    for name, t in zip(_STATE_ARGS, tensors):
        want = torch.int32 if name == "eq_active" else model.dtype
        e = _env_minor_cached(model, name, t, batched, want)
        keep.append(e)
        setattr(st, name, e.data_ptr() if e.numel() else None)
    st.qacc = None
    with torch.cuda.device(qpos.device):
        out = torch.empty((model.data_words, B), dtype=model.dtype, device=qpos.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if mode == MODE_STEP:
            sw = L.mjt_model_scratch_words(h)
            scratch = torch.empty((sw, B), dtype=model.dtype, device=qpos.device) if sw else None
            _check(L.mjt_step(h, ctypes.byref(st), out.data_ptr(), scratch.data_ptr() if sw else None, B, stream))
        else:
            _check(L.mjt_forward(h, ctypes.byref(st), out.data_ptr(), B, stream))
    return out.t() if batched else out[:, 0]


@torch.library.custom_op("mjt::run", mutates_args=(), device_types="cuda")
def run(model_id: int, mode: int, qpos: torch.Tensor, qvel: torch.Tensor, act: torch.Tensor, time: torch.Tensor,
        qacc_warmstart: torch.Tensor, ctrl: torch.Tensor, qfrc_applied: torch.Tensor, xfrc_applied: torch.Tensor,
        mocap_pos: torch.Tensor, mocap_quat: torch.Tensor, sensordata: torch.Tensor, qfrc_constraint: torch.Tensor,
        eq_active: torch.Tensor) -> torch.Tensor:
    out = _launch(model_id, mode, (qpos, qvel, act, time, qacc_warmstart, ctrl, qfrc_applied, xfrc_applied, mocap_pos, mocap_quat,
                                   sensordata, qfrc_constraint, eq_active))
    return out.clone() if out.dim() == 1 else out


@run.register_fake
def _(model_id, mode, qpos, *rest):
    model = _MODELS[model_id]
    if qpos.dim() == 2:
        return qpos.new_empty((model.data_words, qpos.shape[0])).t()
    return qpos.new_empty((model.data_words,))


@run.register_vmap
def _(info, in_dims, model_id, mode, *tensors):
    B = info.batch_size
    moved = []
    for t, dim in zip(tensors, in_dims[2:]):
        if dim is None:
            t = t.unsqueeze(0).expand(B, *t.shape)
        else:
            t = t.movedim(dim, 0)
        moved.append(t)
    inner_batched = moved[0].dim() == 3
    if inner_batched:  # vmap over an already batched Data: fold both batch dims
        b2 = moved[0].shape[1]
        moved = [t.reshape(B * b2, *t.shape[2:]) for t in moved]
    out = run(model_id, mode, *moved)
    if inner_batched:
        out = out.reshape(B, b2, out.shape[-1])
    return out, 0
