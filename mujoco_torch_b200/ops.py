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

# `mode` argument of the op: kind (step / forward) in the low two bits, then the options below
MODE_STEP, MODE_FORWARD = 1, 2
MODE_STATE = 4        # outputs="state": the model's state layout (layout.py) -- only the state fields are produced
MODE_FIXED = 8        # fixed_iterations=True (reference forward.step :463)
MODE_UPTO_SHIFT = 4   # bits 4-5: stop a forward pass after k_smooth (1) / k_collision (2) / k_constraint (3); 0 = all stages
FLAG_FIXED, FLAG_UPTO = 1, {1: 2, 2: 4, 3: 8}   # include/mjt.h MJT_FLAG_*


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
    for fn in ("mjt_model_data_words", "mjt_model_scratch_words", "mjt_model_dtype_size", "mjt_step_launches", "mjt_model_layout"):
        getattr(L, fn).argtypes = [ctypes.c_void_p]
        getattr(L, fn).restype = ctypes.c_int
    L.mjt_model_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.mjt_model_field.restype = ctypes.c_int
    L.mjt_step.argtypes = [ctypes.c_void_p, ctypes.POINTER(StateIn), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.mjt_step.restype = ctypes.c_int
    L.mjt_forward.argtypes = [ctypes.c_void_p, ctypes.POINTER(StateIn), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
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


EXPORTS = ("mjt_version", "mjt_last_error", "mjt_model_create", "mjt_model_destroy", "mjt_model_layout", "mjt_model_data_words",
           "mjt_model_scratch_words", "mjt_model_dtype_size", "mjt_model_field", "mjt_step", "mjt_forward", "mjt_step_launches",
           "mjt_profile_begin", "mjt_profile_end", "mjt_solve_config", "mjt_fma_peak", "mjt_env_obs_dim", "mjt_env_post")


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


def model_handle(model, layout: str = "full"):
    """Device handle of a Model for one output layout ("full" / "state"), created lazily on the model's CUDA device."""
    h = model._handles.get(layout)
    if h is not None:
        return h
    if model.device.type != "cuda":
        raise RuntimeError("mujoco_torch_b200 runs on CUDA only: move the model with .to('cuda') (there is no CPU path)")
    L = lib()
    blob = model.blob_np if layout == "full" else model.state_layout["blob_np"]
    h = ctypes.c_void_p()
    dev = model.device.index if model.device.index is not None else torch.cuda.current_device()
    _check(L.mjt_model_create(blob.ctypes.data_as(ctypes.c_void_p), blob.nbytes, dev, ctypes.byref(h)))
    model._handles[layout] = h
    if layout == "full":
        model._handle = h
    weakref.finalize(model, L.mjt_model_destroy, h)
    assert L.mjt_model_layout(h) == (0 if layout == "full" else 1)
    assert L.mjt_model_data_words(h) == (model.data_words if layout == "full" else model.state_layout["pub_words"])
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


def _launch(model_id: int, mode: int, tensors) -> torch.Tensor:
    model = _MODELS.get(model_id)
    if model is None:
        raise RuntimeError("mjt: unknown model id (model was garbage collected)")
    qpos = tensors[0]
    if not qpos.is_cuda:
        raise RuntimeError("mjt::step is CUDA-only; got a CPU tensor (no CPU fallback exists)")
    if qpos.device != model.device and not (model.device.index is None and model.device.type == "cuda"):
        raise RuntimeError(f"mjt: Data is on {qpos.device} but the Model is on {model.device} (move one of them with .to())")
    batched = qpos.dim() == 2
    B = qpos.shape[0] if batched else 1
    layout = "state" if mode & MODE_STATE else "full"
    h = model_handle(model, layout)
    L = lib()
    keep = []
    st = StateIn()
    for name, t in zip(_STATE_ARGS, tensors):
        want = torch.int32 if name == "eq_active" else model.dtype
        e = _env_minor(t, batched, want)   # no copy when the tensor already is an env-minor view (every output of a step is)
        keep.append(e)
        setattr(st, name, e.data_ptr() if e.numel() else None)
    st.qacc = None
    flags = FLAG_FIXED if mode & MODE_FIXED else 0
    with torch.cuda.device(qpos.device):
        out = torch.empty((L.mjt_model_data_words(h), B), dtype=model.dtype, device=qpos.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        # the work buffer (state layout: the workspace of derived quantities; RK4: stage buffers) goes back to the caching
        # allocator when this call returns: stream-ordered reuse, nothing outlives the step
        sw = L.mjt_model_scratch_words(h)
        work = torch.empty((sw, B), dtype=model.dtype, device=qpos.device) if sw else None
        wp = work.data_ptr() if sw else None
        if mode & 3 == MODE_STEP:
            _check(L.mjt_step(h, ctypes.byref(st), out.data_ptr(), wp, B, flags, stream))
        else:
            upto = (mode >> MODE_UPTO_SHIFT) & 3
            _check(L.mjt_forward(h, ctypes.byref(st), out.data_ptr(), wp, B, flags | FLAG_UPTO.get(upto, 0), stream))
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
    words = model.state_layout["pub_words"] if mode & MODE_STATE else model.data_words
    if qpos.dim() == 2:
        return qpos.new_empty((words, qpos.shape[0]), dtype=model.dtype).t()
    return qpos.new_empty((words,), dtype=model.dtype)


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
