"""CUDA-graph replay of the physics step for launch-bound batch sizes.

The reference reaches its small-batch numbers by handing ``torch.compile`` a whole ``vmap(step)`` (README.md:178-247); below a
few thousand envs its step is bound by the number of kernels it launches.  Here a step already is four launches, and what is
left at small batches is host time per launch sequence: the torch op dispatch, two allocations and the ctypes call.  A
``GraphedStep`` captures ``nsteps`` consecutive ``step()`` calls -- the same Python path, the same kernels -- once into a CUDA
graph and replays them with one ``cudaGraphLaunch``; the state lives in a static env-minor buffer that the last captured node
writes back, so a rollout is ``replay, replay, ...`` with only ``ctrl`` travelling in.

    g = mujoco_torch.GraphedStep(mx, d, nsteps=5)          # e.g. one agent step = FRAME_SKIP physics steps
    for t in range(T):
        d = g(ctrl[t])                                      # Data whose state fields view the static buffer

Bit-identical to calling ``step`` in a loop (tests/test_gpu_api.py::test_graphed_step_equals_the_step_loop).
"""

from __future__ import annotations

import numpy as np
import torch

from . import api
from .data import Data
from .model import Model

# the part of the public state buffer a step reads back in (layout.py P_FIELDS order: these come first and are contiguous)
_CARRIED = ("qpos", "qvel", "act", "time", "qacc_warmstart")


class GraphedStep:
    """``nsteps`` physics steps per call, replayed from one CUDA graph on a fixed batch shape.

    m, d:        Model and a batched Data on the model's device; ``d`` provides the initial state and every input that
                 stays constant over the rollout (qfrc_applied, xfrc_applied, mocap_*, eq_active).
    nsteps:      physics steps per replay; ``ctrl`` is held over them (the reference env layer's FRAME_SKIP, zoo/base.py:266-272).
    outputs:     "state" (default) or "full", as in ``step``; with "full" the returned Data also carries the derived fields of
                 the last step.
    """

    def __init__(self, m: Model, d: Data, nsteps: int = 1, *, outputs: str = "state", fixed_iterations: bool = False, warmup: int = 3):
        if not isinstance(m, Model):
            raise TypeError("first argument must be a mujoco_torch_b200 Model (see device_put)")
        if d.qpos.dim() != 2:
            raise ValueError("GraphedStep needs a Data with one batch dimension")
        if nsteps < 1:
            raise ValueError("nsteps must be >= 1")
        if not d.qpos.is_cuda:
            raise RuntimeError("GraphedStep is CUDA-only (no CPU fallback exists)")
        self.m, self.nsteps, self.outputs, self.fixed = m, int(nsteps), outputs, bool(fixed_iterations)
        dev, B = d.qpos.device, d.qpos.shape[0]
        sizes = [int(np.prod(m.field_shape[k])) if m.field_shape[k] else 1 for k in _CARRIED]
        self._ncarry = sum(sizes)
        # static input: the carried state as views of one env-minor block, ctrl next to it
        self._state = torch.empty((self._ncarry, B), dtype=m.dtype, device=dev)
        self._ctrl = torch.zeros((max(m.nu, 0), B), dtype=m.dtype, device=dev)
        fields, off = {}, 0
        for k, n in zip(_CARRIED, sizes):
            v = self._state[off:off + n].t()
            fields[k] = v[:, 0] if k == "time" else v
            off += n
        fields["ctrl"] = self._ctrl.t()
        self._in = d.replace(**fields)
        self.reset(d)
        # warm-up on a side stream (allocator and lazy module loading settle before capture), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        self.reset(d)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=s):
            self._out = self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        self.reset(d)

    def _body(self) -> Data:
        x = self._in
        for _ in range(self.nsteps):
            x = api.step(self.m, x, fixed_iterations=self.fixed, outputs=self.outputs)
            x = x.replace(ctrl=self._in.ctrl)
        # write the carried state back into the static block: one copy when the step's public buffer holds it contiguously
        B, item = self._state.shape[1], self._state.element_size()
        ptr, packed = x.qpos.data_ptr(), x.qpos.stride() == (1, B)
        for k in _CARRIED:
            v = getattr(x, k)
            n = 1 if k == "time" else v.shape[-1]
            packed = packed and (n == 0 or (v.data_ptr() == ptr and v.stride()[0] == 1))
            ptr += n * B * item
        if packed:
            self._state.copy_(torch.as_strided(x.qpos, (self._ncarry, B), (B, 1)))
        else:
            off = 0
            for k in _CARRIED:
                v = getattr(x, k)
                v = v.unsqueeze(-1) if k == "time" else v
                n = v.shape[-1]
                if n:
                    self._state[off:off + n].copy_(v.t())
                off += n
        return x

    def reset(self, d: Data) -> None:
        """Load the carried state (qpos, qvel, act, time, qacc_warmstart) and ctrl of ``d`` into the static buffers."""
        off = 0
        for k in _CARRIED:
            v = getattr(d, k)
            v = v.unsqueeze(-1) if k == "time" else v
            n = v.shape[-1]
            if n:
                self._state[off:off + n].copy_(v.t())
            off += n
        if self._ctrl.numel():
            self._ctrl.copy_(d.ctrl.t())

    @property
    def state(self) -> Data:
        """The current state (views of the static block): what the next replay starts from."""
        return self._in

    def __call__(self, ctrl: torch.Tensor | None = None) -> Data:
        """Advance ``nsteps`` steps.  ``ctrl`` (B, nu), host or device, is copied into the static ctrl buffer first; None keeps
        the previous one.  The returned Data's fields are valid until the next call."""
        if ctrl is not None and self._ctrl.numel():
            self._ctrl.copy_(ctrl.t(), non_blocking=True)
        self._graph.replay()
        return self._out
