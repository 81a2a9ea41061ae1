"""Host-resident stepping: ``HostPipeline`` keeps the batch state in pinned host memory and runs
``host -> device copy, step(m, d), device -> host copy`` chunk by chunk on separate CUDA streams, so that one chunk's
PCIe transfers overlap another chunk's kernels.

This is the end-to-end form of :func:`mujoco_torch_b200.step` for callers whose simulation state lives on the host
(an external policy or logger that consumes ``qpos`` / ``qvel`` every step).  Environments are independent, so the
chunks are too; the result is identical to one ``step`` over the whole batch.
"""

from __future__ import annotations

import torch

from . import api
from .data import Data
from .model import Model

IN_FIELDS = ("qpos", "qvel", "qacc_warmstart", "ctrl")
OUT_FIELDS = ("qpos", "qvel", "qacc_warmstart")
# State that does not travel by default (activations, time) is carried on the device from step to step.
CARRIED_FIELDS = ("act", "time")


class HostPipeline:
    """``pipe = HostPipeline(mx, d0)`` (``nchunk`` defaults to 2 for batches of 16 384 envs and more, else 1); ``pipe.step()`` advances the host-resident state by one timestep.

    * ``pipe.host[c][name]`` -- pinned ``(n, B_c)`` tensors (env-minor, the layout the kernels read) per chunk ``c``;
      the caller writes ``ctrl`` (and may rewrite any state field) there between steps.
    * ``step()`` copies the inputs of every chunk to the device, steps it (``outputs="state"`` by default: nothing but the
      state is materialised), and copies ``qpos, qvel, qacc_warmstart`` back into the same pinned buffers.  It returns after
      all streams have drained.  Activations and time stay on the device and are carried from step to step.
    """

    def __init__(self, m: Model, d: Data, nchunk: int | None = None, in_fields=IN_FIELDS, out_fields=OUT_FIELDS, outputs: str = "state"):
        if d.qpos.dim() != 2:
            raise ValueError("HostPipeline expects a batched Data with one leading batch dim")
        if not d.qpos.is_cuda:
            raise RuntimeError("HostPipeline needs the template Data on the model's CUDA device (there is no CPU path)")
        self.m, self.in_fields, self.out_fields, self.outputs = m, tuple(in_fields), tuple(out_fields), outputs
        # every state field the step advances is either sent from the host each step or kept current on the device
        self.carried = tuple(k for k in CARRIED_FIELDS if k not in self.in_fields and getattr(d, k).numel() > 0)
        B = d.qpos.shape[0]
        if nchunk is None:   # measured (profiles/README.md): two chunks overlap PCIe with kernels once a chunk still fills the GPU
            nchunk = 2 if B >= 16384 else 1
        nchunk = max(1, min(nchunk, B))
        bounds = [(B * c) // nchunk for c in range(nchunk + 1)]
        self.slices = [slice(bounds[c], bounds[c + 1]) for c in range(nchunk)]
        self.streams = [torch.cuda.Stream(device=d.qpos.device) for _ in range(nchunk)]
        self.base = [d[s] for s in self.slices]       # per-chunk template (fields that never travel stay on the device)
        self.host = []
        for s in self.slices:
            bufs = {}
            for k in set(self.in_fields) | set(self.out_fields):
                v = getattr(d, k)[s]
                bufs[k] = v.reshape(v.shape[0], -1).t().contiguous().cpu().pin_memory()   # (n, B_c)
            self.host.append(bufs)
        self._primed = False
        self.h2d_bytes = sum(b[k].numel() * b[k].element_size() for b in self.host for k in self.in_fields)
        self.d2h_bytes = sum(b[k].numel() * b[k].element_size() for b in self.host for k in self.out_fields)

    def step(self, synchronize: bool = True):
        """One timestep for every chunk.  With ``synchronize=False`` the call only enqueues: each chunk's stream carries
        its own ``H2D -> step -> D2H`` chain in order, so chunk A's next step starts as soon as its own results are back
        on the host, while chunk B is still computing (call :meth:`synchronize` before reading ``host``)."""
        if not self._primed:
            cur = torch.cuda.current_stream()
            for st in self.streams:
                st.wait_stream(cur)
            self._primed = True
        for c, st in enumerate(self.streams):
            with torch.cuda.stream(st):
                dev = {k: self.host[c][k].to(self.base[c].qpos.device, non_blocking=True).t() for k in self.in_fields}
                out = api.step(self.m, self.base[c].replace(**dev), outputs=self.outputs)
                if self.carried:
                    self.base[c] = self.base[c].replace(**{k: getattr(out, k) for k in self.carried})
                for k in self.out_fields:
                    v = getattr(out, k)
                    self.host[c][k].copy_(v.reshape(v.shape[0], -1).t(), non_blocking=True)
        if synchronize:
            self.synchronize()

    def synchronize(self):
        for st in self.streams:
            st.synchronize()

    def gather(self, name: str) -> torch.Tensor:
        """Host state field as one ``(B, n)`` tensor (a copy)."""
        return torch.cat([b[name].t() for b in self.host], dim=0)
