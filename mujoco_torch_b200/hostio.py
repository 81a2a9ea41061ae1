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

# The state a step advances, in the order the state layout stores it (layout.py STATE_PUBLIC), then the per-step input.
STATE_FIELDS = ("qpos", "qvel", "act", "time", "qacc_warmstart")
IN_FIELDS = STATE_FIELDS + ("ctrl",)
OUT_FIELDS = STATE_FIELDS


def _row_blocks(tensors, B):
    """Group consecutive ``(B, n)`` env-minor views that sit back to back in one buffer into single ``(rows, B)`` blocks, so that
    they travel with one copy.  Yields (block tensor, first row, rows); a tensor that is not an env-minor view is its own block
    (copied through a transposed view)."""
    out, row = [], 0
    for t in tensors:
        n = t.shape[1]
        minor = t.numel() == 0 or (t.stride(0) == 1 and (n == 1 or t.stride(1) == B))
        if out and minor and out[-1][3] and t.numel() and t.data_ptr() == out[-1][0].data_ptr() + out[-1][2] * B * t.element_size():
            first, r0, rows, _ = out[-1]
            out[-1] = (first, r0, rows + n, True)
        elif t.numel():
            out.append((t, row, n, minor))
        row += n
    blocks = []
    for first, r0, rows, minor in out:
        blk = torch.as_strided(first, (rows, B), (B, 1), first.storage_offset()) if minor else first.t()
        blocks.append((blk, r0, rows))
    return blocks


class HostPipeline:
    """``pipe = HostPipeline(mx, d0)`` (``nchunk`` defaults to 2 for batches of 16 384 envs and more, else 1); ``pipe.step()``
    advances the host-resident state by one timestep.

    * Per chunk ``c`` one pinned ``(rows, B_c)`` array (env-minor, the layout the kernels read) holds
      ``qpos | qvel | act | time | qacc_warmstart | ctrl``; ``pipe.host[c][name]`` are its row views -- the caller writes ``ctrl``
      (and may rewrite any state field) there between steps.
    * ``resident="device"`` keeps the state on the GPU (carried from step to step): only ``ctrl`` comes from the host array and only
      ``qpos | qvel`` go back to it -- what a host-side policy driving a device-resident simulation moves.
    * ``step()`` sends the whole array to the device with ONE copy, steps the chunk (``outputs="state"`` by default: nothing
      but the state is materialised; the state fields of the result are contiguous) and brings the new state back into the same
      rows with ONE copy.  It returns after all streams have drained (``synchronize=False``: only enqueues).
    """

    def __init__(self, m: Model, d: Data, nchunk: int | None = None, outputs: str = "state", resident: str = "host"):
        if d.qpos.dim() != 2:
            raise ValueError("HostPipeline expects a batched Data with one leading batch dim")
        if not d.qpos.is_cuda:
            raise RuntimeError("HostPipeline needs the template Data on the model's CUDA device (there is no CPU path)")
        if resident not in ("host", "device"):
            raise ValueError("resident must be 'host' or 'device'")
        self.m, self.outputs, self.resident = m, outputs, resident
        # resident="host": the whole state travels both ways every step.  resident="device": the state stays on the GPU and is
        # carried from step to step; only the per-step input (ctrl) comes from the host and only the observation (qpos, qvel)
        # goes back -- the traffic of a host-side policy driving a device-resident simulation.
        self.in_fields = IN_FIELDS if resident == "host" else ("ctrl",)
        self.out_fields = OUT_FIELDS if resident == "host" else ("qpos", "qvel")
        B = d.qpos.shape[0]
        if nchunk is None:   # measured (profiles/README.md): two chunks overlap PCIe with kernels once a chunk still fills the GPU
            nchunk = 2 if B >= 16384 else 1
        nchunk = max(1, min(nchunk, B))
        bounds = [(B * c) // nchunk for c in range(nchunk + 1)]
        self.slices = [slice(bounds[c], bounds[c + 1]) for c in range(nchunk)]
        self.streams = [torch.cuda.Stream(device=d.qpos.device) for _ in range(nchunk)]
        self.base = [d[s] for s in self.slices]       # per-chunk template (fields that never travel stay on the device)
        # host array rows: the fields that come back first (so that a field travelling both ways uses the same rows), then the rest
        order = list(self.out_fields) + [k for k in self.in_fields if k not in self.out_fields]
        widths = {k: (getattr(d, k).numel() // B) for k in order}
        self.rows, r = {}, 0
        for k in order:
            self.rows[k] = (r, r + widths[k])
            r += widths[k]
        self.out_rows = self.rows[self.out_fields[-1]][1]
        self.in_lo = min(self.rows[k][0] for k in self.in_fields)   # the H2D copy covers rows [in_lo, end)
        self.in_rows = r - self.in_lo
        self.host_buf, self.host, self.dev_in = [], [], []
        for s in self.slices:
            n = s.stop - s.start
            buf = torch.empty((r, n), dtype=m.dtype).pin_memory()
            views = {}
            for k in order:
                a, b = self.rows[k]
                v = getattr(d, k)[s]
                buf[a:b].copy_(v.reshape(n, -1).t())
                views[k] = buf[a:b]                    # (n_k, B_c)
            self.host_buf.append(buf)
            self.host.append(views)
            self.dev_in.append(torch.empty((self.in_rows, n), dtype=m.dtype, device=d.qpos.device))
        self._primed = False
        item = torch.empty((), dtype=m.dtype).element_size()
        self.h2d_bytes = self.in_rows * B * item
        self.d2h_bytes = self.out_rows * B * item

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
                din, n = self.dev_in[c], self.dev_in[c].shape[1]
                din.copy_(self.host_buf[c][self.in_lo:], non_blocking=True)         # one H2D: state + ctrl (or ctrl alone)
                tmpl = self.base[c]
                dev = {}
                for k in self.in_fields:
                    a, b = self.rows[k]
                    dev[k] = din[a - self.in_lo:b - self.in_lo].t().reshape(getattr(tmpl, k).shape)   # env-minor views: no transposes before the launch
                out = api.step(self.m, tmpl.replace(**dev), outputs=self.outputs)
                if self.resident == "device":   # the advanced state is the next step's template (views of the step's small output buffer)
                    self.base[c] = tmpl.replace(**{k: getattr(out, k) for k in STATE_FIELDS})
                outs = [getattr(out, k).reshape(n, -1) for k in self.out_fields]
                for blk, r0, rows in _row_blocks(outs, n):                          # state layout: ONE block -> one D2H
                    self.host_buf[c][r0:r0 + rows].copy_(blk, non_blocking=True)
        if synchronize:
            self.synchronize()

    def synchronize(self):
        for st in self.streams:
            st.synchronize()

    def gather(self, name: str) -> torch.Tensor:
        """Host state field as one ``(B, n)`` tensor (a copy)."""
        return torch.cat([b[name].t() for b in self.host], dim=0)
